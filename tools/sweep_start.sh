for st in ${@:-6 8 10 12 99}; do
  echo "== LSTART=$st"
  for wl in flythrough4k ortho4k spherical1080; do
    HMRM_LSTART=$st python tools/profile_frame.py --frames 6 --workload $wl | awk -v w=$wl 'NR>1 {s+=$4; n++} END {printf "   %-14s %.3f ms\n", w, s/n}'
  done
done
