for c in ${@:-1 2 4 8 16 1e30}; do
  echo "== HMRM_CLIMB=$c"
  for wl in flythrough4k ortho4k spherical1080; do
    HMRM_CLIMB=$c python tools/profile_frame.py --frames 8 --workload $wl | awk -v w=$wl 'NR>1 {s+=$4; n++} END {printf "   %-14s %.3f ms\n", w, s/n}'
  done
done
