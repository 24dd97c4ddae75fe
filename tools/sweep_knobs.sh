for cfg in "0 2 4" "-1 2 4" "-2 2 4" "-1 1 4" "0 1 4" "-1 2 2" "-1 2 8" "-2 1 4" "-1 3 4"; do
  set -- $cfg
  echo "== LMIN_BIAS=$1 LSTRIDE=$2 CELL_EXIT=$3"
  for wl in flythrough4k ortho4k spherical1080; do
    HMRM_LMIN_BIAS=$1 HMRM_LSTRIDE=$2 HMRM_CELL_EXIT=$3 python tools/profile_frame.py --frames 3 --workload $wl | awk -v w=$wl '{s+=$4; n++} END {printf "   %-14s %.3f ms\n", w, s/n}'
  done
done
