# usage: bash tools/sweep_knobs.sh ["bias stride exit" ...]  — mean kernel time of frames 1..5 per workload and knob set
if [ $# -eq 0 ]; then set -- "0 2 4" "-1 2 4" "-1 1 4" "-1 2 8" "-1 1 8" "-1 2 16" "-1 1 16" "-2 1 8" "-1 2 32"; fi
for cfg in "$@"; do
  set -- $cfg
  echo "== LMIN_BIAS=$1 LSTRIDE=$2 CELL_EXIT=$3"
  for wl in flythrough4k ortho4k spherical1080; do
    HMRM_LMIN_BIAS=$1 HMRM_LSTRIDE=$2 HMRM_CELL_EXIT=$3 python tools/profile_frame.py --frames 6 --workload $wl | awk -v w=$wl 'NR>1 {s+=$4; n++} END {printf "   %-14s %.3f ms\n", w, s/n}'
  done
done
