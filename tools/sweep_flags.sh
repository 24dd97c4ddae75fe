# usage: bash tools/sweep_flags.sh "<nvcc -D flags>" ...   — rebuilds libhmrm.so with each flag set and times three workloads
for cfg in "$@"; do
  echo "== $cfg"
  HMRM_NVCC_EXTRA="$cfg" python heightmap-ray-marcher_b200/build.py --force > /dev/null 2>&1 || echo build failed
  for wl in flythrough4k ortho4k spherical1080; do
    python tools/profile_frame.py --frames 8 --workload $wl | awk -v w=$wl 'NR>1 {s+=$4; n++} END {printf "   %-14s %.3f ms\n", w, s/n}'
  done
done
