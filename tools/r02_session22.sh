#!/bin/bash
# Round-2 GPU session 22 (1 GPU): GPU suite after the host-side changes (grammar, image ingest), smoke(), a short bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s22_pytest.txt 2>&1; echo "pytest exit $?"; tail -n 2 gpurun_out/s22_pytest.txt
timeout 300 python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 600 python bench.py --steps 60 --warmup 3 > gpurun_out/s22_bench.json 2> gpurun_out/s22_bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/s22_bench.json
