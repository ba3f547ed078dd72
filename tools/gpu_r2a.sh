#!/bin/bash
# round 2, first GPU pass: GPU tests, smoke, default bench, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_smi.txt; nproc >> gpurun_out/r2a_smi.txt; free -g >> gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2a_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2a_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 --profile > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2a_bench.err
