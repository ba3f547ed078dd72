#!/bin/bash
# ncu --set full of the DFA scan launches of workload c4 (after the same command exited 0 without ncu)
python bench.py --workload c4 --emails 262144 --unique 65536 --steps 1 --warmup 3 --skip-cpu-baseline > gpurun_out/ncu_dfa_pre.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:dfa_scan_strided -s 3 -c 3 -o gpurun_out/dfa_r1c -f \
    python bench.py --workload c4 --emails 262144 --unique 65536 --steps 1 --warmup 3 --skip-cpu-baseline > gpurun_out/ncu_dfa.log 2>&1
echo ncu rc=$?
ncu -i gpurun_out/dfa_r1c.ncu-rep --page raw --csv > gpurun_out/dfa_r1c_raw.csv 2>/dev/null
ls -la gpurun_out/dfa_r1c*
