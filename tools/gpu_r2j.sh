#!/bin/bash
# ncu --set full of every kernel of one default step (262144 emails) for profiles/ncu_summary_r2.md and ncu_traffic_r2.json
mkdir -p gpurun_out
CMD="python bench.py --emails 262144 --steps 1 --warmup 3 --skip-cpu-baseline"
$CMD > gpurun_out/r2j_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'rsa_verify|sha256_batch|assemble|frontend_warp|canon_body_staged|dfa_scan|bh_check' -s 40 -c 14 -o gpurun_out/prof_all_r2j $CMD > gpurun_out/r2j_ncu.log 2>&1; echo "ncu rc=$?"
