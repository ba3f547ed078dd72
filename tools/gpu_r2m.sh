#!/bin/bash
# c3 (100 k x 100 KB bodies): where does the lane-per-message SHA-256 lose its pipe rate?
mkdir -p gpurun_out
CMD="python bench.py --workload c3 --steps 3 --warmup 3 --skip-cpu-baseline --skip-extras"
$CMD > gpurun_out/r2m_c3.json 2> gpurun_out/r2m_c3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2m_c3.json') if l.startswith('{')][-1])
print("c3 value %.4g ms/step %.3f" % (d["value"], d["ms_per_step"]), d["kernel_ms"])
PY
ncu --set full --clock-control none --import-source on -k regex:'sha256_batch' -s 4 -c 1 -o gpurun_out/prof_sha_c3 $CMD > gpurun_out/r2m_ncu.log 2>&1; echo "ncu rc=$?"
