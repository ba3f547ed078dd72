#!/bin/bash
# Round-end validation on one B200: GPU tests, smoke, default bench (+ reference arm), other workloads, ncu launch list.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; echo rc=$?
python bench.py --steps 5 --warmup 3 > gpurun_out/final_c2.json 2> gpurun_out/final_c2.err; echo rc=$?
python bench.py --steps 5 --warmup 3 --no-direct --skip-cpu-baseline > gpurun_out/final_c2_pageable.json 2> gpurun_out/final_c2_pageable.err; echo rc=$?
python bench.py --workload c4 --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/final_c4.json 2> gpurun_out/final_c4.err; echo rc=$?
python bench.py --workload c3 --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/final_c3.json 2> gpurun_out/final_c3.err; echo rc=$?
python bench.py --workload c5 --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/final_c5.json 2> gpurun_out/final_c5.err; echo rc=$?
python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/pre_ncu.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/ncu_c.log 2>&1; echo ncu rc=$?
for f in final_reference final_c2 final_c2_pageable final_c4 final_c3 final_c5; do python - "$f" <<'PY'
import json, sys
f = sys.argv[1]
for line in open(f"gpurun_out/{f}.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print(f, "value %.3g" % d["value"], "ms/step %.2f" % d.get("ms_per_step", 0), "e2e", d.get("e2e", {}).get("value"), d.get("kernel_ms"))
PY
done
