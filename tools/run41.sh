#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload c4 --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/c4_flat.json 2> gpurun_out/c4_flat.err; echo rc=$?
ZKB_PROFILE=1 python bench.py --workload c3 --steps 3 --warmup 3 --skip-cpu-baseline > gpurun_out/c3_big.json 2> gpurun_out/c3_big.err; echo rc=$?
grep "zkb profile" gpurun_out/c3_big.err | tail -2
for f in c4_flat c3_big; do python - $f <<'PY'
import json, sys
for line in open("gpurun_out/%s.json" % sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line); print(sys.argv[1], "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], d["kernel_ms"])
PY
done
