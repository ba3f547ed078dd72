#!/bin/bash
mkdir -p gpurun_out
for W in c3 c4 c5; do
  timeout 900 python bench.py --workload $W --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/r2h_$W.json 2> gpurun_out/r2h_$W.err; echo "bench $W rc=$?"; tail -c 300 gpurun_out/r2h_$W.err
  python - $W <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(f'gpurun_out/r2h_{sys.argv[1]}.json') if l.startswith('{')][-1])
    print(sys.argv[1], "value %.4g from_raw %.4g e2e %.4g e2e_reg %.4g" % (d["value"], d["value_from_raw"]["value"], d["e2e"]["value"], d["e2e_registered"]["value"]), d["kernel_ms"], d["value_from_raw"]["kernel_ms"])
except Exception as e: print("no line", e)
PY
done
