#!/bin/bash
# dev aid: e2e throughput for a few chunk sizes (registered inputs)
for C in "$@"; do
  ZKB_PROFILE=1 timeout 400 python bench.py --emails 1000000 --steps 4 --warmup 3 --skip-cpu-baseline --chunk $C > gpurun_out/sweep_$C.json 2> gpurun_out/sweep_$C.err
  python - "$C" <<'PY'
import json,sys
c=sys.argv[1]
d=json.loads(open(f"gpurun_out/sweep_{c}.json").read().strip().splitlines()[-1])
print("chunk",c,"value %.3g"%d["value"],"e2e %.4g"%d["e2e"]["value"],"ms %.1f"%d["e2e"]["ms_per_step"],"h2d",d["e2e"]["h2d_bytes_per_step"])
PY
  grep profile gpurun_out/sweep_$C.err | tail -1
done
