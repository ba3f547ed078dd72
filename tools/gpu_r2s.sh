#!/bin/bash
# c3 (100 KB bodies) and c5 (mixed sizes / key sizes) on NG GPUs: torchrun bench, value + e2e
NG=${NG:-2}
mkdir -p gpurun_out
for W in c3 c5; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $NG --workload $W --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/r2s_${W}_n$NG.json 2> gpurun_out/r2s_${W}_n$NG.err; echo "bench $W n$NG rc=$?"
  python - $W $NG <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(f'gpurun_out/r2s_{sys.argv[1]}_n{sys.argv[2]}.json') if l.startswith('{')][-1])
    print(sys.argv[1], "N=%d value %.4g (%.3f ms) from_raw %s e2e %.4g e2e_reg %s" % (d["n_gpus"], d["value"], d["ms_per_step"], (d.get("value_from_raw") or {}).get("value"), d["e2e"]["value"], (d.get("e2e_registered") or {}).get("value")), d.get("e2e_scaling",{}).get("host_threads_per_rank"))
except Exception as e: print("no line", e)
PY
done
