#!/bin/bash
# 8 GPUs: torchrun bench (value + e2e) with the overlapped record exchange in every step
NG=${NG:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2q_smi.txt; nproc >> gpurun_out/r2q_smi.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $NG --steps 10 --warmup 3 --skip-extras --skip-cpu-baseline > gpurun_out/r2q_n.json 2> gpurun_out/r2q_n.err; echo "bench n$NG rc=$?"; tail -c 500 gpurun_out/r2q_n.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2q_n.json') if l.startswith('{')][-1])
    print("N=%d value %.4g (%.3f ms) e2e %.4g" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e: print("no line", e)
PY
