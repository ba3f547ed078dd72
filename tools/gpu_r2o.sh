#!/bin/bash
# two GPUs: the in-process multi-device engine + the process-per-GPU bench with the library's NCCL all-gather
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2o_smi.txt; nproc >> gpurun_out/r2o_smi.txt
timeout 900 python -m pytest tests/test_multi.py -m gpu -x -q > gpurun_out/r2o_tests.log 2>&1; echo "multi tests rc=$?"; tail -6 gpurun_out/r2o_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus ${NG:-2} --steps 5 --warmup 3 > gpurun_out/r2o_n2.json 2> gpurun_out/r2o_n2.err; echo "bench n2 rc=$?"; tail -c 400 gpurun_out/r2o_n2.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2o_n2.json') if l.startswith('{')][-1])
    print("N=2 value %.4g from_raw %.4g e2e %.4g e2e_reg %.4g regex %.4g" % (d["value"], d["value_from_raw"]["value"], d["e2e"]["value"], d["e2e_registered"]["value"], d["with_regex"]["value"]), d["ms_per_step"])
except Exception as e: print("no line", e)
PY
timeout 600 python tools/multi_bench.py ${NG:-2} 1000000 > gpurun_out/r2o_multi.json 2> gpurun_out/r2o_multi.err; echo "multi bench rc=$?"; cat gpurun_out/r2o_multi.json | tail -3
