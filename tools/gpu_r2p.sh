#!/bin/bash
# N GPUs: communicator tests (run + overlapped exchange) and the torchrun bench with the exchange inside every step
NG=${NG:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi.py -m gpu -x -q > gpurun_out/r2p_tests.log 2>&1; echo "multi tests rc=$?"; tail -6 gpurun_out/r2p_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/r2p_n.json 2> gpurun_out/r2p_n.err; echo "bench n$NG rc=$?"; tail -c 600 gpurun_out/r2p_n.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2p_n.json') if l.startswith('{')][-1])
    print("N=%d value %.4g (%.3f ms) from_raw %.4g e2e %.4g e2e_reg %.4g regex %.4g" % (d["n_gpus"], d["value"], d["ms_per_step"], d["value_from_raw"]["value"], d["e2e"]["value"], d["e2e_registered"]["value"], d["with_regex"]["value"]))
except Exception as e: print("no line", e)
PY
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-extras > gpurun_out/r2p_1.json 2> gpurun_out/r2p_1.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2p_1.json') if l.startswith('{')][-1])
print("N=1 value %.4g (%.3f ms)" % (d["value"], d["ms_per_step"]))
PY
