#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "squaring or rsa" > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2l_tests.log
for V in sqr nosqr; do
  F=""; [ $V = nosqr ] && F="--no-sqr"
  timeout 600 python bench.py --emails 524288 --steps 5 --warmup 3 --skip-cpu-baseline --skip-extras $F > gpurun_out/r2l_$V.json 2> gpurun_out/r2l_$V.err; echo "bench $V rc=$?"
  python - $V <<'PY'
import json,sys
d=json.loads([l for l in open(f'gpurun_out/r2l_{sys.argv[1]}.json') if l.startswith('{')][-1])
print(sys.argv[1], "value %.4g ms/step %.3f" % (d["value"], d["ms_per_step"]), "rsa %.3f sha %.3f" % (d["kernel_ms"]["rsa"], d["kernel_ms"]["sha256"]), "frac %.3f util %.3f" % (d["roofline"]["frac"], d["roofline"]["pipe_utilisation"]))
PY
done
CMD="python bench.py --emails 131072 --steps 2 --warmup 3 --skip-cpu-baseline --skip-extras"
$CMD > gpurun_out/r2l_plain.log 2>&1 && ncu --set full --clock-control none -k regex:'rsa_verify' -s 2 -c 1 -o gpurun_out/prof_rsa_r2l $CMD > gpurun_out/r2l_ncu.log 2>&1; echo "ncu rc=$?"
