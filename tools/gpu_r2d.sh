#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2d_tests.log
for V in sqr nosqr; do
  F=""; [ $V = nosqr ] && F="--no-sqr"
  timeout 600 python bench.py --emails 524288 --steps 5 --warmup 3 --skip-cpu-baseline --skip-extras $F > gpurun_out/r2d_$V.json 2> gpurun_out/r2d_$V.err; echo "bench $V rc=$?"
  python - $V <<'PY'
import json,sys
d=json.loads([l for l in open(f'gpurun_out/r2d_{sys.argv[1]}.json') if l.startswith('{')][-1])
print(sys.argv[1], "value %.4g ms/step %.3f" % (d["value"], d["ms_per_step"]), d["kernel_ms"], "frac %.3f util %.3f" % (d["roofline"]["frac"], d["roofline"]["pipe_utilisation"]))
PY
done
CMD="python bench.py --emails 131072 --steps 2 --warmup 3 --skip-cpu-baseline --skip-extras"
$CMD > gpurun_out/r2d_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'rsa_verify' -s 2 -c 1 -o gpurun_out/prof_rsa_r2d $CMD > gpurun_out/r2d_ncu.log 2>&1; echo "ncu rc=$?"
