#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2i_tests.log
for V in sqr nosqr; do
  F=""; [ $V = nosqr ] && F="--no-sqr"
  timeout 600 python bench.py --emails 524288 --steps 5 --warmup 3 --skip-cpu-baseline --skip-extras $F > gpurun_out/r2i_$V.json 2> gpurun_out/r2i_$V.err; echo "bench $V rc=$?"
  python - $V <<'PY'
import json,sys
d=json.loads([l for l in open(f'gpurun_out/r2i_{sys.argv[1]}.json') if l.startswith('{')][-1])
print(sys.argv[1], "value %.4g ms/step %.3f" % (d["value"], d["ms_per_step"]), "rsa %.3f sha %.3f" % (d["kernel_ms"]["rsa"], d["kernel_ms"]["sha256"]), "frac %.3f util %.3f" % (d["roofline"]["frac"], d["roofline"]["pipe_utilisation"]))
PY
done
timeout 600 python bench.py --workload c5 --steps 3 --warmup 3 --skip-cpu-baseline > gpurun_out/r2i_c5.json 2> gpurun_out/r2i_c5.err; echo "c5 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2i_c5.json') if l.startswith('{')][-1])
print("c5 value %.4g from_raw %.4g" % (d["value"], d["value_from_raw"]["value"]), d["value_from_raw"]["kernel_ms"])
PY
CMD="python bench.py --emails 131072 --steps 2 --warmup 3 --skip-cpu-baseline --skip-extras"
$CMD > gpurun_out/r2i_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'rsa_verify' -s 2 -c 1 -o gpurun_out/prof_rsa_r2i $CMD > gpurun_out/r2i_ncu.log 2>&1; echo "ncu rc=$?"
