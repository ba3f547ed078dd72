#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2f_tests.log
timeout 600 python tools/fuzz_frontend.py 20000 21 > gpurun_out/r2f_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -2 gpurun_out/r2f_fuzz.log
timeout 900 python bench.py --steps 5 --warmup 3 --profile > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2f_bench.json') if l.startswith('{')][-1])
print("value %.4g from_raw %.4g e2e %.4g e2e_reg %.4g regex %.4g regex_e2e %.4g" % (d["value"], d["value_from_raw"]["value"], d["e2e"]["value"], d["e2e_registered"]["value"], d["with_regex"]["value"], d["with_regex"]["e2e"]["value"]))
print(d["value_from_raw"]["kernel_ms"])
PY
grep profile gpurun_out/r2f_bench.err | tail -14 | cut -c1-260
CMD="python bench.py --emails 262144 --steps 2 --warmup 3 --skip-cpu-baseline"
$CMD > gpurun_out/r2f_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2f.csv $CMD > gpurun_out/r2f_ncu1.log 2>&1; echo "ncu list rc=$?"
$CMD > gpurun_out/r2f_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'canon_body' -s 2 -c 1 -o gpurun_out/prof_canon_r2f $CMD > gpurun_out/r2f_ncu2.log 2>&1; echo "ncu canon rc=$?"
