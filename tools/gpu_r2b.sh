#!/bin/bash
# round 2, second GPU pass: tests, bench (with the warp front end), launch list, ncu captures of the front-end kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2b_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 --profile > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2b_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2b_bench.json') if l.startswith('{')][-1])
print("value %.4g from_raw %.4g e2e %.4g e2e_reg %.4g regex %.4g" % (d["value"], d["value_from_raw"]["value"], d["e2e"]["value"], d["e2e_registered"]["value"], d["with_regex"]["value"]))
print(d["value_from_raw"]["kernel_ms"])
PY
timeout 600 python tools/fuzz_frontend.py 20000 7 > gpurun_out/r2b_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -3 gpurun_out/r2b_fuzz.log
CMD="python bench.py --emails 262144 --steps 2 --warmup 3 --skip-cpu-baseline"
$CMD > gpurun_out/r2b_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2b.csv $CMD > gpurun_out/r2b_ncu1.log 2>&1; echo "ncu list rc=$?"
$CMD > gpurun_out/r2b_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'frontend_warp|canon_body' -s 4 -c 4 -o gpurun_out/prof_fe_r2b $CMD > gpurun_out/r2b_ncu2.log 2>&1; echo "ncu fe rc=$?"
