#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2c_tests.log
timeout 600 python tools/fuzz_frontend.py 20000 9 > gpurun_out/r2c_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -2 gpurun_out/r2c_fuzz.log
CMD="python bench.py --emails 262144 --steps 3 --warmup 3 --skip-cpu-baseline"
$CMD > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2c_bench.json') if l.startswith('{')][-1])
print("value %.4g from_raw %.4g e2e %.4g e2e_reg %.4g regex %.4g" % (d["value"], d["value_from_raw"]["value"], d["e2e"]["value"], d["e2e_registered"]["value"], d["with_regex"]["value"]))
print(d["value_from_raw"]["kernel_ms"])
PY
$CMD > gpurun_out/r2c_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'frontend_warp|canon_body' -s 4 -c 2 -o gpurun_out/prof_fe_r2c $CMD > gpurun_out/r2c_ncu2.log 2>&1; echo "ncu fe rc=$?"
