#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2k_tests.log
timeout 900 python tools/fuzz_frontend.py 40000 41 > gpurun_out/r2k_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -1 gpurun_out/r2k_fuzz.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2k_c2.json 2> gpurun_out/r2k_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2k_c2.json') if l.startswith('{')][-1])
print("value %.4g (%.2f ms) from_raw %.4g e2e %.4g e2e_reg %.4g regex %.4g regex_e2e %.4g" % (d["value"], d["ms_per_step"], d["value_from_raw"]["value"], d["e2e"]["value"], d["e2e_registered"]["value"], d["with_regex"]["value"], d["with_regex"]["e2e"]["value"]))
print(d["value_from_raw"]["kernel_ms"])
PY
