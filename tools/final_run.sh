#!/bin/bash
# Round-end validation on one B200: GPU tests, smoke, reference arm + default bench (the driver's command lines), fuzz, launch list.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final_smoke.log
timeout 900 python tools/fuzz_frontend.py 40000 31 > gpurun_out/final_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -1 gpurun_out/final_fuzz.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; echo "reference rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_c2.json 2> gpurun_out/final_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
r=json.loads([l for l in open('gpurun_out/final_reference.json') if l.startswith('{')][-1])
d=json.loads([l for l in open('gpurun_out/final_c2.json') if l.startswith('{')][-1])
print("reference %.4g (with regex %s)" % (r["value"], r.get("with_regex")))
print("value %.4g (%.2f ms) from_raw %.4g e2e %.4g e2e_reg %.4g regex %.4g regex_e2e %.4g warmup %d" % (d["value"], d["ms_per_step"], d["value_from_raw"]["value"], d["e2e"]["value"], d["e2e_registered"]["value"], d["with_regex"]["value"], d["with_regex"]["e2e"]["value"], d["warmup"]))
print(d["value_from_raw"]["kernel_ms"]); print("config equal:", r["config"] == d["config"])
PY
CMD="python bench.py --emails 262144 --steps 2 --warmup 3 --skip-cpu-baseline"
$CMD > gpurun_out/final_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2_final.csv $CMD > gpurun_out/final_ncu1.log 2>&1; echo "ncu list rc=$?"
$CMD > gpurun_out/final_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'rsa_verify|sha256_batch|assemble|frontend_warp|canon_body_staged|dfa_scan|bh_check' -s 40 -c 14 -o gpurun_out/prof_all_final $CMD > gpurun_out/final_ncu2.log 2>&1; echo "ncu all rc=$?"
