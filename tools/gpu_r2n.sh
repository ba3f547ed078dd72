#!/bin/bash
# A/B of the squaring kernel's code-shape variants (ZKB_SQR_VARIANT; needs a library built with
# make -C zkemail.rs_b200/csrc EXTRA=-DZKB_SQR_EXPERIMENTS) against the plain kernel (variant 0 = --no-sqr)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "squaring or rsa" > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2n_tests.log
for V in 0 4 8 104 108 1004 1008 1108 2008; do
  F=""; [ $V = 0 ] && F="--no-sqr"
  ZKB_SQR_VARIANT=$V timeout 600 python bench.py --emails 524288 --steps 5 --warmup 3 --skip-cpu-baseline --skip-extras $F > gpurun_out/r2n_$V.json 2> gpurun_out/r2n_$V.err; rc=$?
  python - $V $rc <<'PY'
import json,sys
d=json.loads([l for l in open(f'gpurun_out/r2n_{sys.argv[1]}.json') if l.startswith('{')][-1])
print("variant", sys.argv[1], "rc", sys.argv[2], "value %.4g ms/step %.3f" % (d["value"], d["ms_per_step"]), "rsa %.3f sha %.3f" % (d["kernel_ms"]["rsa"], d["kernel_ms"]["sha256"]))
PY
done
CMD="python bench.py --emails 131072 --steps 2 --warmup 3 --skip-cpu-baseline --skip-extras"
ZKB_SQR_VARIANT=8 ncu --set full --clock-control none --import-source on -k regex:'rsa_verify' -s 2 -c 1 -o gpurun_out/prof_rsa_r2n_8 $CMD > gpurun_out/r2n_ncu8.log 2>&1; echo "ncu rc=$?"
ZKB_SQR_VARIANT=1008 ncu --set full --clock-control none --import-source on -k regex:'rsa_verify' -s 2 -c 1 -o gpurun_out/prof_rsa_r2n_1008 $CMD > gpurun_out/r2n_ncu1008.log 2>&1; echo "ncu rc=$?"
