#!/bin/bash
# A/B of SHA-256 with 0 .. 3 rotate families issued on the FMA pipe (library built with EXTRA=-DZKB_SHA_EXPERIMENTS)
mkdir -p gpurun_out
for R in 0 1 2 3; do
  ZKB_SHA_ROT=$R timeout 600 python bench.py --emails 524288 --steps 5 --warmup 3 --skip-cpu-baseline --skip-extras > gpurun_out/r2r_$R.json 2> gpurun_out/r2r_$R.err; rc=$?
  ZKB_SHA_ROT=$R timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 --skip-cpu-baseline --skip-extras > gpurun_out/r2r_c3_$R.json 2> gpurun_out/r2r_c3_$R.err
  python - $R $rc <<'PY'
import json,sys
d=json.loads([l for l in open(f'gpurun_out/r2r_{sys.argv[1]}.json') if l.startswith('{')][-1])
c=json.loads([l for l in open(f'gpurun_out/r2r_c3_{sys.argv[1]}.json') if l.startswith('{')][-1])
print("rot", sys.argv[1], "rc", sys.argv[2], "c2 value %.4g sha %.3f rsa %.3f | c3 value %.4g sha %.3f" % (d["value"], d["kernel_ms"]["sha256"], d["kernel_ms"]["rsa"], c["value"], c["kernel_ms"]["sha256"]))
PY
done
