"""e2e phase breakdown with registered inputs (device front end) on the GPU box."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ZKB_PROFILE"] = "1"
import numpy as np
import zkemail_rs_b200 as z
from zkemail_rs_b200.engine import EmailViews
import workload as gen
N = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
kp = gen.KeyPool(256, 0)
mp = gen.MailPool(kp, N, 4096, neg_fraction=0.01)
views = EmailViews.from_arrays(mp.engine_views(), keep=mp)
for chunk, thr in ((32768, 0), (65536, 0), (131072, 0), (32768, 8), (32768, 4)):
    eng = z.Engine(now_unix=1704067200, chunk_emails=chunk, host_threads=thr)
    eng.register_host(mp.raw)
    eng.verify_views(views)
    t = time.perf_counter(); r = eng.verify_views(views); dt = time.perf_counter() - t
    assert int(((r["status"] == 0) != mp.expected_ok()).sum()) == 0
    print(json.dumps({"chunk": chunk, "threads": thr or os.cpu_count(), "emails_per_s": N / dt}), flush=True)
    eng.unregister_host(mp.raw)
    eng.close()
