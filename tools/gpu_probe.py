"""Quick on-GPU probe: integer-pipe peaks, per-kernel timings for a few RSA lane counts, e2e rate.
Development aid; bench.py is the measured contract."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
import zkemail_rs_b200 as z
from zkemail_rs_b200 import synth
from zkemail_rs_b200.engine import EmailViews

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
BODY = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
keys = [synth.KeyPair.generate(2048) for _ in range(8)]
t0 = time.time()
def mk(i):
    rng = np.random.default_rng(1000 + i)
    return synth.make_email(rng, keys[i % 8], f"d{i % 8}.example.com", idx=i, body_len=BODY)
with ThreadPoolExecutor(os.cpu_count()) as ex:
    emails = list(ex.map(mk, range(N)))
print("gen", N, "emails in %.1fs" % (time.time() - t0), "nproc", os.cpu_count(), flush=True)
views = EmailViews.from_emails(emails)
out = {}
for lanes in (4, 8, 16):
    eng = z.Engine(rsa_lanes=lanes, now_unix=1704067200)
    if lanes == 4:
        out["peaks"] = eng.int_pipe_peaks()
        print(json.dumps(out["peaks"]), flush=True)
    pb = eng.prepare(views)
    for _ in range(3): pb.run()
    ts = []
    for _ in range(5):
        pb.run(); ts.append(pb.timing_ms())
    best = min(ts, key=lambda t: t["total"])
    st = pb.stats()
    res = pb.fetch()
    ok = int((res["status"] == 0).sum())
    pb.close()
    t1 = time.time(); r2 = eng.verify_views(views); t2 = time.time()
    r2 = eng.verify_views(views); t3 = time.time()
    print(json.dumps({"lanes": lanes, "timing_ms": best, "ok": ok, "n": N, "emails_per_s_resident": N / (best["total"] * 1e-3),
                      "e2e_emails_per_s": N / (t3 - t2), "e2e_first": N / (t2 - t1), "stats": st}), flush=True)
    eng.close()
