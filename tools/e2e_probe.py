"""e2e phase breakdown on the GPU box (engine flag OPT_PROFILE: host pack / upload / wait / resolve on stderr).

    python tools/e2e_probe.py [N] [--registered] [--threads 0,8,4] [--chunks 32768,65536]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkemail_rs_b200 as z
from zkemail_rs_b200.engine import EmailViews
import workload as gen

ap = argparse.ArgumentParser()
ap.add_argument("n", nargs="?", type=int, default=262144)
ap.add_argument("--registered", action="store_true", help="raw pool in registered (pinned) memory: zero-copy device front end")
ap.add_argument("--threads", default="0")
ap.add_argument("--chunks", default="65536")
a = ap.parse_args()
kp = gen.KeyPool(256, 0)
mp = gen.MailPool(kp, a.n, 4096, neg_fraction=0.01)
views = EmailViews.from_arrays(mp.engine_views(), keep=mp)
for chunk in map(int, a.chunks.split(",")):
    for thr in map(int, a.threads.split(",")):
        eng = z.Engine(now_unix=1704067200, chunk_emails=chunk, host_threads=thr, flags=z.OPT_PROFILE)
        if a.registered:
            eng.register_host(mp.raw)
        eng.verify_views(views)
        t = time.perf_counter(); r = eng.verify_views(views); dt = time.perf_counter() - t
        assert int(((r["status"] == 0) != mp.expected_ok()).sum()) == 0
        print(json.dumps({"chunk": chunk, "threads": thr or os.cpu_count(), "registered": a.registered, "emails_per_s": a.n / dt}), flush=True)
        if a.registered:
            eng.unregister_host(mp.raw)
        eng.close()
