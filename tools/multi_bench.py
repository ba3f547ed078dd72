"""One process, several GPUs (zkb_multi_*): end-to-end and resident throughput of the in-process multi-device engine.

    python tools/multi_bench.py <n_devices> <emails total>

Prints one JSON line: e2e emails/s through zkb_multi_verify_batch (pageable caller memory, one spool), resident
emails/s of zkb_multi_batch_run with and without the NCCL all-gather of the result records."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zkemail_rs_b200 as z
from zkemail_rs_b200.engine import EmailViews
import workload as gen

nd = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
kp = gen.KeyPool(256, 0)
mp = gen.MailPool(kp, n, 4096, neg_fraction=0.01)
views = EmailViews.from_arrays(mp.engine_views(), keep=mp)
exp = mp.expected_ok()
m = z.MultiEngine(n_devices=nd, now_unix=1704067200)
out = {"n_devices": nd, "emails": n, "host_threads": os.cpu_count()}
r = m.verify_views(views)
assert int(((r["status"] == 0) != exp).sum()) == 0
ts = []
for _ in range(3):
    t = time.perf_counter(); r = m.verify_views(views); ts.append(time.perf_counter() - t)
out["e2e_pageable_emails_per_s"] = n / min(ts)
m.register_host(mp.raw)
m.verify_views(views)
ts = []
for _ in range(3):
    t = time.perf_counter(); r = m.verify_views(views); ts.append(time.perf_counter() - t)
assert int(((r["status"] == 0) != exp).sum()) == 0
out["e2e_registered_emails_per_s"] = n / min(ts)
for raw in (False, True):
    mb = m.prepare(views, raw=raw)
    for gather in (False, True):
        for _ in range(3):
            mb.run(gather=gather)
        ms = [mb.run(gather=gather) for _ in range(5)]
        out[f"resident{'_raw' if raw else ''}{'_gather' if gather else ''}_emails_per_s"] = n / (float(np.median(ms)) * 1e-3)
    res = mb.fetch()
    assert int(((res["status"] == 0) != exp).sum()) == 0
    g0 = mb.gathered(nd - 1)
    live = g0["status"] != 0x7fffffff
    assert int(((g0["status"][live] == 0) != exp[live]).sum()) == 0
    mb.close()
m.unregister_host(mp.raw)
m.close()
print(json.dumps(out))
