#!/usr/bin/env python
"""bench.py — DKIM-verified emails/s on B200 (the metric of BASELINE.json), one JSON line.

    python bench.py --gpus N --steps K --warmup W            # this engine (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm

A "step" is one pass of the hot path (verify_email over one batch) over synthetic, locally signed
mail (no DNS / real mail offline).  One JSON line:
  value           device-resident throughput: canonical bytes packed in HBM, one step = SHA-256 + bh= check + RSA
                  (+ DFA scans when the workload has regex parts) + result records, CUDA events on the engine stream
  value_from_raw  the same with the RAW messages resident: one step = device front end (header parsing, preimages,
                  base64) + body canonicalisation + the above
  with_regex      (default workload) the same batch through verify_email_with_regex (configs[3]: two header parts and
                  one body part): resident value, DFA scan time, end-to-end value
  e2e             the batch through the C-ABI call zkb_verify_batch with HOST buffers laid out as the reference's
                  &[Email] is: one separate pageable heap allocation per message (host staging copy, H2D, every
                  kernel, D2H of the result records all inside the timed region)
  e2e_registered  the same call with the messages in ONE caller spool registered with the engine beforehand
                  (zero-copy DMA; the cudaHostRegister cost is reported beside it)
  roofline        the dominant kernel against the roof that binds it (integer FMA pipe for RSA); roofline_hbm is the
                  HBM figure the contract asks for
`cpu_baseline` / `--impl reference` time the CPU restatement of the reference (oracle/, SHA-256 and the RSA public
op through OpenSSL libcrypto, one thread per host core): the reference itself is Rust over un-vendored crates and
cannot be built in this image.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "DKIM-verified emails/sec (RSA-2048+SHA-256+regex)"
NOW = 1704067200
WORKLOADS = {
    "c2": dict(name="configs[1]: 1M synthetic 4 KB-body RSA-2048 DKIM emails, verify_email batch on 1 B200",
               emails=1_000_000, body=4096, regex=False, keys2048=256, keys1024=0),
    "c3": dict(name="configs[2]: 100k emails with 100 KB bodies (large-body batch)",
               emails=100_000, body=102_400, regex=False, keys2048=256, keys1024=0),
    "c4": dict(name="configs[3]: verify_email_with_regex, 1M emails, from/subject/body-token DFAs",
               emails=1_000_000, body=4096, regex=True, keys2048=256, keys1024=0),
    "c5": dict(name="configs[4]: mixed-key sweep, RSA-1024/2048 keys and 1-64 KB bodies",
               emails=1_000_000, body=(1024, 65536), regex=False, keys2048=128, keys1024=128),
}
REGEX_CONFIG = dict(
    header=[r"from:[^\r\n]*@d[0-9]+\.example\.com", r"\r\nsubject:[^\r\n]+\r\n"],
    body=[r"Transaction ID: [A-Z0-9]+"],
)


class ClockSampler:
    """Samples SM clock / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, device: int):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def _run(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.ok:
            self._stop.clear()
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()
            self._t = None

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def build_pool(wl, n_emails, unique, seed, threads, log):
    """Seeded synthetic pool (workload/zk_gen.c).  Returns (MailPool, order) where order repeats the
    unique signed emails (in arena order) up to n_emails when unique < n_emails."""
    import workload as gen
    t0 = time.time()
    keys = gen.KeyPool(wl["keys2048"], wl["keys1024"], threads)
    t_keys = time.time() - t0
    rng = np.random.default_rng(seed)
    def body_lens(n):
        b = wl["body"]
        if isinstance(b, tuple):  # log-uniform
            return np.exp(rng.uniform(np.log(b[0]), np.log(b[1]), size=n)).astype(np.uint32)
        return np.full(n, b, dtype=np.uint32)
    if unique <= 0:  # adaptive: spend at most ~45 s signing
        probe_n = 2048
        t1 = time.time()
        gen.MailPool(keys, probe_n, body_lens(probe_n), seed=seed + 1, token=wl["regex"], threads=threads)
        rate = probe_n / (time.time() - t1)
        unique = int(min(n_emails, max(4096, rate * 45)))
        if unique < n_emails:
            unique = 1 << int(np.log2(unique))
    unique = min(unique, n_emails)
    t1 = time.time()
    pool = gen.MailPool(keys, unique, body_lens(unique), seed=seed, neg_fraction=0.01, token=wl["regex"],
                        qp_percent=10 if wl["regex"] else 0, threads=threads)
    t_gen = time.time() - t1
    if unique < n_emails:
        # the unique pool repeated in arena order: what a caller submitting one spool several times would pass
        # (keeps every chunk's views contiguous, like the all-unique case; the repeat is > L2 either way)
        order = np.arange(n_emails) % unique
    else:
        order = np.arange(n_emails)
    log(f"pool: {unique} unique signed emails ({t_gen:.1f}s, keys {t_keys:.1f}s), batch {n_emails}")
    return pool, order


def oracle_batch(views7: np.ndarray, threads: int, regex_parts=None) -> tuple:
    """Times the CPU restatement (oracle + OpenSSL primitives) over the given zo_email records."""
    import oracle
    L = oracle.lib()
    n = views7.shape[0]
    out = (oracle.Result * max(1, n))()
    hp = bp = None
    nh = nb = 0
    keep = oracle._Keep()
    if regex_parts is not None:
        hp, nh = oracle._marshal_parts(regex_parts[0], keep)
        bp, nb = oracle._marshal_parts(regex_parts[1], keep)
    use_ssl = 1 if L.zo_has_openssl() else 0
    t0 = time.perf_counter()
    rc = L.zo_verify_batch_mt(C.cast(views7.ctypes.data, C.POINTER(oracle._Email)), n, hp, nh, bp, nb, NOW,
                              threads, use_ssl, out)
    dt = time.perf_counter() - t0
    assert rc == 0
    ok = sum(1 for i in range(0, n, max(1, n // 2000)) if out[i].status == 0)
    return dt, out, use_ssl, ok


def workload_config(wl, n_emails, world):
    """The `config` object: identical for this engine's arm and the reference arm."""
    return {"workload": wl["name"], "emails_per_gpu_per_step": n_emails,
            "negatives": "1% (body flip / signature flip / wrong key)", "keys": f"{wl['keys2048']}x2048+{wl['keys1024']}x1024",
            "body_bytes": wl["body"] if isinstance(wl["body"], int) else list(wl["body"]),
            "regex_config": REGEX_CONFIG if (wl["regex"] or wl.get("also_regex")) else None,
            "parallelism": f"shard-by-email x{world}"}


def regex_info():
    import zkemail_rs_b200 as z
    from zkemail_rs_b200.structs import CompiledRegex, RegexInfo
    return RegexInfo([CompiledRegex(z.compile_regex(p), None) for p in REGEX_CONFIG["header"]],
                     [CompiledRegex(z.compile_regex(p), None) for p in REGEX_CONFIG["body"]])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--emails", type=int, default=0, help="emails per GPU per step (default: the workload's)")
    ap.add_argument("--unique", type=int, default=0, help="unique signed emails in the pool (0 = adaptive)")
    ap.add_argument("--rsa-lanes", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--host-threads", type=int, default=0)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-extras", action="store_true", help="only value + e2e (no with_regex / value_from_raw / e2e_registered)")
    ap.add_argument("--profile", action="store_true", help="engine flag ZKB_OPT_PROFILE: per-call host / stream breakdown on stderr")
    ap.add_argument("--no-sqr", action="store_true", help="engine flag ZKB_OPT_NO_SQR: the plain RSA-2048 kernel instead of the one with the dedicated Montgomery squaring (A/B)")
    ap.add_argument("--seed", type=int, default=0xD1C1)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = dict(WORKLOADS[args.workload])
    wl["also_regex"] = args.workload == "c2" and not args.skip_extras   # the default line also carries configs[3]
    n_emails = args.emails or wl["emails"]
    ncpu = os.cpu_count() or 1
    threads = args.host_threads or max(1, ncpu // max(1, world if args.impl == "b200" else 1))
    log = lambda m: print(f"[bench r{rank}] {m}", file=sys.stderr, flush=True)
    W = max(args.warmup, 3)          # timing rule: at least 3 warm-up steps; otherwise exactly what was asked for
    K = max(args.steps, 1)
    pool_wl = dict(wl, regex=wl["regex"] or wl["also_regex"])   # bodies carry the token whenever regex parts may run

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        # same config, same emails per step; large-body workloads are sampled (a 100 KB body costs ~25x a 4 KB one)
        sample = n_emails if (isinstance(wl["body"], int) and wl["body"] <= 8192) else min(n_emails, 40_000)
        pool, order = build_pool(pool_wl, sample, min(args.unique or sample, sample), args.seed, ncpu, log)
        views = pool.oracle_views(order)
        parts = None
        if wl["regex"] or wl["also_regex"]:
            info = regex_info()
            parts = (info.header_parts, info.body_parts)
        main_parts = parts if wl["regex"] else None
        for _ in range(W):
            oracle_batch(views[: max(1, sample // 10)], ncpu, main_parts)
        times = []
        for _ in range(K):
            dt, _, use_ssl, _ = oracle_batch(views, ncpu, main_parts)
            times.append(dt)
        total = sum(times)
        v = sample * K / total
        with_regex = None
        if wl["also_regex"]:
            kk = min(K, 3)
            oracle_batch(views[: max(1, sample // 10)], ncpu, parts)
            dts = [oracle_batch(views, ncpu, parts)[0] for _ in range(kk)]
            with_regex = {"value": sample * kk / sum(dts), "unit": "emails/s", "steps": kk}
        line = {
            "impl": "reference", "metric": METRIC, "value": v, "unit": "emails/s", "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * total / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(wl, n_emails, world),
            "detail": {"emails_per_step": sample, "note": "CPU restatement of zkemail_core::verify_email (oracle/ + OpenSSL libcrypto "
                       "SHA-256/RSA), all host threads; the Rust reference cannot be built in this image"},
            "cpu_baseline": {"value": v, "unit": "emails/s", "cores": ncpu, "kind": "port",
                             "sample": f"{sample} emails/step x {K} steps, openssl={bool(use_ssl)}"},
            "with_regex": with_regex,
            "e2e": {"value": v, "unit": "emails/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ this engine
    import torch
    import zkemail_rs_b200 as z
    from zkemail_rs_b200.engine import EmailViews, RegexSet

    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    pool, order = build_pool(pool_wl, n_emails, args.unique, args.seed + 7919 * rank, threads, log)
    eng = z.Engine(device=local_rank, host_threads=threads, now_unix=NOW, chunk_emails=args.chunk,
                   rsa_lanes=args.rsa_lanes, flags=(z.OPT_PROFILE if args.profile else 0) | (z.OPT_NO_SQR if args.no_sqr else 0))
    views_np = pool.engine_views(order)
    views = EmailViews.from_arrays(views_np, keep=pool)
    exp_ok = pool.expected_ok()[order]
    info = regex_info() if (wl["regex"] or wl["also_regex"]) else None
    regex_set = RegexSet(eng, info) if info is not None else None
    main_regex = regex_set if wl["regex"] else None
    stream = torch.cuda.ExternalStream(eng.lib.zkb_engine_stream(eng.handle), device=dev)
    clocks = ClockSampler(local_rank)
    # multi-GPU: the one exchange of the path is the all-gather of the fixed-size result records, issued by the
    # library itself (ncclSend / ncclRecv per resident chunk beside the next chunk's kernels, csrc/engine_multi.inc); torch.distributed only carries the id
    comm = z.Comm(eng, rank, world) if dist is not None else None

    t0 = time.time()
    eng.register_host(pool.raw)
    t_register = time.time() - t0
    log(f"cudaHostRegister of the raw pool ({pool.raw.nbytes / 1e9:.2f} GB): {t_register:.2f}s")

    def timed_resident(pb, label):
        """W warm-up steps, then K timed steps (CUDA events on the engine stream, barrier + synchronize on both sides,
        max over ranks).  Every step ends with the record all-gather when there are several ranks."""
        def step():
            if comm is not None:
                comm.last = comm.run_allgather(pb)     # the kernels + the record exchange, chunk k's records travelling under chunk k + 1
            else:
                pb.run_async()
        for _ in range(W):
            step()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(K):
            step()
        ev1.record(stream)
        ev1.synchronize()
        barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1))
        res = pb.fetch()   # the verdicts of the timed schedule against the generator's ground truth
        bad = int(((res["status"] == 0) != exp_ok).sum())
        assert bad == 0, f"{label}: {bad} verdicts differ from the generator's ground truth"
        if comm is not None:
            # what the exchange delivered: every rank's slot must carry that rank's verdicts (a checksum per rank,
            # compared through torch.distributed, outside the timed region)
            ptr, slot, rb = comm.last
            counts = comm.rank_records()
            got = torch.empty((world, slot, rb), dtype=torch.uint8, device=dev)
            import ctypes as C
            rt = C.CDLL("libcudart.so.12")
            assert rt.cudaMemcpy(C.c_void_p(got.data_ptr()), C.c_void_p(ptr), C.c_size_t(got.numel()), 3) == 0
            st = got[:, :, 0:4].contiguous().view(torch.int32).squeeze(-1)            # status word of every record
            sums = torch.stack([(st[r, :counts[r]] == 0).sum() for r in range(world)]).to(torch.int64)
            mine = torch.tensor([int((res["status"] == 0).sum())], dtype=torch.int64, device=dev)
            allm = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allm, mine)
            want = torch.cat(allm)
            # records the device declined (status REDO) are finished by the host in fetch(): at most those may differ
            redo = torch.stack([(st[r, :counts[r]] == 0x7fffffff).sum() for r in range(world)]).to(torch.int64)
            assert bool(((want - sums) >= 0).all()) and bool(((want - sums) <= redo).all()), (label, want.tolist(), sums.tolist(), redo.tolist())
        return ms

    # ---- device-resident pass: batch packed + uploaded once ----
    t0 = time.time()
    pb = eng.prepare(views, main_regex, with_captures=False)
    log(f"prepare (pack + H2D): {time.time() - t0:.2f}s")
    stats = pb.stats()
    # per-kernel-family CUDA-event times (synchronous form of the same launches); also takes a fresh box out of its
    # idle power state before the W warm-up steps
    fam = []
    for _ in range(3):
        pb.run()
        fam.append(pb.timing_ms())
    fam_best = {k: float(np.median([f[k] for f in fam])) for k in fam[0]}
    clocks.start()
    dev_ms = timed_resident(pb, "value")
    clocks.stop()
    pb.close()
    total_emails = n_emails * world
    value = total_emails * K / (dev_ms * 1e-3)
    peaks = eng.int_pipe_peaks()

    with_regex = None
    value_from_raw = None
    if wl["also_regex"]:
        pbr = eng.prepare(views, regex_set, with_captures=False)
        st_r = pbr.stats()
        famr = []
        for _ in range(3):
            pbr.run()
            famr.append(pbr.timing_ms())
        ms_r = timed_resident(pbr, "with_regex")
        pbr.close()
        with_regex = {"workload": WORKLOADS["c4"]["name"], "value": total_emails * K / (ms_r * 1e-3), "unit": "emails/s", "ms_per_step": ms_r / K,
                      "dfa_scan": {"launch_ms": float(np.median([f["dfa"] for f in famr])), "bytes_stepped": st_r["dfa_bytes"],
                                   "parts": len(REGEX_CONFIG["header"]) + len(REGEX_CONFIG["body"])},
                      "gpu_launches": st_r["kernel_launches"] * K}
    if not args.skip_extras:
        pbw = eng.prepare(views, main_regex, with_captures=False, raw=True)
        st_w = pbw.stats()
        famw = []
        for _ in range(3):
            pbw.run()
            famw.append(pbw.timing_ms())
        ms_w = timed_resident(pbw, "value_from_raw")
        pbw.close()
        fw = {k: float(np.median([f[k] for f in famw])) for k in famw[0]}
        value_from_raw = {"value": total_emails * K / (ms_w * 1e-3), "unit": "emails/s", "ms_per_step": ms_w / K,
                          "kernel_ms": fw, "gpu_launches": st_w["kernel_launches"] * K,
                          "note": "raw messages resident in HBM; one step = device front end + body canonicalisation + SHA-256 + bh= + RSA"
                                  + (" + DFA scans" if main_regex else "") + " + result records"}

    # ---- end to end through the C ABI with host buffers ----
    from zkemail_rs_b200.engine import RESULT_DTYPE
    out_buf = np.zeros(n_emails, dtype=RESULT_DTYPE)   # the caller's result array, reused across calls like its input buffers

    def timed_e2e(vw, rs, label):
        eng.verify_views(vw, rs, with_captures=False, out=out_buf)   # one untimed call (staging buffers allocated)
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            r2 = eng.verify_views(vw, rs, with_captures=False, out=out_buf)
            if dist is not None:   # every rank learns every shard's verdicts: bitmap all-gather inside the step
                bits = torch.from_numpy(np.packbits(r2["status"] == 0)).to(dev)
                allb = torch.empty(bits.numel() * world, dtype=torch.uint8, device=dev)
                dist.all_gather_into_tensor(allb, bits)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        bad = int(((r2["status"] == 0) != exp_ok).sum())
        assert bad == 0, f"{label}: {bad} verdicts differ"
        b = eng.last_batch_bytes()
        return {"value": total_emails * K / dt, "unit": "emails/s", "h2d_bytes_per_step": b["h2d_bytes"], "d2h_bytes_per_step": b["d2h_bytes"],
                "ms_per_step": 1e3 * dt / K, "host_front_end_emails_per_step": b["host_front_end_emails"]}

    e2e_registered = None
    if not args.skip_extras:
        e2e_registered = timed_e2e(views, main_regex, "e2e_registered")
        e2e_registered["input_memory"] = ("one caller spool registered with zkb_host_register before the timed region "
                                          f"(cudaHostRegister of {pool.raw.nbytes / 1e9:.2f} GB took {t_register:.2f} s, not included): "
                                          "raw messages DMA'd as they are")
    eng.unregister_host(pool.raw)
    # the reference's caller shape: &[Email], one heap Vec<u8> per message -> one separate pageable allocation each
    t0 = time.time()
    raw_off, raw_len = pool.raw_off[order], pool.raw_len[order]
    scattered = [pool.raw[int(o):int(o) + int(l)].tobytes() for o, l in zip(raw_off, raw_len)]
    sv = views_np.copy()
    sv[:, 2] = np.fromiter((C.cast(C.c_char_p(b), C.c_void_p).value for b in scattered), dtype=np.uint64, count=len(scattered))
    views_scattered = EmailViews.from_arrays(sv, keep=(pool, scattered))
    log(f"scattered copy of the batch (one heap allocation per message): {time.time() - t0:.1f}s")
    clocks.start()
    e2e = timed_e2e(views_scattered, main_regex, "e2e")
    clocks.stop()
    e2e["input_memory"] = ("pageable host memory, one separate heap allocation per message (the reference's &[Email] / Vec<u8> shape): "
                           "host threads copy each raw message into pinned staging, everything else on the device")
    if dist is not None:
        e2e["collective"] = "verdict bitmap all-gather (torch.distributed NCCL) after every call"
        e2e["value_per_gpu"] = e2e["value"] / world
    if with_regex is not None:
        er = timed_e2e(views_scattered, regex_set, "with_regex e2e")
        with_regex["e2e"] = {k: er[k] for k in ("value", "unit", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step")}
    del scattered

    # ---- roofline of the dominant kernel ----
    dom = "rsa" if fam_best["rsa"] >= fam_best["sha256"] else "sha256"
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_file):
        try:
            hbm_peak = float(json.load(open(peaks_file))["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    n_rsa = stats["rsa_items_1024"] + stats["rsa_items_2048"] + stats["rsa_items_other"]
    if dom == "rsa":   # per signature: k-byte signature + 32-byte digest + 16-byte item in, 4-byte flag out
        alg_bytes = stats["rsa_items_2048"] * (256 + 32 + 16 + 4) + stats["rsa_items_1024"] * (128 + 32 + 16 + 4)
    else:              # per message: its bytes in + 12-byte descriptor, 32-byte digest out
        alg_bytes = stats["sha_bytes"] + stats["n_sha_messages"] * (12 + 4 + 32)
    dom_ms = fam_best[dom]
    traffic = None
    traffic_src = None
    for name in ("ncu_traffic_r2.json", "ncu_traffic_r1.json"):   # DRAM bytes of the same kernel from the committed ncu capture, per unit
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", name)))
            if dom == "rsa":
                traffic = tr["rsa_verify_kernel"]["dram_bytes_per_launch"] / tr["rsa_verify_kernel"]["signatures_per_launch"] * n_rsa
            else:
                traffic = tr["sha256_batch_kernel"]["dram_bytes_per_launch"] / tr["sha256_batch_kernel"]["message_bytes_per_launch"] * stats["sha_bytes"]
            traffic_src = f"profiles/{name} (ncu --set full of this kernel, scaled per signature / per byte to this launch)"
            break
        except Exception:
            continue
    ach = alg_bytes / (dom_ms * 1e-3) / 1e9
    launches_per_step = -(-n_emails // ((args.chunk or 65536) * 4))
    roofline_hbm = {"kernel": "rsa_verify_kernel" if dom == "rsa" else "sha256_batch_kernel", "bound": "hbm",
                    "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": dom_ms,
                    "note": "not the binding roof: the path is integer-issue bound (SURVEY.md section 8d)"}
    # algorithmic MACs (18 multiplications of 2 l^2 + l) and the MACs the kernel executes
    # (dedicated squarings need fewer): the first grades the step, the second says how busy the pipe is
    rsa_alg = stats["rsa_macs"]
    # executed IMAD.WIDE per signature: a multiplication of the interleaved loop is 2 l^2; a dedicated squaring of the
    # 2048-bit kernel is l(l+1)/2 + l^2 = 6176 (16 of the 18 multiplications when four lanes work on a signature)
    sqr_on = (not args.no_sqr) and (args.rsa_lanes or 4) == 4
    rsa_exec = (stats["rsa_items_2048"] * ((2 * 8192 + 16 * 6176) if sqr_on else 18 * 8192) + stats["rsa_items_1024"] * 18 * 2048
                + stats["rsa_items_other"] * 18 * 32768)
    sha_ops = stats["sha_blocks"] * 1400
    rsa_ms, sha_ms = fam_best["rsa"], fam_best["sha256"]
    roofline_int = {
        "rsa_verify_kernel": {"bound": "int-fma (IMAD.WIDE.U32, FMA-heavy pipe)", "achieved": rsa_alg / (rsa_ms * 1e-3) / 1e9 if rsa_ms > 0 else None,
                              "peak": peaks["imad_wide_gops"], "unit": "G MAC/s", "launch_ms": rsa_ms,
                              "executed_macs_per_launch": rsa_exec, "algorithmic_macs_per_launch": rsa_alg},
        "sha256_batch_kernel": {"bound": "int-alu (SHF/LOP3 on the ALU pipe; additions issued on the FMA pipe as IMAD)",
                                "achieved": sha_ops / (sha_ms * 1e-3) / 1e9, "peak": peaks["iadd3_gops"], "unit": "G instr/s", "launch_ms": sha_ms},
        "dfa_scan": {"launch_ms": fam_best["dfa"], "bytes_stepped": stats["dfa_bytes"]},
        "peak_source": "zkb_int_pipe_peaks (register-only microbenchmark, this run)",
    }
    for k in ("rsa_verify_kernel", "sha256_batch_kernel"):
        a = roofline_int[k]
        a["frac"] = (a["achieved"] / a["peak"]) if a["achieved"] and a["peak"] else None
    rk = roofline_int["rsa_verify_kernel"]
    rk["pipe_utilisation"] = rk["frac"] * rsa_exec / rsa_alg if rk["frac"] else None
    sh = roofline_int["sha256_batch_kernel"]
    # 1400 = algorithmic integer instructions per 64-byte block; 1028 of them (SHF/LOP3) run on the ALU pipe, the
    # additions on the FMA pipe: frac is the algorithmic rate over ONE pipe's peak, pipe_utilisation the ALU pipe's share
    sh["pipe_utilisation"] = sh["frac"] * 1028.0 / 1400.0 if sh["frac"] else None
    d = roofline_int["rsa_verify_kernel" if dom == "rsa" else "sha256_batch_kernel"]
    roofline = {"kernel": "rsa_verify_kernel" if dom == "rsa" else "sha256_batch_kernel",
                "bound": "int-fma" if dom == "rsa" else "int-alu", "achieved": d["achieved"], "peak": d["peak"], "unit": d["unit"],
                "frac": d["frac"], "pipe_utilisation": d["pipe_utilisation"], "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_work_per_launch": rsa_alg if dom == "rsa" else sha_ops, "launch_ms": dom_ms, "launches_per_step": launches_per_step,
                "peak_source": roofline_int["peak_source"],
                "note": "achieved = algorithmic work (SURVEY.md section 8d) / CUDA-event time of the kernel's launches of one step, measured in this run; "
                        "tensor cores and HBM are not the binding roofs of multi-precision integer arithmetic (roofline_hbm carries the HBM figure)"}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        oviews = pool.oracle_views(order)
        parts = (info.header_parts, info.body_parts) if wl["regex"] else None
        probe = min(len(oviews), 20_000 if wl["body"] == 4096 else 2_000)
        dt, _, use_ssl, _ = oracle_batch(oviews[:probe], ncpu, parts)
        sample = int(min(len(oviews), max(probe, probe / dt * 12)))
        dt, _, use_ssl, ok = oracle_batch(oviews[:sample], ncpu, parts)
        cpu_baseline = {"value": sample / dt, "unit": "emails/s", "cores": ncpu, "kind": "port",
                        "sample": f"{sample} emails of the same batch, oracle/ + OpenSSL={bool(use_ssl)}, {ncpu} threads, {dt:.1f}s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "emails/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": workload_config(wl, n_emails, world),
            "detail": {"unique_signed_emails_per_gpu": int(pool.n), "tiling": "unique pool repeated in arena order" if pool.n < n_emails else "none",
                       "l2": f"inputs larger than L2: {stats['arena_bytes'] / 1e9:.2f} GB arena per step", "host_threads": threads,
                       "value_is": "device-resident verify_email" + ("_with_regex" if main_regex else "") + " step: SHA-256 + bh= + RSA"
                                   + (" + DFA scans" if main_regex else "") + " + result records; value_from_raw adds the device front end and canonicalisation",
                       "collective": "none (1 GPU)" if world == 1 else "exchange of the result records (144 B per email; ncclSend / ncclRecv issued by the library per resident chunk, overlapping the next chunk's kernels) inside every step",
                       "rsa_lanes": args.rsa_lanes or 4},
            "roofline": roofline, "roofline_hbm": roofline_hbm, "roofline_int": roofline_int, "cpu_baseline": cpu_baseline,
            "e2e": e2e, "e2e_registered": e2e_registered, "value_from_raw": value_from_raw, "with_regex": with_regex,
            # what bounds the end-to-end numbers as ranks are added on ONE box: the kernels need ~35 ms of
            # stream time per 1 M emails and rank, everything else is the host side the ranks share
            "e2e_scaling": {"ranks": world, "host_threads_per_rank": threads,
                            "e2e_bound": "host copy of every raw message into pinned staging (pageable callers; the box's host threads are split between the ranks)",
                            "e2e_registered_bound": "PCIe Gen5 x16 per GPU (~52 GB/s = ~10 M emails/s of 5.2 KB messages) until the box's host-memory bandwidth is shared by the links",
                            "kernel_stream_ms_per_step": (value_from_raw or {}).get("ms_per_step")},
            "gpu_launches": stats["kernel_launches"] * K,
            "kernel_ms": fam_best, "clocks": clocks.summary(), "int_pipe_peaks": peaks, "nproc": ncpu,
        }
        print(json.dumps(line), flush=True)
    if comm is not None:
        comm.close()
    if regex_set is not None:
        regex_set.close()
    eng.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
