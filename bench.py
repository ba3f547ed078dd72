#!/usr/bin/env python
"""bench.py — DKIM-verified emails/s on B200 (the metric of BASELINE.json), one JSON line.

    python bench.py --gpus N --steps K --warmup W            # this engine (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm

A "step" is one pass of the hot path (verify_email over one batch) over synthetic, locally signed
mail (no DNS / real mail offline).  `value` = device-resident throughput (batch already packed in
HBM; kernels only), `e2e` = the same batch through the C-ABI call zkb_verify_batch with HOST
buffers (host parse + canonicalise + pack, H2D, kernels, D2H, result resolution all inside the
timed region).  `cpu_baseline` / `--impl reference` time the CPU restatement of the reference
(oracle/, SHA-256 and the RSA public op through OpenSSL libcrypto, one thread per host core): the
reference itself is Rust over un-vendored crates and cannot be built in this image.
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
    ap.add_argument("--no-direct", action="store_true", help="pageable inputs: one host copy of each raw message into pinned staging, device front end")
    ap.add_argument("--seed", type=int, default=0xD1C1)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = dict(WORKLOADS[args.workload])
    n_emails = args.emails or wl["emails"]
    ncpu = os.cpu_count() or 1
    threads = args.host_threads or max(1, ncpu // max(1, world if args.impl == "b200" else 1))
    log = lambda m: print(f"[bench r{rank}] {m}", file=sys.stderr, flush=True)
    W = max(args.warmup, 0)
    K = max(args.steps, 1)

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = min(n_emails, 100_000 if not isinstance(wl["body"], int) or wl["body"] <= 8192 else 8_000)
        pool, order = build_pool(wl, sample, min(args.unique or sample, sample), args.seed, ncpu, log)
        views = pool.oracle_views(order)
        regex_parts = None
        if wl["regex"]:
            import zkemail_rs_b200 as z
            from zkemail_rs_b200.structs import CompiledRegex
            regex_parts = ([CompiledRegex(z.compile_regex(p), None) for p in REGEX_CONFIG["header"]],
                           [CompiledRegex(z.compile_regex(p), None) for p in REGEX_CONFIG["body"]])
        for _ in range(W):
            oracle_batch(views[: max(1, sample // 10)], ncpu, regex_parts)
        times = []
        for _ in range(K):
            dt, _, use_ssl, _ = oracle_batch(views, ncpu, regex_parts)
            times.append(dt)
        total = sum(times)
        v = sample * K / total
        line = {
            "impl": "reference", "metric": METRIC, "value": v, "unit": "emails/s", "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * total / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": wl["name"], "emails_per_step": sample,
                       "note": "CPU restatement of zkemail_core::verify_email (oracle/ + OpenSSL libcrypto SHA-256/RSA), "
                               "all host threads; the Rust reference cannot be built in this image"},
            "cpu_baseline": {"value": v, "unit": "emails/s", "cores": ncpu, "kind": "port",
                             "sample": f"{sample} emails/step x {K} steps, openssl={bool(use_ssl)}"},
            "e2e": {"value": v, "unit": "emails/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ this engine
    import torch
    import zkemail_rs_b200 as z
    from zkemail_rs_b200.engine import EmailViews, RegexSet
    from zkemail_rs_b200.structs import CompiledRegex, RegexInfo

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

    pool, order = build_pool(wl, n_emails, args.unique, args.seed + 7919 * rank, threads, log)
    eng = z.Engine(device=local_rank, host_threads=threads, now_unix=NOW, chunk_emails=args.chunk,
                   rsa_lanes=args.rsa_lanes)
    views_np = pool.engine_views(order)
    views = EmailViews.from_arrays(views_np, keep=pool)
    direct = not args.no_direct
    if direct:  # inputs start in pinned host memory (bench contract): zero-copy DMA + device-side canonicalisation
        t0 = time.time()
        eng.register_host(pool.raw)
        log(f"cudaHostRegister of the raw pool ({pool.raw.nbytes / 1e9:.2f} GB): {time.time() - t0:.2f}s")
    regex = None
    if wl["regex"]:
        info = RegexInfo([CompiledRegex(z.compile_regex(p), None) for p in REGEX_CONFIG["header"]],
                         [CompiledRegex(z.compile_regex(p), None) for p in REGEX_CONFIG["body"]])
        regex = RegexSet(eng, info)
    stream = torch.cuda.ExternalStream(eng.lib.zkb_engine_stream(eng.handle), device=dev)
    clocks = ClockSampler(local_rank)

    # ---- device-resident pass: batch packed + uploaded once, K timed launches of the kernels ----
    t0 = time.time()
    pb = eng.prepare(views, regex, with_captures=False)
    log(f"prepare (pack + H2D): {time.time() - t0:.2f}s")
    stats = pb.stats()
    # multi-GPU: the only exchange of the path is an all-gather of the verdict words (NCCL over NVLink),
    # issued on the engine stream right behind the kernels of every step, straight from HBM
    gather = None
    if dist is not None:
        class _DevArray:  # zero-copy view of the engine's device buffer for torch
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 3}
        flags = [torch.as_tensor(_DevArray(p, n), device=dev) for p, n in pb.device_flags() if n]
        sizes = torch.tensor([sum(f.numel() for f in flags)], device=dev)
        mx = sizes.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        pad = torch.zeros(int(mx.item()), dtype=torch.int32, device=dev)
        allv = torch.empty(int(mx.item()) * world, dtype=torch.int32, device=dev)

        def gather():
            with torch.cuda.stream(stream):
                o = 0
                for f in flags:
                    pad[o:o + f.numel()].copy_(f, non_blocking=True)
                    o += f.numel()
                dist.all_gather_into_tensor(allv, pad)

    def step():
        pb.run_async()
        if gather is not None:
            gather()

    # warm-up: W (>= 3) steps plus a fixed 15 more (~0.35 s of device work), so that a fresh box has left its idle
    # power state before the timed region (a 23.2 ms first step-set was seen once against the usual 22.6 ms).  The
    # count is the same on every rank: each step carries a collective.
    n_warm = max(W, 3) + 15
    for _ in range(n_warm):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks.start()
    ev0.record(stream)
    for _ in range(K):
        step()
    ev1.record(stream)
    ev1.synchronize()
    barrier()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    # the verdicts of the timed (two-stream) schedule, checked against the generator's ground truth
    res = pb.fetch()
    exp_ok = pool.expected_ok()[order]
    assert int(((res["status"] == 0) != exp_ok).sum()) == 0, "timed step: verdicts differ from the generator's ground truth"
    # per-kernel-family CUDA-event times on the engine stream (same launches, synchronous form)
    fam = []
    for _ in range(3):
        pb.run()
        fam.append(pb.timing_ms())
    fam_best = {k: float(np.median([f[k] for f in fam])) for k in fam[0]}
    res = pb.fetch()
    exp_ok = pool.expected_ok()[order]
    got_ok = res["status"] == 0
    n_wrong = int((got_ok != exp_ok).sum())
    assert n_wrong == 0, f"{n_wrong} verdicts differ from the generator's ground truth"
    peaks = eng.int_pipe_peaks()
    pb.close()

    # ---- end to end through the C ABI with host buffers (host pack + H2D + kernels + D2H) ----
    for _ in range(min(W, 1) or 1):
        eng.verify_views(views, regex, with_captures=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        r2 = eng.verify_views(views, regex, with_captures=False)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_bytes = eng.last_batch_bytes()
    clocks.stop()
    assert int(((r2["status"] == 0) != exp_ok).sum()) == 0

    # ---- multi-GPU: check the gathered verdict words (every rank sees every shard's RSA/bh bits) ----
    if dist is not None:
        torch.cuda.synchronize()
        per = allv.view(world, -1)
        n_pass = (per & 3).eq(3).sum(dim=1)
        assert int((n_pass > 0).sum()) == world, "all-gather of verdict words failed"
        assert int(n_pass[rank]) == int(got_ok.sum()), (int(n_pass[rank]), int(got_ok.sum()))

    total_emails = n_emails * world
    value = total_emails * K / (dev_ms * 1e-3)
    e2e_value = total_emails * K / e2e_s

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
        int_ops, int_peak, int_unit = stats["rsa_macs"], peaks["imad_wide_gops"], "G IMAD.WIDE/s"
    else:              # per message: its bytes in + 12-byte descriptor, 32-byte digest out
        alg_bytes = stats["sha_bytes"] + stats["n_sha_messages"] * (12 + 4 + 32)
        int_ops, int_peak, int_unit = stats["sha_blocks"] * 1400, peaks["iadd3_gops"], "G ALU instr/s"
    dom_ms = fam_best[dom]
    traffic = None
    try:  # DRAM bytes of the same kernel from the committed ncu capture, scaled to this launch size
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r1.json")))
        if dom == "rsa":
            traffic = tr["rsa_verify_kernel"]["dram_bytes_per_launch"] / tr["rsa_verify_kernel"]["signatures_per_launch"] * n_rsa
        else:
            traffic = tr["sha256_batch_kernel"]["dram_bytes_per_launch"] / tr["sha256_batch_kernel"]["message_bytes_per_launch"] * stats["sha_bytes"]
    except Exception:
        pass
    ach = alg_bytes / (dom_ms * 1e-3) / 1e9
    roofline = {"kernel": "rsa_verify_kernel" if dom == "rsa" else "sha256_batch_kernel", "bound": "hbm",
                "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": dom_ms,
                "launches_per_step": -(-n_emails // ((args.chunk or 65536) * 4)), "note": "integer-issue bound, not HBM bound: see roofline_int; achieved/traffic are "
                "summed over the launches of one step (one per resident chunk)"}
    roofline_int = {
        "rsa_verify_kernel": {"bound": "fma pipe (IMAD.WIDE)", "achieved": stats["rsa_macs"] / (fam_best["rsa"] * 1e-3) / 1e9 if fam_best["rsa"] > 0 else None,
                              "peak": peaks["imad_wide_gops"], "unit": "G MAC/s", "launch_ms": fam_best["rsa"]},
        "sha256_batch_kernel": {"bound": "alu pipe (SHF/LOP3/IADD3)", "achieved": stats["sha_blocks"] * 1400 / (fam_best["sha256"] * 1e-3) / 1e9,
                                "peak": peaks["iadd3_gops"], "unit": "G instr/s", "launch_ms": fam_best["sha256"]},
        "dfa_scan": {"launch_ms": fam_best["dfa"], "bytes_stepped": stats["dfa_bytes"]},
        "peak_source": "zkb_int_pipe_peaks (register-only microbenchmark, this run)",
    }
    for k in ("rsa_verify_kernel", "sha256_batch_kernel"):
        a = roofline_int[k]
        a["frac"] = (a["achieved"] / a["peak"]) if a["achieved"] and a["peak"] else None
    sh = roofline_int["sha256_batch_kernel"]
    # 1400 = algorithmic integer instructions per 64-byte block; the kernel issues 1028 of them on the ALU pipe
    # (SHF/LOP3) and moves the 594 additions to the otherwise idle FMA pipe (IMAD), so the algorithmic rate can
    # exceed the ALU-pipe peak; alu_pipe_frac is the share of the ALU pipe actually used (ncu: 0.87)
    sh["alu_pipe_frac"] = sh["frac"] * 1028.0 / 1400.0 if sh["frac"] else None
    sh["note"] = "frac = algorithmic instr rate / ALU-pipe peak; additions run on the FMA pipe, see alu_pipe_frac"

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        oviews = pool.oracle_views(order)
        probe = min(len(oviews), 20_000 if wl["body"] == 4096 else 2_000)
        dt, _, use_ssl, _ = oracle_batch(oviews[:probe], ncpu, None)
        sample = int(min(len(oviews), max(probe, probe / dt * 12)))
        dt, _, use_ssl, ok = oracle_batch(oviews[:sample], ncpu, None)
        cpu_baseline = {"value": sample / dt, "unit": "emails/s", "cores": ncpu, "kind": "port",
                        "sample": f"{sample} emails of the same batch, oracle/ + OpenSSL={bool(use_ssl)}, {ncpu} threads, {dt:.1f}s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "emails/s", "n_gpus": world, "steps": K, "warmup": n_warm,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": wl["name"], "emails_per_gpu": n_emails, "unique_signed_emails_per_gpu": int(pool.n),
                       "tiling": "unique pool repeated in arena order" if pool.n < n_emails else "none",
                       "negatives": "1% (body flip / signature flip / wrong key)", "keys": f"{wl['keys2048']}x2048+{wl['keys1024']}x1024",
                       "l2": f"inputs larger than L2: {stats['arena_bytes'] / 1e9:.2f} GB arena per step",
                       "host_threads": threads, "parallelism": f"shard-by-email x{world}",
                       "input_memory": "registered (pinned) host memory: raw messages DMA'd as they are; header parsing, preimages, base64 and "
                                       "body canonicalisation on the device (irregular messages fall back to the host front end)" if direct
                                       else "pageable host memory: host threads copy each raw message into pinned staging, everything else on the device",
                       "collective": "none (1 GPU)" if world == 1 else "NCCL all-gather of verdict words per step (inside the timed region)", "rsa_lanes": args.rsa_lanes or 4},
            "roofline": roofline, "roofline_int": roofline_int, "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": "emails/s", "h2d_bytes_per_step": e2e_bytes["h2d_bytes"],
                    "d2h_bytes_per_step": e2e_bytes["d2h_bytes"], "ms_per_step": 1e3 * e2e_s / K,
                    "host_front_end_emails_per_step": e2e_bytes["host_front_end_emails"]},
            "gpu_launches": stats["kernel_launches"] * K,
            "kernel_ms": fam_best, "clocks": clocks.summary(), "int_pipe_peaks": peaks, "nproc": ncpu,
        }
        print(json.dumps(line), flush=True)
    if regex is not None:
        regex.close()
    if direct:
        eng.unregister_host(pool.raw)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
