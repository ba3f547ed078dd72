"""Aggregate pinned host->device bandwidth with all ranks copying at once (torchrun).

Explains the e2e line at N>1: the end-to-end path is PCIe/host-memory bound, so its scaling is set by what
the box can stream to N GPUs concurrently, not by the kernels."""
import os, time, torch, torch.distributed as dist

def main():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    nbytes = 1 << 30
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h.fill_(rank + 1)
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for solo in (True, False):
        res = []
        for r in range(world if solo else 1):
            if world > 1: dist.barrier()
            torch.cuda.synchronize()
            active = (not solo) or r == rank
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            if active:
                for _ in range(4): d.copy_(h, non_blocking=True)
            ev1.record(); torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            if active: res.append(4 * nbytes / ms / 1e6)
            if world > 1: dist.barrier()
        t = torch.tensor([res[0]], device="cuda")
        if world > 1:
            g = [torch.zeros_like(t) for _ in range(world)]; dist.all_gather(g, t); g = [float(x) for x in g]
        else: g = [float(t)]
        if rank == 0:
            print(("solo      " if solo else "concurrent"), "GB/s per rank:", " ".join(f"{x:.1f}" for x in g), " sum", f"{sum(g):.1f}", flush=True)
    if world > 1: dist.destroy_process_group()

if __name__ == "__main__":
    main()
