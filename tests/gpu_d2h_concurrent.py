"""Microbenchmark (not a test): device -> pinned-host copy bandwidth of ONE GPU alone and of all GPUs of the box at once.
Backs the end-to-end scaling figure of bench.py: at N GPUs every step returns m p n / N doubles per rank to the host, and
the ranks share the host's memory system.  Usage:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/gpu_d2h_concurrent.py"""
import json
import os
import time

import torch
import torch.distributed as dist


def bw(dst, src, reps, stream):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        dst.copy_(src, non_blocking=True)
        stream.synchronize()
        e0.record(stream)
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        e1.record(stream)
        stream.synchronize()
    return reps * src.numel() * src.element_size() / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gb = 2
    dev = torch.empty(gb << 27, dtype=torch.float64, device="cuda").normal_()
    host = torch.empty(gb << 27, dtype=torch.float64, pin_memory=True)
    st = torch.cuda.Stream()
    out = {}
    for direction, (d, s) in (("d2h", (host, dev)), ("h2d", (dev, host))):
        alone = None
        if world > 1:
            dist.barrier()
        if rank == 0:
            alone = bw(d, s, 6, st)
        if world > 1:
            dist.barrier()
        t = torch.tensor([bw(d, s, 6, st)], dtype=torch.float64, device="cuda")
        if world > 1:
            allv = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
            per = [float(v.item()) for v in allv]
        else:
            per = [float(t.item())]
        out[direction] = {"rank0_alone_gbs": alone, "concurrent_per_rank_gbs": per, "concurrent_sum_gbs": sum(per)}
    # both directions at once on every rank (what the sampling pipeline does: X chunks in, X-tilde chunks out)
    host2 = torch.empty(gb << 27, dtype=torch.float64, pin_memory=True)
    dev2 = torch.empty_like(dev)
    st2 = torch.cuda.Stream()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        with torch.cuda.stream(st):
            host.copy_(dev, non_blocking=True)
        with torch.cuda.stream(st2):
            dev2.copy_(host2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([4 * dev.numel() * 8 / dt / 1e9], dtype=torch.float64, device="cuda")
    if world > 1:
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        per = [float(v.item()) for v in allv]
    else:
        per = [float(t.item())]
    out["bidirectional_each_way"] = {"concurrent_per_rank_gbs": per, "concurrent_sum_gbs": sum(per)}
    if rank == 0:
        out["n_gpus"] = world
        out["buffer_gb"] = gb
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
