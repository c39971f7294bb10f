#!/usr/bin/env python3
"""Latency of the tiny partial-sum exchange of the N-GPU commit split (one process per GPU, torchrun): the current
pinned-copy + all_reduce + copy-back against all_gather variants.  Prints one line per variant on rank 0."""
import os, time
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world, rank = dist.get_world_size(), dist.get_rank()
n = 24 * 18
res = torch.zeros(n, dtype=torch.int64, device="cuda")
h = torch.zeros(n, dtype=torch.int64).pin_memory()
blk = torch.zeros(3 * 18, dtype=torch.int64, device="cuda")
gath = torch.zeros(world * 3 * 18, dtype=torch.int64, device="cuda")
hb = torch.zeros(3 * 18, dtype=torch.int64).pin_memory()
hg = torch.zeros(world * 3 * 18, dtype=torch.int64).pin_memory()
resf = res.view(torch.float64)

def v_allreduce():
    res.copy_(h, non_blocking=True); dist.all_reduce(res); h.copy_(res, non_blocking=True); torch.cuda.current_stream().synchronize()
def v_allreduce_f64():
    res.copy_(h, non_blocking=True); dist.all_reduce(resf); h.copy_(res, non_blocking=True); torch.cuda.current_stream().synchronize()
def v_allgather():
    blk.copy_(hb, non_blocking=True); dist.all_gather_into_tensor(gath, blk); hg.copy_(gath, non_blocking=True); torch.cuda.current_stream().synchronize()
def v_allgather_only():
    dist.all_gather_into_tensor(gath, blk); torch.cuda.current_stream().synchronize()
def v_barrier_only():
    dist.barrier()
for name, fn in (("all_reduce int64 + copies", v_allreduce), ("all_reduce as f64 + copies", v_allreduce_f64),
                 ("all_gather + copies", v_allgather), ("all_gather only", v_allgather_only), ("barrier", v_barrier_only)):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200): fn()
    dt = (time.perf_counter() - t0) / 200 * 1e6
    # with 2 ms of GPU work on rank-dependent streams in between (as in the prover): does an idle gap change the latency?
    x = torch.zeros(1 << 24, device="cuda")
    ts = []
    for _ in range(30):
        for _ in range(8): x.add_(1.0)
        torch.cuda.synchronize()
        t1 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t1)
    if rank == 0:
        print("%-28s world %d  back-to-back %.1f us   after compute %.1f us (median)" % (name, world, dt, sorted(ts)[len(ts) // 2] * 1e6), flush=True)
dist.destroy_process_group()
