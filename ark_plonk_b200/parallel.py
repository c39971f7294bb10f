"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for the plumbing.

* `ShardedCommitterKey` / `sharded_msm`: the point-split MSM of SURVEY.md section 8(e).  Rank r keeps
  slice r of `powers_of_g` resident (with its precomputed table), multiplies its slice of the
  scalars, and the per-rank partial sums (one normalised point, 144 bytes) are exchanged with a
  single all-gather; every rank folds the G partials on the host (G-1 group additions).
  NCCL has no elliptic-curve reduction, so reduction = gather + local adds.
* A single NTT does not shard (replicas only); independent polynomials / proofs are distributed
  across ranks by the caller (`bench.py --workload prove --gpus N`).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import kzg
from ._lib import Lib, get_lib


def shard_bounds(n: int, world: int, rank: int):
    """contiguous slice [lo, hi) of n items owned by `rank`"""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


class ShardedCommitterKey:
    def __init__(self, curve: int, powers_of_g_mont: np.ndarray, group=None, lib: Lib | None = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.curve = curve
        self.lib = lib or get_lib()
        pts = np.ascontiguousarray(powers_of_g_mont, dtype=np.uint64).reshape(-1, 12)
        self.n = pts.shape[0]
        self.lo, self.hi = shard_bounds(self.n, self.world, self.rank)
        self.local = kzg.CommitterKey(curve, pts[self.lo:self.hi], lib=self.lib)

    def close(self):
        self.local.close()


def sharded_msm(ck: ShardedCommitterKey, scalars: np.ndarray, montgomery: bool = False, device: str | None = None):
    """sum_i scalars[i] * bases[i] over the whole key: each rank multiplies its slice, partials
    are all-gathered and folded.  `scalars`: the full (n, 4) vector (every rank slices its part)."""
    s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    lo, hi = min(ck.lo, s.shape[0]), min(ck.hi, s.shape[0])
    part = kzg.multi_scalar_mul(ck.local, s[lo:hi], montgomery=montgomery)
    if ck.world == 1:
        return part
    dev = device or ("cuda" if dist.get_backend(ck.group) == "nccl" else "cpu")
    mine = torch.from_numpy(part.view(np.int64).copy()).to(dev)
    gathered = torch.empty(ck.world * 18, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gathered, mine, group=ck.group)
    parts = gathered.cpu().numpy().view(np.uint64).reshape(ck.world, 18)
    acc = parts[0]
    for r in range(1, ck.world):
        acc = ck.lib.g1_add(ck.curve, acc, parts[r])
    return acc


class DistributedCommitter:
    """Per-polynomial task split of `PC::commit` over the ranks of a process group (SURVEY.md 8e,
    "batched prove").  Every rank holds the full commitment key resident; rank 0 runs the prover and,
    for each batch of k polynomials, broadcasts the coefficient vectors (k x n x 32 B over
    NVLink), every rank multiplies the polynomials j with j % world == rank, and the normalised
    results (144 B each) are combined with one all-reduce.  Workers sit in `serve()`.
    The transcript, NTTs and pointwise kernels stay on rank 0 (they depend on each commitment).
    """
    OP_STOP, OP_COMMIT = 0, 1

    def __init__(self, curve: int, ck: "kzg.CommitterKey", n_max: int, k_max: int = 8, group=None, device: str = "cuda",
                 lib: Lib | None = None):
        self.curve, self.ck, self.n, self.k_max, self.group, self.device = curve, ck, n_max, k_max, group, device
        self.lib = lib or get_lib()
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.header = torch.zeros(2 + k_max, dtype=torch.int64, device=device)
        self.stage = torch.zeros(k_max * n_max * 4, dtype=torch.int64, device=device)
        self.results = torch.zeros(k_max * max(self.world, 1) * 18, dtype=torch.int64, device=device)
        self._stream = None
        if device == "cuda":
            self._stream = torch.cuda.ExternalStream(self.lib.c.apb_stream())

    def _ctx(self):
        import contextlib
        return torch.cuda.stream(self._stream) if self._stream is not None else contextlib.nullcontext()

    def _tasks(self, k: int, lens):
        """work of this rank for a batch of k polynomials: [(poly j, lo, hi, result slot)], slices per poly S.
        k >= world: whole polynomials round-robin; k < world: every polynomial is cut into world // k
        contiguous coefficient ranges (point split inside the polynomial split)"""
        W, r = self.world, self.rank
        if k >= W:
            return [(j, 0, int(lens[j]), j) for j in range(k) if j % W == r], 1
        S = W // k
        if r >= k * S:
            return [], S
        j, sidx = r % k, r // k
        per = (int(lens[j]) + S - 1) // S
        lo = min(sidx * per, int(lens[j]))
        hi = min(lo + per, int(lens[j]))
        return ([(j, lo, hi, j * S + sidx)] if hi > lo else []), S

    def _local_msms(self, k: int, lens) -> int:
        """MSMs of this rank's share of the staged polynomials -> self.results slots (others zero)"""
        import ctypes as C
        tasks, S = self._tasks(k, lens)
        self.results.zero_()
        if tasks:
            m = len(tasks)
            so = (C.c_size_t * m)(*[j * self.n + lo for j, lo, _, _ in tasks])
            bo = (C.c_size_t * m)(*[lo for _, lo, _, _ in tasks])
            ln = (C.c_size_t * m)(*[hi - lo for _, lo, hi, _ in tasks])
            out = np.zeros((m, 18), dtype=np.uint64)
            self.lib.check(self.lib.c.apb_msm_batch_dev(self.ck._h, m, self.stage.data_ptr(), so, bo, ln, 1, out.ctypes.data))
            host = torch.from_numpy(out.view(np.int64))
            for i, (_, _, _, slot) in enumerate(tasks):
                self.results[slot * 18:(slot + 1) * 18].copy_(host[i])
        return S

    def commit(self, arena, offs, lens) -> np.ndarray:
        """rank 0: commitments of the k polynomials at arena offsets `offs` -> (k, 18) uint64"""
        k = len(offs)
        out = np.zeros((k, 18), dtype=np.uint64)
        for base in range(0, k, self.k_max):
            kk = min(self.k_max, k - base)
            with self._ctx():
                hdr = [self.OP_COMMIT, kk] + [int(x) for x in lens[base:base + kk]] + [0] * (self.k_max - kk)
                self.header.copy_(torch.tensor(hdr, dtype=torch.int64))
                for j in range(kk):
                    self.stage[j * self.n * 4:(j * self.n + int(lens[base + j])) * 4].copy_(arena.view(offs[base + j], int(lens[base + j])))
                dist.broadcast(self.header, src=0, group=self.group)
                dist.broadcast(self.stage[: kk * self.n * 4], src=0, group=self.group)
                if self._stream is not None:
                    self._stream.synchronize()
                S = self._local_msms(kk, lens[base:base + kk])
                dist.all_reduce(self.results, group=self.group)
                parts = self.results[: kk * S * 18].cpu().numpy().view(np.uint64).reshape(kk, S, 18)
                for j in range(kk):
                    acc = parts[j, 0]
                    for sidx in range(1, S):               # fold the point slices of polynomial j (host, 144-byte points)
                        acc = self.lib.g1_add(self.curve, acc, parts[j, sidx])
                    out[base + j] = acc
        return out

    def serve(self) -> int:
        """worker ranks: answer commit requests until rank 0 calls `shutdown`; returns #batches served"""
        served = 0
        while True:
            with self._ctx():
                dist.broadcast(self.header, src=0, group=self.group)
                hdr = self.header.cpu().tolist()
                if hdr[0] == self.OP_STOP:
                    return served
                kk = hdr[1]
                dist.broadcast(self.stage[: kk * self.n * 4], src=0, group=self.group)
                if self._stream is not None:
                    self._stream.synchronize()
                self._local_msms(kk, hdr[2:2 + kk])
                dist.all_reduce(self.results, group=self.group)
            served += 1

    def shutdown(self):
        if self.world > 1 and self.rank == 0:
            with self._ctx():
                self.header.zero_()
                dist.broadcast(self.header, src=0, group=self.group)
