"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for the plumbing.

* `ShardedCommitterKey` / `sharded_msm`: the point-split MSM of SURVEY.md section 8(e).  Rank r keeps
  slice r of `powers_of_g` resident (with its precomputed table), multiplies its slice of the
  scalars, and the per-rank partial sums (one normalised point, 144 bytes) are exchanged with a
  single all-gather; every rank folds the G partials on the host (G-1 group additions).
  NCCL has no elliptic-curve reduction, so reduction = gather + local adds.
* `DistributedCommitter`: one proof over N GPUs, SPMD - every rank runs the prover, the MSM work of
  every commit batch is split evenly, only 144-byte partial sums are exchanged; the independent 4n coset FFTs of
  the quotient round are spread by polynomial and the quotient evaluation by index range, both re-assembled with
  an in-place NVLink all-gather (`all_gather_inplace`).
* A single NTT does not shard (replicas only).
"""
from __future__ import annotations

import time

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


def split_pieces(lens, world: int):
    """Equal-work split of a batch of polynomials over `world` ranks: the coefficient ranges of the k
    polynomials are laid end to end and cut into `world` equal contiguous parts, so every rank multiplies
    the same number of points whatever k is (k = 3 over 2 ranks, k = 2 over 8, ...).  Returns the ordered
    piece list [(poly j, lo, hi, owner rank)]; at most k + world - 1 pieces."""
    lens = [int(x) for x in lens]
    total = sum(lens)
    pieces = []
    if total == 0:
        return pieces
    cuts = [total * r // world for r in range(world + 1)]
    start = 0
    for j, ln in enumerate(lens):
        end = start + ln
        for r in range(world):
            lo, hi = max(start, cuts[r]), min(end, cuts[r + 1])
            if hi > lo:
                pieces.append((j, lo - start, hi - start, r))
        start = end
    return pieces


class DistributedCommitter:
    """SPMD split of `PC::commit` over the ranks of a process group ("batched prove").

    Every rank runs the WHOLE prover on the same witness (same transcript, same challenges) and holds the
    full commitment key resident, so no polynomial ever crosses NVLink: for each batch of k polynomials a
    rank multiplies only its pieces (`split_pieces`: equal point counts per rank, cut across polynomial
    borders), writes the normalised partial sums (144 B each) into its slots of a small result vector,
    and ONE all-reduce of that vector (<= (k + world - 1) x 144 B) gives every rank all partial sums; each
    rank folds the pieces of a polynomial on the host (a handful of group additions) and continues with
    identical commitments.  Round 1 broadcast the coefficient vectors from rank 0 instead (64 MiB per
    8-polynomial batch) and kept NTTs / pointwise work on rank 0 only.
    """

    def __init__(self, curve: int, ck: "kzg.CommitterKey", n_max: int = 0, k_max: int = 16, group=None, device: str = "cuda",
                 lib: Lib | None = None):
        self.curve, self.ck, self.k_max, self.group, self.device = curve, ck, k_max, group, device
        self.lib = lib or get_lib()
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.results = torch.zeros((k_max + max(self.world, 1)) * 18, dtype=torch.int64, device=device)
        self.batches = 0
        # collectives on arena slices are enqueued behind the kernels of the library's stream (no host synchronisation)
        self.lib_stream = torch.cuda.ExternalStream(self.lib.c.apb_stream()) if str(device).startswith("cuda") else None
        self.gathers = 0
        # host staging of the partial sums (pinned on a GPU group): one copy in, one copy out per batch
        self.h_stage = torch.zeros_like(self.results, device="cpu")
        if self.lib_stream is not None:
            self.h_stage = self.h_stage.pin_memory()
        self.exchange_ms = 0.0         # wall time of the partial-sum exchanges (includes waiting for the slowest rank)
        self.fold_ms = 0.0             # host folding of the pieces

    def all_gather_inplace(self, arena, off: int, chunk_elems: int):
        """The `world` chunks of `chunk_elems` Fr elements at arena offset `off`: chunk r is valid on rank r when this
        is called; on return (stream order) every rank holds all of them.  Used by the prover to spread independent
        transforms / index ranges over the ranks: NVLink all-gather of evaluation vectors."""
        if self.world == 1:
            return
        whole = arena.view(off, self.world * chunk_elems)
        mine = arena.view(off + self.rank * chunk_elems, chunk_elems)
        if self.lib_stream is not None:
            with torch.cuda.stream(self.lib_stream):
                dist.all_gather_into_tensor(whole, mine, group=self.group)
        else:
            dist.all_gather_into_tensor(whole, mine, group=self.group)
        self.gathers += 1

    def commit(self, arena, offs, lens) -> np.ndarray:
        """called by EVERY rank with its own copy of the k polynomials at arena element offsets `offs`:
        commitments of all k polynomials -> (k, 18) uint64, identical on every rank"""
        import ctypes as C
        k = len(offs)
        out = np.zeros((k, 18), dtype=np.uint64)
        for base in range(0, k, self.k_max):
            kk = min(self.k_max, k - base)
            pieces = split_pieces(lens[base:base + kk], self.world)
            mine = [(i, p) for i, p in enumerate(pieces) if p[3] == self.rank]
            nres = max(len(pieces), 1) * 18
            res, h = self.results[:nres], self.h_stage[:nres]
            h.zero_()
            if mine:
                m = len(mine)
                so = (C.c_size_t * m)(*[int(offs[base + j]) + lo for _, (j, lo, _, _) in mine])
                bo = (C.c_size_t * m)(*[lo for _, (_, lo, _, _) in mine])
                ln = (C.c_size_t * m)(*[hi - lo for _, (_, lo, hi, _) in mine])
                part = np.zeros((m, 18), dtype=np.uint64)
                self.lib.check(self.lib.c.apb_msm_batch_dev(self.ck._h, m, arena.base, so, bo, ln, 1, part.ctypes.data))
                first = mine[0][0]                      # a rank's pieces are consecutive slots
                h[first * 18:(first + m) * 18] = torch.from_numpy(part.view(np.int64).reshape(-1))
            t0 = time.perf_counter()
            if self.world > 1:
                res.copy_(h, non_blocking=True)
                dist.all_reduce(res, group=self.group)          # slots are disjoint: the sum is a gather
                h.copy_(res, non_blocking=True)
                if self.lib_stream is not None:
                    torch.cuda.current_stream().synchronize()
            parts = h.numpy().view(np.uint64).reshape(-1, 18)
            t1 = time.perf_counter()
            if pieces:                                    # one C call: sums per polynomial, one shared inversion
                out[base:base + kk] = self.lib.g1_fold(self.curve, parts[:len(pieces)], [j for j, _, _, _ in pieces], kk)
            self.exchange_ms += (t1 - t0) * 1e3
            self.fold_ms += (time.perf_counter() - t1) * 1e3
            self.batches += 1
        return out
