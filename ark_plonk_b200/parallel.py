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
