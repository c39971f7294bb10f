"""The N>1 path on CPU: world_size-2 gloo processes run the point-split MSM (per-rank partial
sums via the kernel-logic emulation build, all-gather, host fold) and agree with the oracle."""
import os
import random
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, emu_path, result_q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["APB_MSM_C"] = "8"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ark_plonk_b200 import encoding as enc
    from ark_plonk_b200 import parallel, synth
    from ark_plonk_b200._lib import Lib
    lib = Lib(emu_path)
    lib.init()
    n = 37                                       # ragged: 19 + 18
    pts = synth.progression_bases(0, 11, 5, n)
    ck = parallel.ShardedCommitterKey(0, enc.g1_affine_to_mont(0, pts), lib=lib)
    rnd = random.Random(99)
    s = [rnd.randrange(enc.FR_MODULUS[0]) for _ in range(n)]
    out = parallel.sharded_msm(ck, enc.ints_to_limbs(s, 4))
    exp = synth.progression_expected(0, 11, 5, s)
    ok = enc.g1_from_xyz(0, out) == exp and (ck.lo, ck.hi) == parallel.shard_bounds(n, world, rank)
    # shorter scalar vector than the key (polynomial of lower degree): last rank may get nothing
    s2 = s[:10]
    out2 = parallel.sharded_msm(ck, enc.ints_to_limbs(s2, 4))
    ok = ok and enc.g1_from_xyz(0, out2) == synth.progression_expected(0, 11, 5, s2)
    result_q.put((rank, bool(ok)))
    ck.close()
    dist.destroy_process_group()


def test_point_split_msm_world2(emu_lib):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, emu_lib.path, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


def test_shard_bounds_cover_everything():
    from ark_plonk_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 1000):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
