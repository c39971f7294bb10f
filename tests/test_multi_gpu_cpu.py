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


def _prove_worker(rank, world, port, emu_path, result_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["APB_MSM_C"] = "8"
    os.environ["APB_NTT_MAX_LOG_TILE"] = "4"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hashlib
    from ark_plonk_b200 import bench_circuit as bc, kzg, parallel, plonk as gp
    from ark_plonk_b200._lib import Lib
    import prover_cases
    lib = Lib(emu_path)
    lib.init()
    case = prover_cases.golden_case(0, 5)
    tau = int(case["tau"], 16)
    circ = bc.build(0, 5, [int(b, 16) for b in case["blinders"]])
    ck = kzg.CommitterKey.from_tau(0, tau, circ.n + 1, lib=lib)
    com = parallel.DistributedCommitter(0, ck, group=None, device="cpu", lib=lib)
    # SPMD: every rank runs the whole prover; each commit batch is split evenly, 144-byte partials all-reduced
    # (a torch-backed arena: the coset FFTs / quotient slices of round 4 are exchanged with all-gathers on arena views)
    pr = gp.Prover(0, ck, lib=lib, committer=com, arena_device="cpu")
    pk = pr.preprocess(circ, commit_verifier_key=False)
    blob = pr.prove(pk, gp.wires_to_mont(circ), b"ark")
    ok = hashlib.sha256(blob).hexdigest() == case["proof_sha256"] and com.batches == 5   # 5 batched commit calls per proof
    ok = ok and com.gathers == 2                     # coset-FFT vectors + quotient slices
    result_q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_prove_with_commitments_split_over_two_ranks(emu_lib):
    """batched prove, SPMD: both ranks run the prover, each multiplies half of the points of every commit batch;
    the proof is byte-identical to the golden vector on BOTH ranks"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_prove_worker, args=(r, 2, port, emu_lib.path, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


def _commit_split_worker(rank, world, port, emu_path, result_q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["APB_MSM_C"] = "8"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    from ark_plonk_b200 import encoding as enc, kzg, parallel, synth
    from ark_plonk_b200._lib import Lib
    from ark_plonk_b200.plonk import Arena
    lib = Lib(emu_path)
    lib.init()
    n = 29
    pts = synth.progression_bases(0, 3, 7, n)
    ck = kzg.CommitterKey(0, enc.g1_affine_to_mont(0, pts), lib=lib)
    com = parallel.DistributedCommitter(0, ck, k_max=4, device="cpu", lib=lib)
    ok = True
    arena = Arena(lib, 8 * n)
    rnd = random.Random(5)
    polys = [[rnd.randrange(enc.FR_MODULUS[0]) for _ in range(ln)] for ln in (n, n - 1, 7, 0, n, 3)]
    offs = []
    for q in polys:
        o = arena.alloc(n)
        if q:
            arena.upload(o, enc.fr_to_mont(0, q))
        offs.append(o)
    # k = 1 < world: the polynomial is point-split; k = 3: cut across a polynomial border; k = 6 > k_max: two rounds
    for sel in ([0], [1], [0, 1, 2], [0, 1, 2, 3, 4, 5]):
        out = com.commit(arena, [offs[i] for i in sel], [len(polys[i]) for i in sel])
        for row, i in zip(out, sel):
            ok = ok and enc.g1_from_xyz(0, row) == synth.progression_expected(0, 3, 7, polys[i])
    ok = ok and parallel.split_pieces([10, 10, 10], 2) == [(0, 0, 10, 0), (1, 0, 5, 0), (1, 5, 10, 1), (2, 0, 10, 1)]
    arena.close()
    result_q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_commit_batches_point_split_when_fewer_polys_than_ranks(emu_lib):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_commit_split_worker, args=(r, 2, port, emu_lib.path, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]
