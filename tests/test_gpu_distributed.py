"""The N-GPU paths on real GPUs (skipped on a box with one): one process per visible GPU over NCCL.

  * ONE proof split over all GPUs (`parallel.DistributedCommitter`, SPMD): every rank's proof equals the
    ORACLE's golden proof of the bench instance byte for byte (2^16 - BASELINE config #2 - and the ragged
    2^10 case where pieces cut across polynomial borders),
  * the point-split MSM (`parallel.sharded_msm`): all-gather of 144-byte partial sums, closed-form expected value.
Run with `gpurun --gpus N -- python -m pytest tests -m gpu -k distributed`.
"""
import json
import os
import socket
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import hashlib
    import random

    import numpy as np
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from ark_plonk_b200 import bench_circuit as bc, encoding as enc, kzg, parallel, plonk as gp, synth
    from ark_plonk_b200._lib import get_lib
    lib = get_lib()
    lib.init(rank)
    ok = True
    try:
        for degree in (10, 16):
            case = json.load(open(os.path.join(ROOT, "tests", "golden", "plonk_bench_2p%d.json" % degree)))
            circ = bc.build(0, degree, [int(b, 16) for b in case["blinders"]])
            ck = kzg.CommitterKey.from_tau(0, int(case["tau"], 16), circ.n + 1, lib=lib)
            com = parallel.DistributedCommitter(0, ck, device="cuda", lib=lib)
            pr = gp.Prover(0, ck, lib=lib, committer=com, arena_device="cuda")
            pk = pr.preprocess(circ, commit_verifier_key=False)
            for _ in range(2):
                blob = pr.prove(pk, gp.wires_to_mont(circ), b"ark")
                ok = ok and hashlib.sha256(blob).hexdigest() == case["proof_sha256"]
            ok = ok and com.batches == 10
            pk.arena.close()
            ck.close()
        # point-split MSM: ragged split of 3001 points
        n = 3001
        pts = synth.progression_bases(0, 11, 5, n)
        sck = parallel.ShardedCommitterKey(0, enc.g1_affine_to_mont(0, pts), lib=lib)
        rnd = random.Random(99)
        s = [rnd.randrange(enc.FR_MODULUS[0]) for _ in range(n)]
        out = parallel.sharded_msm(sck, enc.ints_to_limbs(s, 4))
        ok = ok and enc.g1_from_xyz(0, out) == synth.progression_expected(0, 11, 5, s)
        sck.close()
    except Exception as e:                      # report instead of hanging the other ranks in a collective
        ok = False
        print("rank %d: %r" % (rank, e), flush=True)
    q.put((rank, bool(ok), lib.kernel_launches() > 0))
    dist.destroy_process_group()


def test_distributed_prove_and_msm_on_all_gpus():
    import torch
    import torch.multiprocessing as mp
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs at least 2 GPUs (gpurun --gpus N)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(120)
    assert res == [(r, True, True) for r in range(world)]
