"""world_size-2 CPU (gloo) test of the multi-GPU host logic: the reference's contiguous-view sharding rule
(CRF_FeatureStreamManager.cpp:425-464), the per-step share of a minibatch (CRF_Minibatch_GradAccumulator.cpp:229-241)
and "one all-reduce of [gradient | sum numer, sum logZ, n_utt, 0] then / N_active" reproduce the single-process sum.
The per-rank compute is done by the oracle here (no GPU in this container); on the GPU box bench.py runs the
same plumbing with NCCL on the device buffer."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard_range(n, world, rank):
    """stream i owns [i*floor(n/N), (i+1)*floor(n/N)), the last one also takes n mod N"""
    per = n // world
    lo = rank * per
    hi = n if rank == world - 1 else lo + per
    return lo, hi


def minibatch_share(mb, world, rank):
    """floor(mb/N) (+1 if rank < mb mod N)"""
    return mb // world + (1 if rank < mb % world else 0)


def test_sharding_rule():
    assert [shard_range(3696, 8, r) for r in (0, 1, 7)] == [(0, 462), (462, 924), (3234, 3696)]
    assert shard_range(10, 4, 3) == (6, 10) and shard_range(10, 4, 0) == (0, 2)
    assert [minibatch_share(10, 4, r) for r in range(4)] == [3, 3, 2, 2]
    assert sum(minibatch_share(64, 8, r) for r in range(8)) == 64


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle.binding import OracleLib, make_config
    from helpers import synth_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    off, ftrs, labs = synth_batch(rng, 7, 5, 30, 6, 5)
    cfg = make_config("stdseg", n_labs=15, n_base_ftrs=6, max_dur=3, n_actual_labs=5, extract_seg_ftrs=1)
    ora = OracleLib()
    lam = np.random.default_rng(4).uniform(-0.05, 0.05, ora.lambda_len(cfg))
    lo, hi = shard_range(7, world, rank)
    so = (off[lo:hi + 1] - off[lo]).astype(np.uint32)
    g, n, z = ora.fwdbwd(cfg, lam, so, ftrs[off[lo]:off[hi]], labs[off[lo]:off[hi]])
    buf = torch.from_numpy(np.concatenate([g, [n.sum(), z.sum(), hi - lo, 0.0]]))
    dist.all_reduce(buf)                      # the ONE collective per minibatch
    if rank == 0:
        gf, nf, zf = ora.fwdbwd(cfg, lam, off, ftrs, labs)
        q.put((buf.numpy().copy(), gf, nf.sum(), zf.sum()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_process():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    buf, gf, nf, zf = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_allclose(buf[:-4], gf, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(buf[-4:], [nf, zf, 7.0, 0.0], rtol=1e-12)
