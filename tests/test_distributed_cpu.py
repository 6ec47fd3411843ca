"""World-size-2 gloo tests (CPU) of the host-side multi-GPU logic: row sharding, the one all-gather of the
(max, sum-exp) partials, and the ragged gather of per-rank bounds in eval_boed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import aline_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from aline_b200 import spce
        from aline_b200.utils.eval import gather_rows, rank_chunks, _decorrelate_rank_generators
        torch.manual_seed(0)                                  # same histories on every rank
        L, B, T = 1001, 5, 4
        seq = torch.randn(T, L + 1, B) * 20                   # per history point: accumulated log-likelihoods
        lo, hi = spce.shard_rows(L, rank, world)
        mine = seq[:, 1 + lo:1 + hi]                          # this rank's contrastive rows
        m = mine.max(1).values.T.contiguous()                 # [B, T]
        s = torch.exp(mine - mine.max(1, keepdim=True).values).sum(1).T.contiguous()
        mg, sg = spce.all_gather_partials(m, s)
        assert mg.shape == (world, B, T)
        out = O.combine_partials(mg, sg, seq[:, 0].T, L)
        ref_pce = np.log(L + 1) - (seq.logsumexp(1) - seq[:, 0]).T
        ref_nmc = np.log(L) - (seq[:, 1:].logsumexp(1) - seq[:, 0]).T
        ok1 = torch.allclose(out["pce"], ref_pce, rtol=1e-5, atol=1e-5) and torch.allclose(out["nmc"], ref_nmc, rtol=1e-5, atol=1e-5)
        # ragged gather: 7 rollouts over 2 ranks in mini-batches of <= 2 -> rank 0 owns rollouts 0..3, rank 1 owns 4..6
        n_total, bs = 7, 2
        chunks = rank_chunks(n_total, bs, rank, world)
        assert all(sz <= bs for _, sz in chunks)        # 7 rollouts, batch 2: 5 // 4 * 2 = 2, no slack at this size
        t = torch.cat([torch.arange(off, off + sz, dtype=torch.float32).reshape(-1, 1).expand(-1, 3) for off, sz in chunks], 0)
        g = gather_rows(dist, t, n_total)
        ok2 = g.shape == (n_total, 3) and g[:, 0].tolist() == [float(i) for i in range(n_total)]
        # fewer rollouts than ranks: rank 1 owns nothing and contributes an empty block (eval_boed at 8 ranks, M small)
        chunks1 = rank_chunks(1, bs, rank, world)
        t1 = torch.full((1, 3), 7.0) if chunks1 else torch.empty((0, 3))
        g1 = gather_rows(dist, t1, 1)
        ok2 = ok2 and g1.shape == (1, 3) and bool((g1 == 7.0).all())
        # identically seeded ranks are re-seeded so that they do not simulate the same rollouts
        torch.manual_seed(123)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _decorrelate_rank_generators(dist, torch.device("cpu"))
        draw = torch.rand(4)
        both = [torch.empty(4) for _ in range(world)]
        dist.all_gather(both, draw)
        ok2 = ok2 and not torch.equal(both[0], both[1])
        q.put((rank, bool(ok1), bool(ok2)))
    finally:
        dist.destroy_process_group()


def test_rank_chunks_balanced():
    from aline_b200.utils.eval import rank_chunks
    # one rank: the reference's loop -- ceil(M / batch) mini-batches of `batch` rollouts
    assert rank_chunks(2000, 200) == [(200 * i, 200) for i in range(10)]
    # 8 ranks, M = 2000, batch 200: 250 rollouts per rank in ONE mini-batch (within 25 % of the batch size), not whole
    # batches of 200 dealt round-robin; 4 ranks: 500 per rank as 2 x 250
    for r in range(8):
        assert rank_chunks(2000, 200, r, 8) == [(250 * r, 250)]
    assert rank_chunks(2000, 200, 1, 4) == [(500, 250), (750, 250)]
    for n, bs, world in ((2000, 200, 3), (7, 2, 2), (1, 5, 4), (48, 8, 8), (450, 200, 1)):
        seen = []
        for r in range(world):
            for off, sz in rank_chunks(n, bs, r, world):
                assert 1 <= sz <= (bs if world == 1 else max(bs, bs * 5 // 4))
                seen += list(range(off, off + sz))
        assert seen == list(range(n))


def test_shard_rows_partition():
    from aline_b200 import spce
    for L, R in ((10, 3), (1_000_000, 8), (7, 8), (1, 1)):
        cuts = [spce.shard_rows(L, r, R) for r in range(R)]
        assert cuts[0][0] == 0 and cuts[-1][1] == L
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(R - 1))
        sizes = [b - a for a, b in cuts]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_gloo_world2_partials_and_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert all(ok1 and ok2 for _, ok1, ok2 in res), res
