"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous sharding, the descriptor
all-gather (even and uneven shards) and both sharded top-k strategies give the single-process answer.
The local search is the oracle's exact top-k here (no GPU in this tier); on the GPU box the same code
path runs with anyloc_topk (tests/test_dist_gpu.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from anyloc_b200 import dist as adist
from oracle import anyloc_oracle as ao


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _search(db, qu, k, method, norm):
    return ao.top_k(db, qu, k, method, norm)


def _worker(rank, world, port, n_db, n_q, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(11)
        db = torch.randn(n_db, 96, generator=g)
        qu = db[torch.randint(0, n_db, (n_q,), generator=g)] + 0.2 * torch.randn(n_q, 96, generator=g)
        db[7] = db[3]                                            # duplicate rows across / within shards
        db[n_db - 1] = db[0]
        s, e = adist.shard_range(n_db)
        qs, qe = adist.shard_range(n_q)
        gathered = adist.all_gather_descriptors(db[s:e])
        assert torch.equal(gathered, db)
        full_d, full_i = ao.top_k(db, qu, 5)
        for strategy in ("gather_db", "gather_queries"):
            for method in ("cosine", "l2"):
                d, i = adist.sharded_top_k(db[s:e], qu[qs:qe], 5, method, True, strategy, _search)
                rd, ri = ao.top_k(db, qu, 5, method)
                assert torch.equal(i, ri), (strategy, method)
                assert torch.allclose(d, rd, atol=1e-6)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_db,n_q", [(40, 6), (41, 7)])
def test_sharded_pipeline_world2(n_db, n_q):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, n_db, n_q, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)


def _worker_pipelined(rank, world, port, ret):
    """dist.all_gather_into_index over gloo: pieces of the local shard are gathered and placed at rank * n_local + offset
    (the product's FlatIndex runs on the CPU double here: the collective + placement logic is what is under test)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from anyloc_b200 import utilities as u
        from tests.cpu_double import cpu_double
        g = torch.Generator().manual_seed(3)
        n_loc, d = 10, 48
        db = torch.randn(world * n_loc, d, generator=g)
        qu = db[[3, 17, 8]] + 0.1 * torch.randn(3, d, generator=g)
        with cpu_double():
            for chunks in (1, 3, 4, 50):
                ix = u.FlatIndex(d, "cosine", True, capacity=world * n_loc)
                adist.all_gather_into_index(ix, db[rank * n_loc:(rank + 1) * n_loc], chunks=chunks)
                assert ix.ntotal == world * n_loc
                di, ii = ix.search(qu, 4)
                rd, ri = ao.top_k(db, qu, 4)
                assert torch.equal(ii, ri) and torch.allclose(di, rd, atol=1e-6), chunks
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_all_gather_into_index_world2():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_pipelined, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 100000):
        for world in (1, 2, 4, 8):
            spans = [adist.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_merge_candidates_tie_and_padding():
    d = torch.tensor([[0.9, 0.5, 0.9, 0.1]])
    i = torch.tensor([[12, 3, 4, -1]])
    md, mi = adist.merge_candidates(d, i, 3, largest=True)
    assert mi.tolist() == [[4, 12, 3]] and torch.allclose(md, torch.tensor([[0.9, 0.9, 0.5]]))
