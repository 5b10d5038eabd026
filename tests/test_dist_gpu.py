"""Multi-GPU result equality on real GPUs (needs >= 2 GPUs, `gpurun --gpus 2 -- python -m pytest
tests/test_dist_gpu.py -m gpu`): one process per GPU over NCCL, every rank extracts + VLAD-aggregates its contiguous
shard of a synthetic image set, the descriptors go through the pipeline's one collective (all-gather), both sharded
top-k strategies answer the queries -- and everything must equal the single-GPU run of the whole set: descriptors
BITWISE (per-image arithmetic does not depend on the batch an image sits in), top-k indices identical, distances
to 1e-6.  Also the chunked form bench.py uses (async all-gather of step chunks into a replicated database that a
FlatIndex ingests chunk by chunk).  Reference semantics: /root/reference/utilities.py:433-450."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from anyloc_b200 import dist as adist, utilities as u
        from anyloc_b200.vit import random_state_dict
        n_db, n_q, K, D = 26, 7, 8, 384                       # uneven shards on purpose
        sd = random_state_dict("dinov2_vits14", seed=0, device=dev, depth=10)
        ext = u.DinoV2ExtractFeatures("dinov2_vits14", 9, "value", device=dev, weights=sd, precision="f16x3")
        g = torch.Generator(device=dev).manual_seed(1234)    # same images on every rank
        imgs = torch.randn(n_db + n_q, 3, 224, 224, device=dev, generator=g)
        imgs[n_db:] = imgs[:n_q] + 0.05 * imgs[n_db:]         # queries resemble database images 0..n_q-1
        imgs[5] = imgs[2]                                     # a duplicate database image: tie -> lowest index
        vlad = u.VLAD(K)
        vlad.kmeans = u._KMeans(K, mode="cosine")
        c = 0.7 * ext(imgs[:2]).reshape(-1, D)[::61][:K].contiguous()
        dist.broadcast(c, 0)
        vlad.c_centers = vlad.kmeans.centroids = c
        vlad.desc_dim = D

        def describe(x):
            return vlad.generate_multi(ext(x))

        s, e = adist.shard_range(n_db)
        qs, qe = adist.shard_range(n_q)
        db_local, qu_local = describe(imgs[s:e]), describe(imgs[n_db + qs:n_db + qe])
        db_all = adist.all_gather_descriptors(db_local)       # THE collective
        full_db, full_qu = describe(imgs[:n_db]), describe(imgs[n_db:])
        assert torch.equal(db_all, full_db), "sharded descriptors differ from the single-GPU run"
        ref_d, ref_i = u.top_k_search(full_db, full_qu, 5)
        assert ref_i[2, 0].item() == 2 and ref_i[2, 1].item() == 5      # duplicate rows: lowest index first
        for strategy in ("gather_db", "gather_queries"):
            d, i = adist.sharded_top_k(db_local, qu_local, 5, strategy=strategy)
            assert torch.equal(i, ref_i), strategy
            assert torch.allclose(d, ref_d, rtol=1e-6, atol=1e-7), strategy
        # chunked, asynchronous form (bench.py's step): equal-sized step chunks, gathered while the next chunk is built
        B = 4
        chunks = []
        works = []
        index = u.FlatIndex(K * D, "cosine", True, capacity=3 * world * B, device=dev)
        for step in range(3):
            mine = describe(imgs[(step * world + rank) * B:(step * world + rank + 1) * B])
            out = torch.empty(world * B, K * D, device=dev)
            works.append((dist.all_gather_into_tensor(out, mine, async_op=True), out))
        for w, out in works:
            w.wait()
            index.add(out)
            chunks.append(out)
        db_chunked = torch.cat(chunks)
        assert torch.equal(db_chunked, describe(imgs[:3 * world * B])), "chunked all-gather differs from one-GPU build"
        d, i = index.search(full_qu, 5)
        rd, ri = u.top_k_search(db_chunked, full_qu, 5)
        assert torch.equal(i, ri) and torch.equal(d, rd)
        # the pipelined form of the config-4 exchange: all-gather in pieces, each prepared into the index on arrival
        n_eq = 12                                             # equal shards
        loc = full_db[rank * n_eq:(rank + 1) * n_eq].contiguous()
        ix2 = u.FlatIndex(K * D, "cosine", True, capacity=world * n_eq, device=dev)
        adist.all_gather_into_index(ix2, loc, chunks=3)
        d3, i3 = ix2.search(full_qu, 5)
        rd3, ri3 = u.top_k_search(full_db[:world * n_eq].contiguous(), full_qu, 5)
        assert torch.equal(i3, ri3) and torch.equal(d3, rd3)
        # the collective through the C ABI (anyloc_allgather_desc on the process group's ncclComm_t) == torch.distributed's
        torch.cuda.synchronize()
        dist.barrier()
        via_abi = adist.all_gather_descriptors_cabi(loc)
        torch.cuda.synchronize()
        assert torch.equal(via_abi, full_db[:world * n_eq])
        dist.barrier()
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_multi_gpu_equals_single_gpu(cuda, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))
