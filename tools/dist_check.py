"""Multi-GPU result-equality check, run under torchrun on N GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py
Every rank extracts + VLAD-aggregates its contiguous shard of a synthetic image set, the descriptors are
all-gathered (the pipeline's one collective), both sharded top-k strategies answer the queries, and the
results are compared with a single-GPU run of the whole set on rank 0's device (bitwise for the
descriptors -- per-image arithmetic does not depend on the batch it sits in)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from anyloc_b200 import dist as adist, utilities as u
from anyloc_b200.vit import random_state_dict


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_db, n_q, K = 24, 6, 8
    sd = random_state_dict("dinov2_vits14", seed=0, device=dev, depth=10)
    ext = u.DinoV2ExtractFeatures("dinov2_vits14", 9, "value", device=dev, weights=sd)
    g = torch.Generator(device=dev).manual_seed(1234)                    # same images on every rank
    imgs = torch.randn(n_db + n_q, 3, 224, 224, device=dev, generator=g)
    vlad = u.VLAD(K)
    vlad.kmeans = u._KMeans(K, mode="cosine")
    vlad.c_centers = vlad.kmeans.centroids = 0.7 * ext(imgs[:2]).reshape(-1, 384)[::61][:K].contiguous()
    vlad.desc_dim = 384

    def describe(x):
        return vlad.generate_multi(ext(x))

    s, e = adist.shard_range(n_db)
    qs, qe = adist.shard_range(n_q)
    db_local, qu_local = describe(imgs[s:e]), describe(imgs[n_db + qs:n_db + qe])
    db_all = adist.all_gather_descriptors(db_local)
    full_db, full_qu = describe(imgs[:n_db]), describe(imgs[n_db:])
    assert torch.equal(db_all, full_db), "sharded descriptors differ from the single-GPU run"
    ref_d, ref_i = u.top_k_search(full_db, full_qu, 5)
    for strategy in ("gather_db", "gather_queries"):
        d, i = adist.sharded_top_k(db_local, qu_local, 5, strategy=strategy)
        assert torch.equal(i, ref_i), strategy
        assert torch.allclose(d, ref_d, rtol=1e-6, atol=1e-7), strategy
    dist.barrier()
    if rank == 0:
        print(f"dist_check OK: world={world}, descriptors bitwise equal, top-5 identical for both strategies")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
