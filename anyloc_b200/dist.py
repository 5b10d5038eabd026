"""Multi-GPU plumbing for the descriptor pipeline: one process per GPU (`torchrun`), images sharded
by contiguous ranges, replicated weights / vocabulary, and exactly one collective on the data path --
the all-gather of the final [n_local, K*D] descriptors before retrieval (BASELINE.json config 4).
The reference has no multi-GPU path; `torch.distributed` (NCCL over NVLink on the
GPU box, gloo in the CPU tests) is the transport.

`sharded_top_k` offers the two exchange patterns of SURVEY.md 8(e):
  * "gather_db"      : all-gather the database descriptors, every rank answers its own query shard
                       against the full database (the pattern BASELINE config 4 names);
  * "gather_queries" : keep the database sharded, all-gather the (small) query set, local top-k with
                       global index offsets, all-gather the [n_q, k] candidates and merge -- ~1000x
                       less traffic, identical results.
Both return identical (distances, indices) on every rank for the full query set.
"""
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def shard_range(n: int, world: Optional[int] = None, rank: Optional[int] = None) -> Tuple[int, int]:
    """Contiguous [start, end) of `n` items owned by `rank` (first n % world ranks get one extra)."""
    w, r = world_info()
    world = w if world is None else world
    rank = r if rank is None else rank
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_gather_rows(local: torch.Tensor) -> torch.Tensor:
    """Concatenate per-rank row blocks [n_r, ...] in rank order (uneven n_r allowed)."""
    world, _ = world_info()
    if world == 1:
        return local
    local = local.contiguous()
    n = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    n_max = max(counts)
    if all(c == n_max for c in counts):
        out = torch.empty((world * n_max,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(out, local)
        return out
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    pad[:local.shape[0]] = local
    out = torch.empty((world * n_max,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * n_max:r * n_max + c] for r, c in enumerate(counts)])


def all_gather_descriptors(local_desc: torch.Tensor) -> torch.Tensor:
    """The one data-path collective of the pipeline: [n_local, K*D] fp32 -> [n_total, K*D]."""
    return all_gather_rows(local_desc)


def all_gather_descriptors_cabi(local_desc: torch.Tensor) -> torch.Tensor:
    """The same collective through the C ABI (`anyloc_allgather_desc`, include/anyloc_b200.h) on the NCCL communicator
    of the default process group: what a host that is not PyTorch would bind.  Equal shard sizes; enqueued on the
    current stream (the caller must not have other work in flight on the communicator)."""
    from . import _lib
    world, _ = world_info()
    if world == 1:
        return local_desc
    backend = dist.distributed_c10d._get_default_group()._get_backend(local_desc.device)
    comm = backend._comm_ptr()
    local_desc = local_desc.contiguous()
    out = torch.empty((world * local_desc.shape[0], local_desc.shape[1]), device=local_desc.device, dtype=torch.float32)
    with torch.cuda.device(local_desc.device):
        _lib.check(_lib.load().anyloc_allgather_desc(comm, _lib.ptr(local_desc), _lib.ptr(out), local_desc.shape[0],
                                                     local_desc.shape[1], _lib.stream_ptr()), "anyloc_allgather_desc")
    return out


def all_gather_into_index(index, db_local: torch.Tensor, staging: Optional[torch.Tensor] = None, chunks: int = 4):
    """The descriptor all-gather of BASELINE config 4 pipelined with the database preparation: the local shard is
    all-gathered in `chunks` pieces (asynchronously, back to back on the NCCL stream) and every piece is prepared into
    the index (`FlatIndex.add_at`, rank r's rows at r * n_local + offset) as soon as it has landed, so the index build of
    piece c hides behind the transfer of piece c+1.  Equal shard sizes on every rank.  `index` must have capacity
    world * n_local; `staging` ([world * n_local, Dv], optional) receives the raw gathered pieces (piece-major).
    Results are identical to `index.add(all_gather_descriptors(db_local))`."""
    world, rank = world_info()
    n_loc, Dv = db_local.shape
    index.reset()
    if world == 1:
        index.add(db_local)
        return index
    if staging is None:
        staging = torch.empty(world * n_loc, Dv, device=db_local.device, dtype=db_local.dtype)
    chunks = max(1, min(chunks, n_loc))
    bounds = [n_loc * c // chunks for c in range(chunks + 1)]
    works = []
    for c in range(chunks):
        m = bounds[c + 1] - bounds[c]
        out = staging[world * bounds[c]: world * bounds[c + 1]]
        works.append((dist.all_gather_into_tensor(out, db_local[bounds[c]:bounds[c + 1]].contiguous(), async_op=True),
                      out, bounds[c], m))
    for work, out, off, m in works:
        work.wait()
        for r in range(world):
            index.add_at(out[r * m:(r + 1) * m], r * n_loc + off)
    return index


def merge_candidates(dist_c: torch.Tensor, idx_c: torch.Tensor, k: int, largest: bool):
    """k best of per-shard candidates [n_q, C] (distance, global index); ties -> lowest index; -1 = padding."""
    invalid = idx_c < 0
    key = torch.where(invalid, torch.full_like(dist_c, float("-inf") if largest else float("inf")), dist_c)
    big = torch.iinfo(torch.int64).max
    order = torch.sort(torch.where(invalid, torch.full_like(idx_c, big), idx_c), dim=1, stable=True)[1]
    key, idx_s, dist_s = torch.gather(key, 1, order), torch.gather(idx_c, 1, order), torch.gather(dist_c, 1, order)
    order = torch.sort(-key if largest else key, dim=1, stable=True)[1][:, :k]
    return torch.gather(dist_s, 1, order), torch.gather(idx_s, 1, order)


def sharded_top_k(db_local: torch.Tensor, qu_local: torch.Tensor, k: int, method: str = "cosine",
                  norm_descs: bool = True, strategy: str = "gather_db",
                  search: Optional[Callable] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """db_local: this rank's database rows; qu_local: this rank's query rows (both contiguous shards in
    rank order).  Returns (distances [n_q_total, k], global indices [n_q_total, k]) on every rank."""
    if search is None:
        from .utilities import top_k_search as search
    world, rank = world_info()
    if world == 1:
        return search(db_local, qu_local, k, method, norm_descs)
    if strategy == "gather_db":
        db_all = all_gather_descriptors(db_local)
        d, i = search(db_all, qu_local, k, method, norm_descs)
        return all_gather_rows(d), all_gather_rows(i)
    if strategy == "gather_queries":
        qu_all = all_gather_rows(qu_local)
        n_loc = torch.tensor([db_local.shape[0]], device=db_local.device, dtype=torch.int64)
        counts = [torch.zeros_like(n_loc) for _ in range(world)]
        dist.all_gather(counts, n_loc)
        offset = int(sum(int(c.item()) for c in counts[:rank]))
        d, i = search(db_local, qu_all, k, method, norm_descs)
        i = torch.where(i >= 0, i + offset, i)
        d_all = [torch.empty_like(d) for _ in range(world)]
        i_all = [torch.empty_like(i) for _ in range(world)]
        dist.all_gather(d_all, d.contiguous())
        dist.all_gather(i_all, i.contiguous())
        return merge_candidates(torch.cat(d_all, dim=1), torch.cat(i_all, dim=1), k, largest=(method == "cosine"))
    raise ValueError(f"unknown strategy {strategy!r}")
