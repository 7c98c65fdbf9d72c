"""Host-side sharding plan for one 8xB200 box: one process per GPU.

* extraction: frames are independent -> contiguous frame ranges per rank, NO collective;
* Hamming kNN: the descriptor database is split in contiguous row ranges (so a global index is
  base + local row and "lowest index wins" ties stay exact); queries are replicated; per-shard
  (d1, idx1, d2) triples are all-gathered (NCCL on GPUs, gloo in the CPU tests) into a
  [world][3][nq] buffer that orbm_merge_shards_device folds with shard_stride = 3 * nq.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world):
    """Contiguous, ordered, exhaustive ranges: rank r owns [b[r], b[r+1])."""
    return [n * r // world for r in range(world + 1)]


def frame_shard(nframes, rank, world):
    b = shard_bounds(nframes, world)
    return b[rank], b[rank + 1]


def db_shard(ndb, rank, world):
    b = shard_bounds(ndb, world)
    return b[rank], b[rank + 1]


def gather_triples(tri, world):
    """tri: int32 tensor [3][nq] (d1, idx1 with GLOBAL indices, d2) of this rank's shard.
    Returns [world][3][nq] in rank order (= ascending index-range order, as the merge requires)."""
    if world == 1:
        return tri.unsqueeze(0)
    out = torch.empty((world,) + tuple(tri.shape), dtype=tri.dtype, device=tri.device)
    # concatenated-along-dim-0 view: the layout both NCCL and gloo accept for _allgather_base
    dist.all_gather_into_tensor(out.view((world * tri.shape[0],) + tuple(tri.shape[1:])), tri.contiguous())
    return out


def exchange_handles(matcher, max_queries, rank, world):
    """Create this rank's peer-exchange buffer (orbm_exchange_create) and all-gather the 64-byte CUDA IPC handles.
    Returns the handles in rank order, ready for matcher.exchange_open()."""
    h = matcher.exchange_create(max_queries, rank, world)
    if world == 1:
        return [h]
    mine = torch.tensor(list(h), dtype=torch.uint8)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = mine.to(dev)
    allh = torch.empty((world, 64), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allh.view(world * 64), mine)
    return [bytes(allh[r].cpu().tolist()) for r in range(world)]
