"""CPU, world_size 2 over gloo: the N>1 host path (shard plan + all-gather layout + merge) gives the
unsharded result.  The CUDA kernels are replaced by the oracle here (no GPU in this container); the
same plan drives bench.py --gpus N with the real kernels and NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from orbslam_in_practice_b200 import sharding


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q_np, db_np, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    lo, hi = sharding.db_shard(len(db_np), rank, world)
    d1, i1, d2 = O.knn2(q_np, db_np[lo:hi], lo)                   # global indices: base = lo
    tri = torch.from_numpy(np.stack([d1, i1, d2]))
    allt = sharding.gather_triples(tri, world).numpy()             # [world][3][nq]
    merged = O.merge_shards(allt[:, 0], allt[:, 1], allt[:, 2])
    f0, f1 = sharding.frame_shard(37, rank, world)
    frames = torch.zeros(37, dtype=torch.int32); frames[f0:f1] = 1
    dist.all_reduce(frames)
    if rank == 0:
        ret["merged"] = [m.copy() for m in merged]
        ret["frames_cover"] = frames.numpy().copy()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_db_sharded_knn_over_gloo(world):
    from oracle import oracle as O
    from orbslam_in_practice_b200.synth import synth_descriptor_db, synth_queries
    db = synth_descriptor_db(3001, seed=5, dup_frac=0.05); q = synth_queries(db, 257, seed=6)
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), q, db, ret), nprocs=world, join=True)
    full = O.knn2(q, db)
    assert all(np.array_equal(a, b) for a, b in zip(ret["merged"], full))
    assert (ret["frames_cover"] == 1).all()                        # frame ranges: disjoint and exhaustive


def test_shard_bounds_properties():
    for n in (0, 1, 7, 256, 1000003):
        for w in (1, 2, 3, 8):
            b = sharding.shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(w))
            assert max(b[i + 1] - b[i] for i in range(w)) - min(b[i + 1] - b[i] for i in range(w)) <= 1
