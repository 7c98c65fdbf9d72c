"""Worker of tests/test_gpu_sharded_exchange.py (launched under torch.distributed.run, gloo rendezvous): every rank is its own
process with its own CUDA context ON THE SAME GPU, owns one contiguous shard of the database and runs the fused peer-memory
exchange (CUDA IPC mappings, flag handshake, k_merge_peers) -- the path bench.py times at N > 1.  Rank 0 checks the result
against the CPU oracle's unsharded scan (ORBmatcher.cpp:37-67)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from orbslam_in_practice_b200 import _lib, sharding
from orbslam_in_practice_b200.synth import synth_descriptor_db, synth_queries

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
ndb, nq = int(sys.argv[1]), int(sys.argv[2])
db = synth_descriptor_db(ndb, dup_frac=0.03)
q = synth_queries(db, nq)
lo, hi = sharding.db_shard(ndb, rank, world)
m = _lib.Matcher(nq, max(hi - lo, 1), 0)
tq = torch.from_numpy(q).to(dev)
tdb = torch.from_numpy(db[lo:hi].copy()).to(dev)
out = torch.full((4, nq), -7, dtype=torch.int32, device=dev)
m.exchange_open(sharding.exchange_handles(m, nq, rank, world))
ok = True
for rep in range(3):                       # consecutive epochs alternate between the two triple buffers
    out.fill_(-7); torch.cuda.synchronize()
    m.knn2_sharded_device(tq.data_ptr(), nq, tdb.data_ptr(), hi - lo, lo, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                          50, 0.7, out[3].data_ptr(), 0)
    torch.cuda.synchronize()
    m.exchange_status()
    if rank == 0:
        from oracle import oracle as O
        want = O.knn2(q, db, 0, 4)
        wm = O.ratio_select(*want, 50, 0.7)
        got = out.cpu().numpy()
        ok = ok and all(np.array_equal(g, w) for g, w in zip(got, list(want) + [wm]))
    dist.barrier()
flag = torch.tensor([int(ok)])
dist.broadcast(flag, 0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
