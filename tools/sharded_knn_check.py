"""torchrun --nproc-per-node N tools/sharded_knn_check.py : database-sharded kNN over N GPUs, both exchange paths
(NCCL all-gather + merge kernel, and the fused peer-memory kernel) against an unsharded scan on rank 0's GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from orbslam_in_practice_b200 import _lib, sharding
from orbslam_in_practice_b200.synth import synth_descriptor_db, synth_queries

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ndb, nq = 200_000, 20_000
db = synth_descriptor_db(ndb, dup_frac=0.02); q = synth_queries(db, nq)
lo, hi = sharding.db_shard(ndb, rank, world)
m = _lib.Matcher(nq, ndb, local)
tq = torch.from_numpy(q).to(dev); tdb = torch.from_numpy(db[lo:hi]).to(dev)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); s = st.cuda_stream
# path 1: NCCL all-gather + merge kernel
tri = torch.empty((3, nq), dtype=torch.int32, device=dev); out1 = torch.empty((4, nq), dtype=torch.int32, device=dev)
m.knn2_device(tq.data_ptr(), nq, tdb.data_ptr(), hi - lo, lo, tri[0].data_ptr(), tri[1].data_ptr(), tri[2].data_ptr(), s)
g = sharding.gather_triples(tri, world)
m.merge_shards_device(g[0, 0].data_ptr(), g[0, 1].data_ptr(), g[0, 2].data_ptr(), world, nq, out1[0].data_ptr(), out1[1].data_ptr(),
                      out1[2].data_ptr(), s, shard_stride=3 * nq)
m.ratio_select_device(out1[0].data_ptr(), out1[1].data_ptr(), out1[2].data_ptr(), nq, 50, 0.7, out1[3].data_ptr(), s)
# path 2: fused peer-memory exchange
handles = sharding.exchange_handles(m, nq, rank, world)
m.exchange_open(handles)
out2 = torch.empty((4, nq), dtype=torch.int32, device=dev)
for _ in range(4):
    m.knn2_sharded_device(tq.data_ptr(), nq, tdb.data_ptr(), hi - lo, lo, out2[0].data_ptr(), out2[1].data_ptr(), out2[2].data_ptr(),
                          50, 0.7, out2[3].data_ptr(), s)
torch.cuda.synchronize(); m.exchange_status()
# reference: unsharded scan on this GPU
full = torch.from_numpy(db).to(dev); ref = torch.empty((4, nq), dtype=torch.int32, device=dev)
m2 = _lib.Matcher(nq, ndb, local)
m2.knn2_device(tq.data_ptr(), nq, full.data_ptr(), ndb, 0, ref[0].data_ptr(), ref[1].data_ptr(), ref[2].data_ptr(), s)
m2.ratio_select_device(ref[0].data_ptr(), ref[1].data_ptr(), ref[2].data_ptr(), nq, 50, 0.7, ref[3].data_ptr(), s)
torch.cuda.synchronize()
ok1, ok2 = bool(torch.equal(out1, ref)), bool(torch.equal(out2, ref))
# timing of the two exchange paths (scan excluded: tiny shard per rank would hide it anyway)
def timed(fn, n=20):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def nccl_path():
    m.knn2_device(tq.data_ptr(), nq, tdb.data_ptr(), hi - lo, lo, tri[0].data_ptr(), tri[1].data_ptr(), tri[2].data_ptr(), s)
    g = sharding.gather_triples(tri, world)
    m.merge_shards_device(g[0, 0].data_ptr(), g[0, 1].data_ptr(), g[0, 2].data_ptr(), world, nq, out1[0].data_ptr(), out1[1].data_ptr(), out1[2].data_ptr(), s, shard_stride=3 * nq)
    m.ratio_select_device(out1[0].data_ptr(), out1[1].data_ptr(), out1[2].data_ptr(), nq, 50, 0.7, out1[3].data_ptr(), s)
def p2p_path():
    m.knn2_sharded_device(tq.data_ptr(), nq, tdb.data_ptr(), hi - lo, lo, out2[0].data_ptr(), out2[1].data_ptr(), out2[2].data_ptr(), 50, 0.7, out2[3].data_ptr(), s)
t1, t2 = timed(nccl_path), timed(p2p_path)
res = torch.tensor([int(ok1), int(ok2)], device=dev); dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print("world %d: nccl+merge exact=%s  p2p fused exact=%s  step ms: nccl %.3f  p2p %.3f" % (world, bool(res[0]), bool(res[1]), t1, t2))
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if int(res.min()) == 1 else 1)
