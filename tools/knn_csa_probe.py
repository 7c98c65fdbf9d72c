"""Time k_knn2 (1M x 100k) for every carry-save variant; each variant runs in its own process (env read once)."""
import os, sys, subprocess
if len(sys.argv) > 1:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import numpy as np, torch
    from orbslam_in_practice_b200 import _lib
    from orbslam_in_practice_b200.synth import synth_descriptor_db, synth_queries
    ndb, nq = 1_000_000, 100_000
    db = synth_descriptor_db(ndb); q = synth_queries(db, nq)
    m = _lib.Matcher(nq, ndb); m.set_profiling(True)
    tq, tdb = torch.from_numpy(q).cuda(), torch.from_numpy(db).cuda()
    out = torch.empty((3, nq), dtype=torch.int32, device='cuda')
    st = torch.cuda.Stream(); s = st.cuda_stream
    best = 1e9
    for _ in range(4):
        m.knn2_device(tq.data_ptr(), nq, tdb.data_ptr(), ndb, 0, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), s)
        torch.cuda.synchronize(); best = min(best, m.knn2_times()[0])
    print("CSA", os.environ.get("ORBX_KNN_CSA"), "scan ms %.2f" % best, "Gpairs/s %.1f" % (ndb * nq / best / 1e6), "checksum", int(out.sum(dtype=torch.int64)))
else:
    for c in (os.environ.get("CSA_LIST", "0,2,3,4,13,14,15").split(",")):
        subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, ORBX_KNN_CSA=c))
