"""GPU parity: Hamming best-2 kNN / DescriptorDistance / ratio test / shard merge vs the oracle (bit-exact)."""
import numpy as np
import pytest

from orbslam_in_practice_b200.synth import synth_descriptor_db, synth_queries

pytestmark = pytest.mark.gpu


def test_descriptor_distance_pairs(orbx, oracle):
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (1000, 32), dtype=np.uint8); b = rng.integers(0, 256, (1000, 32), dtype=np.uint8)
    b[:10] = a[:10]; b[10:20] = ~a[10:20]
    m = orbx.Matcher(1000, 1000)
    got = m.hamming_pairs(a, b)
    want = np.array([oracle.descriptor_distance(a[i], b[i]) for i in range(1000)], np.int32)
    assert np.array_equal(got, want)
    assert got[:10].max() == 0 and got[10:20].min() == 256


@pytest.mark.parametrize("nq,ndb", [(1, 1), (7, 1), (5, 2), (1000, 127), (1000, 128), (1025, 129), (3000, 5000), (2005, 2005)])
def test_knn2_matches_oracle(orbx, oracle, nq, ndb):
    db = synth_descriptor_db(ndb, seed=ndb, dup_frac=0.05)
    q = synth_queries(db, nq, seed=nq)
    m = orbx.Matcher(nq, ndb)
    d1, i1, d2 = m.knn2_host(q, db, index_base=100)
    o1, oi, o2 = oracle.knn2(q, db, index_base=100, nthreads=8)
    assert np.array_equal(d1, o1) and np.array_equal(i1, oi) and np.array_equal(d2, o2)


def test_knn2_empty_database(orbx):
    m = orbx.Matcher(16, 16)
    q = np.zeros((16, 32), np.uint8)
    d1, i1, d2 = m.knn2_host(q, np.zeros((0, 32), np.uint8))
    assert (d1 == np.iinfo(np.int32).max).all() and (i1 == -1).all() and (d2 == np.iinfo(np.int32).max).all()


def test_knn2_ties_first_index_wins(orbx, oracle):
    db = np.zeros((600, 32), np.uint8)          # every row identical: all distances tie
    q = np.zeros((33, 32), np.uint8); q[:, 0] = 0xff
    m = orbx.Matcher(64, 600)
    d1, i1, d2 = m.knn2_host(q, db)
    assert (i1 == 0).all() and (d1 == 8).all() and (d2 == 8).all()
    o = oracle.knn2(q, db)
    assert np.array_equal(d1, o[0]) and np.array_equal(i1, o[1]) and np.array_equal(d2, o[2])


def test_sharded_merge_and_ratio_select_equal_unsharded(orbx, oracle):
    import torch
    ndb, nq, G = 20000, 3000, 4
    db = synth_descriptor_db(ndb, dup_frac=0.02); q = synth_queries(db, nq)
    m = orbx.Matcher(nq, ndb)
    dev = torch.device("cuda:0")
    tq = torch.from_numpy(q).to(dev); tdb = torch.from_numpy(db).to(dev)
    parts = torch.empty((3, G, nq), dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    bounds = [ndb * g // G for g in range(G + 1)]
    for g in range(G):
        lo, hi = bounds[g], bounds[g + 1]
        m.knn2_device(tq.data_ptr(), nq, tdb[lo:hi].data_ptr(), hi - lo, lo,
                      parts[0, g].data_ptr(), parts[1, g].data_ptr(), parts[2, g].data_ptr(), s)
    out = torch.empty((4, nq), dtype=torch.int32, device=dev)
    m.merge_shards_device(parts[0].data_ptr(), parts[1].data_ptr(), parts[2].data_ptr(), G, nq,
                          out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), s)
    m.ratio_select_device(out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), nq, 50, 0.7, out[3].data_ptr(), s)
    torch.cuda.synchronize()
    d1, i1, d2, match = (out[i].cpu().numpy() for i in range(4))
    o1, oi, o2 = oracle.knn2(q, db, nthreads=8)
    assert np.array_equal(d1, o1) and np.array_equal(i1, oi) and np.array_equal(d2, o2)
    assert np.array_equal(match, oracle.ratio_select(o1, oi, o2, 50, 0.7))
    assert (match >= 0).sum() > nq // 2
    # oracle merge of the per-shard triples agrees too
    p = parts.cpu().numpy()
    om = oracle.merge_shards(p[0], p[1], p[2])
    assert np.array_equal(om[0], o1) and np.array_equal(om[1], oi) and np.array_equal(om[2], o2)


def test_peer_exchange_single_rank_equals_plain_knn(orbx, oracle):
    """The fused wait + peer-read + merge + ratio kernel with world = 1 (own buffer only); N > 1 is exercised by
    tools/sharded_knn_check.py under torchrun (needs several GPUs)."""
    import torch
    ndb, nq = 6000, 1500
    db = synth_descriptor_db(ndb, dup_frac=0.03); q = synth_queries(db, nq)
    m = orbx.Matcher(nq, ndb)
    m.exchange_open([m.exchange_create(nq, 0, 1)])
    dev = torch.device("cuda:0")
    tq, tdb = torch.from_numpy(q).to(dev), torch.from_numpy(db).to(dev)
    out = torch.empty((4, nq), dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):                               # several epochs: both parities of the double buffer
        m.knn2_sharded_device(tq.data_ptr(), nq, tdb.data_ptr(), ndb, 0, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                              50, 0.7, out[3].data_ptr(), s)
    torch.cuda.synchronize()
    m.exchange_status()
    o1, oi, o2 = oracle.knn2(q, db, 0, 8)
    got = out.cpu().numpy()
    assert np.array_equal(got[0], o1) and np.array_equal(got[1], oi) and np.array_equal(got[2], o2)
    assert np.array_equal(got[3], oracle.ratio_select(o1, oi, o2, 50, 0.7))


def test_batched_pairwise_knn2_matches_oracle(orbx, oracle):
    """orbm_knn2_pairs_device: many independent scans (frame a's rows against frame b's rows, the extractor's [F][cap][32] layout,
    counts on the device) in one launch pair -- the left -> right matching of a batch of stereo pairs (BASELINE configs[2]).
    Every pair must equal the oracle's best-2 scan + acceptance (ORBmatcher.cpp:37-67): ragged counts, an empty frame on either
    side, duplicates (ties -> lowest row), counts at capacity, pairs that reuse frames."""
    import torch
    from orbslam_in_practice_b200.synth import synth_descriptor_db, synth_queries
    dev = torch.device("cuda:0")
    for cap, counts in ((300, [300, 257, 0, 1, 129, 128]), (2032, [2005, 2032, 1999, 640, 0, 2031]), (1024, [1024, 1, 1023, 5, 77, 900])):
        F = len(counts)
        desc = np.zeros((F, cap, 32), np.uint8)
        base = synth_descriptor_db(cap, seed=cap, dup_frac=0.05)
        for f, n in enumerate(counts):
            desc[f, :n] = base[:n] if f % 2 == 0 else synth_queries(base, max(n, 1), seed=100 + f)[:n]
            desc[f, n:] = 0xEE                                           # rows past the count must never be read as data
        pa = np.array([0, 1, 2, 3, 1, 4, 5, 0, 3], np.int32); pb = np.array([1, 0, 1, 2, 3, 5, 4, 0, 4], np.int32)
        P = len(pa)
        m = orbx.Matcher(cap, cap, 0)
        t_desc, t_cnt = torch.from_numpy(desc).to(dev), torch.tensor(counts, dtype=torch.int32, device=dev)
        t_pa, t_pb = torch.from_numpy(pa).to(dev), torch.from_numpy(pb).to(dev)
        out = torch.full((4, P, cap), -5, dtype=torch.int32, device=dev)
        wsb = orbx.load().orbm_knn2_pairs_workspace_bytes(cap, P)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        st = torch.cuda.Stream(device=dev); torch.cuda.synchronize()
        m.knn2_pairs_device(t_desc.data_ptr(), t_cnt.data_ptr(), cap, t_pa.data_ptr(), t_pb.data_ptr(), P, out[0].data_ptr(), out[1].data_ptr(),
                            out[2].data_ptr(), 50, 0.7, out[3].data_ptr(), ws.data_ptr(), wsb, st.cuda_stream)
        st.synchronize()
        got = out.cpu().numpy()
        for p in range(P):
            nq, ndb = counts[pa[p]], counts[pb[p]]
            d1, i1, d2 = oracle.knn2(desc[pa[p], :nq], desc[pb[p], :ndb], 0, 2) if nq else (np.zeros(0, np.int32),) * 3
            wm = oracle.ratio_select(d1, i1, d2, 50, 0.7) if nq else np.zeros(0, np.int32)
            assert np.array_equal(got[0, p, :nq], d1) and np.array_equal(got[1, p, :nq], i1) and np.array_equal(got[2, p, :nq], d2), (cap, p)
            assert np.array_equal(got[3, p, :nq], wm), (cap, p)
            assert (got[1, p, nq:] == -1).all() and (got[3, p, nq:] == -1).all() and (got[0, p, nq:] == np.iinfo(np.int32).max).all()
