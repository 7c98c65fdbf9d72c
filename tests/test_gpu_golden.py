"""GPU: the CUDA path against the committed golden fixtures (no oracle call on this path), through the C ABI."""
import numpy as np
import pytest

from _golden import GOLD, ARR, sha, make_image, kp_sha

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(GOLD["cases"]))
def test_cuda_extractor_matches_golden(orbx, name):
    rec = GOLD["cases"][name]
    img = make_image(rec["spec"])
    assert sha(img) == rec["image_sha256"]
    p = rec["params"]
    ex = orbx.Extractor(p.get("nfeatures", 1000), p.get("scale_factor", 1.2), p.get("nlevels", 8), p.get("ini_th", 20),
                        p.get("min_th", 7), img.shape[1], img.shape[0], 1)
    kps, desc, counts = ex.extract_host(img)
    n = int(counts[0])
    assert n == rec["n_keypoints"]
    for l in range(ex.nlevels):
        assert sha(ex.level(0, l)) == rec["level_sha256"][l], "pyramid level %d" % l
        if rec["blur_sha256"][l] is not None:
            assert sha(ex.level(0, l, blurred=True)) == rec["blur_sha256"][l], "blurred level %d" % l
        assert sha(ex.candidates(0, l)) == rec["cand_sha256"][l], "FAST candidates level %d" % l
        assert sha(ex.kept(0, l)) == rec["kept_sha256"][l], "octree survivors level %d" % l
    assert kp_sha(kps[0, :n]) == rec["kp_xy_size_resp_octave_sha256"]
    if name + "_angles" in ARR:
        d = np.abs(kps[0, :n]["angle"] - ARR[name + "_angles"]); d = np.minimum(d, 360 - d)
        assert d.max() <= 1e-3 * 180 / np.pi                       # 1e-3 rad
        bits = int(np.unpackbits(desc[0, :n] ^ ARR[name + "_desc"]).sum())
        assert bits <= 1e-3 * n * 256, "descriptor bits differ: %d" % bits
    # descriptors: >= 99.9 % of bits; in practice the hash matches (no rounding flips on these inputs)
    if n and sha(desc[0, :n]) != rec["desc_sha256"]:
        pytest.xfail("descriptor hash differs (allowed: <= 0.1 % rotated-pattern rounding flips); see parity test")


def test_cuda_knn_matches_golden(orbx):
    from orbslam_in_practice_b200.synth import synth_descriptor_db, synth_queries
    g = GOLD["knn"]
    db = synth_descriptor_db(g["ndb"], dup_frac=0.02); q = synth_queries(db, g["nq"])
    m = orbx.Matcher(g["nq"], g["ndb"])
    d1, i1, d2 = m.knn2_host(q, db)
    assert sha(d1) == g["d1_sha256"] and sha(i1) == g["idx1_sha256"] and sha(d2) == g["d2_sha256"]
    k = GOLD["hamming_kat"]
    got = m.hamming_pairs(np.array(k["a"], np.uint8), np.array(k["b"], np.uint8))
    assert list(got) == k["dist"]


def test_full_size_knn_properties(orbx):
    """BASELINE config 4 size (1M x 100k): size-independent properties instead of an oracle run."""
    import torch
    from orbslam_in_practice_b200.synth import synth_descriptor_db, synth_queries
    ndb, nq = 1_000_000, 100_000
    db = synth_descriptor_db(ndb); q = synth_queries(db, nq)
    m = orbx.Matcher(nq, ndb)
    dev = torch.device("cuda:0")
    tq, tdb = torch.from_numpy(q).to(dev), torch.from_numpy(db).to(dev)
    out = torch.empty((3, nq), dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    m.knn2_device(tq.data_ptr(), nq, tdb.data_ptr(), ndb, 0, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), s)
    # sharded in 8 + merge must be identical (linearity of the best-2 reduction)
    parts = torch.empty((8, 3, nq), dtype=torch.int32, device=dev)
    b = [ndb * g // 8 for g in range(9)]
    for g in range(8):
        m.knn2_device(tq.data_ptr(), nq, tdb[b[g]:b[g + 1]].data_ptr(), b[g + 1] - b[g], b[g],
                      parts[g, 0].data_ptr(), parts[g, 1].data_ptr(), parts[g, 2].data_ptr(), s)
    mo = torch.empty((3, nq), dtype=torch.int32, device=dev)
    m.merge_shards_device(parts[0, 0].data_ptr(), parts[0, 1].data_ptr(), parts[0, 2].data_ptr(), 8, nq,
                          mo[0].data_ptr(), mo[1].data_ptr(), mo[2].data_ptr(), s, shard_stride=3 * nq)
    torch.cuda.synchronize()
    assert torch.equal(out, mo)
    d1, i1, d2 = (out[i].cpu().numpy() for i in range(3))
    assert (d1 <= d2).all() and (i1 >= 0).all() and (i1 < ndb).all()
    # the reported distance is the true distance to the reported row (checksum over all queries)
    true_d = np.unpackbits(q ^ db[i1], axis=1).sum(1)
    assert np.array_equal(true_d, d1)
    # spot-check exactness on a sample of queries against a numpy scan of the whole database
    for qi in (0, 1, 77777, 99999):
        dist = np.unpackbits(q[qi][None] ^ db, axis=1).sum(1)
        order = np.argsort(dist, kind="stable")
        assert i1[qi] == order[0] and d1[qi] == dist[order[0]] and d2[qi] == dist[order[1]]


def test_cuda_matcher_entry_points_match_golden(orbx):
    """The three search instances through the C ABI against the committed golden hashes (no oracle involved)."""
    from orbslam_in_practice_b200.synth import synth_frame
    g = GOLD["search"]
    fa = synth_frame(0); fb = np.roll(np.roll(fa, 5, axis=1), 3, axis=0)
    ex = orbx.Extractor(nfeatures=2000, max_width=640, max_height=480, max_batch=2)
    kps, desc, cnt = ex.extract_host(np.stack([fa, fb]))
    k1, d1, k2, d2 = kps[0][:cnt[0]], desc[0][:cnt[0]], kps[1][:cnt[1]], desc[1][:cnt[1]]
    assert len(k1) == g["n1"] and len(k2) == g["n2"] and sha(d1) == g["desc1_sha256"] and sha(d2) == g["desc2_sha256"]
    m = orbx.Matcher(4096, 4096)
    prev = np.stack([k1["x"], k1["y"]], 1)
    n, m12, p = m.search_init_host(k1, d1, k2, d2, prev, 100, 0.9, True, 640, 480)
    assert n == g["search_init"]["n"] and sha(m12) == g["search_init"]["m12_sha256"] and sha(p) == g["search_init"]["prev_sha256"]
    cen = np.stack([k1["x"] + 5, k1["y"] + 3], 1).astype(np.float32)
    P = orbx.WindowParams.projection(7.0, [float(v) for v in ex.scale_factors], 640, 480, th_dist=100, check_orientation=True)
    n, m12, _ = m.search_window_host(k1, d1, k2, d2, cen, P)
    assert n == g["projection"]["n"] and sha(m12) == g["projection"]["m12_sha256"]
    n, m12 = m.search_groups_host(k1, d1, (d1[:, 0] >> 3).astype(np.uint16), k2, d2, (d2[:, 0] >> 3).astype(np.uint16), 50, 0.7, True)
    assert n == g["bow"]["n"] and sha(m12) == g["bow"]["m12_sha256"]
