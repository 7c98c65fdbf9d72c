"""CPU: the oracle reproduces the committed golden fixtures (tests/golden, made by make_golden.py
from an oracle cross-checked against the reference's own ORBextractor.cpp and real OpenCV)."""
import numpy as np
import pytest

from _golden import GOLD, ARR, sha, make_image, kp_sha


@pytest.mark.parametrize("name", sorted(GOLD["cases"]))
def test_oracle_matches_golden(oracle, name):
    rec = GOLD["cases"][name]
    img = make_image(rec["spec"])
    assert sha(img) == rec["image_sha256"], "synthetic generator changed: fixtures are un-pinned"
    ex = oracle.OracleExtractor(**rec["params"])
    kps, desc = ex(img)
    assert len(kps) == rec["n_keypoints"]
    assert [int(v) for v in ex.features_per_level] == rec["features_per_level"]
    assert [float(v).hex() for v in ex.scale_factors] == rec["scale_factors_hex"]
    assert [int(v) for v in ex.umax] == rec["umax"]
    for l in range(ex.nlevels):
        assert list(ex.level(l).shape[::-1]) == rec["level_dims"][l]
        assert sha(ex.level(l)) == rec["level_sha256"][l]
        b = ex.blurred(l)
        assert (sha(b) if b is not None else None) == rec["blur_sha256"][l]
        assert len(ex.candidates(l)) == rec["n_candidates"][l] and sha(ex.candidates(l)) == rec["cand_sha256"][l]
        assert len(ex.kept(l)) == rec["n_kept"][l] and sha(ex.kept(l)) == rec["kept_sha256"][l]
        assert ex.retries(l) == rec["retries"][l]
    assert kp_sha(kps) == rec["kp_xy_size_resp_octave_sha256"]
    assert sha(desc) == rec["desc_sha256"]
    if name + "_angles" in ARR:
        assert np.array_equal(kps["angle"], ARR[name + "_angles"])


def test_survey_constants(oracle):
    """Values SURVEY.md section 8 lists for (1000, 1.2, 8)."""
    ex = oracle.OracleExtractor()
    assert list(ex.features_per_level) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(ex.umax) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert list(oracle.OracleExtractor(2000).features_per_level) == [434, 362, 302, 251, 209, 175, 145, 122]
    assert list(oracle.OracleExtractor(8000).features_per_level) == [1737, 1448, 1207, 1005, 838, 698, 582, 485]
    ex(np.zeros((480, 640), np.uint8))
    assert [ex.level(l).shape[::-1] for l in range(8)] == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231),
                                                           (257, 193), (214, 161), (179, 134)]


def test_knn_golden_and_kat(oracle):
    from orbslam_in_practice_b200.synth import synth_descriptor_db, synth_queries
    g = GOLD["knn"]
    db = synth_descriptor_db(g["ndb"], dup_frac=0.02); q = synth_queries(db, g["nq"])
    assert sha(db) == g["db_sha256"] and sha(q) == g["q_sha256"]
    d1, i1, d2 = oracle.knn2(q, db, 0, 4)
    assert sha(d1) == g["d1_sha256"] and sha(i1) == g["idx1_sha256"] and sha(d2) == g["d2_sha256"]
    m = oracle.ratio_select(d1, i1, d2, 50, 0.7)
    assert sha(m) == g["match_sha256"] and int((m >= 0).sum()) == g["n_matched"]
    k = GOLD["hamming_kat"]
    a, b = np.array(k["a"], np.uint8), np.array(k["b"], np.uint8)
    assert [oracle.descriptor_distance(a[i], b[i]) for i in range(4)] == k["dist"]
    # popcount cross-check against numpy
    rng = np.random.default_rng(3)
    x = rng.integers(0, 256, (64, 32), dtype=np.uint8); y = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    assert [oracle.descriptor_distance(x[i], y[i]) for i in range(64)] == list(np.unpackbits(x ^ y, axis=1).sum(1))


def test_knn_sharded_merge_equals_unsharded(oracle):
    from orbslam_in_practice_b200.synth import synth_descriptor_db, synth_queries
    db = synth_descriptor_db(5000, seed=9, dup_frac=0.05); q = synth_queries(db, 700, seed=10)
    full = oracle.knn2(q, db)
    for G in (2, 3, 8):
        b = [len(db) * g // G for g in range(G + 1)]
        parts = [oracle.knn2(q, db[b[g]:b[g + 1]], b[g]) for g in range(G)]
        merged = oracle.merge_shards(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]), np.stack([p[2] for p in parts]))
        assert all(np.array_equal(a, c) for a, c in zip(merged, full))


def test_octree_tiebreak_rule_matters(oracle):
    """The defined tie-break is part of the contract: the opposite rule changes the kept set."""
    img = make_image(GOLD["cases"]["vga_seed0"]["spec"])
    a = oracle.OracleExtractor(); b = oracle.OracleExtractor(); b.set_tiebreak(1)
    ka, _ = a(img); kb, _ = b(img)
    sa = set(zip(ka["x"].tolist(), ka["y"].tolist(), ka["octave"].tolist()))
    sb = set(zip(kb["x"].tolist(), kb["y"].tolist(), kb["octave"].tolist()))
    assert sa != sb and len(sa & sb) > 0.95 * len(sa)


def test_fast_score_identity(oracle):
    """corner(t) <=> score0 >= t and score_t == score0 for corners (basis of the one-pass GPU kernel)."""
    img = make_image(GOLD["cases"]["vga_seed7"]["spec"])[100:260, 200:400]
    s0 = oracle.fast9_score0(img)
    for t in (7, 20, 35):
        kp = oracle.fast9_nms(img, t)
        assert len(kp) > 0
        for k in kp[:200]:
            assert s0[k["y"], k["x"]] == k["score"] and k["score"] >= t


def test_matcher_entry_points_golden(oracle):
    """SearchForInitialization (pinned against the reference's ORBmatcher.cpp when the golden was written) and the two
    upstream-form searches, on the golden frame pair."""
    from orbslam_in_practice_b200.synth import synth_frame
    g = GOLD["search"]
    fa = synth_frame(0); fb = np.roll(np.roll(fa, 5, axis=1), 3, axis=0)
    ex = oracle.OracleExtractor(nfeatures=2000)
    k1, d1 = ex(fa); k2, d2 = ex(fb)
    assert len(k1) == g["n1"] and len(k2) == g["n2"] and sha(d1) == g["desc1_sha256"] and sha(d2) == g["desc2_sha256"]
    prev = np.stack([k1["x"], k1["y"]], 1)
    n, m12, p = oracle.search_for_initialization(k1, d1, k2, d2, prev, 100, 0.9, True, 640, 480)
    assert n == g["search_init"]["n"] and sha(m12) == g["search_init"]["m12_sha256"] and sha(p) == g["search_init"]["prev_sha256"]
    sf = [float(v) for v in ex.scale_factors]
    cen = np.stack([k1["x"] + 5, k1["y"] + 3], 1).astype(np.float32)
    wp = oracle.window_params(7.0, sf, (0, 15), 1, 1, gate=1, th_dist=100, nnratio=0.0, check_orientation=True, update_centers=False,
                              width=640, height=480)
    n, m12, _ = oracle.search_window(k1, d1, k2, d2, cen, wp)
    assert n == g["projection"]["n"] and sha(m12) == g["projection"]["m12_sha256"]
    n, m12 = oracle.search_groups(k1, d1, (d1[:, 0] >> 3).astype(np.uint16), k2, d2, (d2[:, 0] >> 3).astype(np.uint16), 50, 0.7, True)
    assert n == g["bow"]["n"] and sha(m12) == g["bow"]["m12_sha256"]
