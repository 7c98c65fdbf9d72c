"""Size-independent properties of the CPU oracle (no GPU): invariants the reference's algorithm implies, checked on
seeded inputs.  They guard the checker itself -- the GPU parity tests compare against this oracle."""
import numpy as np
import pytest

from orbslam_in_practice_b200.synth import synth_frame, synth_descriptor_db, synth_queries


def _brute_knn2(q, db):
    x = np.unpackbits(q[:, None, :] ^ db[None, :, :], axis=2).sum(axis=2).astype(np.int64)
    order = np.argsort(x, axis=1, kind="stable")               # stable: the first minimal index wins (ORBmatcher.cpp:52, strict <)
    d1 = np.take_along_axis(x, order[:, :1], 1)[:, 0]
    d2 = np.take_along_axis(x, order[:, 1:2], 1)[:, 0]
    return d1, order[:, 0], d2


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_knn2_is_the_stable_best_two(oracle, seed):
    rng = np.random.default_rng(seed)
    db = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    db[150:160] = db[10:20]                                    # exact duplicates: tie-break coverage
    q = db[rng.integers(0, 300, 64)] ^ (rng.random((64, 32)) < 0.02).astype(np.uint8)
    d1, i1, d2 = oracle.knn2(q, db)
    b1, bi, b2 = _brute_knn2(q, db)
    assert np.array_equal(d1, b1) and np.array_equal(i1, bi) and np.array_equal(d2, b2)
    assert (d1 <= d2).all()
    # DescriptorDistance (ORBmatcher.cpp:128-144) agrees with a bit count, is symmetric and zero on the diagonal
    for k in range(8):
        assert oracle.descriptor_distance(q[k], db[i1[k]]) == d1[k] == oracle.descriptor_distance(db[i1[k]], q[k])
        assert oracle.descriptor_distance(q[k], q[k]) == 0


def test_ratio_select_is_the_reference_acceptance(oracle):
    d1 = np.array([10, 50, 51, 30, 30, 0], np.int32); d2 = np.array([20, 100, 100, 42, 43, 0], np.int32)
    i1 = np.arange(6, dtype=np.int32)
    m = oracle.ratio_select(d1, i1, d2, 50, 0.7)
    # :65 best <= TH_LOW; :67 best < (float)best2 * ratio, strict (30 < 42 * 0.7 = 29.4 fails, 30 < 43 * 0.7 = 30.1 passes; 0 < 0 fails)
    assert m.tolist() == [0, 1, -1, -1, 4, -1]


@pytest.mark.parametrize("size,nf", [((640, 480), 1000), ((320, 240), 300), ((417, 263), 500)])
def test_extractor_invariants(oracle, size, nf):
    w, h = size
    img = synth_frame(3, w, h)
    ex = oracle.OracleExtractor(nfeatures=nf)
    kps, desc = ex(img)
    kps2, desc2 = ex(img)
    assert kps.tobytes() == kps2.tobytes() and desc.tobytes() == desc2.tobytes()          # deterministic, handle reusable
    assert len(kps) == len(desc) and desc.shape[1] == 32
    # per level: at most N_l + 3 survivors (DistributeOctTree may overshoot by up to three, :621-686), level-major order
    assert (np.diff(kps["octave"]) >= 0).all()
    for l in range(ex.nlevels):
        n = int((kps["octave"] == l).sum())
        assert n == len(ex.kept(l)) and n <= max(int(ex.features_per_level[l]) + 3, 4)
        lv = ex.level(l)
        sel = kps[kps["octave"] == l]
        s = ex.scale_factors[l]
        if l:                                                                             # :1055-1061, float multiply
            x = sel["x"] / s; y = sel["y"] / s
        else:
            x, y = sel["x"], sel["y"]
        # keypoints sit >= EDGE_THRESHOLD (19) px inside their level (:729-732 plus the 3-px FAST margin)
        assert (x > 18.5).all() and (x < lv.shape[1] - 18.5).all() and (y > 18.5).all() and (y < lv.shape[0] - 18.5).all()
        assert (sel["size"] == np.float32(int(31 * s))).all()
        assert ((sel["angle"] >= 0) & (sel["angle"] < 360)).all() and (sel["class_id"] == -1).all()
        # survivors are a subset of the level's FAST candidates, each with its candidate's response
        cand = {(int(c["x"]), int(c["y"])): int(c["score"]) for c in ex.candidates(l)}
        kept = ex.kept(l)
        for k in kept:
            assert cand[(int(k["x"]), int(k["y"]))] == int(k["score"])
        # the output keypoints are the kept candidates shifted by (16, 16) (:801-802), in list order, response = score
        assert np.array_equal(np.rint(x).astype(np.int64), kept["x"].astype(np.int64) + 16)
        assert np.array_equal(np.rint(y).astype(np.int64), kept["y"].astype(np.int64) + 16)
        assert np.array_equal(sel["response"], kept["score"].astype(np.float32))


def test_constant_image_has_no_keypoints(oracle):
    ex = oracle.OracleExtractor()
    kps, desc = ex(np.full((480, 640), 77, np.uint8))
    assert len(kps) == 0 and desc.shape == (0, 32)             # the descriptors.release() path, :1024-1030


def test_pyramid_chain_and_blur_are_the_pinned_primitives(oracle):
    """Level l is the resize of level l-1's rounded output (:1084), the blurred level is gaussian7 of the level (:1045-1046)."""
    img = synth_frame(9, 400, 300)
    ex = oracle.OracleExtractor(nfeatures=400)
    ex(img)
    assert np.array_equal(ex.level(0), img)
    for l in range(1, ex.nlevels):
        prev, cur = ex.level(l - 1), ex.level(l)
        assert np.array_equal(cur, oracle.resize_linear(prev, cur.shape[1], cur.shape[0]))
        b = ex.blurred(l)
        if b is not None:
            assert np.array_equal(b, oracle.gaussian7(cur))


def test_gaussian7_preserves_constants_and_is_symmetric(oracle):
    c = np.full((40, 50), 201, np.uint8)
    assert np.array_equal(oracle.gaussian7(c), c)              # taps sum to 256 in both passes: a constant stays exact
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (33, 47), dtype=np.uint8)
    assert np.array_equal(oracle.gaussian7(a[::-1, ::-1])[::-1, ::-1], oracle.gaussian7(a))   # symmetric kernel + reflect-101


def test_synthetic_db_queries_are_reproducible():
    db = synth_descriptor_db(2048); q = synth_queries(db, 128)
    assert db.tobytes() == synth_descriptor_db(2048).tobytes() and q.tobytes() == synth_queries(db, 128).tobytes()
