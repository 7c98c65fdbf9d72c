"""Seeded random parity hunt for the matcher entry points: CUDA (C ABI) vs the CPU oracle.
usage: python tools/fuzz_match.py [ncases] [seed]      (needs a GPU; test infrastructure, not product)
  * best-2 kNN on random sizes with planted duplicates and near-duplicates (ties, d1 == d2)
  * windowed search (SearchForInitialization / SearchByProjection instances) with random windows, level ranges,
    gates, ratios and orientation checks on keypoints extracted from shifted synthetic frames
  * group-restricted search (SearchByBoW instance) with 1 .. 5000 groups"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np

from oracle import oracle
from orbslam_in_practice_b200 import _lib as orbx
from orbslam_in_practice_b200.synth import synth_frame


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    fails, t0 = 0, time.time()
    m = orbx.Matcher(8192, 8192)
    for case in range(n):
        nq, ndb = int(rng.integers(1, 3000)), int(rng.integers(0, 6000))
        db = rng.integers(0, 256, (ndb, 32), dtype=np.uint8)
        if ndb > 4:
            dup = rng.integers(0, ndb, ndb // 20 + 1); src = rng.integers(0, ndb, ndb // 20 + 1)
            db[dup] = db[src]                                              # exact duplicates: index ties
        q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
        if ndb:
            pick = rng.integers(0, ndb, nq)
            noisy = db[pick] ^ (rng.random((nq, 32, 8)) < 0.04).astype(np.uint8).dot(1 << np.arange(8)).astype(np.uint8)
            take = rng.random(nq) < 0.7
            q[take] = noisy[take]
        base = int(rng.integers(0, 1 << 20))
        g = m.knn2_host(q, db, base)
        o = oracle.knn2(q, db, base)
        if not all(np.array_equal(a, b) for a, b in zip(g, o)):
            fails += 1
            print("FAIL knn2 nq=%d ndb=%d base=%d" % (nq, ndb, base), flush=True)
    print("knn2: %d cases, %d failures, %.1f s" % (n, fails, time.time() - t0), flush=True)

    t0 = time.time()
    ex = orbx.Extractor(nfeatures=2000, max_width=640, max_height=480, max_batch=2)
    sf = [float(v) for v in ex.scale_factors]
    wfails = 0
    for case in range(n):
        sd = int(rng.integers(0, 500)); dx, dy = int(rng.integers(-30, 31)), int(rng.integers(-20, 21))
        a = synth_frame(sd); b = np.roll(np.roll(a, dx, axis=1), dy, axis=0)
        kps, desc, cnt = ex.extract_host(np.stack([a, b]))
        k1, d1, k2, d2 = kps[0][:cnt[0]], desc[0][:cnt[0]], kps[1][:cnt[1]], desc[1][:cnt[1]]
        cen = np.stack([k1["x"] + dx + rng.normal(0, 2, len(k1)), k1["y"] + dy + rng.normal(0, 2, len(k1))], 1).astype(np.float32)
        cen[rng.random(len(k1)) < 0.1, 0] = np.nan
        gate = int(rng.integers(0, 2)); radius = float(rng.choice([3.0, 7.0, 15.0, 40.0, 100.0]))
        below, above = int(rng.integers(-1, 3)), int(rng.integers(-1, 3))
        qlo = int(rng.integers(0, 3)); qhi = int(rng.integers(qlo, 8))
        ratio = float(rng.choice([0.0, 0.6, 0.75, 0.9])); ori = bool(rng.integers(0, 2)); upd = bool(rng.integers(0, 2))
        thd = int(rng.choice([30, 50, 100]))
        scales = sf if rng.integers(0, 2) else None
        Po = oracle.window_params(radius, scales, (qlo, qhi), below, above, gate=gate, th_dist=thd, nnratio=ratio,
                                  check_orientation=ori, update_centers=upd, width=640, height=480)
        Pg = orbx.WindowParams()
        for fld, _ in Pg._fields_:
            setattr(Pg, fld, getattr(Po, fld))
        n_g, m_g, c_g = m.search_window_host(k1, d1, k2, d2, cen, Pg)
        n_o, m_o, c_o = oracle.search_window(k1, d1, k2, d2, cen, Po)
        if n_g != n_o or not np.array_equal(m_g, m_o) or not np.array_equal(c_g, c_o, equal_nan=True):
            wfails += 1
            print("FAIL window seed=%d shift=(%d,%d) gate=%d r=%g lv=(%d,%d) q=(%d,%d) ratio=%g ori=%d upd=%d th=%d: %d vs %d"
                  % (sd, dx, dy, gate, radius, below, above, qlo, qhi, ratio, ori, upd, thd, n_g, n_o), flush=True)
    print("windowed search: %d cases, %d failures, %.1f s" % (n, wfails, time.time() - t0), flush=True)

    t0 = time.time()
    gfails = 0
    for case in range(n):
        sd = int(rng.integers(0, 500)); dx, dy = int(rng.integers(-10, 11)), int(rng.integers(-10, 11))
        a = synth_frame(sd); b = np.roll(np.roll(a, dx, axis=1), dy, axis=0)
        kps, desc, cnt = ex.extract_host(np.stack([a, b]))
        k1, d1, k2, d2 = kps[0][:cnt[0]], desc[0][:cnt[0]], kps[1][:cnt[1]], desc[1][:cnt[1]]
        ng = int(rng.choice([1, 7, 32, 100, 1000, 5000]))
        g1 = (d1[:, :2].view(np.uint16)[:, 0].astype(np.uint32) * ng // 65536).astype(np.uint16)
        g2 = (d2[:, :2].view(np.uint16)[:, 0].astype(np.uint32) * ng // 65536).astype(np.uint16)
        g1[rng.random(len(g1)) < 0.1] = 0xffff; g2[rng.random(len(g2)) < 0.1] = 0xffff
        ratio = float(rng.choice([0.5, 0.6, 0.75, 0.9, 1.5])); ori = bool(rng.integers(0, 2)); thd = int(rng.choice([30, 50, 100, 256]))
        n_g, m_g = m.search_groups_host(k1, d1, g1, k2, d2, g2, thd, ratio, ori)
        n_o, m_o = oracle.search_groups(k1, d1, g1, k2, d2, g2, thd, ratio, ori)
        if n_g != n_o or not np.array_equal(m_g, m_o):
            gfails += 1
            print("FAIL groups seed=%d ng=%d ratio=%g ori=%d th=%d: %d vs %d" % (sd, ng, ratio, ori, thd, n_g, n_o), flush=True)
    print("group search: %d cases, %d failures, %.1f s" % (n, gfails, time.time() - t0), flush=True)
    return 1 if (fails or wfails or gfails) else 0


if __name__ == "__main__":
    sys.exit(main())
