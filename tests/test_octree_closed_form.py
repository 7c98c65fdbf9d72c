"""The closed-form restatement of DistributeOctTree's phase 1 that k_octree runs (csrc/octree.cu) against the CPU oracle
(oracle/orb_oracle.c :415, itself pinned against the reference's own ORBextractor.cpp:489-718): plain form (sort by path code)
and the kernel-shaped form (bins at depth B, XOR-masked prefix scan for the list order).  CPU only."""
import os, sys
import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools", "proto"))
from oracle import oracle as O
import octree_closed_form as P


def _random_problem(rng, t):
    W = int(rng.integers(20, 700)); H = int(rng.integers(20, 500))
    if int(W / H + 0.5) < 1:
        W, H = H + 20, W
    n = int(rng.integers(1, 900))
    if t % 3 == 0:        # clustered candidates: phase 1 runs deep
        cx, cy = rng.integers(0, W), rng.integers(0, H)
        xs = np.clip(rng.normal(cx, W / 8, n).astype(int), 0, W - 1); ys = np.clip(rng.normal(cy, H / 8, n).astype(int), 0, H - 1)
    else:
        xs = rng.integers(0, W, n); ys = rng.integers(0, H, n)
    pts = np.unique(np.stack([ys, xs], 1), axis=0); rng.shuffle(pts)
    c = np.zeros(len(pts), O.CAND_DTYPE); c['x'] = pts[:, 1]; c['y'] = pts[:, 0]; c['score'] = rng.integers(1, 60, len(pts))
    N = int(rng.integers(1, 2 * len(pts) + 2))
    return c, W, H, N


def test_plain_closed_form_equals_oracle():
    rng = np.random.default_rng(5)
    phase2 = 0
    for t in range(150):
        c, W, H, N = _random_problem(rng, t)
        ref = O.distribute_octree(c, W, H, N)
        got, info = P.octree(c, W, H, N)
        assert len(ref) == len(got) and (ref == got).all(), (t, W, H, len(c), N, info)
        phase2 += info['phase2']
    assert phase2 > 20                      # both endings of phase 1 are exercised


def test_bins_form_equals_plain_form():
    rng = np.random.default_rng(6)
    resolved = 0
    for t in range(200):
        c, W, H, N = _random_problem(rng, t)
        r = P.check_bins(c, W, H, N, int(rng.integers(1, 6)))
        assert r is None or r, (t, W, H, len(c), N)
        resolved += r is not None
    assert resolved > 50                    # the rest would take the kernel's sequential fallback


def test_real_candidates_all_levels():
    from orbslam_in_practice_b200.synth import synth_batch
    ex = O.OracleExtractor(1000, 1.2, 8, 20, 7)
    img = synth_batch([3])[0]
    ex(img)
    for l in range(8):
        c, kept = ex.candidates(l), ex.kept(l)
        lw, lh = int(round(img.shape[1] / 1.2 ** l)), int(round(img.shape[0] / 1.2 ** l))
        got, info = P.octree(c, lw - 26, lh - 26, len(kept))
        ref = O.distribute_octree(c, lw - 26, lh - 26, len(kept))
        assert (ref == got).all(), (l, info)
