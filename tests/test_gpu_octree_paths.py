"""k_octree's ways through DistributeOctTree (ORBextractor.cpp:489-718), each forced by the input, against the oracle:
closed-form phase 1 + bins pass (textured frame, quota << candidates), closed form that ends the run (quota above the
candidate count), the sequential fallback (few, clustered candidates: phase 1 runs deeper than the counted prefixes),
and score ties inside nodes (the kept key must be the first in candidate order)."""
import numpy as np
import pytest

from orbslam_in_practice_b200.synth import synth_frame

pytestmark = pytest.mark.gpu


def _check(orbx, oracle, img, nfeatures, nlevels=8, ini=20, mn=7):
    H, W = img.shape
    ex = orbx.Extractor(nfeatures=nfeatures, nlevels=nlevels, ini_th=ini, min_th=mn, max_width=W, max_height=H, max_batch=1)
    oex = oracle.OracleExtractor(nfeatures, 1.2, nlevels, ini, mn)
    kps, desc, counts = ex.extract_host(img)
    ko, do = oex(img)
    assert int(counts[0]) == len(ko)
    total_c = 0
    for l in range(nlevels):
        cg, co = ex.candidates(0, l), oex.candidates(l)
        assert np.array_equal(cg, co), "FAST candidates differ at level %d" % l
        kg, kk = ex.kept(0, l), oex.kept(l)
        assert len(kg) == len(kk) and np.array_equal(kg, kk), "octree survivors differ at level %d (%d candidates, %d kept)" % (l, len(co), len(kk))
        total_c += len(co)
    for fld in ("x", "y", "response", "octave"):
        assert np.array_equal(kps[0, :len(ko)][fld], ko[fld])
    return total_c, len(ko)


def _patch_frame(seed, boxes, W=640, H=480):
    """flat grey frame with textured boxes: candidates only inside them"""
    tex = synth_frame(seed, W, H)
    img = np.full((H, W), 128, np.uint8)
    for (x0, y0, x1, y1) in boxes:
        img[y0:y1, x0:x1] = tex[y0:y1, x0:x1]
    return img


def test_textured_frame_small_quota(orbx, oracle):
    c, k = _check(orbx, oracle, synth_frame(11), 300)
    assert c > 4 * k


def test_quota_above_candidates(orbx, oracle):
    # every node ends with one key: phase 1 alone ends the run (size == prevSize), no phase 2
    img = _patch_frame(12, [(100, 100, 260, 220)])
    c, k = _check(orbx, oracle, img, 20000)
    assert 0 < c and k <= c
    # a fully textured frame under the same quota: every candidate survives, the tree is as deep as it gets
    c, k = _check(orbx, oracle, synth_frame(12), 20000)
    assert k > 3000


@pytest.mark.parametrize("boxes", [[(300, 200, 340, 240)], [(40, 40, 100, 90), (560, 400, 620, 450)], [(0, 0, 640, 60)]])
def test_clustered_candidates(orbx, oracle, boxes):
    # all candidates in a corner of the region: the tree is deep and narrow, phase 1 outruns the counted prefixes
    img = _patch_frame(13, boxes)
    for nf in (200, 1000, 4000):
        _check(orbx, oracle, img, nf)


def test_score_ties_inside_nodes(orbx, oracle):
    # a periodic pattern: many candidates with identical scores in every node, so the choice rests on the candidate order
    y, x = np.mgrid[0:480, 0:640]
    img = (((x // 5 + y // 5) % 2) * 90 + 60).astype(np.uint8)
    img[::7, ::11] += 40
    for nf in (150, 1000):
        _check(orbx, oracle, img, nf)


def test_tiny_quota_and_single_level(orbx, oracle):
    _check(orbx, oracle, synth_frame(14), 10, nlevels=1)
    _check(orbx, oracle, synth_frame(15), 37, nlevels=3)


def test_sequential_form_in_a_fresh_process():
    """ORBX_OCT_CLOSED=0 (read once per process) switches the closed form off: the sequential partition passes, which inputs
    whose phase 1 outruns the counted prefixes fall back to, must give the same survivors on ordinary frames too."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from oracle import oracle\n"
        "from orbslam_in_practice_b200 import _lib as orbx\n"
        "from orbslam_in_practice_b200.synth import synth_frame\n"
        "for seed, nf in ((21, 1000), (22, 300)):\n"
        "    img = synth_frame(seed)\n"
        "    ex = orbx.Extractor(nfeatures=nf, max_width=640, max_height=480, max_batch=1)\n"
        "    oex = oracle.OracleExtractor(nf, 1.2, 8, 20, 7)\n"
        "    kps, desc, counts = ex.extract_host(img); ko, do = oex(img)\n"
        "    assert int(counts[0]) == len(ko)\n"
        "    for l in range(8):\n"
        "        assert np.array_equal(ex.kept(0, l), oex.kept(l)), (seed, l)\n"
        "print('sequential ok')\n" % root)
    env = dict(os.environ, ORBX_OCT_CLOSED="0")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "sequential ok" in out.stdout, out.stdout + out.stderr
