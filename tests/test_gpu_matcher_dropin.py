"""GPU: ORBSlam::ORBmatcher as a drop-in -- the reference's matcher call sites compiled VERBATIM against the repo's
cpp/ORBmatcher.h over a stand-in for the reference's Frame / KeyFrame / MapPoint public interface (tests/cpp/frame_shim.h):
  src/Tracking.cpp:181-189  SearchForInitialization(Frame&, Frame&, ...)  -> byte-identical to the reference's own
                            src/ORBmatcher.cpp compiled unmodified (oracle/_ref/ref_match)
  src/Tracking.cpp:298,344  SearchByBoW(KeyFrame*, Frame, ...) / SearchByProjection(Frame&, const Frame&, ...): empty in the
                            reference (parity unpinned) -> equal to the oracle's restatement of the upstream loops
  ComputeThreeMaxima (src/ORBmatcher.cpp:147-188), DescriptorDistance (:128-144)
"""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import ref as R
from orbslam_in_practice_b200.synth import synth_frame

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(accessor):
    from orbslam_in_practice_b200 import build as b
    b.build(); b.build_cpp()
    out_dir = os.path.join(ROOT, "tests", "cpp", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "matcher_dropin_acc" if accessor else "matcher_dropin")
    pkg = os.path.join(ROOT, "orbslam_in_practice_b200")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-I", os.path.join(pkg, "cpp")] + (["-DFRAME_SHIM_WITH_BOUNDS_ACCESSOR"] if accessor else []) +
                          ["-o", exe, os.path.join(ROOT, "tests", "cpp", "matcher_dropin_main.cpp"), "-L", pkg, "-lorbslam_frontend", "-lorbx",
                           "-Wl,-rpath," + pkg])
    return exe


def _pair(orbx, seed, shift):
    a = synth_frame(seed); b = np.roll(np.roll(a, shift[0], axis=1), shift[1], axis=0)
    ex = orbx.Extractor(nfeatures=2000, max_width=640, max_height=480, max_batch=2)
    kps, desc, cnt = ex.extract_host(np.stack([a, b]))
    return kps[0][:cnt[0]].copy(), desc[0][:cnt[0]].copy(), kps[1][:cnt[1]].copy(), desc[1][:cnt[1]].copy(), ex


def _write_input(path, k1, d1, k2, d2, prev, window, ratio, ori, bug):
    with open(path, "wb") as f:
        f.write(struct.pack("<7if", len(k1), len(k2), 640, 480, window, int(ori), int(bug), ratio))
        f.write(k1.tobytes()); f.write(d1.tobytes()); f.write(k2.tobytes()); f.write(d2.tobytes())
        f.write(np.ascontiguousarray(prev, np.float32).tobytes())


@pytest.mark.skipif(not R.match_available(), reason="oracle/_ref/ref_match not built")
@pytest.mark.parametrize("accessor", [False, True])
@pytest.mark.parametrize("ratio,ori,bug", [(0.9, True, False), (0.7, False, False), (0.9, True, True)])
def test_tracking_call_site_equals_reference_matcher(orbx, tmp_path, accessor, ratio, ori, bug):
    exe = _build(accessor)
    k1, d1, k2, d2, _ = _pair(orbx, 21, (6, -4))
    prev = np.stack([k1["x"], k1["y"]], 1)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _write_input(fin, k1, d1, k2, d2, prev, 100, ratio, ori, bug)
    subprocess.check_call([exe, fin, fout, "init"])
    ours = open(fout, "rb").read()
    ref = R.run_search_for_initialization(k1, d1, k2, d2, prev, 100, ratio, ori, 640, 480, bug, raw_output=True)
    assert ours == ref, "the GPU matcher behind the reference call site differs from the reference's own ORBmatcher.cpp"
    n = struct.unpack_from("<i", ours, 0)[0]
    assert (n == 0) if bug else (n > 100)


def test_explicit_grid_bounds_on_the_device(orbx, oracle):
    """orbm_window_params.use_bounds (FindimageBound of a distorted lens) through the C ABI against the oracle."""
    k1, d1, k2, d2, _ = _pair(orbx, 22, (-7, 5))
    prev = np.stack([k1["x"], k1["y"]], 1)
    bounds = (40.5, 600.25, 30.75, 440.5)
    P = orbx.WindowParams()
    P.radius = 100.0
    for i in range(16):
        P.level_scale[i] = 1.0
    P.query_level_min = P.query_level_max = P.level_below = P.level_above = 0
    P.gate, P.th_dist, P.nnratio, P.check_orientation, P.update_centers = 0, 50, 0.9, 1, 1
    P.width, P.height, P.use_bounds = 1, 1, 1
    P.min_x, P.max_x, P.min_y, P.max_y = bounds
    m = orbx.Matcher(4096, 4096)
    n, m12, cen = m.search_window_host(k1, d1, k2, d2, prev, P)
    PO = oracle.window_params(100.0, None, (0, 0), 0, 0, gate=0, th_dist=50, nnratio=0.9, check_orientation=True,
                              update_centers=True, width=640, height=480, bounds=bounds)
    n_o, m_o, c_o = oracle.search_window(k1, d1, k2, d2, prev, PO)
    assert n == n_o and np.array_equal(m12, m_o) and np.array_equal(cen, c_o)
    assert n > 100


def test_typed_bow_and_projection_call_sites(orbx, oracle, tmp_path):
    exe = _build(False)
    k1, d1, k2, d2, ex = _pair(orbx, 23, (4, 3))
    prev = np.stack([k1["x"], k1["y"]], 1)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _write_input(fin, k1, d1, k2, d2, prev, 100, 0.9, True, False)
    subprocess.check_call([exe, fin, fout, "typed"])
    raw = open(fout, "rb").read(); off = 0
    n_proj, n2 = struct.unpack_from("<2i", raw, off); off += 8
    assert n2 == len(k2)
    typed_proj = np.frombuffer(raw, np.int32, n2, off); off += 4 * n2
    n_bow = struct.unpack_from("<i", raw, off)[0]; off += 4
    typed_bow = np.frombuffer(raw, np.int32, n2, off); off += 4 * n2
    ind1, ind2, ind3, dd = struct.unpack_from("<4i", raw, off); off += 16
    sizes = np.frombuffer(raw, np.int32, 30, off)

    # the driver gives every keypoint of frame 1 a map point 2 m in front of the identity-pose camera (every third one
    # has none, every 17th is bad, #5 is an outlier) -- replay that set-up
    has_mp = (np.arange(len(k1)) % 3) != 2
    bad = has_mp & ((np.arange(len(k1)) % 17) == 0)
    fx, fy, cx, cy = 500.0, 500.0, 320.0, 240.0
    z = 2.0
    # same arithmetic as the driver: float subtraction and division, then double
    X = ((k1["x"] - np.float32(cx)) / np.float32(fx)).astype(np.float64) * z
    Y = ((k1["y"] - np.float32(cy)) / np.float32(fy)).astype(np.float64) * z
    u = (np.float32(fx) * X / z + np.float32(cx)).astype(np.float32)
    v = (np.float32(fy) * Y / z + np.float32(cy)).astype(np.float32)
    cen = np.stack([u, v], 1)
    visible = has_mp.copy(); visible[5] = False                     # upstream skips outliers; bad points are still projected there
    cen[~visible] = np.nan
    sf = [float(s) for s in ex.scale_factors]
    P = oracle.window_params(7.0, sf, (0, 15), 1, 1, gate=1, th_dist=100, nnratio=0.0, check_orientation=True,
                             update_centers=False, width=640, height=480)
    n_o, m_o, _ = oracle.search_window(k1, d1, k2, d2, cen, P)
    want = np.full(n2, -1, np.int32)
    for i1 in np.flatnonzero(m_o >= 0):
        want[m_o[i1]] = i1
    assert n_proj == n_o and np.array_equal(typed_proj, want)
    assert n_proj > 200

    # SearchByBoW: no vocabulary in the reference -> one node for everything; keyframe keypoints without a good map point sit out
    g1 = np.where(has_mp & ~bad, 0, 0xffff).astype(np.uint16)
    g2 = np.zeros(len(k2), np.uint16)
    n_b, m_b = oracle.search_groups(k1, d1, g1, k2, d2, g2, 50, 0.7, True)
    want = np.full(n2, -1, np.int32)
    for i1 in np.flatnonzero(m_b >= 0):
        want[m_b[i1]] = i1
    assert n_bow == n_b and np.array_equal(typed_bow, want)
    assert n_bow > 100

    # ComputeThreeMaxima / DescriptorDistance against plain restatements
    order = sorted(range(30), key=lambda i: (-int(sizes[i]), i))
    e1, e2, e3 = order[0], order[1], order[2]
    if sizes[e2] < 0.1 * sizes[e1]:
        e2 = e3 = -1
    elif sizes[e3] < 0.1 * sizes[e1]:
        e3 = -1
    assert (ind1, ind2, ind3) == (e1, e2, e3)
    assert dd == oracle.descriptor_distance(d1[0], d1[1])
