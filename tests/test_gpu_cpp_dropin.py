"""GPU: the C++ drop-in classes (ORBSlam::ORBextractor / ORBmatcher, orbslam_in_practice_b200/cpp) called the way the
reference's Frame calls them (src/Frame.cpp:75-78), checked against the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from orbslam_in_practice_b200.synth import synth_batch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_driver():
    from orbslam_in_practice_b200 import build as b
    b.build(); b.build_cpp()
    out_dir = os.path.join(ROOT, "tests", "cpp", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "dropin_main")
    pkg = os.path.join(ROOT, "orbslam_in_practice_b200")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-I", os.path.join(pkg, "cpp"), "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp"), "-L", pkg, "-lorbslam_frontend", "-lorbx",
                           "-Wl,-rpath," + pkg])
    return exe


def test_cpp_classes_match_oracle(orbx, oracle, tmp_path):
    exe = _build_driver()
    imgs = synth_batch([30, 31], 640, 480)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<7if", 640, 480, 2, 1000, 8, 20, 7, 1.2)); f.write(imgs.tobytes())
    subprocess.check_call([exe, fin, fout])
    raw = open(fout, "rb").read(); off = 0
    oex = oracle.OracleExtractor()
    last = None
    frames_out = []
    for fidx in range(2):
        n = struct.unpack_from("<i", raw, off)[0]; off += 4
        kps = np.frombuffer(raw, oracle.KEYPOINT_DTYPE, n, off); off += 28 * n
        desc = np.frombuffer(raw, np.uint8, 32 * n, off).reshape(n, 32); off += 32 * n
        ko, do = oex(imgs[fidx])
        assert n == len(ko)
        for fld in ("x", "y", "size", "response", "octave", "class_id"):
            assert np.array_equal(kps[fld], ko[fld]), fld
        d = np.abs(kps["angle"] - ko["angle"]); d = np.minimum(d, 360 - d)
        assert d.max() <= 1e-3 * 180 / np.pi
        assert np.unpackbits(desc ^ do).sum() <= 1e-3 * do.size * 8
        last = desc
        frames_out.append((kps.copy(), desc.copy()))
    levels = struct.unpack_from("<i", raw, off)[0]; off += 4
    assert levels == 8
    for l in range(levels):
        w, h = struct.unpack_from("<2i", raw, off); off += 8
        s, bs = struct.unpack_from("<2Q", raw, off); off += 16
        lv = oex.level(l)
        assert (w, h) == lv.shape[::-1]
        assert s == int(lv.astype(np.uint64).sum()), "mvImagePyramid[%d] pixels" % l
        assert bs == int(oracle.reflect101_border(lv, 19).astype(np.uint64).sum()), "mvImagePyramid[%d] border" % l
    sf = np.frombuffer(raw, np.float32, 8, off); off += 32
    assert np.array_equal(sf, oex.scale_factors)
    d01, nm, _ = struct.unpack_from("<3i", raw, off); off += 12
    assert d01 == oracle.descriptor_distance(last[0], last[1])
    d1, i1, d2 = oracle.knn2(last, last)
    assert nm == int((oracle.ratio_select(d1, i1, d2, 50, 0.7) >= 0).sum())
    # ORBmatcher::SearchForInitialization through the C++ class, on the keypoints the C++ extractor produced
    n_si, cnt = struct.unpack_from("<2i", raw, off); off += 8
    m12 = np.frombuffer(raw, np.int32, cnt, off)
    (k1, dd1), (k2, dd2) = frames_out
    prev = np.stack([k1["x"], k1["y"]], 1)
    n_o, m_o, _ = oracle.search_for_initialization(k1, dd1, k2, dd2, prev, 100, 0.9, True, 640, 480)
    assert n_si == n_o and np.array_equal(m12, m_o)
    off += 4 * cnt
    # ORBmatcher::SearchByProjection (plain-container form) against the oracle's windowed search
    n_p, cntp = struct.unpack_from("<2i", raw, off); off += 8
    mp = np.frombuffer(raw, np.int32, cntp, off)
    P = oracle.window_params(15.0, [float(v) for v in oex.scale_factors], (0, 15), 1, 1, gate=1, th_dist=100, nnratio=0.0,
                             check_orientation=True, update_centers=False, width=640, height=480)
    n_po, m_po, _ = oracle.search_window(k1, dd1, k2, dd2, prev, P)
    assert n_p == n_po and np.array_equal(mp, m_po)
    off += 4 * cntp
    # ORBmatcher::SearchByBoW (plain-container form, node = first descriptor byte / 8) against the oracle's group search
    n_b, cntb = struct.unpack_from("<2i", raw, off); off += 8
    mb = np.frombuffer(raw, np.int32, cntb, off)
    n_bo, m_bo = oracle.search_groups(k1, dd1, (dd1[:, 0] >> 3).astype(np.uint16), k2, dd2, (dd2[:, 0] >> 3).astype(np.uint16), 50, 0.7, True)
    assert n_b == n_bo and np.array_equal(mb, m_bo)
