"""Runner for oracle/_ref/ref_orb: the reference's own ORBextractor.cpp compiled unmodified from
/root/reference against oracle/ref_shim (see oracle/Makefile).  TEST INFRASTRUCTURE ONLY.

The binary travels to the GPU box prebuilt (oracle/_ref is git-ignored, not gpurun-ignored);
/root/reference itself is never read at run time.
"""
import os
import struct
import subprocess
import tempfile

import numpy as np

from .oracle import KEYPOINT_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(_HERE, "_ref", "ref_orb")


def available():
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def build():
    """(Re)build when the reference tree is present; no-op otherwise."""
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_ref/ref_orb"])
    return available()


def run(frames, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, allocator="bump", repeat=1):
    """frames: (F, H, W) u8.  -> (list of (keypoints, descriptors), seconds_per_frame)"""
    frames = np.ascontiguousarray(frames, np.uint8)
    if frames.ndim == 2:
        frames = frames[None]
    F, H, W = frames.shape
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("<7if", W, H, F, nfeatures, nlevels, ini_th, min_th, scale_factor))
            f.write(frames.tobytes())
        out = subprocess.check_output([REF_BIN, fin, fout, allocator, str(repeat)], text=True)
        spf = float(out.split()[-1])
        raw = open(fout, "rb").read()
    res, off = [], 0
    for _ in range(F):
        n = struct.unpack_from("<i", raw, off)[0]; off += 4
        kps = np.frombuffer(raw, KEYPOINT_DTYPE, n, off).copy(); off += 28 * n
        desc = np.frombuffer(raw, np.uint8, 32 * n, off).reshape(n, 32).copy(); off += 32 * n
        res.append((kps, desc))
    return res, spf


MATCH_BIN = os.path.join(_HERE, "_ref", "ref_match")


def match_available():
    return os.path.exists(MATCH_BIN) and os.access(MATCH_BIN, os.X_OK)


def build_match():
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_ref/ref_match"])
    return match_available()


def run_search_for_initialization(kp1, desc1, kp2, desc2, prev_matched, window=100, nnratio=0.9, check_orientation=True,
                                  width=640, height=480, literal_bug=False, bounds=None, raw_output=False):
    """The reference's own ORBmatcher::SearchForInitialization (src/ORBmatcher.cpp:9-126, compiled unmodified).
    bounds = (minX, maxX, minY, maxY) of the Frame grid (default: the zero-distortion bounds of width x height)."""
    kp1 = np.ascontiguousarray(kp1, KEYPOINT_DTYPE); kp2 = np.ascontiguousarray(kp2, KEYPOINT_DTYPE)
    desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
    prev = np.ascontiguousarray(prev_matched, np.float32)
    n1, n2 = len(kp1), len(kp2)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("<7if", n1, n2, width, height, window, int(check_orientation), int(literal_bug), nnratio))
            f.write(kp1.tobytes()); f.write(desc1.tobytes()); f.write(kp2.tobytes()); f.write(desc2.tobytes()); f.write(prev.tobytes())
        subprocess.check_call([MATCH_BIN, fin, fout] + ([repr(float(np.float32(b))) for b in bounds] if bounds is not None else []))
        raw = open(fout, "rb").read()
    if raw_output:
        return raw
    n = struct.unpack_from("<i", raw, 0)[0]
    m12 = np.frombuffer(raw, np.int32, n1, 4).copy()
    prev_out = np.frombuffer(raw, np.float32, 2 * n1, 4 + 4 * n1).reshape(n1, 2).copy()
    return n, m12, prev_out
