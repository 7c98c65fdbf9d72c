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
