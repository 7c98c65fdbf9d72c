"""Deterministic synthetic inputs, integer arithmetic only.

The survey's probe generator used cv2's float GaussianBlur and rng.normal; both can change bits
with the host's SIMD path / libm, which would un-pin the golden fixtures.  This restatement keeps
the same structure (flat 128 background, W*H/600 random rectangles, a small blur, additive noise)
but uses only numpy integer ops, so a seed gives the same bytes on every box.
"""
import numpy as np


def synth_frame(seed, width=640, height=480):
    """Grayscale u8 test frame; seed = frame index."""
    rng = np.random.default_rng(int(seed))
    W, H = int(width), int(height)
    img = np.full((H, W), 128, np.int32)
    n = W * H // 600
    for _ in range(n):
        x0 = int(rng.integers(0, W)); y0 = int(rng.integers(0, H))
        w = int(rng.integers(6, 80)); h = int(rng.integers(6, 80))
        val = int(rng.integers(0, 256))
        img[y0:y0 + h, x0:x0 + w] = val
    # 3x3 binomial blur, replicate edges, exact integer rounding
    p = np.pad(img, 1, mode="edge")
    hsum = p[:, :-2] + 2 * p[:, 1:-1] + p[:, 2:]
    v = hsum[:-2, :] + 2 * hsum[1:-1, :] + hsum[2:, :]
    img = (v + 8) >> 4
    img = img + rng.integers(-3, 4, (H, W))
    return np.clip(img, 0, 255).astype(np.uint8)


def synth_batch(seeds, width=640, height=480):
    return np.stack([synth_frame(s, width, height) for s in seeds])


def adversarial_frame(kind, width=640, height=480, seed=0):
    """constant: no keypoints; noise: saturated cells; checker: equal-score NMS ties."""
    W, H = int(width), int(height)
    if kind == "constant":
        return np.full((H, W), 77, np.uint8)
    if kind == "noise":
        return np.random.default_rng(seed).integers(0, 256, (H, W), dtype=np.uint8)
    if kind == "checker":
        yy, xx = np.mgrid[0:H, 0:W]
        return (((yy // 8 + xx // 8) & 1) * 200 + 20).astype(np.uint8)
    raise ValueError(kind)


def synth_descriptor_db(ndb, seed=1234, dup_frac=0.01):
    """Uniform random 32-byte rows, `dup_frac` of them exact copies of earlier rows (tie-break)."""
    rng = np.random.default_rng(seed)
    db = rng.integers(0, 256, (ndb, 32), dtype=np.uint8)
    ndup = int(ndb * dup_frac)
    if ndup and ndb > 1:
        dst = rng.integers(1, ndb, ndup)
        src = (rng.random(ndup) * dst).astype(np.int64)  # src < dst
        db[dst] = db[src]
    return db


def synth_queries(db, nq, seed=5678, match_frac=0.8, flip_p=0.06):
    """80 % perturbed db rows (each bit flipped w.p. flip_p), 20 % uniform random rows."""
    rng = np.random.default_rng(seed)
    ndb = len(db)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    nm = int(nq * match_frac)
    rows = rng.integers(0, ndb, nm)
    flips = (rng.random((nm, 256)) < flip_p)
    q[:nm] = db[rows] ^ np.packbits(flips, axis=1, bitorder="little")
    perm = rng.permutation(nq)
    return q[perm]
