"""Pin the C oracle against REAL OpenCV (cv2 4.13.0) -- test infrastructure only.

The reference's arithmetic lives in OpenCV, which it neither vendors nor pins.
This module is a second, independent restatement ("Tier A") of ORBextractor.cpp's control flow in
Python that calls the real cv2 primitives (resize, GaussianBlur, FastFeatureDetector, fastAtan2);
`compare()` checks the C oracle against it stage by stage.  Used by tests/test_oracle_cv2_pin.py
(skipped when cv2 is not importable) and runnable by hand:  python -m oracle.pin_cv2
"""
import math

import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

from . import oracle as O


def cv_round(v):
    return int(np.rint(v))


def _pattern():
    import os, re
    txt = open(os.path.join(os.path.dirname(__file__), "orb_pattern_31.inc")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return np.array([int(t) for t in re.findall(r"-?\d+", txt)], np.int32).reshape(256, 4)


class _Node:
    __slots__ = ("x0", "y0", "x1", "y1", "keys", "seq")


def distribute_octree_py(cands, width, height, N, tiebreak=0):
    """ORBextractor.cpp:489-718 with a Python list as the std::list (index 0 = front)."""
    n_ini = int(np.float32(width) / np.float32(height) + np.float32(0.5))  # round(), positive
    if n_ini <= 0:
        return []
    hx = np.float32(width) / np.float32(n_ini)
    seq = [0]

    def mk(x0, y0, x1, y1):
        nd = _Node(); nd.x0, nd.y0, nd.x1, nd.y1 = x0, y0, x1, y1; nd.keys = []; nd.seq = seq[0]; seq[0] += 1
        return nd

    roots = [mk(int(hx * np.float32(i)), 0, int(hx * np.float32(i + 1)), height) for i in range(n_ini)]
    for k, c in enumerate(cands):
        roots[int(np.float32(c["x"]) / hx)].keys.append(k)
    lst = [r for r in roots if r.keys]

    def divide(nd):
        half_x = int(math.ceil(np.float32(nd.x1 - nd.x0) / np.float32(2)))
        half_y = int(math.ceil(np.float32(nd.y1 - nd.y0) / np.float32(2)))
        sx, sy = nd.x0 + half_x, nd.y0 + half_y
        ch = [mk(nd.x0, nd.y0, sx, sy), mk(sx, nd.y0, nd.x1, sy), mk(nd.x0, sy, sx, nd.y1), mk(sx, sy, nd.x1, nd.y1)]
        for k in nd.keys:
            c = cands[k]
            if c["x"] < sx:
                (ch[0] if c["y"] < sy else ch[2]).keys.append(k)
            else:
                (ch[1] if c["y"] < sy else ch[3]).keys.append(k)
        return ch

    def split_into(lst, nd, pending):
        # children creation consumes seq numbers only for non-empty ones in the reference (list
        # nodes are allocated at push_front); renumber to keep seq = creation order of list nodes
        for c in divide(nd):
            if c.keys:
                lst.insert(0, c)
                if len(c.keys) > 1:
                    pending.append(c)

    finish = False
    while not finish:
        prev_size = len(lst)
        pending = []
        snapshot = list(lst)
        for nd in snapshot:
            if len(nd.keys) == 1:
                continue
            split_into(lst, nd, pending)
            lst.remove(nd)
        n_to_expand = len(pending)
        if len(lst) >= N or len(lst) == prev_size:
            finish = True
        elif len(lst) + 3 * n_to_expand > N:
            while not finish:
                prev_size = len(lst)
                prev = pending
                pending = []
                sgn = -1 if tiebreak else 1
                prev.sort(key=lambda nd: (len(nd.keys), sgn * nd.seq))
                for nd in reversed(prev):
                    split_into(lst, nd, pending)
                    lst.remove(nd)
                    if len(lst) >= N:
                        break
                if len(lst) >= N or len(lst) == prev_size:
                    finish = True
    out = []
    for nd in lst:
        best = nd.keys[0]
        for k in nd.keys[1:]:
            if cands[k]["score"] > cands[best]["score"]:
                best = k
        out.append(best)
    return out


class TierAExtractor:
    """ORBextractor.cpp:1001-1065 with real cv2 primitives."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.o = O.OracleExtractor(nfeatures, scale_factor, nlevels, ini_th, min_th)  # tables only
        self.nlevels, self.ini_th, self.min_th = nlevels, ini_th, min_th
        self.pattern = _pattern()

    def __call__(self, image):
        o = self.o
        pyr = []
        for l in range(self.nlevels):
            s = o.inv_scale_factors[l]
            w = cv_round(np.float32(image.shape[1]) * s); h = cv_round(np.float32(image.shape[0]) * s)
            pyr.append(image.copy() if l == 0 else cv2.resize(pyr[l - 1], (w, h), interpolation=cv2.INTER_LINEAR))
        self.pyr = pyr
        self.cands, self.kept, self.angles, self.blur = [], [], [], []
        kps, descs = [], []
        f_ini = cv2.FastFeatureDetector_create(self.ini_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        f_min = cv2.FastFeatureDetector_create(self.min_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        for l in range(self.nlevels):
            img = pyr[l]
            min_bx = min_by = 16
            max_bx, max_by = img.shape[1] - 16, img.shape[0] - 16
            width, height = np.float32(max_bx - min_bx), np.float32(max_by - min_by)
            n_cols, n_rows = int(width / np.float32(30)), int(height / np.float32(30))
            w_cell = int(math.ceil(width / np.float32(n_cols))); h_cell = int(math.ceil(height / np.float32(n_rows)))
            cand = []
            for i in range(n_rows):
                ini_y = min_by + i * h_cell; max_y = ini_y + h_cell + 6
                if ini_y >= max_by - 3:
                    continue
                max_y = min(max_y, max_by)
                for j in range(n_cols):
                    ini_x = min_bx + j * w_cell; max_x = ini_x + w_cell + 6
                    if ini_x >= max_bx - 6:
                        continue
                    max_x = min(max_x, max_bx)
                    cell = np.ascontiguousarray(img[ini_y:max_y, ini_x:max_x])
                    kp = f_ini.detect(cell)
                    if not kp:
                        kp = f_min.detect(cell)
                    for k in kp:
                        cand.append((int(k.pt[0]) + j * w_cell, int(k.pt[1]) + i * h_cell, int(k.response)))
            cand = np.array(cand, O.CAND_DTYPE) if cand else np.zeros(0, O.CAND_DTYPE)
            self.cands.append(cand)
            keep = distribute_octree_py(cand, max_bx - min_bx, max_by - min_by, int(o.features_per_level[l]))
            kept = cand[keep] if len(keep) else np.zeros(0, O.CAND_DTYPE)
            self.kept.append(kept)
        for l in range(self.nlevels):
            ang = []
            img = pyr[l].astype(np.int64)
            for c in self.kept[l]:
                x, y = int(c["x"]) + 16, int(c["y"]) + 16
                m10 = m01 = 0
                for v in range(-15, 16):
                    d = int(o.umax[abs(v)])
                    row = img[y + v, x - d:x + d + 1]
                    m10 += int((np.arange(-d, d + 1) * row).sum()); m01 += v * int(row.sum())
                ang.append(cv2.fastAtan2(float(m01), float(m10)))
            self.angles.append(np.array(ang, np.float32))
        for l in range(self.nlevels):
            if len(self.kept[l]) == 0:
                self.blur.append(None)
                continue
            b = cv2.GaussianBlur(pyr[l].copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
            self.blur.append(b)
            factor_pi = np.float32(np.pi / np.float64(np.float32(180.0)))
            for c, a_deg in zip(self.kept[l], self.angles[l]):
                x, y = int(c["x"]) + 16, int(c["y"]) + 16
                ang = np.float32(a_deg) * factor_pi
                a, bb = np.float32(math.cos(float(ang))), np.float32(math.sin(float(ang)))
                px0, py0, px1, py1 = (self.pattern[:, i].astype(np.float32) for i in range(4))
                r0 = np.rint(px0 * bb + py0 * a).astype(np.int64); c0 = np.rint(px0 * a - py0 * bb).astype(np.int64)
                r1 = np.rint(px1 * bb + py1 * a).astype(np.int64); c1 = np.rint(px1 * a - py1 * bb).astype(np.int64)
                bits = (b[y + r0, x + c0] < b[y + r1, x + c1]).astype(np.uint8)
                descs.append(np.packbits(bits, bitorder="little"))
                sc = o.scale_factors[l]
                fx, fy = np.float32(x), np.float32(y)
                if l:
                    fx, fy = fx * sc, fy * sc
                kps.append((fx, fy, float(int(np.float32(31) * sc)), a_deg, float(c["score"]), l, -1))
        kps = np.array(kps, O.KEYPOINT_DTYPE) if kps else np.zeros(0, O.KEYPOINT_DTYPE)
        descs = np.stack(descs) if descs else np.zeros((0, 32), np.uint8)
        return kps, descs


def compare(image, **params):
    """Run Tier A (cv2) and the C oracle on one image; return a dict of mismatch counts."""
    a = TierAExtractor(**params)
    c = O.OracleExtractor(**params)
    ka, da = a(image)
    kc, dc = c(image)
    rep = {"pyramid_px": 0, "blur_px": 0, "cand_mismatch": 0, "kept_mismatch": 0, "angle_max_abs_deg": 0.0,
           "desc_bits_diff": 0, "desc_bits": 0, "n_tier_a": len(ka), "n_oracle": len(kc)}
    for l in range(a.nlevels):
        rep["pyramid_px"] += int((a.pyr[l] != c.level(l)).sum())
        if a.blur[l] is not None:
            rep["blur_px"] += int((a.blur[l] != c.blurred(l)).sum())
        ca, cc = a.cands[l], c.candidates(l)
        rep["cand_mismatch"] += 0 if (len(ca) == len(cc) and (ca == cc).all()) else 1
        kka, kkc = a.kept[l], c.kept(l)
        rep["kept_mismatch"] += 0 if (len(kka) == len(kkc) and (kka == kkc).all()) else 1
    if len(ka) == len(kc) and len(ka):
        d = np.abs(ka["angle"] - kc["angle"]); d = np.minimum(d, 360 - d)
        rep["angle_max_abs_deg"] = float(d.max())
        rep["desc_bits_diff"] = int(np.unpackbits(da ^ dc).sum()); rep["desc_bits"] = int(da.size * 8)
        rep["xy_equal"] = bool((ka["x"] == kc["x"]).all() and (ka["y"] == kc["y"]).all()
                               and (ka["size"] == kc["size"]).all() and (ka["response"] == kc["response"]).all()
                               and (ka["octave"] == kc["octave"]).all())
    return rep


if __name__ == "__main__":  # pragma: no cover
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from orbslam_in_practice_b200.synth import synth_frame, adversarial_frame
    for name, img, p in [("synth0 640x480", synth_frame(0), {}),
                         ("synth1 1241x376 N=2000", synth_frame(1, 1241, 376), {"nfeatures": 2000}),
                         ("checker", adversarial_frame("checker"), {}),
                         ("noise 320x240", adversarial_frame("noise", 320, 240), {})]:
        print(name, compare(img, **p))
