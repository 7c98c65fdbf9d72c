#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the CPU oracle.

Run in the build container (where /root/reference and cv2 exist).  Before writing, every fixture
case is cross-checked against (a) the reference's own ORBextractor.cpp (oracle/_ref/ref_orb, built
from /root/reference against the header shim, monotonic allocator) and (b) the cv2-based Tier-A
restatement (real OpenCV 4.13.0 primitives).  A mismatch aborts: the fixtures are only written
from an oracle that agrees with both.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as O, ref as R, pin_cv2 as P          # noqa: E402
from orbslam_in_practice_b200.synth import (synth_frame, adversarial_frame, synth_descriptor_db,  # noqa: E402
                                             synth_queries)

CASES = [
    ("vga_seed0", dict(kind="synth", seed=0, w=640, h=480), dict(nfeatures=1000)),
    ("vga_seed7", dict(kind="synth", seed=7, w=640, h=480), dict(nfeatures=1000)),
    ("kitti_seed1", dict(kind="synth", seed=1, w=1241, h=376), dict(nfeatures=2000)),
    ("checker", dict(kind="checker", w=640, h=480), dict(nfeatures=1000)),
    ("noise_qvga", dict(kind="noise", w=320, h=240), dict(nfeatures=1000)),
    ("constant", dict(kind="constant", w=640, h=480), dict(nfeatures=1000)),
    ("small_params", dict(kind="synth", seed=11, w=512, h=384), dict(nfeatures=500, scale_factor=1.5, nlevels=5, ini_th=30, min_th=10)),
]


def make_image(spec):
    if spec["kind"] == "synth":
        return synth_frame(spec["seed"], spec["w"], spec["h"])
    return adversarial_frame(spec["kind"], spec["w"], spec["h"])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def case_record(img, params):
    ex = O.OracleExtractor(**params)
    kps, desc = ex(img)
    nl = ex.nlevels
    rec = {
        "image_sha256": sha(img), "shape": list(img.shape), "params": params,
        "n_keypoints": int(len(kps)),
        "level_dims": [list(ex.level(l).shape[::-1]) for l in range(nl)],
        "level_sha256": [sha(ex.level(l)) for l in range(nl)],
        "blur_sha256": [sha(ex.blurred(l)) if ex.blurred(l) is not None else None for l in range(nl)],
        "n_candidates": [int(len(ex.candidates(l))) for l in range(nl)],
        "cand_sha256": [sha(ex.candidates(l)) for l in range(nl)],
        "n_kept": [int(len(ex.kept(l))) for l in range(nl)],
        "kept_sha256": [sha(ex.kept(l)) for l in range(nl)],
        "retries": [int(ex.retries(l)) for l in range(nl)],
        "kp_xy_size_resp_octave_sha256": sha(np.stack([kps["x"], kps["y"], kps["size"], kps["response"],
                                                        kps["octave"].astype(np.float32)], 1)) if len(kps) else sha(np.zeros(0)),
        "desc_sha256": sha(desc),
        "features_per_level": [int(v) for v in ex.features_per_level],
        "scale_factors_hex": [float(v).hex() for v in ex.scale_factors],
        "umax": [int(v) for v in ex.umax],
    }
    return rec, kps, desc


def main():
    assert R.build(), "oracle/_ref/ref_orb must be buildable here (needs /root/reference)"
    out = {"version": 1, "cases": {}, "knn": {}}
    arrays = {}
    for name, spec, params in CASES:
        img = make_image(spec)
        rec, kps, desc = case_record(img, params)
        # (a) the reference's own translation unit
        (kr, dr), = R.run(img, params.get("nfeatures", 1000), params.get("scale_factor", 1.2), params.get("nlevels", 8),
                          params.get("ini_th", 20), params.get("min_th", 7))[0]
        assert len(kr) == len(kps) and all(np.array_equal(kr[f], kps[f]) for f in kr.dtype.names), name + ": oracle != reference TU"
        assert np.array_equal(dr, desc), name + ": descriptors != reference TU"
        # (b) real OpenCV primitives
        rep = P.compare(img, **params)
        assert rep["pyramid_px"] == 0 and rep["blur_px"] == 0 and rep["cand_mismatch"] == 0 and rep["kept_mismatch"] == 0, (name, rep)
        assert rep["desc_bits_diff"] == 0 and rep["angle_max_abs_deg"] < 1e-4, (name, rep)
        rec["spec"] = spec
        out["cases"][name] = rec
        if name in ("vga_seed0", "small_params"):
            arrays[name + "_angles"] = kps["angle"].astype(np.float32)
            arrays[name + "_desc"] = desc
            arrays[name + "_xy"] = np.stack([kps["x"], kps["y"]], 1).astype(np.float32)
        print(name, rec["n_keypoints"], rec["n_candidates"], rec["retries"])
    # kNN goldens
    db = synth_descriptor_db(20000, dup_frac=0.02); q = synth_queries(db, 3000)
    d1, i1, d2 = O.knn2(q, db, 0, 8)
    m = O.ratio_select(d1, i1, d2, 50, 0.7)
    out["knn"] = {"ndb": 20000, "nq": 3000, "db_sha256": sha(db), "q_sha256": sha(q), "d1_sha256": sha(d1),
                  "idx1_sha256": sha(i1), "d2_sha256": sha(d2), "match_sha256": sha(m), "n_matched": int((m >= 0).sum())}
    # matcher entry points on a frame pair (frame 0 and a shifted copy), extracted by the oracle with nfeatures = 2000
    fa = synth_frame(0); fb = np.roll(np.roll(fa, 5, axis=1), 3, axis=0)
    exm = O.OracleExtractor(nfeatures=2000)
    k1, dd1 = exm(fa); k2, dd2 = exm(fb)
    prev = np.stack([k1["x"], k1["y"]], 1)
    n_si, m_si, p_si = O.search_for_initialization(k1, dd1, k2, dd2, prev, 100, 0.9, True, 640, 480)
    assert R.build_match(), "oracle/_ref/ref_match must be buildable here"
    n_rf, m_rf, p_rf = R.run_search_for_initialization(k1, dd1, k2, dd2, prev, 100, 0.9, True, 640, 480, False)
    assert n_si == n_rf and np.array_equal(m_si, m_rf) and np.array_equal(p_si, p_rf), "SearchForInitialization: oracle != reference TU"
    sf = [float(v) for v in exm.scale_factors]
    cen = np.stack([k1["x"] + 5, k1["y"] + 3], 1).astype(np.float32)
    wp = O.window_params(7.0, sf, (0, 15), 1, 1, gate=1, th_dist=100, nnratio=0.0, check_orientation=True, update_centers=False,
                         width=640, height=480)
    n_w, m_w, _ = O.search_window(k1, dd1, k2, dd2, cen, wp)
    g1 = (dd1[:, 0] >> 3).astype(np.uint16); g2 = (dd2[:, 0] >> 3).astype(np.uint16)
    n_g, m_g = O.search_groups(k1, dd1, g1, k2, dd2, g2, 50, 0.7, True)
    out["search"] = {"pair": "synth seed 0 vs the same frame rolled by (5, 3), nfeatures 2000",
                     "n1": int(len(k1)), "n2": int(len(k2)), "desc1_sha256": sha(dd1), "desc2_sha256": sha(dd2),
                     "search_init": {"window": 100, "ratio": 0.9, "ori": True, "n": int(n_si), "m12_sha256": sha(m_si), "prev_sha256": sha(p_si),
                                     "pinned": "equal to the reference's src/ORBmatcher.cpp (oracle/_ref/ref_match) at generation time"},
                     "projection": {"th": 7.0, "levels": [1, 1], "th_dist": 100, "n": int(n_w), "m12_sha256": sha(m_w),
                                    "pinned": "unpinned by the reference (empty body); oracle restatement of upstream"},
                     "bow": {"group": "first descriptor byte >> 3", "th_dist": 50, "ratio": 0.7, "n": int(n_g), "m12_sha256": sha(m_g),
                             "pinned": "unpinned by the reference (empty body); oracle restatement of upstream"}}
    print("search", n_si, n_w, n_g)
    # known-answer vectors for DescriptorDistance
    a = np.zeros((4, 32), np.uint8); b = np.zeros((4, 32), np.uint8)
    b[1] = 0xff; a[2, :4] = [0x0f, 0xf0, 0xaa, 0x55]; a[3] = np.arange(32); b[3] = np.arange(32)[::-1]
    out["hamming_kat"] = {"a": a.tolist(), "b": b.tolist(), "dist": [int(O.descriptor_distance(a[i], b[i])) for i in range(4)]}
    assert out["hamming_kat"]["dist"][:3] == [0, 256, 16]
    json.dump(out, open(os.path.join(HERE, "golden_v1.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "golden_v1_arrays.npz"), **arrays)
    print("wrote golden_v1.json, golden_v1_arrays.npz")


if __name__ == "__main__":
    main()
