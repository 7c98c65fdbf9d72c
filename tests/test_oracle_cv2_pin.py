"""CPU: pin the C oracle's OpenCV restatements against REAL OpenCV (cv2), the unpinned third-party
dependency that owns the reference's arithmetic."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from orbslam_in_practice_b200.synth import synth_frame, adversarial_frame


@pytest.mark.parametrize("sw,sh,dw,dh", [(640, 480, 533, 400), (533, 400, 444, 333), (1241, 376, 1034, 313),
                                         (179, 134, 149, 112), (64, 48, 53, 40), (100, 80, 120, 96), (37, 29, 31, 24)])
def test_resize_linear_matches_cv2(oracle, sw, sh, dw, dh):
    img = np.random.default_rng(sw * 7 + sh).integers(0, 256, (sh, sw), dtype=np.uint8)
    assert np.array_equal(oracle.resize_linear(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("w,h", [(640, 480), (179, 134), (33, 21), (8, 8), (7, 40)])
def test_gaussian_blur_matches_cv2(oracle, w, h):
    img = np.random.default_rng(w + h).integers(0, 256, (h, w), dtype=np.uint8)
    assert np.array_equal(oracle.gaussian7(img), cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))


@pytest.mark.parametrize("th", [7, 20, 40])
def test_fast_cells_match_cv2(oracle, th):
    det = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    img = synth_frame(3)
    rng = np.random.default_rng(th)
    for _ in range(40):
        x0, y0 = int(rng.integers(0, 590)), int(rng.integers(0, 430))
        cw, ch = int(rng.integers(7, 45)), int(rng.integers(7, 45))
        cell = np.ascontiguousarray(img[y0:y0 + ch, x0:x0 + cw])
        got = oracle.fast9_nms(cell, th)
        want = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det.detect(cell)]
        assert [(int(k["x"]), int(k["y"]), int(k["score"])) for k in got] == want
    noise = adversarial_frame("noise", 64, 64)
    got = oracle.fast9_nms(noise, th)
    want = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det.detect(noise)]
    assert [(int(k["x"]), int(k["y"]), int(k["score"])) for k in got] == want


def test_fast_atan2_matches_cv2(oracle):
    rng = np.random.default_rng(0)
    ys = rng.integers(-200000, 200000, 4000); xs = rng.integers(-200000, 200000, 4000)
    err = max(abs(oracle.fast_atan2(float(y), float(x)) - cv2.fastAtan2(float(y), float(x))) for x, y in zip(xs, ys))
    assert err <= 1e-4          # degrees; SURVEY.md A4: <= 3.05e-5 (1 ulp), tolerance is 0.0573 deg
    assert oracle.fast_atan2(0.0, 0.0) == 0.0


def test_border_matches_cv2(oracle):
    img = synth_frame(2, 97, 61)
    assert np.array_equal(oracle.reflect101_border(img, 19), cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))


@pytest.mark.parametrize("case", ["vga", "kitti", "checker"])
def test_full_extractor_matches_cv2_tier_a(oracle, case):
    from oracle import pin_cv2 as P
    img, params = {"vga": (synth_frame(5), {}), "kitti": (synth_frame(2, 1241, 376), {"nfeatures": 2000}),
                   "checker": (adversarial_frame("checker", 320, 240), {})}[case]
    rep = P.compare(img, **params)
    assert rep["pyramid_px"] == 0 and rep["blur_px"] == 0 and rep["cand_mismatch"] == 0 and rep["kept_mismatch"] == 0
    assert rep["n_tier_a"] == rep["n_oracle"] and rep.get("xy_equal", True)
    assert rep["angle_max_abs_deg"] <= 1e-4 and rep["desc_bits_diff"] <= 1e-3 * max(rep["desc_bits"], 1)


@pytest.mark.parametrize("code,channels,rgb", [("COLOR_BGR2GRAY", 3, False), ("COLOR_RGB2GRAY", 3, True),
                                               ("COLOR_BGRA2GRAY", 4, False), ("COLOR_RGBA2GRAY", 4, True)])
def test_cvt_gray_matches_cv2(oracle, code, channels, rgb):
    """The conversion in front of the extractor (src/Tracking.cpp:57-70), SURVEY.md 8f-2."""
    img = np.random.default_rng(channels + rgb).integers(0, 256, (97, 131, channels), dtype=np.uint8)
    assert np.array_equal(oracle.cvt_gray(img, rgb), cv2.cvtColor(img, getattr(cv2, code)))


def test_undistort_keypoints_matches_cv2(oracle):
    """Frame::UndistortedKeyPoints (src/Frame.cpp:80-109), SURVEY.md 8f-3: cv::undistortPoints(pts, K, dist, Mat(), K)."""
    rng = np.random.default_rng(0)
    K = np.array([[517.3, 0, 318.6], [0, 516.5, 255.3], [0, 0, 1]], np.float32)
    dist = np.array([0.2624, -0.9531, -0.0054, 0.0026, 1.1633], np.float32)          # TUM fr1-like
    kps = np.zeros(2000, oracle.KEYPOINT_DTYPE)
    kps["x"] = rng.random(2000) * 640; kps["y"] = rng.random(2000) * 480; kps["octave"] = 3
    want = cv2.undistortPoints(np.stack([kps["x"], kps["y"]], 1).reshape(-1, 1, 2), K, dist, None, K).reshape(-1, 2)
    got = oracle.undistort_keypoints(kps, [K[0, 0], K[1, 1], K[0, 2], K[1, 2]], dist)
    assert np.array_equal(got["x"], want[:, 0]) and np.array_equal(got["y"], want[:, 1])
    assert (got["octave"] == 3).all()
    bug = oracle.undistort_keypoints(kps, [K[0, 0], K[1, 1], K[0, 2], K[1, 2]], dist, literal_bug=True)
    assert np.array_equal(bug["y"], want[:, 0])                                        # Frame.cpp:106 as written
    same = oracle.undistort_keypoints(kps, [K[0, 0], K[1, 1], K[0, 2], K[1, 2]], [0, 0.1, 0, 0, 0])
    assert np.array_equal(same["x"], kps["x"]) and np.array_equal(same["y"], kps["y"])  # Frame.cpp:82-86
