"""GPU parity: CUDA extractor (through the C ABI) vs the CPU oracle, stage by stage.

Bars (BASELINE.json north_star): pyramid pixels, FAST scores/candidates, kept keypoint sets:
bit-exact; angles within 1e-3 rad; >= 99.9 % descriptor bits identical.
"""
import numpy as np
import pytest

from orbslam_in_practice_b200.synth import synth_frame, synth_batch, adversarial_frame

pytestmark = pytest.mark.gpu

ANGLE_TOL_DEG = 1e-3 * 180.0 / np.pi   # 1e-3 rad


def _compare_frame(O, ex, oex, img, frame, kps, desc, counts, check_stages=True):
    ko, do = oex(img)
    n = int(counts[frame])
    assert n == len(ko), "keypoint count differs: gpu %d oracle %d" % (n, len(ko))
    kg, dg = kps[frame, :n], desc[frame, :n]
    if check_stages:
        for l in range(ex.nlevels):
            assert np.array_equal(ex.level(frame, l), oex.level(l)), "pyramid level %d pixels differ" % l
            cg, co = ex.candidates(frame, l), oex.candidates(l)
            assert len(cg) == len(co) and np.array_equal(cg, co), "FAST candidates differ at level %d" % l
            kg_l, ko_l = ex.kept(frame, l), oex.kept(l)
            assert len(kg_l) == len(ko_l) and np.array_equal(kg_l, ko_l), "octree survivors differ at level %d" % l
            ob = oex.blurred(l)
            if ob is not None:
                assert np.array_equal(ex.level(frame, l, blurred=True), ob), "blurred level %d differs" % l
    for fld in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(kg[fld], ko[fld]), "keypoint field %s differs" % fld
    if n:
        d = np.abs(kg["angle"] - ko["angle"]); d = np.minimum(d, 360.0 - d)
        assert d.max() <= ANGLE_TOL_DEG, "angle error %g deg" % d.max()
        bits = np.unpackbits(dg ^ do).sum()
        assert bits <= 1e-3 * dg.size * 8, "descriptor bits differing: %d of %d" % (bits, dg.size * 8)
    return n


def test_single_vga_frame_all_stages(orbx, oracle):
    img = synth_frame(0)
    ex = orbx.Extractor(max_width=640, max_height=480, max_batch=1)
    oex = oracle.OracleExtractor()
    kps, desc, counts = ex.extract_host(img)
    n = _compare_frame(oracle, ex, oex, img, 0, kps, desc, counts)
    assert n > 900
    assert np.array_equal(ex.features_per_level, oex.features_per_level)
    assert np.array_equal(ex.scale_factors, oex.scale_factors)
    assert np.array_equal(ex.umax, oex.umax)


def test_pyramid_border_is_reflect101(orbx, oracle):
    img = synth_frame(5)
    ex = orbx.Extractor(max_width=640, max_height=480, max_batch=1)
    ex.extract_host(img)
    oex = oracle.OracleExtractor(); oex(img)
    for l in range(8):
        assert np.array_equal(ex.level(0, l, border=19), oracle.reflect101_border(oex.level(l), 19))


def test_batch_of_frames(orbx, oracle):
    seeds = list(range(8))
    imgs = synth_batch(seeds)
    ex = orbx.Extractor(max_width=640, max_height=480, max_batch=8)
    oex = oracle.OracleExtractor()
    kps, desc, counts = ex.extract_host(imgs)
    for f in range(len(seeds)):
        _compare_frame(oracle, ex, oex, imgs[f], f, kps, desc, counts, check_stages=(f in (0, 7)))


@pytest.mark.parametrize("kind", ["constant", "noise", "checker"])
def test_adversarial_frames(orbx, oracle, kind):
    img = adversarial_frame(kind, 320, 240) if kind == "noise" else adversarial_frame(kind)
    H, W = img.shape
    ex = orbx.Extractor(max_width=W, max_height=H, max_batch=1)
    oex = oracle.OracleExtractor()
    kps, desc, counts = ex.extract_host(img)
    n = _compare_frame(oracle, ex, oex, img, 0, kps, desc, counts)
    if kind == "constant":
        assert n == 0


def test_kitti_sized_stereo_pair(orbx, oracle):
    imgs = synth_batch([0, 1], 1241, 376)
    ex = orbx.Extractor(nfeatures=2000, max_width=1241, max_height=376, max_batch=2)
    oex = oracle.OracleExtractor(nfeatures=2000)
    kps, desc, counts = ex.extract_host(imgs)
    for f in range(2):
        _compare_frame(oracle, ex, oex, imgs[f], f, kps, desc, counts)


def test_other_parameters_and_smaller_frame_on_same_handle(orbx, oracle):
    ex = orbx.Extractor(nfeatures=500, scale_factor=1.5, nlevels=5, ini_th=30, min_th=10,
                        max_width=640, max_height=480, max_batch=2)
    oex = oracle.OracleExtractor(500, 1.5, 5, 30, 10)
    for (w, h, seed) in [(640, 480, 11), (512, 384, 12), (640, 480, 13)]:
        img = synth_frame(seed, w, h)
        kps, desc, counts = ex.extract_host(img)
        _compare_frame(oracle, ex, oex, img, 0, kps, desc, counts)


def test_empty_image_returns_zero_counts(orbx):
    ex = orbx.Extractor(max_width=64, max_height=64, max_batch=1)
    counts = np.full(1, 7, np.int32)
    ex.extract_host_ptr(None, 0, 0, 0, 0, 1, None, None, counts.ctypes.data)
    assert counts[0] == 0


def test_capacity_errors(orbx):
    ex = orbx.Extractor(max_width=320, max_height=240, max_batch=1)
    with pytest.raises(orbx.OrbxError):
        ex.extract_host(np.zeros((480, 640), np.uint8))
    with pytest.raises(orbx.OrbxError):
        ex.extract_host(np.zeros((2, 240, 320), np.uint8))
    ok = orbx.Extractor(nfeatures=2300, nlevels=1, max_width=640, max_height=480, max_batch=1)
    assert ok.extract_host(np.zeros((480, 640), np.uint8))[2][0] == 0


@pytest.mark.parametrize("low_latency", [True, False])
def test_level_quota_beyond_one_sm(orbx, oracle, low_latency):
    """A per-level quota whose octree node tables do not fit one SM's shared memory (> ~2 400 features on one level) runs
    with the tables in global memory: same keypoints as the oracle, through the graph path and the stream path."""
    imgs = np.stack([adversarial_frame("noise", 640, 480, seed=3), synth_frame(9)])
    params = dict(nfeatures=3900, nlevels=1)
    ex = orbx.Extractor(max_width=640, max_height=480, max_batch=2, **params)
    ex.set_low_latency(low_latency)
    oex = oracle.OracleExtractor(**params)
    kps, desc, counts = ex.extract_host(imgs)
    assert _compare_frame(oracle, ex, oex, imgs[0], 0, kps, desc, counts) > 3000
    _compare_frame(oracle, ex, oex, imgs[1], 1, kps, desc, counts)
    # two levels, the first one beyond the limit
    params = dict(nfeatures=5000, nlevels=2, scale_factor=1.5)
    ex = orbx.Extractor(max_width=640, max_height=480, max_batch=2, **params)
    ex.set_low_latency(low_latency)
    oex = oracle.OracleExtractor(**params)
    kps, desc, counts = ex.extract_host(imgs)
    _compare_frame(oracle, ex, oex, imgs[0], 0, kps, desc, counts)
    _compare_frame(oracle, ex, oex, imgs[1], 1, kps, desc, counts)


def test_4k_frame_nfeatures_8000(orbx, oracle):
    """BASELINE configs[4] frame size: 3840x2160, nfeatures=8000 (one frame here; the oracle needs ~1 s for it)."""
    img = synth_frame(0, 3840, 2160)
    ex = orbx.Extractor(nfeatures=8000, max_width=3840, max_height=2160, max_batch=2)
    oex = oracle.OracleExtractor(nfeatures=8000)
    imgs = np.stack([img, img[::-1].copy()])
    kps, desc, counts = ex.extract_host(imgs)
    n = _compare_frame(oracle, ex, oex, imgs[0], 0, kps, desc, counts)
    assert n >= 7900
    _compare_frame(oracle, ex, oex, imgs[1], 1, kps, desc, counts, check_stages=False)


def test_stereo_left_right_matching(orbx, oracle):
    """BASELINE configs[2]: KITTI-sized pair, 2000 features per side, left->right brute-force best-2 + ratio 0.7."""
    imgs = synth_batch([0, 1], 1241, 376)
    imgs[1] = np.roll(imgs[0], 7, axis=1)                      # a shifted copy so that matches exist
    ex = orbx.Extractor(nfeatures=2000, max_width=1241, max_height=376, max_batch=2)
    kps, desc, counts = ex.extract_host(imgs)
    nl, nr = int(counts[0]), int(counts[1])
    m = orbx.Matcher(nl, nr)
    d1, i1, d2 = m.knn2_host(desc[0, :nl], desc[1, :nr])
    o1, oi, o2 = oracle.knn2(desc[0, :nl], desc[1, :nr], 0, 4)
    assert np.array_equal(d1, o1) and np.array_equal(i1, oi) and np.array_equal(d2, o2)
    match = oracle.ratio_select(d1, i1, d2, 50, 0.7)
    assert (match >= 0).sum() > nl // 4


@pytest.mark.parametrize("channels,rgb", [(3, False), (3, True), (4, False)])
def test_colour_input_fused_gray_conversion(orbx, oracle, channels, rgb):
    """cvtColor(..2GRAY) (src/Tracking.cpp:57-70) fused into the level-0 kernel."""
    rng = np.random.default_rng(channels)
    base = synth_batch([3, 4])
    col = np.repeat(base[..., None], channels, axis=3).astype(np.int32) + rng.integers(-20, 21, base.shape + (channels,))
    col = np.clip(col, 0, 255).astype(np.uint8)
    ex = orbx.Extractor(max_width=640, max_height=480, max_batch=2)
    ex.set_input_format(channels, rgb)
    kps, desc, counts = ex.extract_host(col)
    oex = oracle.OracleExtractor()
    for f in range(2):
        gray = oracle.cvt_gray(col[f], rgb)
        assert np.array_equal(ex.level(f, 0), gray), "level 0 is not the OpenCV fixed-point gray image"
        _compare_frame(oracle, ex, oex, gray, f, kps, desc, counts, check_stages=(f == 0))   # blurred levels: the 4-px border too


@pytest.mark.parametrize("channels,w,h", [(3, 333, 250), (4, 333, 250), (3, 656, 100), (4, 48, 40)])
def test_colour_input_other_widths(orbx, oracle, channels, w, h):
    """colour rows that are not 16-byte aligned (3 x 333 bytes) take the byte-wise conversion, aligned ones the 16-byte-load form;
    the last chunk of a 656-px row also writes the reflected border.  Blurred levels are compared, so the border counts."""
    rng = np.random.default_rng(channels + w)
    base = synth_frame(9, w, h)
    col = np.clip(np.repeat(base[..., None], channels, axis=2).astype(np.int32) + rng.integers(-20, 21, base.shape + (channels,)), 0, 255).astype(np.uint8)
    ex = orbx.Extractor(nlevels=4 if min(w, h) < 100 else 8, max_width=w, max_height=h, max_batch=1)
    ex.set_input_format(channels, True)
    kps, desc, counts = ex.extract_host(np.ascontiguousarray(col))
    oex = oracle.OracleExtractor(1000, 1.2, ex.nlevels, 20, 7)
    gray = oracle.cvt_gray(col, True)
    assert np.array_equal(ex.level(0, 0), gray)
    _compare_frame(oracle, ex, oex, gray, 0, kps, desc, counts, check_stages=True)


def test_undistort_keypoints(orbx, oracle):
    """Frame::UndistortedKeyPoints on the device vs the oracle (pinned against cv2.undistortPoints)."""
    import torch
    img = synth_frame(9)
    ex = orbx.Extractor(max_width=640, max_height=480, max_batch=1)
    kps, desc, counts = ex.extract_host(img)
    n = int(counts[0])
    cam = [517.3, 516.5, 318.6, 255.3]; dist = [0.2624, -0.9531, -0.0054, 0.0026, 1.1633]
    d_in = torch.from_numpy(kps[0, :n].view(np.uint8).reshape(n, 28)).cuda()
    d_out = torch.zeros_like(d_in)
    ex.undistort_keypoints_device(d_in.data_ptr(), d_out.data_ptr(), n, cam, dist, False, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(oracle.KEYPOINT_DTYPE).reshape(n)
    want = oracle.undistort_keypoints(kps[0, :n].astype(oracle.KEYPOINT_DTYPE), cam, dist)
    assert np.abs(got["x"] - want["x"]).max() <= 1e-3 and np.abs(got["y"] - want["y"]).max() <= 1e-3   # px; observed: identical
    for fld in ("size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(got[fld], want[fld])
    ex.undistort_keypoints_device(d_in.data_ptr(), d_out.data_ptr(), n, cam, [0, 0, 0, 0, 0])
    torch.cuda.synchronize()
    assert torch.equal(d_in, d_out)


@pytest.mark.parametrize("w,h,nf", [(333, 257, 600), (801, 603, 1500), (1023, 767, 1000), (752, 480, 1200), (97, 65, 100)])
def test_odd_frame_sizes(orbx, oracle, w, h, nf):
    img = synth_frame(w + h, w, h)
    ex = orbx.Extractor(nfeatures=nf, max_width=w, max_height=h, max_batch=1)
    oex = oracle.OracleExtractor(nfeatures=nf)
    kps, desc, counts = ex.extract_host(img)
    _compare_frame(oracle, ex, oex, img, 0, kps, desc, counts)


def test_strided_host_input_and_two_handles(orbx, oracle):
    """Row pitch > width and a frame stride that is not rows * pitch; two extractors alive at once
    (the reference keeps an nFeatures and a 2*nFeatures extractor, src/Tracking.cpp:47-48)."""
    big = np.zeros((3, 500, 700), np.uint8)
    frames = synth_batch([40, 41, 42])
    big[:, 10:490, 30:670] = frames
    view = big[::2, 10:490, 30:670]                                  # frames 0 and 2, pitch 700, stride 2*500*700
    ex1 = orbx.Extractor(nfeatures=1000, max_width=640, max_height=480, max_batch=2)
    ex2 = orbx.Extractor(nfeatures=2000, max_width=640, max_height=480, max_batch=2)
    k1, d1, c1 = ex1.extract_host(view)
    k2, d2, c2 = ex2.extract_host(view)
    o1, o2 = oracle.OracleExtractor(1000), oracle.OracleExtractor(2000)
    for slot, f in enumerate((0, 2)):
        _compare_frame(oracle, ex1, o1, frames[f], slot, k1, d1, c1, check_stages=False)
        _compare_frame(oracle, ex2, o2, frames[f], slot, k2, d2, c2, check_stages=False)
    # eager 19-px border mode gives the same keypoints and a device-resident border
    ex1.set_pyramid_border(True)
    k3, d3, c3 = ex1.extract_host(view)
    assert np.array_equal(c1, c3) and np.array_equal(d1, d3)
    oex = oracle.OracleExtractor(1000); oex(frames[0])
    assert np.array_equal(ex1.level(0, 3, border=19), oracle.reflect101_border(oex.level(3), 19))


@pytest.mark.parametrize("params", [dict(nfeatures=50, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7),
                                    dict(nfeatures=3000, scale_factor=1.1, nlevels=12, ini_th=12, min_th=5),
                                    dict(nfeatures=700, scale_factor=2.0, nlevels=3, ini_th=40, min_th=40),
                                    dict(nfeatures=400, scale_factor=1.2, nlevels=1, ini_th=7, min_th=20),
                                    dict(nfeatures=300, scale_factor=2.9, nlevels=3, ini_th=20, min_th=7),
                                    dict(nfeatures=300, scale_factor=3.6, nlevels=2, ini_th=20, min_th=7)])
def test_parameter_corners(orbx, oracle, params):
    """Tiny and huge quotas (phase-2-only and never-phase-2 octrees), many / one level, iniTh < minTh, and scale
    factors at / beyond the staged resize kernel's limit of 3 (3.6 runs the gather kernel)."""
    img = synth_frame(77)
    ex = orbx.Extractor(max_width=640, max_height=480, max_batch=1, **params)
    oex = oracle.OracleExtractor(**params)
    kps, desc, counts = ex.extract_host(img)
    _compare_frame(oracle, ex, oex, img, 0, kps, desc, counts)


def test_wide_panorama_many_octree_roots(orbx, oracle):
    """Aspect ratio 8:1 -> nIni = 8 root nodes (ORBextractor.cpp:493): two sweeps of the 4-roots-per-sweep partition."""
    img = synth_frame(5, 1600, 200)
    ex = orbx.Extractor(nfeatures=1500, max_width=1600, max_height=200, max_batch=1)
    oex = oracle.OracleExtractor(nfeatures=1500)
    kps, desc, counts = ex.extract_host(img)
    _compare_frame(oracle, ex, oex, img, 0, kps, desc, counts)


def test_portrait_frame_without_octree_roots(orbx, oracle):
    """width/height < 0.5 -> nIni = round(...) = 0: the reference divides by zero there; oracle and device both
    return no keypoints for such levels (documented deviation)."""
    img = synth_frame(6, 200, 640)
    ex = orbx.Extractor(nfeatures=500, max_width=200, max_height=640, max_batch=1)
    oex = oracle.OracleExtractor(nfeatures=500)
    kps, desc, counts = ex.extract_host(img)
    _compare_frame(oracle, ex, oex, img, 0, kps, desc, counts)


def test_begin_end_pipelining_across_two_handles(orbx, oracle):
    """orbx_extract_host_begin/_end: two batches in flight on two handles give the same results as blocking calls."""
    import torch
    batches = [synth_batch([50, 51, 52]), synth_batch([53, 54, 55])]
    exs = [orbx.Extractor(max_width=640, max_height=480, max_batch=3) for _ in range(2)]
    cap = exs[0].capacity
    bufs = []
    for b in batches:
        hf = torch.from_numpy(b).pin_memory()
        bufs.append((hf, torch.zeros((3, cap, 7)).pin_memory(), torch.zeros((3, cap, 32), dtype=torch.uint8).pin_memory(),
                     torch.zeros(3, dtype=torch.int32).pin_memory()))
    for rep in range(3):
        for e, (hf, hk, hd, hc) in zip(exs, bufs):
            e.extract_host_begin(hf.data_ptr(), 640, 640 * 480, 640, 480, 3, hk.data_ptr(), hd.data_ptr(), hc.data_ptr())
        for e in exs:
            e.extract_host_end()
    oex = oracle.OracleExtractor()
    for b, e, (hf, hk, hd, hc) in zip(batches, exs, bufs):
        kps = hk.numpy().view(np.float32).reshape(3, cap, 7)
        ks = np.zeros((3, cap), orbx.KEYPOINT_DTYPE)
        for i, fld in enumerate(("x", "y", "size", "angle", "response")):
            ks[fld] = kps[:, :, i]
        ks["octave"] = kps[:, :, 5].view(np.int32); ks["class_id"] = kps[:, :, 6].view(np.int32)
        for f in range(3):
            _compare_frame(oracle, e, oex, b[f], f, ks, hd.numpy(), hc.numpy(), check_stages=False)


def _random_cases(n, seed):
    rng = np.random.default_rng(seed)
    cases = []
    while len(cases) < n:
        w, h = int(rng.integers(70, 901)), int(rng.integers(70, 701))
        sf = float(rng.choice([1.1, 1.2, 1.25, 1.3, 1.5, 1.7, 2.0, 2.5]))
        nl = int(rng.integers(1, 11))
        # the reference needs one 30-px FAST cell in every level (it divides by zero otherwise)
        if min(w, h) / sf ** (nl - 1) < 70:
            continue
        ini = int(rng.integers(5, 41)); mn = int(rng.integers(2, ini + 1))
        cases.append((w, h, int(rng.integers(30, 2501)), sf, nl, ini, mn, int(rng.integers(0, 1 << 30))))
    return cases


@pytest.mark.parametrize("case", _random_cases(14, 20260101), ids=lambda c: "%dx%d_n%d_s%g_l%d_t%d-%d" % c[:7])
def test_random_geometry_sweep(orbx, oracle, case):
    """Seeded random frame sizes / parameters, eager and lazy border: every stage bit-exact against the oracle."""
    w, h, nf, sf, nl, ini, mn, seed = case
    img = synth_frame(seed % 1000, w, h)
    params = dict(nfeatures=nf, scale_factor=sf, nlevels=nl, ini_th=ini, min_th=mn)
    ex = orbx.Extractor(max_width=w, max_height=h, max_batch=2, **params)
    oex = oracle.OracleExtractor(**params)
    if seed & 1:
        ex.set_pyramid_border(True)
    kps, desc, counts = ex.extract_host(np.stack([img, img[::-1].copy()]))
    _compare_frame(oracle, ex, oex, img, 0, kps, desc, counts)
    _compare_frame(oracle, ex, oex, img[::-1].copy(), 1, kps, desc, counts)
    oex(img)
    assert np.array_equal(ex.level(0, nl - 1, border=19), oracle.reflect101_border(oex.level(nl - 1), 19))


def test_device_entry_sub_batches_give_identical_results(orbx, oracle):
    """orbx_extract_device on >= 32 resident frames runs as independent sub-batches on two streams
    (orbx_set_device_split): outputs must not depend on the split, and equal the oracle's."""
    import torch
    F = 40
    imgs = synth_batch(range(100, 100 + F), 320, 240)
    dev = torch.device("cuda:0")
    d_img = torch.from_numpy(imgs).to(dev)
    ex = orbx.Extractor(nfeatures=500, max_width=320, max_height=240, max_batch=F)
    cap = ex.capacity
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    outs = []
    for split in (1, 2, 3):
        ex.set_device_split(split)
        k = torch.zeros((F, cap, 7), dtype=torch.float32, device=dev)
        d = torch.zeros((F, cap, 32), dtype=torch.uint8, device=dev)
        c = torch.zeros(F, dtype=torch.int32, device=dev)
        ex.extract_device(d_img.data_ptr(), 320, 320 * 240, 320, 240, F, k.data_ptr(), d.data_ptr(), c.data_ptr(), st.cuda_stream)
        st.synchronize()
        outs.append((k.cpu().numpy().view(np.int32), d.cpu().numpy(), c.cpu().numpy()))   # bit patterns (class_id -1 is a NaN as float)
    for k, d, c in outs[1:]:
        assert np.array_equal(c, outs[0][2])
        for f in range(F):
            n = int(c[f])
            assert np.array_equal(k[f, :n], outs[0][0][f, :n]) and np.array_equal(d[f, :n], outs[0][1][f, :n])
    oex = oracle.OracleExtractor(nfeatures=500)
    for f in (0, 19, 20, 39):                       # both sides of the sub-batch boundaries
        ko, do = oex(imgs[f])
        n = int(outs[1][2][f])
        assert n == len(ko)
        assert np.array_equal(outs[1][0][f, :n, 0].view(np.float32), ko["x"]) and np.array_equal(outs[1][0][f, :n, 1].view(np.float32), ko["y"])
        assert np.unpackbits(outs[1][1][f, :n] ^ do).sum() <= 1e-3 * do.size * 8


def test_low_latency_graph_path_equals_stream_path(orbx, oracle):
    """Small batches replay the kernels as a CUDA graph with per-level FAST/octree branches (orbx_set_low_latency, the default);
    the stream path is the same kernels in one chain.  Both must give the oracle's result: pageable and pinned caller
    buffers, a padded row pitch, 1..8 frames per call, repeated calls (graph replay) and a change of batch size (re-capture)."""
    import torch
    imgs = synth_batch(range(40, 48))
    oex = oracle.OracleExtractor()
    want = [oex(im) for im in imgs]
    ex = orbx.Extractor(max_width=640, max_height=480, max_batch=8)
    ref = orbx.Extractor(max_width=640, max_height=480, max_batch=8)
    ref.set_low_latency(False)

    def check(kps, desc, counts, frames):
        for i, f in enumerate(frames):
            ko, do = want[f]
            n = int(counts[i])
            assert n == len(ko)
            for fld in ("x", "y", "size", "response", "octave", "class_id"):
                assert np.array_equal(kps[i, :n][fld], ko[fld]), fld
            d = np.abs(kps[i, :n]["angle"] - ko["angle"]); d = np.minimum(d, 360.0 - d)
            assert d.max() <= ANGLE_TOL_DEG
            assert np.unpackbits(desc[i, :n] ^ do).sum() <= 1e-3 * do.size * 8

    for n in (1, 1, 3, 8, 2, 1):                          # repeats replay the cached graph; size changes re-capture it
        fr = list(range(n))
        out = ex.extract_host(imgs[:n]); check(*out, fr)
        out2 = ref.extract_host(imgs[:n]); check(*out2, fr)
        for f in range(n):
            c = int(out[2][f])
            assert c == int(out2[2][f]) and out[0][f, :c].tobytes() == out2[0][f, :c].tobytes() and np.array_equal(out[1][f, :c], out2[1][f, :c])
    # padded row pitch (a cv::Mat ROI), pageable
    wide = np.zeros((2, 480, 704), np.uint8); wide[:, :, :640] = imgs[4:6]
    check(*ex.extract_host(wide[:, :, :640]), [4, 5])
    # pinned caller buffers
    pinned = torch.from_numpy(imgs[6:8].copy()).pin_memory()
    check(*ex.extract_host(pinned.numpy()), [6, 7])
    # the intermediate buffers are the ones the stage accessors read
    kps, desc, counts = ex.extract_host(imgs[:1])
    _compare_frame(oracle, ex, oex, imgs[0], 0, kps, desc, counts)
    # one handle, alternating frame sizes: the geometry tables are re-uploaded and the graph re-captured each time
    small = synth_batch([50, 51], 320, 240)
    oex_small = [oex(im) for im in small]
    for rep in range(2):
        k2, d2, c2 = ex.extract_host(small[rep:rep + 1])
        ko, do = oex_small[rep]
        n = int(c2[0])
        assert n == len(ko) and np.array_equal(k2[0, :n]["x"], ko["x"]) and np.array_equal(k2[0, :n]["y"], ko["y"])
        assert np.unpackbits(d2[0, :n] ^ do).sum() <= 1e-3 * do.size * 8
        check(*ex.extract_host(imgs[rep:rep + 1]), [rep])
    # device pointers: the second graph slot
    dev = torch.device("cuda:0")
    d_f = torch.from_numpy(imgs[:2].copy()).to(dev); cap = ex.capacity
    d_k = torch.zeros((2, cap, 7), dtype=torch.float32, device=dev); d_d = torch.zeros((2, cap, 32), dtype=torch.uint8, device=dev)
    d_c = torch.zeros(2, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(device=dev)
    for _ in range(3):
        ex.extract_device(d_f.data_ptr(), 640, 640 * 480, 640, 480, 2, d_k.data_ptr(), d_d.data_ptr(), d_c.data_ptr(), st.cuda_stream)
    st.synchronize()
    kd = d_k.cpu().numpy().view(np.uint8).reshape(2, cap, 28).copy().view(oracle.KEYPOINT_DTYPE).reshape(2, cap)
    check(kd, d_d.cpu().numpy(), d_c.cpu().numpy(), [0, 1])


@pytest.mark.parametrize("offset,pitch_pad,w,h", [(1, 0, 640, 480), (2, 3, 333, 250), (3, 5, 1241, 376), (5, 0, 96, 80), (0, 1, 640, 480)])
def test_device_input_at_any_alignment(orbx, oracle, offset, pitch_pad, w, h):
    """k_level0 reads source rows that are not 16-byte aligned as aligned words + funnel shifts: a device image at a byte offset
    and with a row pitch that is no multiple of 4 must give the same pyramid (incl. the reflected border) and keypoints."""
    import torch
    img = synth_frame(77 + offset, w, h)
    pitch = w + pitch_pad
    buf = np.zeros(offset + pitch * h + 64, np.uint8)
    buf[offset:offset + pitch * h].reshape(h, pitch)[:, :w] = img
    buf[offset:offset + pitch * h].reshape(h, pitch)[:, w:] = 255           # the padding must never be read as pixels
    d = torch.from_numpy(buf).cuda()
    ex = orbx.Extractor(max_width=w, max_height=h, max_batch=1)
    cap = ex.capacity
    d_k = torch.zeros((1, cap, 7), dtype=torch.float32, device="cuda"); d_d = torch.zeros((1, cap, 32), dtype=torch.uint8, device="cuda")
    d_c = torch.zeros(1, dtype=torch.int32, device="cuda")
    ex.extract_device(d.data_ptr() + offset, pitch, pitch * h, w, h, 1, d_k.data_ptr(), d_d.data_ptr(), d_c.data_ptr(), 0)
    torch.cuda.synchronize()
    oex = oracle.OracleExtractor()
    ko, do = oex(img)
    assert int(d_c.item()) == len(ko)
    assert np.array_equal(ex.level(0, 0, border=19), oracle.reflect101_border(oex.level(0), 19))
    for l in range(1, ex.nlevels):
        assert np.array_equal(ex.level(0, l), oex.level(l)), "pyramid level %d differs" % l
    kp = np.zeros(cap, orbx.KEYPOINT_DTYPE)
    kp.view(np.uint8).reshape(cap, 28)[:] = d_k.cpu().numpy().view(np.uint8).reshape(cap, 28)
    n = len(ko)
    for fld in ("x", "y", "response", "octave"):
        assert np.array_equal(kp[:n][fld], ko[fld])
    assert np.array_equal(d_d.cpu().numpy()[0, :n], do)
