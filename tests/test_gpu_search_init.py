"""GPU parity, SURVEY.md 8f-1: ORBmatcher::SearchForInitialization + the Frame grid (AssignFeaturesToGrid /
GetFeaturesInArea) on the device vs the oracle (which is pinned against the reference's own ORBmatcher.cpp)."""
import numpy as np
import pytest

from orbslam_in_practice_b200.synth import synth_frame

pytestmark = pytest.mark.gpu


def _frames():
    a = synth_frame(0); b = np.roll(np.roll(a, 5, axis=1), 3, axis=0)
    c = synth_frame(1); d = np.roll(np.roll(c, -12, axis=1), 9, axis=0)
    return np.stack([a, b, c, d])


@pytest.mark.parametrize("ratio,ori,bug", [(0.9, True, False), (0.7, False, False), (0.9, True, True)])
def test_search_for_initialization_matches_oracle(orbx, oracle, ratio, ori, bug):
    import torch
    imgs = _frames()
    F = len(imgs)
    ex = orbx.Extractor(nfeatures=2000, max_width=640, max_height=480, max_batch=F)
    cap = ex.capacity
    dev = torch.device("cuda:0")
    d_img = torch.from_numpy(imgs).to(dev)
    d_kps = torch.zeros((F, cap, 7), dtype=torch.float32, device=dev)
    d_desc = torch.zeros((F, cap, 32), dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros(F, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st); s = st.cuda_stream
    ex.extract_device(d_img.data_ptr(), 640, 640 * 480, 640, 480, F, d_kps.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr(), s)
    pairs = [(0, 1), (1, 0), (2, 3), (0, 2)]
    P = len(pairs)
    pa = torch.tensor([p[0] for p in pairs], dtype=torch.int32, device=dev)
    pb = torch.tensor([p[1] for p in pairs], dtype=torch.int32, device=dev)
    prev = torch.stack([d_kps[p[0], :, :2] for p in pairs]).contiguous()      # Tracking.cpp:170-176: prev = F1 keypoints
    prev0 = prev.clone()
    m12 = torch.full((P, cap), -7, dtype=torch.int32, device=dev)
    nm = torch.zeros(P, dtype=torch.int32, device=dev)
    m = orbx.Matcher(cap, cap)
    wsb = orbx.load().orbm_search_init_workspace_bytes(cap, P)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    m.search_init_device(d_kps.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr(), cap, pa.data_ptr(), pb.data_ptr(), P,
                         prev.data_ptr(), m12.data_ptr(), nm.data_ptr(), 100, ratio, ori, 640, 480, ws.data_ptr(), wsb, s, bug)
    torch.cuda.synchronize()
    kps = d_kps.cpu().numpy().view(np.float32).reshape(F, cap, 7)
    kps_s = np.zeros((F, cap), oracle.KEYPOINT_DTYPE)
    for i, fld in enumerate(("x", "y", "size", "angle", "response")):
        kps_s[fld] = kps[:, :, i]
    kps_s["octave"] = kps[:, :, 5].view(np.int32); kps_s["class_id"] = kps[:, :, 6].view(np.int32)
    desc = d_desc.cpu().numpy(); cnt = d_cnt.cpu().numpy()
    got_m, got_n, got_prev = m12.cpu().numpy(), nm.cpu().numpy(), prev.cpu().numpy()
    for p, (a, b) in enumerate(pairs):
        n1, n2 = int(cnt[a]), int(cnt[b])
        n_o, m_o, p_o = oracle.search_for_initialization(kps_s[a, :n1], desc[a, :n1], kps_s[b, :n2], desc[b, :n2],
                                                         prev0[p, :n1].cpu().numpy(), 100, ratio, ori, 640, 480, bug)
        assert got_n[p] == n_o, "pair %s: nmatches %d vs %d" % ((a, b), got_n[p], n_o)
        assert np.array_equal(got_m[p, :n1], m_o)
        assert np.array_equal(got_prev[p, :n1], p_o)
        if bug:
            assert n_o == 0
        elif (a, b) in ((0, 1), (1, 0), (2, 3)):
            assert n_o > 50


def test_search_init_workspace_too_small_is_reported(orbx):
    import torch
    imgs = _frames()[:2]
    ex = orbx.Extractor(nfeatures=500, max_width=640, max_height=480, max_batch=2)
    cap = ex.capacity
    dev = torch.device("cuda:0")
    d_img = torch.from_numpy(imgs).to(dev)
    d_kps = torch.zeros((2, cap, 7), dtype=torch.float32, device=dev); d_desc = torch.zeros((2, cap, 32), dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros(2, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    ex.extract_device(d_img.data_ptr(), 640, 640 * 480, 640, 480, 2, d_kps.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr(), s)
    torch.cuda.synchronize()
    pa = torch.tensor([0], dtype=torch.int32, device=dev); pb = torch.tensor([1], dtype=torch.int32, device=dev)
    prev = d_kps[0:1, :, :2].contiguous(); m12 = torch.zeros((1, cap), dtype=torch.int32, device=dev); nm = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = torch.empty(64, dtype=torch.uint8, device=dev)
    m = orbx.Matcher(cap, cap)
    m.search_init_device(d_kps.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr(), cap, pa.data_ptr(), pb.data_ptr(), 1,
                         prev.data_ptr(), m12.data_ptr(), nm.data_ptr(), 100, 0.9, True, 640, 480, ws.data_ptr(), 64, s)
    torch.cuda.synchronize()
    assert int(nm[0]) == -1


def test_search_init_host_entry(orbx, oracle):
    imgs = _frames()[:2]
    ex = orbx.Extractor(nfeatures=2000, max_width=640, max_height=480, max_batch=2)
    kps, desc, cnt = ex.extract_host(imgs)
    n1, n2 = int(cnt[0]), int(cnt[1])
    k1, d1, k2, d2 = kps[0, :n1], desc[0, :n1], kps[1, :n2], desc[1, :n2]
    prev = np.stack([k1["x"], k1["y"]], 1)
    m = orbx.Matcher(max(n1, n2), max(n1, n2))
    got = m.search_init_host(k1, d1, k2, d2, prev, 100, 0.9, True, 640, 480)
    want = oracle.search_for_initialization(k1.astype(oracle.KEYPOINT_DTYPE), d1, k2.astype(oracle.KEYPOINT_DTYPE), d2, prev, 100, 0.9, True, 640, 480)
    assert got[0] == want[0] and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    assert got[0] > 50
    # empty second frame / empty first frame
    assert m.search_init_host(k1, d1, k2[:0], d2[:0], prev)[0] == 0
    assert m.search_init_host(k1[:0], d1[:0], k2, d2, prev[:0])[0] == 0


@pytest.mark.parametrize("th,below,above,ori", [(7.0, 1, 1, True), (15.0, -1, 0, False), (7.0, 0, -1, True)])
def test_projection_style_windowed_search_matches_oracle(orbx, oracle, th, below, above, ori):
    """per-query windows (r = th * scaleFactor[octave]), octave ranges, first-come gate, best <= TH_HIGH.
    The three cases are upstream SearchByProjection's default / bBackward / bForward level ranges."""
    a = synth_frame(5); b = np.roll(np.roll(a, 4, axis=1), -3, axis=0)
    ex = orbx.Extractor(nfeatures=2000, max_width=640, max_height=480, max_batch=2)
    kps, desc, cnt = ex.extract_host(np.stack([a, b]))
    k1, d1, k2, d2 = kps[0][:cnt[0]], desc[0][:cnt[0]], kps[1][:cnt[1]], desc[1][:cnt[1]]
    sf = [float(v) for v in ex.scale_factors]
    cen = np.stack([k1["x"] + 4, k1["y"] - 3], 1).astype(np.float32)
    cen[::7, 0] = np.nan
    P = orbx.WindowParams.projection(th, sf, 640, 480, th_dist=100, check_orientation=ori, level_below=below, level_above=above)
    Po = oracle.window_params(th, sf, (0, 15), below, above, gate=1, th_dist=100, nnratio=0.0, check_orientation=ori,
                              update_centers=False, width=640, height=480)
    m = orbx.Matcher(4096, 4096)
    n_g, m_g, c_g = m.search_window_host(k1, d1, k2, d2, cen, P)
    n_o, m_o, c_o = oracle.search_window(k1, d1, k2, d2, cen, Po)
    assert n_g == n_o and np.array_equal(m_g, m_o) and np.array_equal(c_g, c_o, equal_nan=True)
    assert n_o > 300


@pytest.mark.parametrize("ngroups,ratio,ori", [(32, 0.7, True), (1, 0.6, False), (200, 0.9, True)])
def test_group_restricted_search_matches_oracle(orbx, oracle, ngroups, ratio, ori):
    """SURVEY 8f-4, SearchByBoW form: candidates = the other frame's keypoints of the same group (vocabulary node)."""
    a = synth_frame(6); b = np.roll(np.roll(a, 3, axis=1), 2, axis=0)
    ex = orbx.Extractor(nfeatures=2000, max_width=640, max_height=480, max_batch=2)
    kps, desc, cnt = ex.extract_host(np.stack([a, b]))
    k1, d1, k2, d2 = kps[0][:cnt[0]], desc[0][:cnt[0]], kps[1][:cnt[1]], desc[1][:cnt[1]]
    g1 = (d1[:, 0].astype(np.uint32) * ngroups // 256).astype(np.uint16)
    g2 = (d2[:, 0].astype(np.uint32) * ngroups // 256).astype(np.uint16)
    g1[::11] = 0xffff; g2[5::13] = 0xffff
    m = orbx.Matcher(4096, 4096)
    n_g, m_g = m.search_groups_host(k1, d1, g1, k2, d2, g2, 50, ratio, ori)
    n_o, m_o = oracle.search_groups(k1, d1, g1, k2, d2, g2, 50, ratio, ori)
    assert n_g == n_o and np.array_equal(m_g, m_o)
    assert n_o > 100
