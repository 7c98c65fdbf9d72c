"""CPU: SearchForInitialization.  The oracle restatement against the reference's OWN
src/ORBmatcher.cpp compiled unmodified over a Frame shim (oracle/_ref/ref_match)."""
import numpy as np
import pytest

from oracle import ref as R
from orbslam_in_practice_b200.synth import synth_frame

pytestmark = pytest.mark.skipif(not (R.match_available() or R.build_match()), reason="oracle/_ref/ref_match not built")


def _pair(oracle, seed, shift, nfeatures=2000):
    a = synth_frame(seed)
    b = np.roll(np.roll(a, shift[0], axis=1), shift[1], axis=0)
    ex = oracle.OracleExtractor(nfeatures)
    k1, d1 = ex(a); k2, d2 = ex(b)
    return k1, d1, k2, d2


@pytest.mark.parametrize("seed,shift,ratio,ori,bug", [(0, (5, 3), 0.9, True, False), (1, (12, -7), 0.9, True, False),
                                                     (2, (0, 0), 0.7, False, False), (3, (40, 25), 0.9, True, False),
                                                     (4, (5, 3), 0.9, True, True)])
def test_oracle_equals_reference_matcher_tu(oracle, seed, shift, ratio, ori, bug):
    k1, d1, k2, d2 = _pair(oracle, seed, shift)
    prev = np.stack([k1["x"], k1["y"]], 1)                       # Tracking.cpp:170-176: prev matched = frame-1 keypoints
    n_o, m_o, p_o = oracle.search_for_initialization(k1, d1, k2, d2, prev, 100, ratio, ori, 640, 480, bug)
    n_r, m_r, p_r = R.run_search_for_initialization(k1, d1, k2, d2, prev, 100, ratio, ori, 640, 480, bug)
    assert n_o == n_r and np.array_equal(m_o, m_r) and np.array_equal(p_o, p_r)
    if bug:
        assert n_r == 0            # Frame.cpp:164 as written: nothing ever lands in the grid
    else:
        assert n_r > 50
        # second call with the updated prev-matched (the tracker's retry loop)
        n_o2, m_o2, _ = oracle.search_for_initialization(k1, d1, k2, d2, p_o, 100, ratio, ori, 640, 480, bug)
        n_r2, m_r2, _ = R.run_search_for_initialization(k1, d1, k2, d2, p_r, 100, ratio, ori, 640, 480, bug)
        assert n_o2 == n_r2 and np.array_equal(m_o2, m_r2)


@pytest.mark.parametrize("seed,shift,ratio,ori", [(0, (5, 3), 0.9, True), (3, (40, 25), 0.7, False)])
def test_windowed_search_with_init_parameters_is_search_for_initialization(oracle, seed, shift, ratio, ori):
    """The generalised windowed search (gate 0, octave 0, fixed window, TH_LOW, ratio) must reproduce the pinned function."""
    k1, d1, k2, d2 = _pair(oracle, seed, shift)
    prev = np.stack([k1["x"], k1["y"]], 1)
    n_a, m_a, p_a = oracle.search_for_initialization(k1, d1, k2, d2, prev, 100, ratio, ori, 640, 480, False)
    P = oracle.window_params(100.0, None, (0, 0), 0, 0, gate=0, th_dist=50, nnratio=ratio, check_orientation=ori,
                             update_centers=True, width=640, height=480)
    n_b, m_b, p_b = oracle.search_window(k1, d1, k2, d2, prev, P)
    assert n_a == n_b and np.array_equal(m_a, m_b) and np.array_equal(p_a, p_b)


def test_projection_style_search_properties(oracle):
    """Gate 1 (upstream SearchByProjection; reference body empty -> unpinned): checked through its defining properties."""
    k1, d1, k2, d2 = _pair(oracle, 5, (4, -3))
    ex = oracle.OracleExtractor(2000)
    sf = [float(v) for v in ex.scale_factors]
    cen = np.stack([k1["x"] + 4, k1["y"] - 3], 1).astype(np.float32)     # the "projection": the true shift
    cen[::7, 0] = np.nan                                                  # points without a projection
    P = oracle.window_params(7.0, sf, (0, 15), 1, 1, gate=1, th_dist=100, nnratio=0.0, check_orientation=True,
                             update_centers=False, width=640, height=480)
    n, m12, cen_out = oracle.search_window(k1, d1, k2, d2, cen, P)
    assert np.array_equal(cen_out, cen, equal_nan=True)                  # centres untouched
    assert np.all(m12[::7] == -1)
    hit = np.flatnonzero(m12 >= 0)
    assert n == len(hit) and n > 300
    assert len(np.unique(m12[hit])) == len(hit)                          # one-to-one by construction of gate 1
    oct1, oct2 = k1["octave"][hit], k2["octave"][m12[hit]]
    assert np.all(np.abs(oct1 - oct2) <= 1)
    r = 7.0 * np.asarray(sf, np.float32)[oct1]
    assert np.all(np.abs(k2["x"][m12[hit]] - cen[hit, 0]) < r) and np.all(np.abs(k2["y"][m12[hit]] - cen[hit, 1]) < r)
    dist = np.unpackbits(d1[hit] ^ d2[m12[hit]], axis=1).sum(1)
    assert np.all(dist <= 100)


def test_group_search_properties(oracle):
    """SearchByBoW-style search (reference body empty -> unpinned): matches stay inside a group, are one-to-one, pass the
    acceptance, and with a single all-embracing group + generous ratio the first query takes the global best match."""
    k1, d1, k2, d2 = _pair(oracle, 6, (3, 2))
    g1 = (d1[:, 0] >> 3).astype(np.uint16); g2 = (d2[:, 0] >> 3).astype(np.uint16)
    g1[::11] = 0xffff
    n, m12 = oracle.search_groups(k1, d1, g1, k2, d2, g2, 50, 0.7, True)
    hit = np.flatnonzero(m12 >= 0)
    assert n == len(hit) and n > 200
    assert np.all(g1[hit] == g2[m12[hit]]) and np.all(g1[hit] != 0xffff)
    assert len(np.unique(m12[hit])) == len(hit)
    assert np.all(np.unpackbits(d1[hit] ^ d2[m12[hit]], axis=1).sum(1) <= 50)
    one = np.zeros(len(k1), np.uint16); two = np.zeros(len(k2), np.uint16)
    n1, m1 = oracle.search_groups(k1[:1], d1[:1], one[:1], k2, d2, two, 256, 10.0, False)
    dist = np.unpackbits(d1[:1] ^ d2, axis=1).sum(1)
    assert n1 == 1 and m1[0] == int(np.argmin(dist))


BOUNDS = (40.5, 600.25, 30.75, 440.5)        # FindimageBound of a distorted lens (src/Frame.cpp:121-141): min over the undistorted corners, can cut into the image


@pytest.mark.parametrize("seed,shift,ratio,ori", [(0, (5, 3), 0.9, True), (7, (-9, 14), 0.7, False)])
def test_explicit_grid_bounds_equal_reference_matcher_tu(oracle, seed, shift, ratio, ori):
    """Float grid bounds (a distorted lens): the windowed oracle with explicit bounds against the reference's own
    ORBmatcher.cpp over a Frame whose statics hold those bounds; and bounds = {0, w, 0, h} is the default case."""
    k1, d1, k2, d2 = _pair(oracle, seed, shift)
    prev = np.stack([k1["x"], k1["y"]], 1)
    mk = lambda b: oracle.window_params(100.0, None, (0, 0), 0, 0, gate=0, th_dist=50, nnratio=ratio, check_orientation=ori,
                                        update_centers=True, width=640, height=480, bounds=b)
    n_b, m_b, p_b = oracle.search_window(k1, d1, k2, d2, prev, mk(BOUNDS))
    n_r, m_r, p_r = R.run_search_for_initialization(k1, d1, k2, d2, prev, 100, ratio, ori, 640, 480, False, bounds=BOUNDS)
    assert n_b == n_r and np.array_equal(m_b, m_r) and np.array_equal(p_b, p_r)
    assert n_r > 50
    n_0, m_0, p_0 = oracle.search_window(k1, d1, k2, d2, prev, mk(None))
    n_1, m_1, p_1 = oracle.search_window(k1, d1, k2, d2, prev, mk((0.0, 640.0, 0.0, 480.0)))
    assert n_0 == n_1 and np.array_equal(m_0, m_1) and np.array_equal(p_0, p_1)
    assert not np.array_equal(m_0, m_b)          # the bounds do change the grid (otherwise this test pins nothing)
