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
