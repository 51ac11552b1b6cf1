"""GPU parity of the RANSAC inlier scoring against the oracle: identical
per-hypothesis inlier counts, identical choice (first maximum), identical inlier sets, for
identical hypotheses — several pairs in one call, ragged sizes, empty pairs."""
import numpy as np
import pytest

from tests.ransac_util import ransac_case

pytestmark = pytest.mark.gpu


def test_ransac_scoring_bit_exact(engine, orc):
    cases = [ransac_case(931, 1500, 11), ransac_case(8, 40, 12), ransac_case(2500, 300, 13), ransac_case(77, 1, 14)]
    got = engine.ransac_score_pairs(cases)
    for k, (case, g) in enumerate(zip(cases, got)):
        want = orc.ransac_score(*case)
        assert g[0] == want[0] and g[1] == want[1], (k, g[:2], want[:2])
        assert np.array_equal(g[2], want[2]), k
        assert np.array_equal(g[3], want[3]), k
        assert g[1] == int(g[3].sum())
    assert got[0][1] > 500


def test_ransac_scoring_degenerate(engine, orc):
    kp1, kp2, homos, thres = ransac_case(20, 5, 21)
    empty_h = np.zeros((0, 9))
    empty_m = np.zeros((0, 2))
    got = engine.ransac_score_pairs([(kp1, kp2, empty_h, thres), (empty_m, empty_m, homos, thres), (kp1, kp2, homos, thres)])
    assert got[0][0] == -1 and got[0][1] == 0 and not got[0][3].any()
    assert got[1][1] == 0 and got[1][0] == 0                  # every hypothesis has 0 inliers: the first one wins
    want = orc.ransac_score(kp1, kp2, homos, thres)
    assert got[2][0] == want[0] and np.array_equal(got[2][3], want[3])
    # a hypothesis that sends points to infinity / behind the camera must not crash or count
    bad = homos.copy()
    bad[0] = [0, 0, 0, 0, 0, 0, 0, 0, 0]
    bad[1] = [1, 0, 0, 0, 1, 0, 0, 0, -1e-300]
    g = engine.ransac_score_pairs([(kp1, kp2, bad, thres)])[0]
    w = orc.ransac_score(kp1, kp2, bad, thres)
    assert g[0] == w[0] and np.array_equal(g[2], w[2]) and np.array_equal(g[3], w[3])
