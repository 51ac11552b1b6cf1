"""GPU parity of the bundle-adjustment Jacobian assembly against the oracle and
the golden fixture made by the reference's own calcJacobianSymbolic: every J row and every
J^T J entry bit for bit, at fixture size, at panorama size (38 cameras) and at the size the
reference's comment names ("J.rows() could reach 700000")."""
import numpy as np
import pytest

from tests import golden_util as gu
from tests.ba_util import ba_case, numpy_pair_mats

pytestmark = pytest.mark.gpu


def test_ba_jacobian_matches_golden(engine):
    g = gu.load("ba_5cams.npz")
    cams, pairs, pts = ba_case(5, 40, 5, extra_pairs=3)
    rows, jtj = engine.ba_jacobian(5, [(f, t, n, m) for (f, t, n), m in zip(pairs, g["mats"])], pts[:, :2])
    assert gu.same_bits(rows, g["rows"]) and gu.same_bits(jtj, g["jtj"])


@pytest.mark.parametrize("n_cam,per_pair,seed,extra", [(3, 1, 3, 0), (38, 400, 7, 60), (24, 9000, 8, 16)])
def test_ba_jacobian_bit_exact(engine, orc, n_cam, per_pair, seed, extra):
    cams, pairs, pts = ba_case(n_cam, per_pair, seed, extra_pairs=extra)
    mats = numpy_pair_mats(cams, pairs)       # any 13 matrices are valid inputs of the per-point code
    want_rows, want_jtj = orc.ba_jacobian(n_cam, pairs, mats, pts[:, :2])
    rows, jtj = engine.ba_jacobian(n_cam, [(f, t, n, m) for (f, t, n), m in zip(pairs, mats)], pts[:, :2])
    assert gu.same_bits(rows, want_rows)
    assert gu.same_bits(jtj, want_jtj)
    if per_pair >= 9000:
        assert 2 * len(pts) > 600000         # rows of J
    # J^T J without the rows coming back
    _, jtj2 = engine.ba_jacobian(n_cam, [(f, t, n, m) for (f, t, n), m in zip(pairs, mats)], pts[:, :2], want_rows=False)
    assert gu.same_bits(jtj2, want_jtj)


def test_ba_jacobian_degenerate(engine, orc):
    """No pairs: J^T J is all zeros; a pair without matches contributes nothing; bad slots are refused."""
    _, jtj = engine.ba_jacobian(3, [], np.zeros((0, 2)))
    assert jtj.shape == (18, 18) and not jtj.any()
    cams, pairs, pts = ba_case(4, 30, 9)
    mats = numpy_pair_mats(cams, pairs)
    f, t, n = pairs[1]
    cut = sum(p[2] for p in pairs[:1])
    pairs2 = [pairs[0], (f, t, 0)] + pairs[2:]
    pts2 = np.concatenate([pts[:cut], pts[cut + n:]], 0)
    want = orc.ba_jacobian(4, pairs2, mats, pts2[:, :2])
    got = engine.ba_jacobian(4, [(a, b, c, m) for (a, b, c), m in zip(pairs2, mats)], pts2[:, :2])
    assert gu.same_bits(got[0], want[0]) and gu.same_bits(got[1], want[1])
    from openpano_b200.capi import PanoError
    with pytest.raises(PanoError):
        engine.ba_jacobian(2, [(0, 5, 1, mats[0])], pts[:1, :2])
