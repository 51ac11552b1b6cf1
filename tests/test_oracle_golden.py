"""CPU: the plain-C oracle against the golden vectors generated from the
reference's own translation units (tests/golden/make_golden.py).  Bit-exact."""
import numpy as np
import pytest

from openpano_b200._abi import default_params
from tests import golden_util as gu


def test_fixture_inputs_unchanged():
    """The synthetic generator must reproduce the arrays the fixtures were made from."""
    assert gu.sha(gu.sift_input()) == str(gu.load("sift_240x180.npz")["input_sha"])
    assert gu.sha(gu.warp_input()) == str(gu.load("cyl_warp_120x80.npz")["input_sha"])
    assert gu.sha(*gu.blend_inputs()[0]) == str(gu.load("blend_3x120x80.npz")["input_sha"])
    g = gu.load("match.npz")
    imgs = gu.match_inputs()
    assert gu.sha(imgs[0], imgs[1], g["a"], g["b"], g["c"]) == str(g["input_sha"])


def test_oracle_sift_matches_golden(orc):
    tr = orc.sift_trace(gu.sift_input())
    gu.check_sift_trace(tr, gu.load("sift_240x180.npz"))
    tr.close()


def test_oracle_match_matches_golden(orc):
    g = gu.load("match.npz")
    imgs = gu.match_inputs()
    d0 = orc.sift_detect(imgs[0])[1]
    d1 = orc.sift_detect(imgs[1])[1]
    assert gu.same_bits(d0, g["d0"]) and gu.same_bits(d1, g["d1"])
    assert np.array_equal(orc.match(d0, d1), g["pairs_01"])
    assert np.array_equal(orc.match(d1, d0), g["pairs_10"])
    for x, y, key in (("a", "b", "pairs_ab"), ("b", "a", "pairs_ba"), ("a", "c", "pairs_ac"), ("c", "a", "pairs_ca")):
        assert np.array_equal(orc.match(g[x], g[y]), g[key]), key


def test_oracle_cyl_warp_matches_golden(orc):
    g = gu.load("cyl_warp_120x80.npz")
    img = gu.warp_input()
    assert tuple(g["shape"]) == orc.cyl_warp_shape(120, 80)
    out, kk = orc.cyl_warp(img, g["kpts_in"], 1.0)
    assert gu.same_bits(out, g["out"]) and gu.same_bits(kk, g["kpts_out"])
    out, kk = orc.cyl_warp(img, g["kpts_in"], 0.9)
    assert gu.same_bits(out, g["out_h09"]) and gu.same_bits(kk, g["kpts_out_h09"])


@pytest.mark.parametrize("key,lazy,ordered,bands", [
    ("linear_lazy0_ord0", 0, 0, 0), ("linear_lazy0_ord1", 0, 1, 0), ("linear_lazy1_ord0", 1, 0, 0),
    ("linear_lazy1_ord1", 1, 1, 0), ("multiband_1", 1, 0, 1), ("multiband_3", 1, 0, 3), ("multiband_5", 1, 0, 5)])
def test_oracle_blend_matches_golden(orc, key, lazy, ordered, bands):
    g = gu.load("blend_3x120x80.npz")
    imgs, items, geom = gu.blend_inputs()
    out = orc.blend(imgs, items, geom, bands, default_params(lazy_read=lazy, ordered_input=ordered))
    assert gu.same_bits(out, g[key])


def test_oracle_imgio_matches_golden(orc):
    """read_img / crop / write_rgb restatements against the reference's lib/imgio.cc + imgproc.cc output."""
    g = gu.load("imgio.npz")
    pix, grey, mos = gu.imgio_inputs()
    assert gu.sha(pix, grey, mos) == str(g["input_sha"])
    assert gu.same_bits(orc.read_img_rgb8(pix), g["read_rgb"])
    assert gu.same_bits(orc.read_img_rgb8(grey), g["read_grey"])
    rect, cropped = orc.crop(mos)
    assert np.array_equal(rect[2:], g["crop_wh"])
    assert gu.same_bits(cropped, g["cropped"])
    x0, y0, cw, ch = rect
    assert gu.same_bits(cropped, mos[y0:y0 + ch, x0:x0 + cw])
    assert gu.same_bits(orc.write_rgb8(mos), g["write_full"])
    assert gu.same_bits(orc.write_rgb8(cropped), g["write_cropped"])


def test_oracle_ba_jacobian_matches_golden(orc):
    """J rows and J^T J of the reference's calcJacobianSymbolic (fixture made by its own TU)."""
    from tests.ba_util import ba_case
    g = gu.load("ba_5cams.npz")
    cams, pairs, pts = ba_case(5, 40, 5, extra_pairs=3)
    assert gu.sha(cams, np.array(pairs), pts) == str(g["input_sha"])
    rows, jtj = orc.ba_jacobian(5, pairs, g["mats"], pts[:, :2])
    assert gu.same_bits(rows, g["rows"]) and gu.same_bits(jtj, g["jtj"])
