"""GPU: the CUDA engine, through the C ABI, against the committed golden vectors
(generated from the reference's own translation units).  Bit-exact."""
import numpy as np
import pytest

from openpano_b200._abi import default_params
from tests import golden_util as gu

pytestmark = pytest.mark.gpu


def test_engine_sift_matches_golden(engine):
    tr = engine.sift_trace(gu.sift_input())
    gu.check_sift_trace(tr, gu.load("sift_240x180.npz"))
    tr.close()


def test_engine_match_matches_golden(engine):
    g = gu.load("match.npz")
    imgs = gu.match_inputs()
    fs = engine.sift_detect_batch(imgs)
    d0, d1 = fs.download(0)[1], fs.download(1)[1]
    assert gu.same_bits(d0, g["d0"]) and gu.same_bits(d1, g["d1"])
    m01, m10 = engine.match_pairs(fs, [(0, 1), (1, 0)])
    fs.free()
    assert np.array_equal(m01, g["pairs_01"]) and np.array_equal(m10, g["pairs_10"])
    for x, y, key in (("a", "b", "pairs_ab"), ("b", "a", "pairs_ba"), ("a", "c", "pairs_ac"), ("c", "a", "pairs_ca")):
        assert np.array_equal(engine.match_bruteforce(g[x], g[y]), g[key]), key


def test_engine_cyl_warp_matches_golden(engine):
    g = gu.load("cyl_warp_120x80.npz")
    img = gu.warp_input()
    assert tuple(g["shape"]) == engine.cyl_warp_shape(120, 80)
    out, kk = engine.cyl_warp(img, g["kpts_in"], 1.0)
    assert gu.same_bits(out, g["out"]) and gu.same_bits(kk, g["kpts_out"])
    out, kk = engine.cyl_warp(img, g["kpts_in"], 0.9)
    assert gu.same_bits(out, g["out_h09"]) and gu.same_bits(kk, g["kpts_out_h09"])


@pytest.mark.parametrize("key,lazy,ordered,bands", [
    ("linear_lazy0_ord0", 0, 0, 0), ("linear_lazy0_ord1", 0, 1, 0), ("linear_lazy1_ord0", 1, 0, 0),
    ("linear_lazy1_ord1", 1, 1, 0), ("multiband_1", 1, 0, 1), ("multiband_3", 1, 0, 3), ("multiband_5", 1, 0, 5)])
def test_engine_blend_matches_golden(engine, key, lazy, ordered, bands):
    g = gu.load("blend_3x120x80.npz")
    imgs, items, geom = gu.blend_inputs()
    out = engine.blend(imgs, items, geom, bands, default_params(lazy_read=lazy, ordered_input=ordered))
    assert gu.same_bits(out, g[key])


def test_stitcher_end_to_end_equals_stagewise(engine, orc):
    """The Stitcher mirror (features -> match -> blend in one build()) must equal
    the stage-wise oracle results."""
    from openpano_b200 import synth
    from openpano_b200.stitcher import Stitcher, ordered_pairs
    imgs, org = synth.make_stack(4, 320, 240, 110, 13)
    items, geom = synth.translation_blend_setup(org, 320, 240)
    p = default_params(ordered_input=1)
    st = Stitcher(engine, p)
    pairs = ordered_pairs(4)
    matches, mosaic = st.build_numpy(imgs, pairs, items, geom, 0)
    st.close()
    descs = [orc.sift_detect(im, p)[1] for im in imgs]
    for (i, j), m in zip(pairs, matches):
        assert np.array_equal(m, orc.match(descs[i], descs[j], p))
    assert gu.same_bits(mosaic, orc.blend(imgs, items, geom, 0, p))
