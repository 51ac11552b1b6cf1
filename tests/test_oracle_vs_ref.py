"""CPU: the plain-C oracle against the reference's own translation units
(oracle/_ref, compiled from /root/reference/src) on inputs beyond the committed
fixtures: other shapes, up-sampling, edge cases.  Skipped when oracle/_ref was
not built (it needs the reference tree)."""
import numpy as np
import pytest

from openpano_b200 import synth
from openpano_b200._abi import default_params
from tests import golden_util as gu


def _same_trace(a, b, noct=4, nscale=7):
    assert a.working_size() == b.working_size()
    assert gu.same_bits(a.plane(0), b.plane(0))
    for o in range(noct):
        assert a.octave_size(o) == b.octave_size(o)
        for l in range(nscale):
            assert gu.same_bits(a.plane(1, o, l), b.plane(1, o, l)), (o, l)
        for l in range(nscale - 1):
            assert gu.same_bits(a.plane(2, o, l), b.plane(2, o, l)), (o, l)
        for l in range(1, nscale):
            assert gu.same_bits(a.plane(3, o, l), b.plane(3, o, l)), ("mag", o, l)
            assert gu.same_bits(a.plane(4, o, l), b.plane(4, o, l)), ("ort", o, l)
    for st in range(3):
        assert a.points(st).tobytes() == b.points(st).tobytes(), st
    ca, da = a.descriptors()
    cb, db = b.descriptors()
    assert gu.same_bits(ca, cb) and gu.same_bits(da, db)
    return len(da)


@pytest.mark.parametrize("w,h,seed", [(300, 200, 7), (200, 300, 8), (157, 211, 9)])
def test_sift_every_stage(orc, ref, w, h, seed):
    img = synth.make_canvas(h, w, seed)
    a, b = orc.sift_trace(img), ref.sift_trace(img)
    assert _same_trace(a, b) > 20
    a.close(); b.close()


def test_sift_other_params(orc, ref):
    img = synth.make_canvas(180, 260, 31)
    p = default_params(num_octave=3, num_scale=6, contrast_thres=3e-2, edge_ratio=10.0, sift_working_size=300)
    a, b = orc.sift_trace(img, p), ref.sift_trace(img, p)
    _same_trace(a, b, noct=3, nscale=6)
    a.close(); b.close()


@pytest.mark.parametrize("hist_scale,ori_radius", [(8, 4.5), (17, 9.0)])
def test_sift_wide_windows(orc, ref, hist_scale, ori_radius):
    """The parameter sets tests/test_gpu_sift.py::test_sift_wide_descriptor_windows runs on the GPU (descriptor
    windows wider than one interval-table block of the kernel, wide orientation windows): the restatement
    against the reference's own TUs, so that the GPU-vs-oracle result there is pinned to the reference too."""
    img = synth.make_canvas(200, 280, 41)
    p = default_params(desc_hist_scale_factor=hist_scale, ori_radius=ori_radius)
    (ca, da), (cb, db) = orc.sift_detect(img, p), ref.sift_detect(img, p)
    assert len(da) > 100
    assert gu.same_bits(ca, cb) and gu.same_bits(da, db)


def test_sift_flat_image(orc, ref):
    img = np.full((120, 160, 3), 0.25, np.float32)
    assert len(orc.sift_detect(img)[1]) == 0 and len(ref.sift_detect(img)[1]) == 0


def test_match_ragged_and_ties(orc, ref):
    rng = np.random.RandomState(3)
    a = synth.rootsift_like(260, 5)
    b = np.concatenate([a[:80], a[:80], a[150:]])                   # exact duplicates: zero-distance ties
    noisy = (a[rng.permutation(260)][:200] + rng.randn(200, 128).astype(np.float32) * 36).astype(np.float32)
    for x, y in ((a, b), (b, a), (a, noisy), (noisy, a), (a[:1], b), (b, a[:1]), (a[:2], a[:2])):
        assert np.array_equal(orc.match(x, y), ref.match(x, y))


@pytest.mark.parametrize("w,h,hf", [(200, 140, 1.0), (141, 173, 0.8), (160, 120, 1.3)])
def test_cyl_warp(orc, ref, w, h, hf):
    img = synth.make_canvas(h, w, 61)
    k = np.array([[3.5, -2.25], [-60.0, 40.0]])
    assert orc.cyl_warp_shape(w, h, hf) == ref.cyl_warp_shape(w, h, hf)
    oa, ka = orc.cyl_warp(img, k, hf)
    ob, kb = ref.cyl_warp(img, k, hf)
    assert gu.same_bits(oa, ob) and gu.same_bits(ka, kb)


@pytest.mark.parametrize("projection", [0, 1, 2])
@pytest.mark.parametrize("bands", [0, 2])
def test_blend_projections(orc, ref, projection, bands, capfd):
    import math
    imgs, org = synth.make_stack(3, 160, 110, 60, 71)
    items = []
    for k, (x, y) in enumerate(org):
        if projection == 0:
            th = 0.003 * (k - 1)
            H = np.array([[math.cos(th), -math.sin(th), x - 80], [math.sin(th), math.cos(th), 2 * k], [1e-5 * k, -2e-5, 1.0]])
        else:
            f = 300.0
            H = np.array([[1 / f, 0, (x - 80) / f], [0, 1 / f, 0.004 * k], [0, 0, 1]])
        items.append((k * 60, 0, k * 60 + 160, 115, list(np.linalg.inv(H).ravel())))
    res = 1.0 if projection == 0 else 1 / 300.0
    pmin = (-80.0, -55.0) if projection == 0 else (-0.3, -0.2)
    geom = dict(projection=projection, res_x=res, res_y=res, proj_min_x=pmin[0], proj_min_y=pmin[1])
    a = orc.blend(imgs, items, geom, bands)
    b = ref.blend(imgs, items, geom, bands)
    assert (a >= 0).mean() > 0.3
    assert gu.same_bits(a, b)


def _random_mosaic(rng, h, w):
    m = rng.rand(h, w, 3).astype(np.float32)
    for _ in range(rng.randint(0, 6)):
        y0, x0 = rng.randint(0, h), rng.randint(0, w)
        m[y0:y0 + rng.randint(1, 8), x0:x0 + rng.randint(1, 10)] = -1
    if rng.rand() < 0.4:
        m[:rng.randint(0, 4)] = -1
        m[:, :rng.randint(0, 5)] = -1
    return m


def test_imgio_read_write(orc, ref):
    """read_img / write_rgb through the reference's imgio.cc (lossless PNM files) vs the restatement."""
    rng = np.random.RandomState(11)
    allv = np.arange(256, dtype=np.uint8).reshape(16, 16)
    for pix in (rng.randint(0, 256, (37, 53, 3)).astype(np.uint8), np.stack([allv] * 3, -1), allv,
                rng.randint(0, 256, (9, 31)).astype(np.uint8)):
        assert np.array_equal(orc.read_img_rgb8(pix), ref.read_img_rgb8(pix))
    m = _random_mosaic(rng, 40, 60)
    m[0, 0] = (1.0, 0.0, 0.999999)
    m[0, 1] = np.float32(1.0) / np.float32(255.0) * np.arange(1, 4, dtype=np.float32)
    assert np.array_equal(orc.write_rgb8(m), ref.write_rgb8(m))


def test_imgio_crop(orc, ref):
    rng = np.random.RandomState(12)
    cases = [_random_mosaic(rng, rng.randint(5, 60), rng.randint(5, 90)) for _ in range(25)]
    cases.append(-np.ones((6, 7, 3), np.float32))                  # nothing valid: 0 x 1 result
    cases.append(rng.rand(8, 9, 3).astype(np.float32))              # everything valid
    for m in cases:
        ro, co = orc.crop(m)
        rr, cr = ref.crop(m)
        assert np.array_equal(ro[2:], rr[2:])
        assert np.array_equal(co, cr)
        x0, y0, cw, ch = ro
        assert np.array_equal(co, m[y0:y0 + ch, x0:x0 + cw])


_RATIO_SCRIPT = """
import sys
sys.path.insert(0, {root!r})
import numpy as np
from openpano_b200 import synth
from openpano_b200._abi import default_params
from tests.checker import get_checker
orc, ref = get_checker('orc'), get_checker('ref')
rng = np.random.RandomState(21)
a = synth.rootsift_like(220, 22)
b = (a[rng.permutation(220)][:190] + rng.randn(190, 128).astype(np.float32) * 30).astype(np.float32)
p = default_params(match_reject_next_ratio={ratio})
ab, ba = orc.match(a, b, p), orc.match(b, a, p)
assert np.array_equal(ab, ref.match(a, b, p)) and np.array_equal(ba, ref.match(b, a, p))
print(len(ab))
"""


@pytest.mark.parametrize("ratio,lo,hi", [(0.6, 1, 60), (0.95, 150, 190)])
def test_match_other_ratios(ref, ratio, lo, hi):
    """MATCH_REJECT_NEXT_RATIO is a config value (config.cfg:33).  The reference squares it into a
    function-local `static const` on the FIRST call (matcher.cc:16, :91), i.e. it is frozen per
    process exactly like the config file is — so every ratio is pinned in a fresh process."""
    import subprocess
    import sys
    from pathlib import Path
    root = str(Path(__file__).resolve().parent.parent)
    out = subprocess.run([sys.executable, "-c", _RATIO_SCRIPT.format(root=root, ratio=ratio)], capture_output=True,
                         text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-1500:]
    assert lo <= int(out.stdout.strip().splitlines()[-1]) <= hi


@pytest.mark.parametrize("lazy,ordered", [(0, 0), (1, 1)])
def test_blend_scaled_resolution(orc, ref, lazy, ordered):
    """MAX_OUTPUT_SIZE shrinks the canvas through `resolution` (stitcher_image.cc:108-119): the
    blend map then samples the sources at a stride > 1."""
    imgs, org = synth.make_stack(3, 200, 150, 80, 73)
    items, geom = synth.translation_blend_setup(org, 200, 150, max_output_size=170)
    assert geom["res_x"] > 2.0
    p = default_params(lazy_read=lazy, ordered_input=ordered)
    a, b = orc.blend(imgs, items, geom, 0, p), ref.blend(imgs, items, geom, 0, p)
    assert gu.same_bits(a, b) and (a >= 0).mean() > 0.5
    a, b = orc.blend(imgs, items, geom, 3, p), ref.blend(imgs, items, geom, 3, p)
    assert gu.same_bits(a, b)


def test_cyl_warp_other_focal(orc, ref):
    img = synth.make_canvas(120, 180, 62)
    p = default_params(focal_length=24.0)
    k = np.array([[10.0, 5.0], [-80.0, -50.0]])
    assert orc.cyl_warp_shape(180, 120, 1.0, p) == ref.cyl_warp_shape(180, 120, 1.0, p)
    oa, ka = orc.cyl_warp(img, k, 1.0, p)
    ob, kb = ref.cyl_warp(img, k, 1.0, p)
    assert gu.same_bits(oa, ob) and gu.same_bits(ka, kb)


def test_oracle_mt_equals_oracle(orc):
    """oracle/liboracle_mt.so (independent loops under OpenMP, used by the BASELINE-size GPU
    parity tests) must be bit-identical to the single-thread restatement."""
    from tests.checker import get_checker, have
    if not have("orc_mt"):
        pytest.skip("oracle/liboracle_mt.so not built")
    mt = get_checker("orc_mt")
    rng = np.random.RandomState(21)
    a = synth.rootsift_like(900, 4)
    b = a[rng.permutation(900)][:700] + rng.randn(700, 128).astype(np.float32) * 25.0
    for x, y in ((a, b), (b, a), (a[:1], b), (a, a)):
        assert np.array_equal(orc.match(x, y), mt.match(x, y))
    imgs, org = synth.make_stack(5, 260, 200, 90, 77, rows=2, step_y=70)
    items, geom = synth.translation_blend_setup(org, 260, 200)
    for bands in (0, 2, 5):
        for lazy in (0, 1):
            p = default_params(lazy_read=lazy, multiband=bands)
            assert gu.same_bits(orc.blend(imgs, items, geom, bands, p), mt.blend(imgs, items, geom, bands, p)), (bands, lazy)


@pytest.mark.parametrize("n_match,n_hyp,seed", [(300, 200, 1), (8, 50, 2), (1200, 64, 3)])
def test_ransac_scoring(orc, ref, n_match, n_hyp, seed):
    """TransformEstimation::get_inliers itself (the reference TU, reached through the shim) against
    the restatement: per-hypothesis counts, first-maximum selection, inlier flags."""
    from tests.ransac_util import ransac_case
    kp1, kp2, homos, thres = ransac_case(n_match, n_hyp, seed)
    a, b = orc.ransac_score(kp1, kp2, homos, thres), ref.ransac_score(kp1, kp2, homos, thres)
    assert a[0] == b[0] and a[1] == b[1]
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    assert a[1] > 0.5 * n_match


@pytest.mark.parametrize("n_cam,per_pair,seed,extra", [(4, 50, 1, 2), (8, 300, 2, 6), (3, 1, 3, 0), (16, 1000, 4, 30)])
def test_ba_jacobian(orc, ref, n_cam, per_pair, seed, extra):
    """IncrementalBundleAdjuster::calcJacobianSymbolic itself (the reference TU, compiled where it lies
    by oracle/refshim/ref_ba.cc) against the restatement: every J row and every J^T J entry, bit for bit —
    including reversed pairs, repeated camera pairs and the identity camera's small-angle branch."""
    from tests.ba_util import ba_case
    cams, pairs, pts = ba_case(n_cam, per_pair, seed, extra_pairs=extra)
    mats = ref.ba_pair_mats(cams, pairs)
    r_rows, r_jtj = ref.ba_jacobian_ref(cams, pairs, pts)
    o_rows, o_jtj = orc.ba_jacobian(n_cam, pairs, mats, pts[:, :2])
    assert gu.same_bits(r_rows, o_rows)
    assert gu.same_bits(r_jtj, o_jtj)
    assert np.isfinite(r_rows).all() and (r_jtj != 0).any()
