"""GPU: the 8-bit boundary kernels (read_img conversion, crop rectangle, write_rgb
conversion) against the oracle and the reference-made fixture.
Bit-exact."""
import numpy as np
import pytest

from tests import golden_util as gu

pytestmark = pytest.mark.gpu


def test_engine_imgio_matches_golden(engine):
    g = gu.load("imgio.npz")
    pix, grey, mos = gu.imgio_inputs()
    assert gu.same_bits(engine.read_img_rgb8(pix), g["read_rgb"])
    assert gu.same_bits(engine.read_img_rgb8(grey), g["read_grey"])
    rect, out = engine.crop_write_rgb8(mos, crop=True)
    assert np.array_equal(rect[2:], g["crop_wh"])
    assert gu.same_bits(out, g["write_cropped"])
    _, full = engine.crop_write_rgb8(mos, crop=False)
    assert gu.same_bits(full, g["write_full"])


@pytest.mark.parametrize("h,w,ch", [(37, 53, 3), (16, 16, 3), (9, 31, 1), (480, 641, 3), (3, 3, 3)])
def test_read_img_sizes(engine, orc, h, w, ch):
    rng = np.random.RandomState(h * 1000 + w)
    pix = rng.randint(0, 256, (h, w, ch) if ch == 3 else (h, w)).astype(np.uint8)
    assert gu.same_bits(engine.read_img_rgb8(pix), orc.read_img_rgb8(pix))


def _mosaic(rng, h, w, holes, border=True):
    m = rng.rand(h, w, 3).astype(np.float32)
    for _ in range(holes):
        y0, x0 = rng.randint(0, h), rng.randint(0, w)
        m[y0:y0 + rng.randint(1, 8), x0:x0 + rng.randint(1, 10)] = -1
    if border:
        yy, xx = np.mgrid[0:h, 0:w]
        m[(yy < 3 + 0.02 * xx) | (yy > h - 4 - 0.01 * xx) | (xx < 2) | (xx > w - 5)] = -1
    return m


@pytest.mark.parametrize("h,w,holes,border", [
    (40, 60, 4, False), (23, 91, 0, True), (64, 1100, 6, True), (97, 2500, 3, True), (33, 32, 2, False),
    (200, 1024, 30, True), (5, 4000, 0, False), (12, 6000, -1, False)])
def test_crop_and_write(engine, orc, h, w, holes, border):
    rng = np.random.RandomState(h + w)
    m = _mosaic(rng, h, w, max(holes, 0), border)
    if holes < 0:      # per-column noise at the top: thousands of height runs per line (per-column search path)
        for k in range(w):
            m[:rng.randint(0, 6), k] = -1
    want_rect, want_px = orc.crop(m)
    rect, out = engine.crop_write_rgb8(m, crop=True)
    assert np.array_equal(rect, want_rect)
    assert gu.same_bits(out, orc.write_rgb8(want_px))


def test_crop_degenerate(engine, orc):
    none = -np.ones((6, 7, 3), np.float32)
    rect, out = engine.crop_write_rgb8(none, crop=True)
    assert np.array_equal(rect, orc.crop(none)[0]) and out.size == 0
    full = np.random.RandomState(3).rand(8, 9, 3).astype(np.float32)
    rect, out = engine.crop_write_rgb8(full, crop=True)
    assert np.array_equal(rect, [0, 0, 9, 8])
    assert gu.same_bits(out, orc.write_rgb8(full))
    # ties: two equal-area rectangles — the first in (line, column) order wins
    m = np.zeros((10, 20, 3), np.float32)
    m[:, 9:11] = -1
    rect, _ = engine.crop_write_rgb8(m, crop=True)
    assert np.array_equal(rect, orc.crop(m)[0])


@pytest.mark.parametrize("crop", [True, False])
def test_pipelined_stitcher_rgb8_boundary(orc, crop):
    """8-bit pixels in, (cropped) 8-bit mosaic out through the pipelined stitcher ==
    read_img -> SIFT -> match -> blend -> crop -> write_rgb chained on the oracle.
    Three jobs in flight to exercise the slot ring."""
    import ctypes as C
    from openpano_b200 import synth
    from openpano_b200._abi import default_params
    from openpano_b200.capi import Engine
    from openpano_b200.stitcher import PipelinedStitcher, ordered_pairs, unpack_rgb8_mosaic
    p = default_params(ordered_input=1)
    pairs = ordered_pairs(3)
    jobs = []
    for seed in (21, 22, 23):
        imgs, org = synth.make_stack(3, 300, 200, 100, seed)
        pix = [(im * 255.0 + 0.5).astype(np.uint8) for im in imgs]
        items, geom = synth.translation_blend_setup(org, 300, 200)
        if seed == 22:                       # leave a hole so that crop has work to do
            items = [items[0], items[2]]
            pix, pairs_j = [pix[0], pix[2]], [(0, 1)]
        else:
            pairs_j = pairs
        jobs.append((pix, items, geom, pairs_j))
    ps = PipelinedStitcher(0, p, depth=2, rgb8=True, crop=crop)
    outs, handles = [], []
    for pix, items, geom, pairs_j in jobs:
        ow, oh = max(it[2] for it in items), max(it[3] for it in items)
        out = np.zeros(ps.out_bytes((ow, oh)), np.uint8)
        k = ps.stage([a.ctypes.data for a in pix], [a.shape[:2] for a in pix], (ow, oh))
        handles.append(ps.run(k, pairs_j, items, geom, out.ctypes.data))
        outs.append((out, (ow, oh)))
    got = [ps.wait(hd) for hd in handles]
    ps.close()
    for (pix, items, geom, pairs_j), matches, (out, wh) in zip(jobs, got, outs):
        f32 = [orc.read_img_rgb8(a) for a in pix]
        descs = [orc.sift_detect(im, p)[1] for im in f32]
        for (i, j), m in zip(pairs_j, matches):
            assert np.array_equal(m, orc.match(descs[i], descs[j], p))
        mosaic = orc.blend(f32, items, geom, 0, p)
        rect, px = unpack_rgb8_mosaic(out, wh, cropped=crop)
        if crop:
            want_rect, want = orc.crop(mosaic)
            assert np.array_equal(rect, want_rect)
            assert gu.same_bits(px, orc.write_rgb8(want))
        else:
            assert gu.same_bits(px, orc.write_rgb8(mosaic))


def test_stitch_lanes_two_concurrent_jobs(orc):
    """Two pipelined lanes (two host threads) sharing the GPU return, for every job, exactly what
    the oracle chain returns — jobs differ, so a cross-talk between lanes would be seen."""
    from openpano_b200 import synth
    from openpano_b200._abi import default_params
    from openpano_b200.stitcher import StitchLanes, ordered_pairs, unpack_rgb8_mosaic
    p = default_params(ordered_input=1)
    jobs, keep, expect = [], [], []
    lanes = StitchLanes(0, p, lanes=2, depth=2, rgb8=True, crop=True)
    for seed in (41, 42, 43, 44, 45):
        n = 2 + seed % 2
        imgs, org = synth.make_stack(n, 280, 190, 95, seed)
        pix = [(im * 255.0 + 0.5).astype(np.uint8) for im in imgs]
        items, geom = synth.translation_blend_setup(org, 280, 190)
        pairs = ordered_pairs(n) if n > 2 else [(0, 1)]
        ow, oh = max(it[2] for it in items), max(it[3] for it in items)
        out = np.zeros(lanes.out_bytes((ow, oh)), np.uint8)
        keep.append((pix, out))
        jobs.append(([a.ctypes.data for a in pix], [a.shape[:2] for a in pix], (ow, oh), pairs, items, geom,
                     out.ctypes.data, 0))
        expect.append((pix, pairs, items, geom, (ow, oh), out))
    got = lanes.map(jobs)
    lanes.close()
    for matches, (pix, pairs, items, geom, wh, out) in zip(got, expect):
        f32 = [orc.read_img_rgb8(a) for a in pix]
        descs = [orc.sift_detect(im, p)[1] for im in f32]
        for (i, j), m in zip(pairs, matches):
            assert np.array_equal(m, orc.match(descs[i], descs[j], p))
        want_rect, want = orc.crop(orc.blend(f32, items, geom, 0, p))
        rect, px = unpack_rgb8_mosaic(out, wh)
        assert np.array_equal(rect, want_rect)
        assert gu.same_bits(px, orc.write_rgb8(want))
