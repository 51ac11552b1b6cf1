"""GPU parity on the BASELINE.json stacks themselves (configs 2-5), whole stacks through
the public call chain (Stitcher / match_pairs / blend over the C ABI), compared bit for
bit with the oracle — plus the product-path branches small inputs never reach (generic
blur windows, > 64 images on one tile, the matcher's full re-scan fallback, keypoint-list
growth, two devices in one process).

The oracle runs with its independent loops on all host threads (oracle/liboracle_mt.so,
bit-identical to the single-thread build: tests/test_oracle_vs_ref.py) because the
BASELINE sizes take minutes on one core."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from openpano_b200 import synth
from openpano_b200._abi import default_params

pytestmark = pytest.mark.gpu

NTHREADS = max(1, min(64, len(os.sched_getaffinity(0))))


@pytest.fixture(scope="module")
def omt():
    from tests.checker import get_checker
    return get_checker("orc_mt")


def bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a.view(np.uint64) if a.dtype == np.float64 else a


def orc_features(orc, imgs, params=None):
    """orc.sift_detect on every image, images in parallel (ctypes releases the GIL; the
    oracle keeps no global state)."""
    with ThreadPoolExecutor(NTHREADS) as ex:
        return list(ex.map(lambda im: orc.sift_detect(im, params), imgs))


def quantised(views):
    """The stack as read_img delivers it: 8-bit pixels -> (float)v / 255.0 (imgio.cc:79-81)."""
    return [((v * 255.0 + 0.5).astype(np.uint8).astype(np.float32).astype(np.float64) / 255.0).astype(np.float32)
            for v in views]


def check_features(fs, want):
    total = 0
    for i, (co, do) in enumerate(want):
        c, d = fs.download(i)
        assert len(d) == len(do), f"image {i}: {len(d)} vs {len(do)} descriptors"
        assert np.array_equal(bits(c), bits(co)), f"image {i}: coordinates differ"
        assert np.array_equal(bits(d), bits(do)), f"image {i}: descriptors differ"
        total += len(d)
    return total


def check_matches(omt, got, pairs, want_feats, params=None):
    n = 0
    for (i, j), m in zip(pairs, got):
        want = omt.match(want_feats[i][1], want_feats[j][1], params)
        assert np.array_equal(m, want), f"pair ({i},{j}): {len(m)} vs {len(want)} matches"
        n += len(m)
    return n


# ------------------------------------------------------------------ config 1: 8 x 600x400, cylinder mode
def test_config1_cmu0_cylinder_hot_path(engine, orc, omt):
    """BASELINE config 1 (CMU0, cylinder mode) at its shape: the hot-path stages CylinderStitcher::build chains
    (cylstitcher.cc:20-87) — SIFT on the 8 views, the 7 adjacent-pair matches (:40-41), CylinderWarper::warp of every
    image and of its keypoints (:65-67, one batched launch over device-resident images), LinearBlender over the
    warped images (flat projection, :24-27) — each bit-identical to the oracle.  The geometry between them
    (update_h_factor, transform estimation, perspective_correction) is host code outside the path: the blend
    uses generator-known translations."""
    views, org = synth.config_stack("cmu0_8x600x400")
    imgs = quantised(views)
    n, (h, w) = len(imgs), imgs[0].shape[:2]
    params = default_params(ordered_input=1)
    pairs = [(k, k + 1) for k in range(n - 1)]
    want = orc_features(orc, imgs, params)
    fs = engine.sift_detect_batch(imgs, params)
    assert check_features(fs, want) > 3000
    matches = engine.match_pairs(fs, pairs, params)
    fs.free()
    assert check_matches(omt, matches, pairs, want, params) > 500
    # warp: images on the device, keypoints (image-centred) in place
    ow, oh, _, _ = engine.cyl_warp_shape(w, h, 1.0, params)
    assert (ow, oh) == orc.cyl_warp_shape(w, h, 1.0, params)[:2]
    src = [engine.dev_alloc(im.nbytes) for im in imgs]
    dst = [engine.dev_alloc(oh * ow * 12) for _ in imgs]
    for d, im in zip(src, imgs):
        engine.dev_upload(d, im)
    kp = [np.ascontiguousarray(c).copy() for c, _ in want]      # Descriptor::coor is image-centred already
    kp_in = [k.copy() for k in kp]
    engine.cyl_warp_batch_dev(src, [(h, w)] * n, dst, kp, 1.0, params)
    warped = []
    for k in range(n):
        out = np.empty((oh, ow, 3), np.float32)
        engine.dev_download(out, dst[k])
        o_img, o_kp = orc.cyl_warp(imgs[k], kp_in[k], 1.0, params)
        assert np.array_equal(bits(out), bits(o_img)), f"warped image {k} differs"
        assert np.array_equal(bits(kp[k]), bits(o_kp)), f"warped keypoints of image {k} differ"
        warped.append(out)
    # composite of the warped images, straight from the device buffers
    items, geom = synth.translation_blend_setup(org, ow, oh)
    tw, th = max(it[2] for it in items), max(it[3] for it in items)
    d_out = engine.dev_alloc(tw * th * 12)
    engine.blend_dev(dst, [(oh, ow)] * n, items, geom, d_out, tw, th, 0, params)
    mosaic = np.empty((th, tw, 3), np.float32)
    engine.dev_download(mosaic, d_out)
    assert np.array_equal(bits(mosaic), bits(omt.blend(warped, items, geom, 0, params))), "mosaic differs"
    for d in src + dst + [d_out]:
        engine.dev_free(d)


# ------------------------------------------------------------------ config 2: 13 ordered 1500x1112
def test_config2_ordered13_whole_stack(engine, orc, omt):
    from openpano_b200.stitcher import Stitcher
    views, org = synth.config_stack("ordered_13x1500x1112")
    imgs = quantised(views)
    params = default_params(ordered_input=1)
    items, geom = synth.translation_blend_setup(org, 1500, 1112)
    pairs = synth.ordered_pairs(13)
    want = orc_features(orc, imgs, params)
    st = Stitcher(engine, params)
    matches, mosaic = st.build_numpy(imgs, pairs, items, geom, bands=0)
    fs = engine.sift_detect_batch(imgs, params)
    n_feat = check_features(fs, want)
    fs.free()
    n_match = check_matches(omt, matches, pairs, want, params)
    assert n_feat > 20000 and n_match > 5000
    ref_lin = omt.blend(imgs, items, geom, 0, params)
    assert np.array_equal(bits(mosaic), bits(ref_lin)), "linear mosaic differs"
    # the same stack through MultiBandBlender{5}
    p5 = default_params(ordered_input=1, multiband=5)
    _, mosaic5 = st.build_numpy(imgs, pairs, items, geom, bands=5)
    st.close()
    ref5 = omt.blend(imgs, items, geom, 5, p5)
    assert np.array_equal(bits(mosaic5), bits(ref5)), "5-band mosaic differs"


# ------------------------------------------------------------------ config 3: 38 unordered, 703 pairs
def test_config3_unordered38_all_pairs(engine, orc, omt):
    views, org = synth.config_stack("unordered_38x1300x867")
    imgs = quantised(views)
    params = default_params()
    pairs = synth.all_pairs(38)
    assert len(pairs) == 703
    want = orc_features(orc, imgs, params)
    fs = engine.sift_detect_batch(imgs, params)
    n_feat = check_features(fs, want)
    got = engine.match_pairs(fs, pairs, params)
    fs.free()
    # 703 oracle matches: pairs in parallel on the host threads (each call single-threaded inside)
    with ThreadPoolExecutor(NTHREADS) as ex:
        wants = list(ex.map(lambda ij: orc.match(want[ij[0]][1], want[ij[1]][1], params), pairs))
    n_match = 0
    for (i, j), m, w in zip(pairs, got, wants):
        assert np.array_equal(m, w), f"pair ({i},{j}): {len(m)} vs {len(w)}"
        n_match += len(m)
    assert n_feat > 50000 and n_match > 20000
    items, geom = synth.translation_blend_setup(org, 1300, 867)
    mosaic = engine.blend(imgs, items, geom, 0, params)
    assert np.array_equal(bits(mosaic), bits(omt.blend(imgs, items, geom, 0, params))), "linear mosaic differs"


# ------------------------------------------------------------------ config 4: brute-force match sweep
def sweep_sets(n, seed=4):
    """RootSIFT-like rows; B = permuted A, half of it with noise (a true mutual match survives),
    half replaced by unrelated rows."""
    rng = np.random.RandomState(seed)
    a = synth.rootsift_like(n, seed)
    b = a[rng.permutation(n)].copy()
    half = n // 2
    b[:half] += rng.randn(half, 128).astype(np.float32) * 10.0
    b[half:] = synth.rootsift_like(n - half, seed + 1)
    return a, b


@pytest.mark.parametrize("n,lazy", [(10000, None), (10000, "1"), (50000, None), (50000, "0")])
def test_config4_sweep_vs_oracle(engine, omt, monkeypatch, n, lazy):
    """lazy=None: the matcher's own choice (both sides through the first pass at 10 k, columns on
    demand at 50 k); "1"/"0" force the other way at each size."""
    a, b = sweep_sets(n)
    if lazy is not None:
        monkeypatch.setenv("PANO_MATCH_LAZY", lazy)
    got = engine.match_bruteforce(a, b)
    monkeypatch.delenv("PANO_MATCH_LAZY", raising=False)
    want = omt.match(a, b)
    assert np.array_equal(got, want), (len(got), len(want))
    assert 0.3 * n < len(got) < 0.7 * n


def test_config4_sweep_100k_tensor_equals_exact_path(engine, monkeypatch):
    """Above the sizes the CPU oracle finishes in seconds the all-fp32 CUDA-core path of the engine
    (PANO_MATCH_PATH=exact, itself pinned to the oracle at 10k/50k below) is the cross-check."""
    a, b = sweep_sets(100000)
    monkeypatch.delenv("PANO_MATCH_PATH", raising=False)
    got = engine.match_bruteforce(a, b)
    monkeypatch.setenv("PANO_MATCH_PATH", "exact")
    want = engine.match_bruteforce(a, b)
    monkeypatch.delenv("PANO_MATCH_PATH", raising=False)
    assert np.array_equal(got, want), (len(got), len(want))
    assert 30000 < len(got) < 70000
    # and the result is orientation-symmetric (matcher.cc:127-128 reverses the pairs)
    rev = engine.match_bruteforce(b, a)
    assert np.array_equal(np.sort(rev[:, ::-1].copy().view("i4,i4").ravel()), np.sort(got.copy().view("i4,i4").ravel()))


def test_config4_exact_path_vs_oracle(engine, omt, monkeypatch):
    a, b = sweep_sets(10000, seed=14)
    monkeypatch.setenv("PANO_MATCH_PATH", "exact")
    got = engine.match_bruteforce(a, b)
    monkeypatch.delenv("PANO_MATCH_PATH", raising=False)
    assert np.array_equal(got, omt.match(a, b))


# ------------------------------------------------------------------ config 5: UAV views, MULTIBAND 5, LAZY_READ 0
def test_config5_uav16_multiband5(engine, orc, omt):
    """16 of the 64 4000x3000 views (two columns of the 8x8 grid), MAX_OUTPUT_SIZE 8000: the canvas
    is scaled exactly as get_final_resolution does; SIFT on every view, 5-band mosaic, all bit-exact."""
    views, org = synth.config_stack("uav_64x4000x3000", n=16)
    imgs = quantised(views)
    del views
    params = default_params(multiband=5, lazy_read=0)
    want = orc_features(orc, imgs, params)
    fs = engine.sift_detect_batch(imgs, params)
    n_feat = check_features(fs, want)
    pairs = [(i, i + 1) for i in range(15)]
    got = engine.match_pairs(fs, pairs, params)
    fs.free()
    check_matches(omt, got, pairs, want, params)
    assert n_feat > 8000
    items, geom = synth.translation_blend_setup(org, 4000, 3000, max_output_size=8000)
    tw, th = max(it[2] for it in items), max(it[3] for it in items)
    assert max(tw, th) == 8000
    mosaic = engine.blend(imgs, items, geom, 5, params)
    ref5 = omt.blend(imgs, items, geom, 5, params)
    assert np.array_equal(bits(mosaic), bits(ref5)), "5-band UAV mosaic differs"
    lin = engine.blend(imgs, items, geom, 0, params)
    assert np.array_equal(bits(lin), bits(omt.blend(imgs, items, geom, 0, params))), "linear UAV mosaic differs"


# ------------------------------------------------------------------ branches of the product path
@pytest.mark.parametrize("factor", [4, 8])
def test_generic_blur_windows(engine, orc, factor):
    """GAUSS_WINDOW_FACTOR other than 6 gives window widths outside {7, 13}: the generic
    k_blur_dog runs instead of the register-blocked TMA kernel."""
    from tests.test_gpu_sift import compare_trace
    p = default_params(gauss_window_factor=factor)
    img = synth.make_canvas(300, 420, 61)
    g, o = engine.sift_trace(img, p), orc.sift_trace(img, p)
    assert compare_trace(g, o) > 20
    g.close(); o.close()


def test_blend_more_than_64_images_on_one_tile(engine, omt):
    """TileList holds 64 entries: 70 images stacked on the same canvas area take the
    fallback loop over all images (config 5 has exactly 64 views)."""
    base, _ = synth.make_stack(1, 96, 80, 0, 71)
    rng = np.random.RandomState(3)
    imgs, items = [], []
    for k in range(70):
        im = np.clip(base[0] + rng.randn(80, 96, 3).astype(np.float32) * 0.02, 0, 1).astype(np.float32)
        imgs.append(np.ascontiguousarray(im))
        dx, dy = k % 7, k // 10
        items.append((dx, dy, dx + 96, dy + 80, [1.0, 0.0, -(dx - 3.0), 0.0, 1.0, -(dy - 3.0), 0.0, 0.0, 1.0]))
    geom = dict(projection=0, res_x=1.0, res_y=1.0, proj_min_x=-51.0, proj_min_y=-43.0)
    for bands, lazy in ((0, 1), (0, 0), (3, 0)):
        p = default_params(lazy_read=lazy, multiband=bands)
        a = engine.blend(imgs, items, geom, bands, p)
        b = omt.blend(imgs, items, geom, bands, p)
        assert (a[..., 0] >= 0).mean() > 0.5
        assert np.array_equal(bits(a), bits(b)), (bands, lazy)


@pytest.mark.parametrize("lazy", ["0", "1"])
def test_match_full_rescan_fallback(engine, orc, monkeypatch, lazy):
    """PANO_MATCH_BLOCK_CAP=1 leaves one gather block for the second tensor pass: every other
    undecided side goes through k_exact_rows (the whole-row fp32 re-scan)."""
    rng = np.random.RandomState(78)
    a = synth.rootsift_like(1500, 8)
    b = a[rng.permutation(1500)][:1300] + rng.randn(1300, 128).astype(np.float32) * 30.0
    monkeypatch.setenv("PANO_MATCH_BLOCK_CAP", "1")
    monkeypatch.setenv("PANO_MATCH_LAZY", lazy)
    got = engine.match_bruteforce(a, b)
    monkeypatch.delenv("PANO_MATCH_BLOCK_CAP", raising=False)
    monkeypatch.delenv("PANO_MATCH_LAZY", raising=False)
    assert np.array_equal(got, orc.match(a, b))


def test_keypoint_lists_grow(orc, monkeypatch):
    """PANO_SIFT_CAP=256 starts the per-image lists far too small: the batch must come back
    complete (re-run with doubled lists at the first count query), not with PANO_ERR_CAPACITY."""
    from openpano_b200.capi import Engine
    monkeypatch.setenv("PANO_SIFT_CAP", "256")
    eng = Engine(0)
    try:
        imgs, _ = synth.make_stack(3, 480, 360, 160, 43)
        fs = eng.sift_detect_batch(imgs)
        counts = [fs.count(i) for i in range(3)]
        assert max(counts) > 256
        for i, im in enumerate(imgs):
            c, d = fs.download(i)
            co, do = orc.sift_detect(im)
            assert np.array_equal(bits(c), bits(co)) and np.array_equal(bits(d), bits(do))
        m = eng.match_pairs(fs, [(0, 1), (1, 2)])
        assert np.array_equal(m[0], orc.match(fs.download(0)[1], fs.download(1)[1]))
        fs.free()
        # the grown capacity is this context's new starting point: the next batch does not retry
        fs2 = eng.sift_detect_batch(imgs[:1])
        assert fs2.count(0) == counts[0]
        fs2.free()
    finally:
        eng.close()


def test_two_devices_in_one_process(orc):
    """Function attributes (dynamic shared memory limits) are per device: a second context on
    another GPU of the same process must match and blend like the first."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from openpano_b200.capi import Engine
    a = synth.rootsift_like(600, 9)
    rng = np.random.RandomState(9)
    b = a[rng.permutation(600)][:500] + rng.randn(500, 128).astype(np.float32) * 20.0
    want = orc.match(a, b)
    engines = [Engine(0), Engine(1)]
    try:
        for eng in engines:
            assert np.array_equal(eng.match_bruteforce(a, b), want)
        imgs, org = synth.make_stack(3, 300, 200, 100, 7)
        items, geom = synth.translation_blend_setup(org, 300, 200)
        outs = [eng.blend(imgs, items, geom, 3) for eng in engines]
        assert np.array_equal(bits(outs[0]), bits(outs[1]))
        fss = [eng.sift_detect_batch(imgs) for eng in engines]
        for i in range(3):
            assert np.array_equal(bits(fss[0].download(i)[1]), bits(fss[1].download(i)[1]))
        for fs in fss:
            fs.free()
    finally:
        for eng in engines:
            eng.close()
