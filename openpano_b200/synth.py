"""Deterministic synthetic image stacks of the BASELINE.json shapes.

The reference ships no example data (it downloads it, run_test.py:42-46) and the
GPU box has no network, so every config is fed by this generator (SURVEY.md
§8d): a wide textured canvas — 5 octaves of bilinear value noise (cell 128→8 px,
amplitude 0.375/2^o per channel) plus about one flat-coloured disc or square
(radius 2…16 px) per 900 px² — from which each view is a crop.  Pure numpy,
seeded `numpy.random.RandomState` (MT19937), so the same arrays come out here and
on the GPU box.
"""
from __future__ import annotations

import numpy as np


def ordered_pairs(n: int):
    """linear_pairwise_match task list: (i, (i+1) % n) (stitcher.cc:116-123)."""
    return [(i, (i + 1) % n) for i in range(n)]


def all_pairs(n: int):
    """pairwise_match task list (stitcher.cc:98-100)."""
    return [(i, j) for i in range(n) for j in range(i + 1, n)]


def _noise_rows(grid: np.ndarray, r0: int, r1: int, w: int, cell: int, out: np.ndarray, scale: float):
    """Rows [r0, r1) of one octave of bilinear value noise, added to `out` as
    (v - 0.5) * scale.  Element arithmetic (float32, left to right):
    g00*(1-fy)*(1-fx) + g01*(1-fy)*fx + g10*fy*(1-fx) + g11*fy*fx."""
    ys = np.arange(r0, r1, dtype=np.float32) / cell
    xs = np.arange(w, dtype=np.float32) / cell
    y0 = ys.astype(np.int64)
    fy = (ys - y0)[:, None, None]
    fx = (xs - xs.astype(np.int64))[None, :, None]
    nx = (w - 1) // cell + 1
    top, bot = grid[y0], grid[y0 + 1]                      # (rows, gw, 3)
    # x0 = i // cell runs in blocks of `cell` equal indices: a repeat, not a gather
    g00 = np.repeat(top[:, :nx], cell, axis=1)[:, :w]
    g01 = np.repeat(top[:, 1:nx + 1], cell, axis=1)[:, :w]
    g10 = np.repeat(bot[:, :nx], cell, axis=1)[:, :w]
    g11 = np.repeat(bot[:, 1:nx + 1], cell, axis=1)[:, :w]
    v = g00 * (1 - fy) * (1 - fx) + g01 * (1 - fy) * fx + g10 * fy * (1 - fx) + g11 * fy * fx
    out[r0:r1] += (v - 0.5) * scale


def _add_value_noise(rng: np.random.RandomState, img: np.ndarray, cell: int, scale: float, pool=None) -> None:
    h, w = img.shape[:2]
    gh, gw = h // cell + 2, w // cell + 2
    grid = rng.rand(gh, gw, 3).astype(np.float32)
    step = 128
    blocks = [(r, min(h, r + step)) for r in range(0, h, step)]
    if pool is None or len(blocks) == 1:
        for r0, r1 in blocks:
            _noise_rows(grid, r0, r1, w, cell, img, scale)
    else:                                                   # numpy releases the GIL inside these array ops
        list(pool.map(lambda b: _noise_rows(grid, b[0], b[1], w, cell, img, scale), blocks))


def _value_noise(rng: np.random.RandomState, h: int, w: int, cell: int) -> np.ndarray:
    """One octave on its own (h×w×3 float32 in [0,1])."""
    img = np.zeros((h, w, 3), np.float32)
    gh, gw = h // cell + 2, w // cell + 2
    grid = rng.rand(gh, gw, 3).astype(np.float32)
    ys = np.arange(h, dtype=np.float32) / cell
    xs = np.arange(w, dtype=np.float32) / cell
    y0 = ys.astype(np.int64)
    x0 = xs.astype(np.int64)
    fy = (ys - y0)[:, None, None]
    fx = (xs - x0)[None, :, None]
    g00 = grid[y0][:, x0]
    g01 = grid[y0][:, x0 + 1]
    g10 = grid[y0 + 1][:, x0]
    g11 = grid[y0 + 1][:, x0 + 1]
    return (g00 * (1 - fy) * (1 - fx) + g01 * (1 - fy) * fx + g10 * fy * (1 - fx) + g11 * fy * fx)


def make_canvas(h: int, w: int, seed: int) -> np.ndarray:
    """H×W×3 float32 in [0,1]."""
    rng = np.random.RandomState(seed)
    img = np.full((h, w, 3), 0.5, np.float32)
    pool = None
    if h * w > (1 << 21):
        import os
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max(1, min(32, len(os.sched_getaffinity(0)))))
    for o, cell in enumerate((128, 64, 32, 16, 8)):
        amp = 0.375 / (2 ** o)
        _add_value_noise(rng, img, cell, 2 * amp, pool)
    if pool is not None:
        pool.shutdown()
    n_shapes = (h * w) // 900
    cy = rng.randint(0, h, n_shapes)
    cx = rng.randint(0, w, n_shapes)
    rad = rng.randint(2, 17, n_shapes)
    col = rng.rand(n_shapes, 3).astype(np.float32)
    square = rng.rand(n_shapes) < 0.5
    for i in range(n_shapes):
        r = int(rad[i])
        y0, y1 = max(0, cy[i] - r), min(h, cy[i] + r + 1)
        x0, x1 = max(0, cx[i] - r), min(w, cx[i] + r + 1)
        if square[i]:
            img[y0:y1, x0:x1] = col[i]
        else:
            yy, xx = np.ogrid[y0:y1, x0:x1]
            m = (yy - cy[i]) ** 2 + (xx - cx[i]) ** 2 <= r * r
            img[y0:y1, x0:x1][m] = col[i]
    np.clip(img, 0.0, 1.0, out=img)
    return img


def make_stack(n: int, w: int, h: int, step_x: int, seed: int, rows: int = 1, step_y: int = 0):
    """n views of w×h cut from one canvas.  rows>1 lays them on a serpentine.

    Returns (list of H×W×3 float32 C-contiguous arrays, list of (x, y) crop
    origins on the canvas)."""
    per_row = (n + rows - 1) // rows
    cw = w + step_x * (per_row - 1)
    ch = h + step_y * (rows - 1)
    canvas = make_canvas(ch, cw, seed)
    imgs, origins = [], []
    for k in range(n):
        r, c = divmod(k, per_row)
        if r % 2 == 1:
            c = per_row - 1 - c
        x, y = c * step_x, r * step_y
        imgs.append(np.ascontiguousarray(canvas[y:y + h, x:x + w]))
        origins.append((x, y))
    return imgs, origins


# BASELINE.json configs → concrete stacks.
CONFIGS = {
    "cmu0_8x600x400": dict(n=8, w=600, h=400, step_x=200, seed=1),
    "ordered_13x1500x1112": dict(n=13, w=1500, h=1112, step_x=500, seed=2),
    "unordered_38x1300x867": dict(n=38, w=1300, h=867, step_x=430, seed=3, rows=2, step_y=290),
    "uav_64x4000x3000": dict(n=64, w=4000, h=3000, step_x=2000, seed=5, rows=8, step_y=1500),
}


def config_stack(name: str, n: int | None = None):
    cfg = dict(CONFIGS[name])
    if n is not None:
        cfg["n"] = n
    return make_stack(**cfg)


def translation_blend_setup(origins, w: int, h: int, max_output_size: int | None = None):
    """Generator-known geometry for the blend stage of a crop stack: image k is
    the canvas translated by its origin, so with the flat projection the inverse
    homography is a pure translation.

    Returns (list of (x0, y0, x1, y1, homo_inv[9]), geom dict) in the form
    ConnectedImages::blend builds (stitcher_image.cc:116-155): ranges are the
    projected corner ranges min-shifted by proj_min, divided by the resolution and
    truncated to int.  With max_output_size the resolution is raised so that the
    longer canvas edge is that many pixels (get_final_resolution,
    stitcher_image.cc:108-111: `resolution *= max_edge / MAX_OUTPUT_SIZE`)."""
    ox = min(o[0] for o in origins)
    oy = min(o[1] for o in origins)
    proj_min = (ox - w / 2.0, oy - h / 2.0)
    proj_max = (max(o[0] for o in origins) + w / 2.0, max(o[1] for o in origins) + h / 2.0)
    res = 1.0
    if max_output_size is not None:
        max_edge = max(proj_max[0] - proj_min[0], proj_max[1] - proj_min[1])
        if max_edge > max_output_size:
            res = float(np.float32(max_edge / max_output_size))      # `float ratio` in the reference
    items = []
    for (x, y) in origins:
        # homo maps image-centred pixel -> canvas-centred coordinate: + (x, y)
        cx, cy = float(x), float(y)
        homo_inv = [1.0, 0.0, -cx, 0.0, 1.0, -cy, 0.0, 0.0, 1.0]
        rmin = (cx - w / 2.0, cy - h / 2.0)
        rmax = (cx + w / 2.0, cy + h / 2.0)
        x0 = int((rmin[0] - proj_min[0]) / res)
        y0 = int((rmin[1] - proj_min[1]) / res)
        x1 = int((rmax[0] - proj_min[0]) / res)
        y1 = int((rmax[1] - proj_min[1]) / res)
        items.append((x0, y0, x1, y1, homo_inv))
    geom = dict(projection=0, res_x=res, res_y=res, proj_min_x=proj_min[0], proj_min_y=proj_min[1])
    return items, geom


def rootsift_like(n: int, seed: int) -> np.ndarray:
    """n×128 float32 RootSIFT-like rows (L2 norm 512): draw U(0,1)^4,
    L1-normalise, sqrt, ×512 (config 4 of BASELINE.json)."""
    rng = np.random.RandomState(seed)
    x = rng.rand(n, 128).astype(np.float32) ** 4
    x /= x.sum(axis=1, keepdims=True)
    return (np.sqrt(x) * 512.0).astype(np.float32)
