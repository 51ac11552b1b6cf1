"""ctypes front-end to the CHECKERS (oracle/liboracle.so = orc_*, the plain-C
restatement; oracle/_ref/libopenpano_ref.so = ref_*, the reference's own
translation units).  Test infrastructure: never imported by the product."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from openpano_b200._abi import (PanoBaPair, PanoBlendGeom, PanoBlendImage, PanoParams, PanoSSPoint,
                                default_params)

ROOT = Path(__file__).resolve().parent.parent
ORACLE_SO = ROOT / "oracle" / "liboracle.so"
ORACLE_MT_SO = ROOT / "oracle" / "liboracle_mt.so"
REF_SO = ROOT / "oracle" / "_ref" / "libopenpano_ref.so"
REF_FAST_SO = ROOT / "oracle" / "_ref" / "libopenpano_ref_fast.so"

SSPOINT_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("real_x", "<f8"), ("real_y", "<f8"),
                          ("pyr_id", "<i4"), ("scale_id", "<i4"), ("dir", "<f4"),
                          ("scale_factor", "<f4")], align=True)
assert SSPOINT_DTYPE.itemsize == C.sizeof(PanoSSPoint)

_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _f(a):
    return a.ctypes.data_as(_fp)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def make_blend_images(imgs, items):
    """imgs: list of HxWx3 float32 arrays; items: (x0,y0,x1,y1,homo_inv[9])."""
    arr = (PanoBlendImage * len(imgs))()
    for k, (im, it) in enumerate(zip(imgs, items)):
        assert im.dtype == np.float32 and im.flags.c_contiguous
        arr[k].rgb_hwc = im.ctypes.data
        arr[k].h, arr[k].w = im.shape[0], im.shape[1]
        arr[k].x0, arr[k].y0, arr[k].x1, arr[k].y1 = it[0], it[1], it[2], it[3]
        for q in range(9):
            arr[k].homo_inv[q] = it[4][q]
    return arr


def make_geom(g):
    return PanoBlendGeom(projection=g["projection"], res_x=g["res_x"], res_y=g["res_y"],
                         proj_min_x=g["proj_min_x"], proj_min_y=g["proj_min_y"])


def blend_target_size(items):
    return max(it[2] for it in items), max(it[3] for it in items)


class SiftTrace:
    """One image's full SIFT chain with every intermediate (both checkers and
    the CUDA engine expose the same getters)."""

    def __init__(self, lib, prefix, handle):
        self._lib, self._p, self._h = lib, prefix, handle

    def _fn(self, name):
        return getattr(self._lib, f"{self._p}_sift_{name}")

    def working_size(self):
        w, h = C.c_int(), C.c_int()
        self._fn("working_size")(self._h, C.byref(w), C.byref(h))
        return w.value, h.value

    def octave_size(self, o):
        w, h = C.c_int(), C.c_int()
        rc = self._fn("octave_size")(self._h, o, C.byref(w), C.byref(h))
        assert rc == 0
        return w.value, h.value

    def plane(self, kind, octave=0, level=0):
        if kind == 0:
            w, h = self.working_size()
            out = np.empty((h, w, 3), np.float32)
        else:
            w, h = self.octave_size(octave)
            out = np.empty((h, w), np.float32)
        rc = self._fn("plane")(self._h, kind, octave, level, _f(out))
        assert rc == 0, (kind, octave, level)
        return out

    def points(self, stage):
        n = self._fn("points")(self._h, stage, 0, None)
        out = np.zeros(n, SSPOINT_DTYPE)
        if n:
            self._fn("points")(self._h, stage, n, out.ctypes.data_as(C.POINTER(PanoSSPoint)))
        return out

    def descriptors(self):
        n = self._fn("descriptors")(self._h, 0, None, None)
        coor = np.zeros((n, 2), np.float64)
        desc = np.zeros((n, 128), np.float32)
        if n:
            self._fn("descriptors")(self._h, n, _d(coor), _f(desc))
        return coor, desc

    def close(self):
        if self._h:
            self._fn("free")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Checker:
    def __init__(self, path: Path, prefix: str):
        if not Path(path).exists():
            raise FileNotFoundError(path)
        self.lib = C.CDLL(str(path), mode=os.RTLD_LOCAL)
        self.p = prefix
        L, P = self.lib, prefix
        fn = lambda n: getattr(L, f"{P}_{n}")
        fn("sift_run").restype = C.c_void_p
        fn("sift_run").argtypes = [_fp, C.c_int, C.c_int, C.POINTER(PanoParams)]
        fn("sift_working_size").argtypes = [C.c_void_p, _ip, _ip]
        fn("sift_working_size").restype = None
        fn("sift_octave_size").argtypes = [C.c_void_p, C.c_int, _ip, _ip]
        fn("sift_plane").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _fp]
        fn("sift_points").argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(PanoSSPoint)]
        fn("sift_descriptors").argtypes = [C.c_void_p, C.c_int, _dp, _fp]
        fn("sift_free").argtypes = [C.c_void_p]
        fn("sift_free").restype = None
        fn("sift_detect").argtypes = [_fp, C.c_int, C.c_int, C.POINTER(PanoParams), C.c_int, _dp, _fp]
        fn("match").argtypes = [_fp, C.c_int, _fp, C.c_int, C.POINTER(PanoParams), _ip, _ip]
        fn("cyl_warp_shape").argtypes = [C.c_int, C.c_int, C.c_double, C.POINTER(PanoParams),
                                         _ip, _ip, _dp, _dp]
        fn("cyl_warp").argtypes = [_fp, C.c_int, C.c_int, C.c_double, C.POINTER(PanoParams), _fp,
                                   C.c_int, C.c_int, _dp, C.c_int]
        fn("blend").argtypes = [C.c_int, C.POINTER(PanoBlendImage), C.POINTER(PanoBlendGeom),
                                C.c_int, C.POINTER(PanoParams), _fp, C.c_int, C.c_int]
        fn("hotpath").argtypes = [C.c_int, C.POINTER(C.c_void_p), _ip, _ip, C.c_int, _ip, C.c_int,
                                  C.POINTER(PanoBlendImage), C.POINTER(PanoBlendGeom), C.c_int,
                                  C.POINTER(PanoParams), _fp, C.c_int, C.c_int, _ip, _ip, _dp]
        _ubp = C.POINTER(C.c_ubyte)
        fn("ransac_score").argtypes = [C.c_int, _dp, _dp, C.c_int, _dp, C.c_float, _ip, _ip, _ip, _ubp]
        fn("num_threads").restype = C.c_int
        _up = C.POINTER(C.c_ubyte)
        fn("read_img_rgb8").argtypes = [_up, C.c_int, C.c_int, C.c_int, _fp]
        fn("crop").argtypes = [_fp, C.c_int, C.c_int, _ip, _fp]
        fn("write_rgb8").argtypes = [_fp, C.c_int, C.c_int, _up]
        self._fn = fn

    # ---- 8-bit boundary (read_img / crop / write_rgb)
    def read_img_rgb8(self, pix):
        """pix: H×W×3 or H×W uint8 -> H×W×3 float32 as read_img produces it."""
        pix = np.ascontiguousarray(pix, np.uint8)
        h, w = pix.shape[:2]
        ch = 1 if pix.ndim == 2 else pix.shape[2]
        out = np.empty((h, w, 3), np.float32)
        rc = self._fn("read_img_rgb8")(pix.ctypes.data_as(C.POINTER(C.c_ubyte)), w, h, ch, _f(out))
        assert rc == 0
        return out

    def crop(self, mat):
        """Returns (rect [x0,y0,w,h] — x0,y0 are -1 from the ref_ build — and the cropped pixels)."""
        mat = np.ascontiguousarray(mat, np.float32)
        h, w = mat.shape[:2]
        rect = np.zeros(4, np.int32)
        out = np.empty((h, w, 3), np.float32)
        rc = self._fn("crop")(_f(mat), w, h, _i(rect), _f(out))
        assert rc == 0
        cw, ch = int(rect[2]), int(rect[3])
        return rect, out.reshape(-1)[:cw * ch * 3].reshape(ch, cw, 3).copy()

    def write_rgb8(self, mat):
        mat = np.ascontiguousarray(mat, np.float32)
        h, w = mat.shape[:2]
        out = np.empty((h, w, 3), np.uint8)
        rc = self._fn("write_rgb8")(_f(mat), w, h, out.ctypes.data_as(C.POINTER(C.c_ubyte)))
        assert rc == 0
        return out

    def ransac_score(self, kp1, kp2, homos, thres):
        """Returns (best_hyp, best_count, hyp_counts, inlier_flags) as TransformEstimation would pick them."""
        kp1 = np.ascontiguousarray(kp1, np.float64).reshape(-1, 2)
        kp2 = np.ascontiguousarray(kp2, np.float64).reshape(-1, 2)
        homos = np.ascontiguousarray(homos, np.float64).reshape(-1, 9)
        counts = np.zeros(max(len(homos), 1), np.int32)
        flags = np.zeros(max(len(kp1), 1), np.uint8)
        best, bcnt = C.c_int(), C.c_int()
        rc = self._fn("ransac_score")(len(kp1), _d(kp1), _d(kp2), len(homos), _d(homos), thres, _i(counts),
                                      C.byref(best), C.byref(bcnt), flags.ctypes.data_as(C.POINTER(C.c_ubyte)))
        assert rc == 0, rc
        return best.value, bcnt.value, counts[:len(homos)], flags[:len(kp1)]

    def num_threads(self):
        return self._fn("num_threads")()

    # ---- bundle-adjustment Jacobian (incremental_bundle_adjuster.cc:276-385)
    @staticmethod
    def _ba_pairs(pairs, mats=None):
        arr = (PanoBaPair * max(len(pairs), 1))()
        begin = 0
        for k, (f, t, nm) in enumerate(pairs):
            arr[k].from_, arr[k].to, arr[k].match_begin, arr[k].n_match = f, t, begin, nm
            if mats is not None:
                C.memmove(arr[k].m, np.ascontiguousarray(mats[k], np.float64).ctypes.data, 13 * 9 * 8)
            begin += nm
        return arr, begin

    def ba_pair_mats(self, cams, pairs):
        """ref_ only: the 13 per-pair matrices, evaluated with the reference's own Homography / Camera
        operations.  cams [n, 12] = focal, ppx, ppy, R; pairs = [(from, to, n_match)].  -> [n_pair, 13, 9]."""
        cams = np.ascontiguousarray(cams, np.float64).reshape(-1, 12)
        arr, _ = self._ba_pairs(pairs)
        fn = self.lib.ref_ba_pair_mats
        fn.argtypes = [C.c_int, _dp, C.c_int, C.POINTER(PanoBaPair)]
        assert fn(len(cams), _d(cams), len(pairs), arr) == 0
        return np.array([np.ctypeslib.as_array(arr[k].m).reshape(13, 9).copy() for k in range(len(pairs))])

    def ba_jacobian_ref(self, cams, pairs, pts):
        """ref_ only: the reference's own calcJacobianSymbolic.  pts [n_match, 4] = to.x, to.y, from.x, from.y.
        -> (j_rows [n_match, 24], jtj [6n, 6n])."""
        cams = np.ascontiguousarray(cams, np.float64).reshape(-1, 12)
        pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 4)
        arr, total = self._ba_pairs(pairs)
        assert total == len(pts)
        rows = np.zeros((max(total, 1), 24), np.float64)
        jtj = np.full((6 * len(cams), 6 * len(cams)), np.nan, np.float64)
        fn = self.lib.ref_ba_jacobian
        fn.argtypes = [C.c_int, _dp, C.c_int, C.POINTER(PanoBaPair), _dp, _dp, _dp]
        assert fn(len(cams), _d(cams), len(pairs), arr, _d(pts), _d(rows), _d(jtj)) == 0
        return rows[:total], jtj

    def ba_jacobian(self, n_cam, pairs, mats, pts_to):
        """orc_ only: the restatement, from the per-pair matrices.  -> (j_rows, jtj)."""
        pts_to = np.ascontiguousarray(pts_to, np.float64).reshape(-1, 2)
        arr, total = self._ba_pairs(pairs, mats)
        assert total == len(pts_to)
        rows = np.zeros((max(total, 1), 24), np.float64)
        jtj = np.full((6 * n_cam, 6 * n_cam), np.nan, np.float64)
        fn = self.lib.orc_ba_jacobian
        fn.argtypes = [C.c_int, C.c_int, C.POINTER(PanoBaPair), _dp, _dp, _dp]
        assert fn(n_cam, len(pairs), arr, _d(pts_to), _d(rows), _d(jtj)) == 0
        return rows[:total], jtj

    def sift_trace(self, img, params=None) -> SiftTrace:
        params = params or default_params()
        img = np.ascontiguousarray(img, np.float32)
        h = self._fn("sift_run")(_f(img), img.shape[1], img.shape[0], C.byref(params))
        assert h
        return SiftTrace(self.lib, self.p, h)

    def sift_detect(self, img, params=None, cap=65536):
        params = params or default_params()
        img = np.ascontiguousarray(img, np.float32)
        coor = np.zeros((cap, 2), np.float64)
        desc = np.zeros((cap, 128), np.float32)
        n = self._fn("sift_detect")(_f(img), img.shape[1], img.shape[0], C.byref(params), cap,
                                    _d(coor), _f(desc))
        assert n >= 0
        return coor[:n].copy(), desc[:n].copy()

    def match(self, a, b, params=None):
        params = params or default_params()
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        pairs = np.zeros((max(1, min(len(a), len(b))), 2), np.int32)
        n = C.c_int()
        rc = self._fn("match")(_f(a), len(a), _f(b), len(b), C.byref(params), _i(pairs), C.byref(n))
        assert rc == 0
        return pairs[:n.value].copy()

    def cyl_warp_shape(self, w, h, h_factor=1.0, params=None):
        params = params or default_params()
        ow, oh, ox, oy = C.c_int(), C.c_int(), C.c_double(), C.c_double()
        rc = self._fn("cyl_warp_shape")(w, h, h_factor, C.byref(params), C.byref(ow), C.byref(oh),
                                        C.byref(ox), C.byref(oy))
        assert rc == 0
        return ow.value, oh.value, ox.value, oy.value

    def cyl_warp(self, img, kpts=None, h_factor=1.0, params=None):
        params = params or default_params()
        img = np.ascontiguousarray(img, np.float32)
        ow, oh, _, _ = self.cyl_warp_shape(img.shape[1], img.shape[0], h_factor, params)
        out = np.empty((oh, ow, 3), np.float32)
        k = np.ascontiguousarray(kpts if kpts is not None else np.zeros((0, 2)), np.float64).copy()
        rc = self._fn("cyl_warp")(_f(img), img.shape[1], img.shape[0], h_factor, C.byref(params),
                                  _f(out), ow, oh, _d(k), len(k))
        assert rc == 0
        return out, k

    def hotpath(self, imgs, pairs, items, geom, bands=0, params=None, use_flann=True):
        """One CPU pass of SIFT + match + blend. Returns (n_feat, n_match, mosaic, seconds[3])."""
        params = params or default_params()
        imgs = [np.ascontiguousarray(im, np.float32) for im in imgs]
        n = len(imgs)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        ws = (C.c_int * n)(*[im.shape[1] for im in imgs])
        hs = (C.c_int * n)(*[im.shape[0] for im in imgs])
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        arr = make_blend_images(imgs, items)
        g = make_geom(geom)
        ow, oh = blend_target_size(items)
        out = np.empty((oh, ow, 3), np.float32)
        n_feat = np.zeros(n, np.int32)
        n_match = np.zeros(max(len(pairs), 1), np.int32)
        secs = np.zeros(3, np.float64)
        rc = self._fn("hotpath")(n, ptrs, ws, hs, len(pairs), _i(pairs), 1 if use_flann else 0, arr,
                                 C.byref(g), bands, C.byref(params), _f(out), ow, oh, _i(n_feat),
                                 _i(n_match), _d(secs))
        assert rc == 0, rc
        return n_feat, n_match[:len(pairs)], out, secs

    def blend(self, imgs, items, geom, bands=0, params=None):
        params = params or default_params()
        arr = make_blend_images(imgs, items)
        g = make_geom(geom)
        ow, oh = blend_target_size(items)
        out = np.empty((oh, ow, 3), np.float32)
        rc = self._fn("blend")(len(imgs), arr, C.byref(g), bands, C.byref(params), _f(out), ow, oh)
        assert rc == 0
        return out


_cache = {}


def get_checker(kind: str) -> Checker:
    """kind: 'orc' (C restatement), 'orc_mt' (the same with independent loops under
    OpenMP: bit-identical, for BASELINE-size inputs), 'ref' (reference TUs, parity flags) or
    'ref_fast' (reference TUs, perf flags + OpenMP)."""
    if kind not in _cache:
        if kind == "orc":
            _cache[kind] = Checker(ORACLE_SO, "orc")
        elif kind == "orc_mt":
            _cache[kind] = Checker(ORACLE_MT_SO, "orc")
        elif kind == "ref":
            _cache[kind] = Checker(REF_SO, "ref")
        elif kind == "ref_fast":
            _cache[kind] = Checker(REF_FAST_SO, "ref")
        else:
            raise KeyError(kind)
    return _cache[kind]


def have(kind: str) -> bool:
    return {"orc": ORACLE_SO, "orc_mt": ORACLE_MT_SO, "ref": REF_SO, "ref_fast": REF_FAST_SO}[kind].exists()
