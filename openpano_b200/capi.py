"""ctypes binding of libpano_b200.so (include/pano_b200.h).

There is no CPU fallback: if the CUDA library is missing this module raises at
import, and if no B200 is visible `Engine()` raises PanoError (PANO_ERR_NO_DEVICE).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from ._abi import (PanoBaPair, PanoBlendGeom, PanoCylJob, PanoBlendImage, PanoMatches, PanoParams, PanoRansacPair, PanoSSPoint,
                   default_params)

LIB_PATH = Path(__file__).resolve().parent / "libpano_b200.so"

_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vpp = C.POINTER(C.c_void_p)

SSPOINT_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("real_x", "<f8"), ("real_y", "<f8"),
                          ("pyr_id", "<i4"), ("scale_id", "<i4"), ("dir", "<f4"),
                          ("scale_factor", "<f4")], align=True)


class PanoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pano_b200 error {code}: {msg}")
        self.code = code


def _load():
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). openpano_b200 has no CPU fallback.")
    # PANO_B200_LIB: development override to A/B a differently tuned build of the same library
    lib = C.CDLL(os.environ.get("PANO_B200_LIB", str(LIB_PATH)), mode=os.RTLD_LOCAL)
    P = C.POINTER(PanoParams)
    sig = {
        "pano_params_default": (None, [P]),
        "pano_create": (C.c_int, [_vpp, C.c_int, C.c_void_p]),
        "pano_destroy": (None, [C.c_void_p]),
        "pano_last_error": (C.c_char_p, [C.c_void_p]),
        "pano_sync": (C.c_int, [C.c_void_p]),
        "pano_stream": (C.c_void_p, [C.c_void_p]),
        "pano_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
        "pano_profile_reset": (C.c_int, [C.c_void_p]),
        "pano_profile_read": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, _ip, _dp]),
        "pano_launch_count": (C.c_longlong, [C.c_void_p]),
        "pano_trim": (C.c_int, [C.c_void_p]),
        "pano_match_last_exact_rows": (C.c_int, [C.c_void_p]),
        "pano_match_last_nominated_rows": (C.c_int, [C.c_void_p]),
        "pano_sift_detect_batch": (C.c_int, [C.c_void_p, C.c_int, _vpp, _ip, _ip, P, _vpp]),
        "pano_sift_detect_batch_dev": (C.c_int, [C.c_void_p, C.c_int, _vpp, _ip, _ip, P, _vpp]),
        "pano_sift_detect": (C.c_int, [C.c_void_p, _fp, C.c_int, C.c_int, P, _vpp]),
        "pano_featureset_upload": (C.c_int, [C.c_void_p, C.c_int, _ip, _vpp, _vpp, _vpp]),
        "pano_featureset_num_images": (C.c_int, [C.c_void_p]),
        "pano_featureset_count": (C.c_int, [C.c_void_p, C.c_int]),
        "pano_featureset_download": (C.c_int, [C.c_void_p, C.c_int, _dp, _fp]),
        "pano_featureset_download_real": (C.c_int, [C.c_void_p, C.c_int, _dp]),
        "pano_featureset_free": (None, [C.c_void_p]),
        "pano_match_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, _ip, P, C.POINTER(PanoMatches)]),
        "pano_matches_free": (None, [C.POINTER(PanoMatches)]),
        "pano_match_pairs_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, _ip, P, _ip]),
        "pano_match_pairs_shard": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, _ip, P, C.c_int, C.c_int, C.POINTER(PanoMatches)]),
        "pano_match_pairs_dev_shard": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, _ip, P, C.c_int, C.c_int, _ip]),
        "pano_match_bruteforce": (C.c_int, [C.c_void_p, _fp, C.c_int, _fp, C.c_int, P, _ip, _ip]),
        "pano_comm_unique_id": (C.c_int, [C.c_char_p]),
        "pano_comm_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p, _vpp]),
        "pano_comm_adopt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, _vpp]),
        "pano_comm_destroy": (None, [C.c_void_p]),
        "pano_comm_world": (C.c_int, [C.c_void_p]),
        "pano_comm_rank": (C.c_int, [C.c_void_p]),
        "pano_comm_allgather_features": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, _vpp]),
        "pano_comm_allgather_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
        "pano_ransac_score_pairs": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(PanoRansacPair), _ip, _ip, _vpp, _vpp]),
        "pano_ba_jacobian": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(PanoBaPair), _dp, _dp, _dp]),
        "pano_cyl_warp_shape": (C.c_int, [C.c_int, C.c_int, C.c_double, P, _ip, _ip, _dp, _dp]),
        "pano_cyl_warp": (C.c_int, [C.c_void_p, _fp, C.c_int, C.c_int, C.c_double, P, _fp, C.c_int,
                                    C.c_int, _dp, C.c_int]),
        "pano_cyl_warp_batch_dev": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(PanoCylJob), C.c_double, P]),
        "pano_blend_target_size": (C.c_int, [C.c_int, C.POINTER(PanoBlendImage), _ip, _ip]),
        "pano_blend": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(PanoBlendImage), C.POINTER(PanoBlendGeom),
                                 C.c_int, P, _fp, C.c_int, C.c_int]),
        "pano_blend_dev": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(PanoBlendImage),
                                     C.POINTER(PanoBlendGeom), C.c_int, P, C.c_void_p, C.c_int, C.c_int]),
        "pano_blend_rows_dev": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(PanoBlendImage), C.POINTER(PanoBlendGeom),
                                          C.c_int, P, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
        "pano_featureset_import_dev": (C.c_int, [C.c_void_p, C.c_int, _ip, _vpp, _vpp, _vpp]),
        "pano_featureset_export_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
        "pano_featureset_export_all_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
        "pano_rgb8_to_mat32f_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
        "pano_rgb8_to_mat32f_batch_dev": (C.c_int, [C.c_void_p, C.c_int, _vpp, _ip, _ip, _ip, _vpp]),
        "pano_crop_rect_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
        "pano_mat32f_to_rgb8_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
        "pano_dev_alloc": (C.c_int, [C.c_void_p, C.c_size_t, _vpp]),
        "pano_dev_free": (C.c_int, [C.c_void_p, C.c_void_p]),
        "pano_dev_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
        "pano_dev_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
        "pano_dev_upload_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
        "pano_dev_download_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
        "pano_host_alloc": (C.c_int, [C.c_size_t, _vpp]),
        "pano_host_free": (C.c_int, [C.c_void_p]),
        "pano_event_create": (C.c_int, [C.c_void_p, _vpp]),
        "pano_event_record": (C.c_int, [C.c_void_p, C.c_void_p]),
        "pano_event_wait": (C.c_int, [C.c_void_p, C.c_void_p]),
        "pano_event_sync": (C.c_int, [C.c_void_p]),
        "pano_event_destroy": (None, [C.c_void_p]),
        "pano_sift_trace_run": (C.c_int, [C.c_void_p, _fp, C.c_int, C.c_int, P, _vpp]),
        "pano_sift_trace_working_size": (C.c_int, [C.c_void_p, _ip, _ip]),
        "pano_sift_trace_octave_size": (C.c_int, [C.c_void_p, C.c_int, _ip, _ip]),
        "pano_sift_trace_plane": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _fp]),
        "pano_sift_trace_points": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(PanoSSPoint)]),
        "pano_sift_trace_descriptors": (C.c_int, [C.c_void_p, C.c_int, _dp, _fp]),
        "pano_sift_trace_free": (None, [C.c_void_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sig)


LIB, EXPORTED = _load()


def _f(a):
    return a.ctypes.data_as(_fp)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class FeatureSet:
    """Device-resident descriptors + coordinates of a batch of images."""

    def __init__(self, eng, handle):
        self.eng, self._h = eng, handle

    @property
    def n_images(self):
        return LIB.pano_featureset_num_images(self._h)

    def count(self, i):
        n = LIB.pano_featureset_count(self._h, i)
        if n < 0:
            self.eng._raise(n)
        return n

    def download(self, i):
        n = self.count(i)
        coor = np.zeros((n, 2), np.float64)
        desc = np.zeros((n, 128), np.float32)
        self.eng._check(LIB.pano_featureset_download(self._h, i, _d(coor), _f(desc)))
        return coor, desc

    def export_all_dev(self, d_coor, d_desc):
        """Every image's rows packed back to back into device buffers (one launch)."""
        self.eng._check(LIB.pano_featureset_export_all_dev(self._h, C.c_void_p(d_coor or 0), C.c_void_p(d_desc or 0)))

    def export_dev(self, i, d_coor, d_desc):
        """Device-to-device copy of image i's rows (coordinates n×2 f64, descriptors n×128 f32)."""
        self.eng._check(LIB.pano_featureset_export_dev(self._h, i, C.c_void_p(d_coor or 0), C.c_void_p(d_desc or 0)))

    def free(self):
        if self._h:
            LIB.pano_featureset_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class GpuSiftTrace:
    def __init__(self, eng, handle):
        self.eng, self._h = eng, handle

    def working_size(self):
        w, h = C.c_int(), C.c_int()
        LIB.pano_sift_trace_working_size(self._h, C.byref(w), C.byref(h))
        return w.value, h.value

    def octave_size(self, o):
        w, h = C.c_int(), C.c_int()
        self.eng._check(LIB.pano_sift_trace_octave_size(self._h, o, C.byref(w), C.byref(h)))
        return w.value, h.value

    def plane(self, kind, octave=0, level=0):
        if kind == 0:
            w, h = self.working_size()
            out = np.empty((h, w, 3), np.float32)
        else:
            w, h = self.octave_size(octave)
            out = np.empty((h, w), np.float32)
        self.eng._check(LIB.pano_sift_trace_plane(self._h, kind, octave, level, _f(out)))
        return out

    def points(self, stage):
        n = LIB.pano_sift_trace_points(self._h, stage, 0, None)
        if n < 0:
            self.eng._raise(n)
        out = np.zeros(n, SSPOINT_DTYPE)
        if n:
            LIB.pano_sift_trace_points(self._h, stage, n, out.ctypes.data_as(C.POINTER(PanoSSPoint)))
        return out

    def descriptors(self):
        n = LIB.pano_sift_trace_descriptors(self._h, 0, None, None)
        if n < 0:
            self.eng._raise(n)
        coor = np.zeros((n, 2), np.float64)
        desc = np.zeros((n, 128), np.float32)
        if n:
            LIB.pano_sift_trace_descriptors(self._h, n, _d(coor), _f(desc))
        return coor, desc

    def close(self):
        if self._h:
            LIB.pano_sift_trace_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One pano_ctx: a CUDA device + stream.  `stream` is a raw cudaStream_t
    (e.g. torch.cuda.current_stream().cuda_stream) or None."""

    def __init__(self, device: int = 0, stream: int | None = None):
        h = C.c_void_p()
        rc = LIB.pano_create(C.byref(h), device, C.c_void_p(stream) if stream else None)
        if rc != 0:
            raise PanoError(rc, LIB.pano_last_error(None).decode())
        self._h = h
        self.device = device

    # -- plumbing
    def _raise(self, rc):
        raise PanoError(rc, LIB.pano_last_error(self._h).decode())

    def _check(self, rc):
        if rc != 0:
            self._raise(rc)

    def close(self):
        if self._h:
            LIB.pano_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(LIB.pano_sync(self._h))

    @property
    def stream(self):
        return LIB.pano_stream(self._h)

    def trim(self):
        """Hand the freed device blocks this context keeps for reuse back to its pool."""
        self._check(LIB.pano_trim(self._h))

    def launch_count(self):
        return LIB.pano_launch_count(self._h)

    def match_last_exact_rows(self):
        return LIB.pano_match_last_exact_rows(self._h)

    def match_last_nominated_rows(self):
        return LIB.pano_match_last_nominated_rows(self._h)

    def profile(self, on: bool):
        self._check(LIB.pano_profile_enable(self._h, 1 if on else 0))

    def profile_reset(self):
        self._check(LIB.pano_profile_reset(self._h))

    def profile_read(self):
        cap = 64
        names = C.create_string_buffer(cap * 64)
        launches = (C.c_int * cap)()
        ms = (C.c_double * cap)()
        n = LIB.pano_profile_read(self._h, cap, names, launches, ms)
        out = {}
        for i in range(min(n, cap)):
            nm = names.raw[i * 64:(i + 1) * 64].split(b"\0", 1)[0].decode()
            out[nm] = (launches[i], ms[i])
        return out

    # -- device memory helpers (bench / tests)
    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(LIB.pano_dev_alloc(self._h, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, ptr):
        LIB.pano_dev_free(self._h, C.c_void_p(ptr))

    def dev_upload(self, ptr, arr):
        arr = np.ascontiguousarray(arr)
        self._check(LIB.pano_dev_upload(self._h, C.c_void_p(ptr), arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def dev_download(self, arr, ptr):
        assert arr.flags.c_contiguous
        self._check(LIB.pano_dev_download(self._h, arr.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), arr.nbytes))

    def dev_upload_async(self, d_ptr, h_ptr, nbytes):
        self._check(LIB.pano_dev_upload_async(self._h, C.c_void_p(d_ptr), C.c_void_p(h_ptr), nbytes))

    def dev_download_async(self, h_ptr, d_ptr, nbytes):
        self._check(LIB.pano_dev_download_async(self._h, C.c_void_p(h_ptr), C.c_void_p(d_ptr), nbytes))

    @staticmethod
    def host_alloc(nbytes):
        p = C.c_void_p()
        if LIB.pano_host_alloc(nbytes, C.byref(p)) != 0:
            raise PanoError(-1, "pano_host_alloc failed")
        return p.value

    @staticmethod
    def host_free(ptr):
        LIB.pano_host_free(C.c_void_p(ptr))

    # -- events (cross-context ordering)
    def event_create(self):
        e = C.c_void_p()
        self._check(LIB.pano_event_create(self._h, C.byref(e)))
        return e

    def event_record(self, ev):
        self._check(LIB.pano_event_record(self._h, ev))

    def event_wait(self, ev):
        self._check(LIB.pano_event_wait(self._h, ev))

    @staticmethod
    def event_sync(ev):
        if LIB.pano_event_sync(ev) != 0:
            raise PanoError(-1, "pano_event_sync failed")

    @staticmethod
    def event_destroy(ev):
        LIB.pano_event_destroy(ev)

    # -- features
    def sift_detect_batch(self, imgs, params=None) -> FeatureSet:
        params = params or default_params()
        imgs = [np.ascontiguousarray(im, np.float32) for im in imgs]
        n = len(imgs)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        ws = (C.c_int * n)(*[im.shape[1] for im in imgs])
        hs = (C.c_int * n)(*[im.shape[0] for im in imgs])
        out = C.c_void_p()
        self._check(LIB.pano_sift_detect_batch(self._h, n, ptrs, ws, hs, C.byref(params), C.byref(out)))
        return FeatureSet(self, out)

    def sift_detect_batch_ptr(self, ptrs, ws, hs, params=None, device=False) -> FeatureSet:
        """Raw-pointer variant (pinned host or device pointers), no numpy copies."""
        params = params or default_params()
        n = len(ptrs)
        cp = (C.c_void_p * n)(*ptrs)
        cw = (C.c_int * n)(*ws)
        ch = (C.c_int * n)(*hs)
        out = C.c_void_p()
        fn = LIB.pano_sift_detect_batch_dev if device else LIB.pano_sift_detect_batch
        self._check(fn(self._h, n, cp, cw, ch, C.byref(params), C.byref(out)))
        return FeatureSet(self, out)

    def sift_detect(self, img, params=None):
        fs = self.sift_detect_batch([img], params)
        try:
            return fs.download(0)
        finally:
            fs.free()

    def sift_trace(self, img, params=None) -> GpuSiftTrace:
        params = params or default_params()
        img = np.ascontiguousarray(img, np.float32)
        out = C.c_void_p()
        self._check(LIB.pano_sift_trace_run(self._h, _f(img), img.shape[1], img.shape[0], C.byref(params),
                                            C.byref(out)))
        return GpuSiftTrace(self, out)

    def featureset_upload(self, descs, coors=None) -> FeatureSet:
        descs = [np.ascontiguousarray(d, np.float32).reshape(-1, 128) for d in descs]
        n = len(descs)
        cnt = (C.c_int * n)(*[len(d) for d in descs])
        dp = (C.c_void_p * n)(*[d.ctypes.data for d in descs])
        cp = None
        if coors is not None:
            coors = [np.ascontiguousarray(c, np.float64).reshape(-1, 2) for c in coors]
            cp = (C.c_void_p * n)(*[c.ctypes.data for c in coors])
        out = C.c_void_p()
        self._check(LIB.pano_featureset_upload(self._h, n, cnt, dp, cp, C.byref(out)))
        return FeatureSet(self, out)

    def featureset_import_dev(self, counts, d_descs, d_coors=None) -> FeatureSet:
        """Featureset from device pointers (stream-ordered, no host sync)."""
        n = len(counts)
        cnt = (C.c_int * n)(*counts)
        dp = (C.c_void_p * n)(*d_descs)
        cp = (C.c_void_p * n)(*d_coors) if d_coors is not None else None
        out = C.c_void_p()
        self._check(LIB.pano_featureset_import_dev(self._h, n, cnt, dp, cp, C.byref(out)))
        return FeatureSet(self, out)

    def blend_rows_dev(self, ptrs, shapes, items, geom, d_out_rows, out_w, out_h, row0, row1, bands=0, params=None):
        """Rows [row0, row1) of the mosaic into a (row1-row0)×out_w×3 device buffer."""
        params = params or default_params()
        arr, g = self._blend_args(ptrs, shapes, items, geom)
        self._check(LIB.pano_blend_rows_dev(self._h, len(ptrs), arr, C.byref(g), bands, C.byref(params),
                                            C.c_void_p(d_out_rows), out_w, out_h, row0, row1))

    # -- matching
    def match_pairs(self, fs: FeatureSet, pairs, params=None, shard=(0, 1)):
        """shard=(s, S): decide only share s of S of every pair's smaller set (row-sharded multi-GPU
        match); the shares' lists concatenated in shard order are the unsharded lists."""
        params = params or default_params()
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        m = PanoMatches()
        self._check(LIB.pano_match_pairs_shard(self._h, fs._h, len(pairs), _i(pairs), C.byref(params), int(shard[0]),
                                               int(shard[1]), C.byref(m)))
        n = m.n_pairs
        offs = np.ctypeslib.as_array(m.offset, shape=(n + 1,)).copy() if n else np.zeros(1, np.int32)
        total = int(offs[n]) if n else 0
        idx = (np.ctypeslib.as_array(m.idx, shape=(2 * total,)).reshape(-1, 2).copy() if total
               else np.zeros((0, 2), np.int32))
        LIB.pano_matches_free(C.byref(m))
        return [idx[offs[k]:offs[k + 1]] for k in range(n)]

    def match_pairs_dev(self, fs: FeatureSet, pairs, params=None, shard=(0, 1)) -> int:
        params = params or default_params()
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        tot = C.c_int()
        self._check(LIB.pano_match_pairs_dev_shard(self._h, fs._h, len(pairs), _i(pairs), C.byref(params), int(shard[0]),
                                                   int(shard[1]), C.byref(tot)))
        return tot.value

    def match_bruteforce(self, a, b, params=None):
        params = params or default_params()
        a = np.ascontiguousarray(a, np.float32).reshape(-1, 128)
        b = np.ascontiguousarray(b, np.float32).reshape(-1, 128)
        pairs = np.zeros((max(1, min(len(a), len(b))), 2), np.int32)
        n = C.c_int()
        self._check(LIB.pano_match_bruteforce(self._h, _f(a), len(a), _f(b), len(b), C.byref(params),
                                              _i(pairs), C.byref(n)))
        return pairs[:n.value].copy()

    # -- RANSAC inlier scoring
    def ransac_score_pairs(self, pairs):
        """pairs: list of (kp1_xy [n,2] f64, kp2_xy [n,2] f64, homos [m,9] f64, inlier_thres).
        Returns per pair (best_hyp, best_count, hyp_counts int32[m], inlier_flags uint8[n])."""
        n = len(pairs)
        arr = (PanoRansacPair * max(n, 1))()
        keep, counts, flags = [], [], []
        for k, (a, b, h, thr) in enumerate(pairs):
            a = np.ascontiguousarray(a, np.float64).reshape(-1, 2)
            b = np.ascontiguousarray(b, np.float64).reshape(-1, 2)
            h = np.ascontiguousarray(h, np.float64).reshape(-1, 9)
            keep.append((a, b, h))
            arr[k].n_match, arr[k].kp1_xy, arr[k].kp2_xy = len(a), a.ctypes.data, b.ctypes.data
            arr[k].n_hyp, arr[k].homos, arr[k].inlier_thres = len(h), h.ctypes.data, thr
            counts.append(np.zeros(max(len(h), 1), np.int32))
            flags.append(np.zeros(max(len(a), 1), np.uint8))
        best, bcnt = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
        cp = (C.c_void_p * max(n, 1))(*[c.ctypes.data for c in counts])
        fp = (C.c_void_p * max(n, 1))(*[f.ctypes.data for f in flags])
        self._check(LIB.pano_ransac_score_pairs(self._h, n, arr, _i(best), _i(bcnt), cp, fp))
        return [(int(best[k]), int(bcnt[k]), counts[k][:len(keep[k][2])], flags[k][:len(keep[k][0])]) for k in range(n)]

    # -- bundle-adjustment Jacobian assembly
    def ba_jacobian(self, n_cam, pairs, pts_to, want_rows=True):
        """pairs: list of (from_slot, to_slot, n_match, mats [13, 9] f64) in match order; pts_to: [n_match_total, 2].
        Returns (j_rows [n_match_total, 24] or None, jtj [6 n_cam, 6 n_cam])."""
        n = len(pairs)
        arr = (PanoBaPair * max(n, 1))()
        begin = 0
        for k, (f, t, nm, mats) in enumerate(pairs):
            arr[k].from_, arr[k].to, arr[k].match_begin, arr[k].n_match = f, t, begin, nm
            C.memmove(arr[k].m, np.ascontiguousarray(mats, np.float64).ctypes.data, 13 * 9 * 8)
            begin += nm
        pts_to = np.ascontiguousarray(pts_to, np.float64).reshape(-1, 2)
        assert len(pts_to) == begin
        rows = np.zeros((max(begin, 1), 24), np.float64) if want_rows else None
        jtj = np.full((6 * n_cam, 6 * n_cam), np.nan, np.float64)
        self._check(LIB.pano_ba_jacobian(self._h, n_cam, n, arr, _d(pts_to), _d(rows) if want_rows else None, _d(jtj)))
        return (rows[:begin] if want_rows else None), jtj

    # -- cylinder warp
    @staticmethod
    def cyl_warp_shape(w, h, h_factor=1.0, params=None):
        params = params or default_params()
        ow, oh, ox, oy = C.c_int(), C.c_int(), C.c_double(), C.c_double()
        rc = LIB.pano_cyl_warp_shape(w, h, h_factor, C.byref(params), C.byref(ow), C.byref(oh), C.byref(ox),
                                     C.byref(oy))
        if rc:
            raise PanoError(rc, "pano_cyl_warp_shape")
        return ow.value, oh.value, ox.value, oy.value

    def cyl_warp(self, img, kpts=None, h_factor=1.0, params=None):
        params = params or default_params()
        img = np.ascontiguousarray(img, np.float32)
        ow, oh, _, _ = self.cyl_warp_shape(img.shape[1], img.shape[0], h_factor, params)
        out = np.empty((oh, ow, 3), np.float32)
        k = np.ascontiguousarray(kpts if kpts is not None else np.zeros((0, 2)), np.float64).copy()
        self._check(LIB.pano_cyl_warp(self._h, _f(img), img.shape[1], img.shape[0], h_factor, C.byref(params),
                                      _f(out), ow, oh, _d(k), len(k)))
        return out, k

    def cyl_warp_batch_dev(self, src_ptrs, shapes, dst_ptrs, kpts=None, h_factor=1.0, params=None):
        """Device pointers in, device pointers out (out sizes from cyl_warp_shape), asynchronous; kpts: optional
        list of [n, 2] float64 arrays rewritten in place."""
        params = params or default_params()
        n = len(src_ptrs)
        arr = (PanoCylJob * max(n, 1))()
        for k in range(n):
            h, w = shapes[k]
            ow, oh, _, _ = self.cyl_warp_shape(w, h, h_factor, params)
            arr[k].d_rgb_hwc, arr[k].w, arr[k].h = src_ptrs[k], w, h
            arr[k].d_out_hwc, arr[k].out_w, arr[k].out_h = dst_ptrs[k], ow, oh
            if kpts is not None and kpts[k] is not None and len(kpts[k]):
                assert kpts[k].dtype == np.float64 and kpts[k].flags["C_CONTIGUOUS"]
                arr[k].kpts_xy, arr[k].n_kpts = kpts[k].ctypes.data, len(kpts[k])
        self._check(LIB.pano_cyl_warp_batch_dev(self._h, n, arr, h_factor, C.byref(params)))

    # -- blend
    @staticmethod
    def _blend_args(imgs_or_ptrs, shapes, items, geom):
        n = len(items)
        arr = (PanoBlendImage * n)()
        for k in range(n):
            arr[k].rgb_hwc = imgs_or_ptrs[k]
            arr[k].h, arr[k].w = shapes[k]
            arr[k].x0, arr[k].y0, arr[k].x1, arr[k].y1 = items[k][:4]
            for q in range(9):
                arr[k].homo_inv[q] = items[k][4][q]
        g = PanoBlendGeom(projection=geom["projection"], res_x=geom["res_x"], res_y=geom["res_y"],
                          proj_min_x=geom["proj_min_x"], proj_min_y=geom["proj_min_y"])
        return arr, g

    def blend(self, imgs, items, geom, bands=0, params=None):
        """imgs: list of HxWx3 float32; items: (x0,y0,x1,y1,homo_inv[9]) per image."""
        params = params or default_params()
        imgs = [np.ascontiguousarray(im, np.float32) for im in imgs]
        arr, g = self._blend_args([im.ctypes.data for im in imgs], [im.shape[:2] for im in imgs], items, geom)
        ow, oh = C.c_int(), C.c_int()
        LIB.pano_blend_target_size(len(imgs), arr, C.byref(ow), C.byref(oh))
        out = np.empty((oh.value, ow.value, 3), np.float32)
        self._check(LIB.pano_blend(self._h, len(imgs), arr, C.byref(g), bands, C.byref(params), _f(out),
                                   ow.value, oh.value))
        return out

    def blend_dev(self, ptrs, shapes, items, geom, d_out, out_w, out_h, bands=0, params=None):
        params = params or default_params()
        arr, g = self._blend_args(ptrs, shapes, items, geom)
        self._check(LIB.pano_blend_dev(self._h, len(ptrs), arr, C.byref(g), bands, C.byref(params),
                                       C.c_void_p(d_out), out_w, out_h))

    # ---- 8-bit boundary: read_img / crop / write_rgb formats (device pointers)
    def rgb8_to_mat32f_batch_dev(self, d_pix, ws, hs, channels, d_out):
        """u8 -> f32 for n images in one launch (read_img's conversion, imgio.cc:75-88)."""
        n = len(d_pix)
        src = (C.c_void_p * n)(*d_pix)
        dst = (C.c_void_p * n)(*d_out)
        self._check(LIB.pano_rgb8_to_mat32f_batch_dev(self._h, n, src, (C.c_int * n)(*ws), (C.c_int * n)(*hs),
                                                      (C.c_int * n)(*channels), dst))

    def rgb8_to_mat32f_dev(self, d_pix, w, h, channels, d_out):
        self._check(LIB.pano_rgb8_to_mat32f_dev(self._h, C.c_void_p(d_pix), w, h, channels, C.c_void_p(d_out)))

    def crop_rect_dev(self, d_mat, w, h, d_rect):
        """crop()'s rectangle (imgproc.cc:200-235) into device int[4] {x0,y0,w,h}."""
        self._check(LIB.pano_crop_rect_dev(self._h, C.c_void_p(d_mat), w, h, C.c_void_p(d_rect)))

    def mat32f_to_rgb8_dev(self, d_mat, w, h, d_rect, d_out):
        """write_rgb's conversion (imgio.cc:98-113) of the rectangle d_rect (0/None = whole image)."""
        self._check(LIB.pano_mat32f_to_rgb8_dev(self._h, C.c_void_p(d_mat), w, h, C.c_void_p(d_rect or 0),
                                                C.c_void_p(d_out)))

    # numpy conveniences for tests
    def read_img_rgb8(self, pix):
        pix = np.ascontiguousarray(pix, np.uint8)
        h, w = pix.shape[:2]
        ch = 1 if pix.ndim == 2 else pix.shape[2]
        d_in = self.dev_alloc(max(pix.nbytes, 256))
        d_out = self.dev_alloc(h * w * 12)
        out = np.empty((h, w, 3), np.float32)
        try:
            self.dev_upload(d_in, pix)
            self.rgb8_to_mat32f_dev(d_in, w, h, ch, d_out)
            self.dev_download(out, d_out)
        finally:
            self.dev_free(d_in)
            self.dev_free(d_out)
        return out

    def crop_write_rgb8(self, mat, crop=True):
        """Returns (rect or None, 8-bit pixels of the (cropped) mosaic)."""
        mat = np.ascontiguousarray(mat, np.float32)
        h, w = mat.shape[:2]
        d_mat = self.dev_alloc(mat.nbytes)
        d_rect = self.dev_alloc(256)
        d_out = self.dev_alloc(max(h * w * 3, 256))
        rect = np.zeros(4, np.int32)
        out = np.empty(h * w * 3, np.uint8)
        try:
            self.dev_upload(d_mat, mat)
            if crop:
                self.crop_rect_dev(d_mat, w, h, d_rect)
            self.mat32f_to_rgb8_dev(d_mat, w, h, d_rect if crop else 0, d_out)
            self.dev_download(out, d_out)
            if crop:
                self.dev_download(rect, d_rect)
        finally:
            for p_ in (d_mat, d_rect, d_out):
                self.dev_free(p_)
        if not crop:
            return None, out.reshape(h, w, 3)
        cw, ch = int(rect[2]), int(rect[3])
        return rect, out[:cw * ch * 3].reshape(ch, cw, 3).copy()
