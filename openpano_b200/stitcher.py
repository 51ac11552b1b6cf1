"""Host-side mirror of the reference's orchestration for the hot path.

`Stitcher.build()` follows Stitcher::build() (stitch/stitcher.cc:32-64):
calc_feature() (stitcherbase.cc:9-27) -> linear/pairwise match
(stitcher.cc:96-136) -> ConnectedImages::blend() (stitcher_image.cc:116-155).
The geometry in between (RANSAC, camera estimation, bundle adjustment) is host
code outside the hot path (SURVEY.md §8 scope): the caller supplies the
per-image inverse homographies and ranges it produced, exactly the numbers the
reference's blend lambda closes over.

Everything heavy is one C-ABI call into libpano_b200.so; this class only owns
the device buffers so that images are uploaded once per build and reused by the
feature and blend stages.
"""
from __future__ import annotations

import numpy as np

from ._abi import default_params
from .capi import Engine


def ordered_pairs(n: int):
    """linear_pairwise_match task list: (i, (i+1) % n) (stitcher.cc:116-123)."""
    return [(i, (i + 1) % n) for i in range(n)]


def all_pairs(n: int):
    """pairwise_match task list (stitcher.cc:98-100)."""
    return [(i, j) for i in range(n) for j in range(i + 1, n)]


class Stitcher:
    def __init__(self, engine: Engine, params=None):
        self.eng = engine
        self.params = params or default_params()
        self._d_imgs = None      # device block holding all input images
        self._d_out = None
        self._shapes = None
        self._offs = None
        self._out_shape = None

    # -- device buffers, (re)allocated only when the workload shape changes
    def _ensure(self, shapes, out_wh):
        if self._shapes != shapes:
            self.release_images()
            offs, total = [], 0
            for (h, w) in shapes:
                offs.append(total)
                total += (h * w * 3 * 4 + 255) // 256 * 256
            self._d_imgs = self.eng.dev_alloc(max(total, 256))
            self._shapes, self._offs = list(shapes), offs
        if self._out_shape != out_wh:
            if self._d_out:
                self.eng.dev_free(self._d_out)
            self._d_out = self.eng.dev_alloc(max(out_wh[0] * out_wh[1] * 3 * 4, 256))
            self._out_shape = out_wh

    def release_images(self):
        if self._d_imgs:
            self.eng.dev_free(self._d_imgs)
            self._d_imgs = None
            self._shapes = None

    def close(self):
        self.release_images()
        if self._d_out:
            self.eng.dev_free(self._d_out)
            self._d_out = None
            self._out_shape = None

    def image_ptrs(self):
        return [self._d_imgs + o for o in self._offs]

    # -- stages on device-resident inputs (bench `value` leg)
    def upload(self, host_ptrs, shapes, out_wh):
        """host_ptrs: raw pointers of PINNED H×W×3 float32 buffers."""
        self._ensure(list(shapes), tuple(out_wh))
        for p, o, (h, w) in zip(host_ptrs, self._offs, shapes):
            self.eng.dev_upload_async(self._d_imgs + o, p, h * w * 3 * 4)

    def run_device(self, pairs, items, geom, bands=0, want_matches=False):
        """SIFT + match + blend on the images already in HBM.  Returns
        (featureset, matches-or-total)."""
        shapes = self._shapes
        ptrs = self.image_ptrs()
        fs = self.eng.sift_detect_batch_ptr(ptrs, [s[1] for s in shapes], [s[0] for s in shapes], self.params,
                                            device=True)
        if want_matches:
            m = self.eng.match_pairs(fs, pairs, self.params)
        else:
            m = self.eng.match_pairs_dev(fs, pairs, self.params)
        self.eng.blend_dev(ptrs, shapes, items, geom, self._d_out, self._out_shape[0], self._out_shape[1], bands,
                           self.params)
        return fs, m

    # -- the end-to-end call a user makes: host images in, host mosaic + matches out
    def build(self, host_ptrs, shapes, pairs, items, geom, out_host_ptr, bands=0):
        out_w = max(it[2] for it in items)
        out_h = max(it[3] for it in items)
        self.upload(host_ptrs, shapes, (out_w, out_h))
        fs, matches = self.run_device(pairs, items, geom, bands, want_matches=True)
        self.eng.dev_download_async(out_host_ptr, self._d_out, out_w * out_h * 3 * 4)
        self.eng.sync()
        fs.free()
        return matches

    def build_numpy(self, imgs, pairs, items, geom, bands=0):
        """Convenience for tests: numpy in/out (pageable memory; slower)."""
        imgs = [np.ascontiguousarray(im, np.float32) for im in imgs]
        out_w = max(it[2] for it in items)
        out_h = max(it[3] for it in items)
        out = np.empty((out_h, out_w, 3), np.float32)
        m = self.build([im.ctypes.data for im in imgs], [im.shape[:2] for im in imgs], pairs, items, geom,
                       out.ctypes.data, bands)
        return m, out
