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


from .synth import all_pairs, ordered_pairs  # noqa: F401  (task lists live with the workload generator)


class Stitcher:
    def __init__(self, engine: Engine, params=None, overlap_blend: bool = True):
        """overlap_blend: the composite depends on the images and the caller's geometry only, not
        on the features (the geometry in between is host code), so it runs on a second context /
        stream of the same device next to SIFT + matching: its bandwidth-bound kernels fill the
        gaps of the issue-bound descriptor kernel.  Ordered with events; results are unchanged."""
        self.eng = engine
        self.params = params or default_params()
        self._aux = None
        self._overlap = overlap_blend
        self._ev_in = self._ev_blend = None
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
        if self._aux is not None:
            self._aux.sync()
            Engine.event_destroy(self._ev_in)
            Engine.event_destroy(self._ev_blend)
            self._aux.close()
            self._aux = None

    def _aux_engine(self):
        if self._aux is None:
            self._aux = Engine(self.eng.device)
            self._ev_in = self.eng.event_create()
            self._ev_blend = self._aux.event_create()
        return self._aux

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
        if self._overlap:
            aux = self._aux_engine()
            self.eng.event_record(self._ev_in)          # images (and the previous reader of d_out) are done on the main stream
            aux.event_wait(self._ev_in)
            aux.blend_dev(ptrs, shapes, items, geom, self._d_out, self._out_shape[0], self._out_shape[1], bands, self.params)
            aux.event_record(self._ev_blend)
        fs = self.eng.sift_detect_batch_ptr(ptrs, [s[1] for s in shapes], [s[0] for s in shapes], self.params,
                                            device=True)
        if want_matches:
            m = self.eng.match_pairs(fs, pairs, self.params)
        else:
            m = self.eng.match_pairs_dev(fs, pairs, self.params)
        if self._overlap:
            self.eng.event_wait(self._ev_blend)         # the mosaic is complete in main-stream order
        else:
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


class PipelinedStitcher:
    """Throughput form of Stitcher.build(): consecutive stitch jobs overlap on the
    device.  Three contexts share the GPU — upload, compute and download, each with
    its own stream — ordered by events, so job i+1's images cross PCIe while job i
    runs SIFT/match/blend and job i-1's mosaic streams back (PCIe is full duplex).

        slot = ps.stage(host_ptrs, shapes, out_wh)     # enqueue the H2D copies
        job  = ps.run(slot, pairs, items, geom, out_host_ptr)   # compute + enqueue D2H
        matches = ps.wait(job)                          # host mosaic + matches ready

    Call stage() for the NEXT job before run() of the current one: run() blocks the
    host while the match lists come back, and that is when the next upload flies.
    """

    def __init__(self, device: int, params=None, depth: int = 2, rgb8: bool = False, crop: bool = True):
        """rgb8: the host side speaks the reference's FILE formats instead of Mat32f —
        decoded 8-bit pixels in (what read_img starts from, imgio.cc:72) and the 8-bit
        mosaic out (what write_rgb saves, imgio.cc:98-113, after crop() when `crop`,
        main.cc:226-229) — so 3 B/px cross PCIe each way instead of 12.  The conversions
        run on the device with the reference's arithmetic.  The output buffer then
        holds a 256-byte header (int32 x0, y0, width, height of the crop rectangle)
        followed by height*width*3 packed bytes."""
        self.params = params or default_params()
        self.up = Engine(device)
        self.cmp = Engine(device)
        self.dn = Engine(device)
        self.depth = depth
        self.rgb8 = rgb8
        self.crop = crop
        self.slots = [dict(imgs=None, out=None, shapes=None, offs=None, out_wh=None, pix=None, pix_offs=None, out8=None,
                           ev_up=self.up.event_create(), ev_cmp=self.cmp.event_create(),
                           ev_dn=self.dn.event_create(), busy=False) for _ in range(depth)]
        self._next = 0

    def _ensure(self, s, shapes, out_wh):
        realloc = False
        if s["shapes"] != shapes:
            realloc = True
            if s["imgs"]:
                self.cmp.dev_free(s["imgs"])
            offs, total = [], 0
            for (h, w) in shapes:
                offs.append(total)
                total += (h * w * 3 * 4 + 255) // 256 * 256
            s["imgs"] = self.cmp.dev_alloc(max(total, 256))
            s["shapes"], s["offs"] = list(shapes), offs
            if self.rgb8:
                if s["pix"]:
                    self.cmp.dev_free(s["pix"])
                poffs, ptotal = [], 0
                for (h, w) in shapes:
                    poffs.append(ptotal)
                    ptotal += (h * w * 3 + 255) // 256 * 256
                s["pix"] = self.cmp.dev_alloc(max(ptotal, 256))
                s["pix_offs"] = poffs
        if s["out_wh"] != out_wh:
            if s["out"]:
                self.cmp.dev_free(s["out"])
            s["out"] = self.cmp.dev_alloc(max(out_wh[0] * out_wh[1] * 3 * 4, 256))
            if self.rgb8:
                if s["out8"]:
                    self.cmp.dev_free(s["out8"])
                s["out8"] = self.cmp.dev_alloc(self.out_bytes(out_wh))
            s["out_wh"] = tuple(out_wh)
            realloc = True
        if realloc:
            # The buffers come stream-ordered from the compute context's pool, but the upload
            # stream writes them first: a block the pool hands out may still be in use by work
            # queued on cmp (the other slot's featureset, a SIFT arena freed in stream order).
            # Drain cmp so that the allocation is complete and the block idle before `up` touches it.
            self.cmp.sync()

    RGB8_HEADER = 256

    def out_bytes(self, out_wh) -> int:
        """Size of the host buffer run() fills for a canvas of out_wh."""
        if self.rgb8:
            return self.RGB8_HEADER + out_wh[0] * out_wh[1] * 3
        return out_wh[0] * out_wh[1] * 3 * 4

    def in_bytes(self, shapes) -> int:
        return sum(h * w * 3 * (1 if self.rgb8 else 4) for (h, w) in shapes)

    def stage(self, host_ptrs, shapes, out_wh) -> int:
        k = self._next
        self._next = (self._next + 1) % self.depth
        s = self.slots[k]
        if s["busy"]:
            Engine.event_sync(s["ev_dn"])          # the slot's previous job has fully left the device
            s["busy"] = False
        self._ensure(s, list(shapes), tuple(out_wh))
        self.up.event_wait(s["ev_cmp"])            # its previous compute no longer reads these images
        if self.rgb8:
            for p, o, (h, w) in zip(host_ptrs, s["pix_offs"], shapes):
                self.up.dev_upload_async(s["pix"] + o, p, h * w * 3)
        else:
            for p, o, (h, w) in zip(host_ptrs, s["offs"], shapes):
                self.up.dev_upload_async(s["imgs"] + o, p, h * w * 3 * 4)
        self.up.event_record(s["ev_up"])
        return k

    def run(self, k: int, pairs, items, geom, out_host_ptr, bands: int = 0):
        s = self.slots[k]
        shapes = s["shapes"]
        ptrs = [s["imgs"] + o for o in s["offs"]]
        self.cmp.event_wait(s["ev_up"])
        if self.rgb8:
            self.cmp.rgb8_to_mat32f_batch_dev([s["pix"] + o for o in s["pix_offs"]], [q[1] for q in shapes],
                                              [q[0] for q in shapes], [3] * len(shapes), ptrs)
        fs = self.cmp.sift_detect_batch_ptr(ptrs, [q[1] for q in shapes], [q[0] for q in shapes], self.params,
                                            device=True)
        matches = self.cmp.match_pairs(fs, pairs, self.params)       # host waits here; copies keep flowing
        self.cmp.event_wait(s["ev_dn"])                               # previous mosaic of this slot is out
        ow, oh = s["out_wh"]
        self.cmp.blend_dev(ptrs, shapes, items, geom, s["out"], ow, oh, bands, self.params)
        if self.rgb8:
            if self.crop:
                self.cmp.crop_rect_dev(s["out"], ow, oh, s["out8"])
            self.cmp.mat32f_to_rgb8_dev(s["out"], ow, oh, s["out8"] if self.crop else 0, s["out8"] + self.RGB8_HEADER)
        self.cmp.event_record(s["ev_cmp"])
        fs.free()
        self.dn.event_wait(s["ev_cmp"])
        if self.rgb8:
            self.dn.dev_download_async(out_host_ptr, s["out8"], self.out_bytes((ow, oh)))
        else:
            self.dn.dev_download_async(out_host_ptr, s["out"], ow * oh * 3 * 4)
        self.dn.event_record(s["ev_dn"])
        s["busy"] = True
        return (k, matches)

    def wait(self, job):
        k, matches = job
        Engine.event_sync(self.slots[k]["ev_dn"])
        self.slots[k]["busy"] = False
        return matches

    def close(self):
        for e in (self.up, self.cmp, self.dn):
            try:
                e.sync()
            except Exception:
                pass
        for s in self.slots:
            if s["imgs"]:
                self.cmp.dev_free(s["imgs"])
            if s["out"]:
                self.cmp.dev_free(s["out"])
            for key in ("pix", "out8"):
                if s[key]:
                    self.cmp.dev_free(s[key])
            for key in ("ev_up", "ev_cmp", "ev_dn"):
                Engine.event_destroy(s[key])
        self.slots = []
        for e in (self.up, self.cmp, self.dn):
            e.close()


def unpack_rgb8_mosaic(buf: np.ndarray, out_wh, cropped: bool = True):
    """View of the 8-bit mosaic PipelinedStitcher(rgb8=True).run() wrote into `buf`
    (uint8, out_bytes long).  Returns (rect (x0, y0, w, h), H×W×3 uint8 view)."""
    hdr = PipelinedStitcher.RGB8_HEADER
    if cropped:
        x0, y0, w, h = (int(v) for v in buf[:16].view(np.int32))
    else:
        x0, y0, w, h = 0, 0, out_wh[0], out_wh[1]
    return (x0, y0, w, h), buf[hdr:hdr + w * h * 3].reshape(h, w, 3)


class StitchLanes:
    """Several PipelinedStitcher lanes on ONE GPU, one host thread per lane.

    A single lane leaves the compute stream idle whenever its host thread is between
    two dependent phases of a job (feature counts -> match plan -> match lists ->
    blend launch) and whenever a tiny metadata move waits on a busy PCIe link.  With
    two lanes the other job's kernels fill those gaps (ctypes releases the GIL inside
    every engine call, so the lanes' host threads run concurrently).

        lanes = StitchLanes(device, params, lanes=2, rgb8=True)
        results = lanes.map(jobs)     # jobs[i] = (host_ptrs, shapes, out_wh, pairs, items, geom, out_host_ptr, bands)
    """

    def __init__(self, device: int, params=None, lanes: int = 2, depth: int = 2, rgb8: bool = False, crop: bool = True):
        self.lanes = [PipelinedStitcher(device, params, depth=depth, rgb8=rgb8, crop=crop) for _ in range(lanes)]
        self.done_times = []       # perf_counter() at which each job of the last map() came back (diagnostics)

    def out_bytes(self, out_wh):
        return self.lanes[0].out_bytes(out_wh)

    def in_bytes(self, shapes):
        return self.lanes[0].in_bytes(shapes)

    @staticmethod
    def _run_lane(ps, jobs, results, errors, done=None):
        import time
        try:
            if not jobs:
                return
            it = iter(jobs)
            idx, job = next(it)
            slot = ps.stage(job[0], job[1], job[2])
            pending = None
            while job is not None:
                nxt = next(it, None)
                nslot = ps.stage(nxt[1][0], nxt[1][1], nxt[1][2]) if nxt is not None else None
                handle = ps.run(slot, job[3], job[4], job[5], job[6], job[7])
                if pending is not None:
                    results[pending[0]] = ps.wait(pending[1])
                    if done is not None:
                        done[pending[0]] = time.perf_counter()
                pending = (idx, handle)
                if nxt is None:
                    break
                (idx, job), slot = nxt, nslot
            results[pending[0]] = ps.wait(pending[1])
            if done is not None:
                done[pending[0]] = time.perf_counter()
        except Exception as ex:  # surfaced by map()
            errors.append(ex)

    def map(self, jobs):
        """Runs the jobs (in order within a lane, job i on lane i mod L); returns their match lists."""
        import threading
        jobs = list(jobs)
        results = [None] * len(jobs)
        errors = []
        done = [0.0] * len(jobs)
        L = len(self.lanes)
        threads = [threading.Thread(target=self._run_lane, args=(self.lanes[q], [(i, j) for i, j in enumerate(jobs) if i % L == q],
                                                                 results, errors, done)) for q in range(L)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        self.done_times = done
        return results

    def close(self):
        for ps in self.lanes:
            ps.close()
        self.lanes = []
