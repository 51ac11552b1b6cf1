"""CPU: host-side logic around the kernels (synthetic inputs, geometry set-up,
byte model of bench.py)."""
import numpy as np

from openpano_b200 import synth
from openpano_b200._abi import default_params


def test_synth_is_deterministic_and_textured():
    a = synth.make_canvas(120, 160, 5)
    b = synth.make_canvas(120, 160, 5)
    assert a.dtype == np.float32 and a.shape == (120, 160, 3)
    assert np.array_equal(a, b)
    assert a.min() >= 0 and a.max() <= 1 and a.std() > 0.1
    assert not np.array_equal(a, synth.make_canvas(120, 160, 6))


def test_stack_views_are_crops_of_one_canvas():
    imgs, org = synth.make_stack(4, 100, 80, 30, 9)
    assert org == [(0, 0), (30, 0), (60, 0), (90, 0)]
    assert np.array_equal(imgs[0][:, 30:], imgs[1][:, :70])
    imgs2, org2 = synth.make_stack(6, 50, 40, 20, 3, rows=2, step_y=15)
    assert len(imgs2) == 6 and org2[3][1] == 15                 # serpentine second row


def test_baseline_config_shapes():
    for name, cfg in synth.CONFIGS.items():
        assert cfg["n"] > 0 and cfg["w"] > 0 and cfg["h"] > 0
    c = synth.CONFIGS["ordered_13x1500x1112"]
    assert (c["n"], c["w"], c["h"]) == (13, 1500, 1112)


def test_translation_blend_setup_matches_reference_convention(orc):
    """The closed-form inverse map must send target pixel t to image pixel t - origin."""
    imgs, org = synth.make_stack(3, 90, 60, 30, 11)
    items, geom = synth.translation_blend_setup(org, 90, 60)
    assert [it[:4] for it in items] == [(0, 0, 90, 60), (30, 0, 120, 60), (60, 0, 150, 60)]
    out = orc.blend(imgs, items, geom, 0, default_params(lazy_read=0))
    canvas = synth.make_canvas(60, 150, 11)
    cov = out[..., 0] >= 0
    assert cov.mean() > 0.9 and np.abs(out[cov] - canvas[:out.shape[0], :out.shape[1]][cov]).max() < 1e-5


def test_octave_dims_match_the_oracle(orc):
    import bench
    for (w, h) in ((600, 400), (1500, 1112), (1300, 867), (4000, 3000), (333, 517)):
        img = np.zeros((h, w, 3), np.float32)
        tr = orc.sift_trace(img)
        want = [tr.octave_size(o) for o in range(4)]
        tr.close()
        assert bench.octave_dims(w, h, default_params()) == want, (w, h)


def test_survey_octave_sizes():
    """exact working/octave sizes of the BASELINE configs."""
    import bench
    assert bench.octave_dims(1500, 1112, default_params()) == [(918, 681), (650, 482), (459, 341), (325, 241)]
    assert bench.octave_dims(600, 400, default_params()) == [(960, 640), (679, 453), (480, 320), (340, 227)]


def test_algorithmic_bytes_follow_survey_formula():
    import bench
    imgs = [np.zeros((1112, 1500, 3), np.float32)]
    items = [(0, 0, 1500, 1112, [1, 0, 0, 0, 1, 0, 0, 0, 1])]
    ab = bench.algorithmic_bytes(imgs, items, default_params(), [1800])
    sp = 918 * 681 + 650 * 482 + 459 * 341 + 325 * 241
    assert ab["k_blur_dog"] == sp * 4 * 13                      # grey in, 6 levels + 6 |DoG| out
    assert ab["k_extrema_scan"] == sp * 24
    assert ab["k_descriptor"] == 1800 * 528


class _FakeLane:
    """Stands in for PipelinedStitcher: records the call order, returns the job's tag as its matches."""

    def __init__(self, depth=2):
        self.depth, self.log, self._next, self._inflight = depth, [], 0, {}

    def stage(self, host_ptrs, shapes, out_wh):
        k = self._next
        self._next = (self._next + 1) % self.depth
        if k in self._inflight and self._inflight[k] != "done":
            # the real stage() blocks here until the slot's previous job has left the device
            assert ("run", self._inflight[k]) in self.log, "slot reused before its job even ran"
            self.log.append(("sync", self._inflight[k]))
        self._inflight[k] = host_ptrs            # the tag travels in place of the pointers
        self.log.append(("stage", host_ptrs))
        return k

    def run(self, k, pairs, items, geom, out_ptr, bands=0):
        tag = self._inflight[k]
        self.log.append(("run", tag))
        return (k, tag)

    def wait(self, job):
        k, tag = job
        if self._inflight.get(k) == tag:
            self._inflight[k] = "done"
        self.log.append(("wait", tag))
        return [tag]

    def close(self):
        pass


def test_stitch_lanes_schedule_every_job_once_and_in_order():
    """StitchLanes.map: job i runs on lane i mod L, jobs of a lane run in order, the NEXT job of a
    lane is staged before the current one runs (so its upload overlaps), every job is waited for
    exactly once and results come back in job order."""
    from openpano_b200.stitcher import StitchLanes
    for n_lanes, n_jobs in ((1, 1), (2, 7), (3, 10), (3, 2)):
        lanes = StitchLanes.__new__(StitchLanes)
        lanes.lanes = [_FakeLane() for _ in range(n_lanes)]
        jobs = [(f"job{i}", None, None, None, None, None, None, 0) for i in range(n_jobs)]
        res = lanes.map(jobs)
        assert res == [[f"job{i}"] for i in range(n_jobs)]
        for q, lane in enumerate(lanes.lanes):
            mine = [f"job{i}" for i in range(n_jobs) if i % n_lanes == q]
            assert [t for op, t in lane.log if op == "run"] == mine
            assert [t for op, t in lane.log if op == "wait"] == mine
            assert [t for op, t in lane.log if op == "stage"] == mine
            for a, b in zip(mine[:-1], mine[1:]):          # stage(next) precedes run(current)
                assert lane.log.index(("stage", b)) < lane.log.index(("run", a))


def test_stitch_lanes_surface_worker_errors():
    from openpano_b200.stitcher import StitchLanes

    class Boom(_FakeLane):
        def run(self, *a, **k):
            raise RuntimeError("lane failed")

    lanes = StitchLanes.__new__(StitchLanes)
    lanes.lanes = [_FakeLane(), Boom()]
    jobs = [(f"job{i}", None, None, None, None, None, None, 0) for i in range(4)]
    try:
        lanes.map(jobs)
    except RuntimeError as ex:
        assert "lane failed" in str(ex)
    else:
        raise AssertionError("the worker's exception was swallowed")


def test_descriptor_column_intervals_cover_the_window(tmp_path):
    """csrc/desc_interval.h (the window enumeration of k_descriptor) compiled for the host with the parity flags:
    for random keypoints — axis-aligned and near-axis orientations included — every position the reference's
    calc_descriptor accepts (sift.cc:107-124, brute force over the whole window) lies inside its column's interval."""
    import re
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    exe = tmp_path / "desc_interval_check"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-o", str(exe), str(root / "tools" / "probes" / "desc_interval_check.cc"), "-lm"],
                   check=True)
    out = subprocess.run([str(exe), "30000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    m = re.search(r"accepted (\d+)\s+listed (\d+).*missing (\d+)", out.stdout)
    assert m and int(m.group(3)) == 0
    accepted, listed = int(m.group(1)), int(m.group(2))
    assert accepted > 10_000_000 and listed < 1.06 * accepted      # a superset, and a tight one


def test_match_lists_wire_format_round_trip():
    """DistributedStitcher's list gather: every rank packs [counts of its dealt tasks | pairs], padded to the
    largest rank total; rank 0 unpacks the concatenation.  Uneven deals, empty lists and an idle rank included."""
    from openpano_b200.parallel import deal_pairs, pack_match_lists, unpack_match_lists
    rng = np.random.RandomState(3)
    for world, n_img in ((2, 7), (4, 9), (8, 5), (3, 2)):
        pairs = synth.all_pairs(n_img)
        counts = [int(c) for c in rng.randint(50, 400, n_img)]
        dealt = deal_pairs(pairs, counts, world)
        assert sorted(t for d in dealt for t in d) == list(range(len(pairs)))
        truth = {}
        for t, (i, j) in enumerate(pairs):
            c = 0 if t % 5 == 0 else int(rng.randint(0, min(counts[i], counts[j])))
            truth[t] = rng.randint(0, 1000, (c, 2)).astype(np.int32)
        ntask = max(max(len(d) for d in dealt), 1)
        tots = [sum(len(truth[t]) for t in d) for d in dealt]
        pad = ntask + 2 * max(max(tots), 1)
        wire = np.full(world * pad, -7, np.int32)                      # stale contents of the staging buffers
        for r, d in enumerate(dealt):
            buf = wire[r * pad:(r + 1) * pad]
            tot = pack_match_lists([truth[t] for t in d], ntask, buf)
            assert tot == tots[r]
            buf[ntask + 2 * tot:] = 0
        full = unpack_match_lists(wire, dealt, ntask, pad, len(pairs))
        for t in range(len(pairs)):
            assert full[t].shape == truth[t].shape and np.array_equal(full[t], truth[t]), (world, t)
