"""Shared loaders for the fixtures under tests/golden/ (generated from the
reference's own translation units by tests/golden/make_golden.py)."""
import hashlib
from pathlib import Path

import numpy as np

from openpano_b200 import synth

GOLDEN = Path(__file__).resolve().parent / "golden"


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def plane_crc(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], np.uint64).copy()


def load(name):
    return np.load(GOLDEN / name, allow_pickle=False)


def sift_input():
    return synth.make_canvas(180, 240, 101)


def match_inputs():
    return synth.make_stack(2, 240, 180, 90, 102)[0]


def warp_input():
    return synth.make_canvas(80, 120, 103)


def blend_inputs():
    imgs, org = synth.make_stack(3, 120, 80, 40, 104)
    items, geom = synth.translation_blend_setup(org, 120, 80)
    return imgs, items, geom


def imgio_inputs():
    """(8-bit colour image, 8-bit grey image, f32 mosaic with Color::NO outside a slanted quadrilateral
    and a few holes) for the read_img / crop / write_rgb fixtures."""
    pix = (synth.make_canvas(60, 90, 105) * 255.0 + 0.5).astype(np.uint8)
    grey = pix[..., 1].copy()
    mos = synth.make_canvas(70, 110, 106)
    yy, xx = np.mgrid[0:70, 0:110]
    inside = (yy > 4 + 0.08 * xx) & (yy < 64 - 0.05 * xx) & (xx > 3 + 0.1 * yy) & (xx < 105 - 0.07 * yy)
    mos[~inside] = -1.0
    mos[30:33, 50:54] = -1.0
    mos[10:12, 80:81] = -1.0
    return pix, grey, mos


def same_bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    return a.tobytes() == b.tobytes()


def check_sift_trace(tr, g):
    """tr: any SIFT trace (oracle, reference shim or GPU engine)."""
    assert tuple(g["working_size"]) == tr.working_size()
    for o in range(4):
        assert tuple(g["octave_sizes"][o]) == tr.octave_size(o)
        for l in range(7):
            assert np.array_equal(plane_crc(tr.plane(1, o, l)), g[f"gauss_o{o}_l{l}_crc"]), f"gaussian[{o}][{l}]"
        for l in range(6):
            assert np.array_equal(plane_crc(tr.plane(2, o, l)), g[f"dog_o{o}_l{l}_crc"]), f"dog[{o}][{l}]"
    assert same_bits(tr.plane(2, 0, 3)[40:60, 50:90], g["dog_o0_l3"])
    for stage, key, fields in ((0, "raw", ("x", "y", "pyr_id", "scale_id")),
                               (1, "refined", ("x", "y", "pyr_id", "scale_id", "real_x", "real_y", "scale_factor")),
                               (2, "oriented", ("x", "y", "pyr_id", "scale_id", "real_x", "real_y", "scale_factor", "dir"))):
        p = tr.points(stage)
        assert len(p) == len(g[key]), (key, len(p), len(g[key]))
        for f in fields:
            assert same_bits(p[f], g[key][f]), f"{key}.{f}"
    coor, desc = tr.descriptors()
    assert same_bits(coor, g["coor"])
    assert same_bits(desc, g["desc"])
