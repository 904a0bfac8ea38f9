"""The sparse, order-dependent steps of FindStars as the library runs them on host threads (nl_star_*_host: no device
needed) against the oracle's restatement of the reference: bad-pixel rejection with the device/host split of the batched
path (border candidates replayed with the carried-over gather buffer), the unstable quicksort run on keys, and the
overlap filter on a fine grid instead of the reference's 256-pixel bins."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402

import nightlight_b200 as nl  # noqa: E402

fp = C.POINTER(C.c_float)


def _olib():
    L = O.lib()
    sp = C.POINTER(O.Star)
    L.nlo_reject_bad_pixels.restype = C.c_int
    L.nlo_reject_bad_pixels.argtypes = [sp, C.c_int, fp, C.c_int32, C.c_int32, C.c_float, C.c_float]
    L.nlo_qsort_stars_desc.restype = None
    L.nlo_qsort_stars_desc.argtypes = [sp, C.c_int]
    L.nlo_filter_out_overlaps.restype = C.c_int
    L.nlo_filter_out_overlaps.argtypes = [sp, C.c_int, C.c_int32, C.c_int32, C.c_int32]
    return L, sp


def _stars(n):
    assert O.STAR_DTYPE.itemsize == nl.STAR_DTYPE.itemsize == C.sizeof(O.Star)
    return np.zeros(max(n, 1), dtype=nl.STAR_DTYPE)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_reject_bad_pixels_interior_on_their_own_border_candidates_replayed(seed):
    lib = nl.load_library()
    L, sp = _olib()
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(8, 90)), int(rng.integers(4, 60))
    data = (rng.standard_normal(w * h) * 10 + 100).astype(np.float32)
    # bright pixels everywhere, and plenty in the first / last row and at both ends of the frame (runs of them, too)
    hot = rng.integers(0, w * h, 60)
    data[hot] += rng.choice([30.0, 200.0, 3000.0], hot.size).astype(np.float32)
    for row in (0, h - 1):
        cols = rng.integers(0, w, max(3, w // 3))
        data[row * w + cols] += rng.choice([40.0, 500.0], cols.size).astype(np.float32)
    data[0] += 900.0
    data[-1] += 900.0
    if seed % 2:
        data[rng.integers(0, w * h, 5)] = np.nan
    cand = O.find_bright_pixels(data, w, 125.0, int(rng.integers(0, 3)))
    assert len(cand) > 10
    border = sum(1 for c in cand if c["index"] - w - 1 < 0 or c["index"] + w + 1 >= w * h)
    assert border >= 3
    for sigma, mds in ((5.0, 4.0), (1.0, 10.0), (50.0, 30.0)):
        a = cand.copy()
        b = _stars(len(cand))
        b[:len(cand)] = cand
        ka = L.nlo_reject_bad_pixels(a.ctypes.data_as(sp), len(a), data.ctypes.data_as(fp), w * h, w, sigma, mds)
        kb = C.c_int32()
        nl.binding.check(lib.nl_star_reject_bad_pixels_host(b.ctypes.data_as(C.c_void_p), len(cand), data.ctypes.data_as(fp), w * h, w, sigma, mds,
                                                            C.byref(kb)))
        assert ka == kb.value, (seed, sigma, ka, kb.value)
        assert a[:ka].tobytes() == b[:ka].tobytes()


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_sort_by_mass_with_ties_is_the_reference_permutation(seed):
    lib = nl.load_library()
    L, sp = _olib()
    rng = np.random.default_rng(seed)
    for n in (0, 1, 2, 3, 17, 256, 5000):
        a = _stars(n)
        a["index"] = np.arange(a.size)
        a["mass"] = rng.choice(np.array([1.0, 2.0, 2.0, 3.5, 65535.0, -0.0, 0.0], np.float32), a.size) if seed % 2 else \
            np.round(rng.standard_normal(a.size) * 3).astype(np.float32)
        b = a.copy()
        L.nlo_qsort_stars_desc(a.ctypes.data_as(sp), n)
        nl.binding.check(lib.nl_star_sort_desc_host(b.ctypes.data_as(C.c_void_p), n))
        assert a[:n].tobytes() == b[:n].tobytes(), (seed, n)


@pytest.mark.parametrize("seed", [21, 22, 23, 24, 25])
def test_overlap_filter_on_the_fine_grid_gives_the_reference_verdicts(seed):
    """dense candidate lists, every radius class (tiny, typical, the largest the fine grid takes, beyond it: the
    256-pixel bins), star centres outside the image on every side, coordinates exactly on cell and bin borders"""
    lib = nl.load_library()
    L, sp = _olib()
    rng = np.random.default_rng(seed)
    for radius in (0, 1, 4, 16, 50, 126, 127, 128, 300):
        w, h = int(rng.integers(300, 1500)), int(rng.integers(200, 1100))
        n = int(rng.integers(500, 6000))
        a = _stars(n)
        a["x"] = rng.uniform(-1.5 * 256, w + 1.5 * 256, a.size).astype(np.float32)
        a["y"] = rng.uniform(-1.5 * 256, h + 1.5 * 256, a.size).astype(np.float32)
        inside = rng.random(a.size) < 0.8
        a["x"][inside] = rng.uniform(0, w, int(inside.sum())).astype(np.float32)
        a["y"][inside] = rng.uniform(0, h, int(inside.sum())).astype(np.float32)
        snap = rng.random(a.size) < 0.2                      # integer and half-integer coordinates, multiples of the cell sizes
        a["x"][snap] = (np.round(a["x"][snap] / 16) * 16 + rng.choice([0.0, 0.5, -0.5], int(snap.sum()))).astype(np.float32)
        a["y"][snap] = (np.round(a["y"][snap] / 16) * 16 + rng.choice([0.0, 0.5, -0.5], int(snap.sum()))).astype(np.float32)
        a["index"] = np.arange(a.size)
        a["mass"] = rng.random(a.size).astype(np.float32)
        b = a.copy()
        ka = L.nlo_filter_out_overlaps(a.ctypes.data_as(sp), n, w, h, radius)
        kb = C.c_int32()
        nl.binding.check(lib.nl_star_filter_overlaps_host(b.ctypes.data_as(C.c_void_p), n, w, h, radius, C.byref(kb)))
        assert ka == kb.value, (seed, radius, ka, kb.value)
        assert a[:ka].tobytes() == b[:ka].tobytes(), (seed, radius)
