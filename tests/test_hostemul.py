"""The product's per-pixel device routines (nightlight_b200/csrc/nl_column.cuh), compiled for the CPU
with element stride 1, against the oracle: the flattened quick-select must leave exactly the
reference's permutation, every reducer must return the oracle's bits and clip counts."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import oracle as O  # noqa: E402
from util import MODE_ID, bits_equal, first_mismatch, mode_cases, weights_for  # noqa: E402

fp = C.POINTER(C.c_float)


def emul_stack(L, frames, mode, sl=2.75, sh=2.75, w=None, ref_loc=0.0):
    frames = [np.ascontiguousarray(f, dtype=np.float32) for f in frames]
    n, p = len(frames), frames[0].size
    ptrs = (fp * n)(*[f.ctypes.data_as(fp) for f in frames])
    res = np.empty(p, np.float32)
    cl, ch = C.c_longlong(), C.c_longlong()
    wp = None
    if w is not None:
        w = np.ascontiguousarray(w, dtype=np.float32)
        wp = w.ctypes.data_as(fp)
    rc = L.emul_stack(MODE_ID[mode], ptrs, n, p, wp, ref_loc, sl, sh, res.ctypes.data_as(fp), C.byref(cl), C.byref(ch))
    assert rc == 0
    return res, cl.value, ch.value


def test_qselect_permutation(hostemul):
    rng = np.random.default_rng(7)
    for n in list(range(1, 70)) + [127, 128, 255, 256, 1000]:
        for kind in range(3):
            if kind == 0:
                a = rng.standard_normal(n).astype(np.float32)
            elif kind == 1:
                a = rng.integers(0, 4, n).astype(np.float32)        # many ties
            else:
                a = np.sort(rng.standard_normal(n).astype(np.float32))[::-1].copy()
            want_med, want_perm = O.qselect_median(a)
            b = a.copy()
            got = hostemul.emul_qselect_median(b.ctypes.data_as(fp), n)
            assert np.float32(got) == want_med, (n, kind)
            assert bits_equal(b, want_perm), (n, kind)


def test_sorts(hostemul):
    """the bitonic network with virtual +inf padding, for every length and a longer warp maximum"""
    rng = np.random.default_rng(3)
    for n in list(range(1, 40)) + [63, 64, 65, 100, 255, 256, 257, 1000]:
        for kind in range(2):
            a = rng.standard_normal(n).astype(np.float32) if kind == 0 else rng.integers(0, 4, n).astype(np.float32)
            for nmax in (n, n + 37):
                b = a.copy()
                hostemul.emul_sort(b.ctypes.data_as(fp), n, nmax)
                assert np.array_equal(b, np.sort(a)), (n, kind, nmax)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 16, 25, 64, 256])
def test_reducers_match_oracle(hostemul, n):
    frames = O.synth_frames(n, 1000 * n, 2048)
    w = weights_for(n)
    for mode, weighted in mode_cases():
        a = O.stack(frames, mode, weights=w if weighted else None)
        b = emul_stack(hostemul, frames, mode, w=w if weighted else None)
        assert bits_equal(a[0], b[0]), (mode, weighted, first_mismatch(a[0], b[0]))
        assert a[1:] == b[1:], (mode, weighted)


def test_reducers_random_data_and_sigmas(hostemul):
    rng = np.random.default_rng(11)
    for n in (6, 15, 30, 100):
        frames = (rng.standard_normal((n, 512)) * 50 + 1000).astype(np.float32)
        frames[rng.random(frames.shape) < 0.02] = np.nan
        frames[rng.random(frames.shape) < 0.03] += 2000
        frames[:, 5] = np.nan                       # an all-NaN pixel
        frames[:, 6] = 7.0                          # a constant pixel (sigma 0)
        w = rng.random(n).astype(np.float32) + np.float32(0.1)
        for sl, sh in ((2.75, 2.75), (1.0, 3.0), (-1.0, -1.0), (0.5, 0.5)):
            for mode, weighted in mode_cases():
                a = O.stack(frames, mode, sl, sh, weights=w if weighted else None, ref_loc=0.25)
                b = emul_stack(hostemul, frames, mode, sl, sh, w if weighted else None, ref_loc=0.25)
                assert bits_equal(a[0], b[0]), (n, sl, mode, weighted, first_mismatch(a[0], b[0]))
                assert a[1:] == b[1:], (n, sl, mode, weighted)


@pytest.mark.parametrize("n", [7, 16, 40, 256])
def test_signed_zero_ties_every_mode(hostemul, n):
    rng = np.random.default_rng(100 + n)
    frames = np.round(rng.standard_normal((n, 900)) * 0.6).astype(np.float32)
    frames[rng.random(frames.shape) < 0.02] = np.nan
    for mode, weighted in mode_cases():
        w = weights_for(n) if weighted else None
        for sl, sh in ((2.75, 2.75), (0.5, 1.0)):
            a = O.stack(frames, mode, sl, sh, weights=w)
            b = emul_stack(hostemul, frames, mode, sl, sh, w)
            assert bits_equal(a[0], b[0]), (n, mode, weighted, sl, first_mismatch(a[0], b[0]))
            assert a[1:] == b[1:]


@pytest.mark.parametrize("seed", [20261017, 1, 2, 3])
def test_fuzz_shapes_modes_signed_zeros(hostemul, seed):
    """seeded fuzz (the same generator as the GPU fuzz test): frame counts, NaN densities, outliers, ties
    including -0.0 / +0.0 mixes, negative sigmas, every mode -- device routines (host build) == oracle"""
    rng = np.random.default_rng(seed)
    cases = mode_cases()
    for it in range(60):
        n = int(rng.choice([2, 3, 4, 7, 9, 14, 17, 24, 26, 31, 33, 48, 65, 127, 129, 200, 255, 256, 257, 300]))
        p = int(rng.integers(1, 700))
        scale = float(rng.choice([1e-3, 1.0, 50.0, 4e4]))
        frames = (rng.standard_normal((n, p)) * scale + float(rng.choice([0.0, 1000.0, -3.0]))).astype(np.float32)
        if rng.random() < 0.7:
            frames[rng.random(frames.shape) < float(rng.choice([0.001, 0.02, 0.3]))] = np.nan
        if rng.random() < 0.7:
            frames[rng.random(frames.shape) < 0.03] += np.float32(20 * scale)
        if rng.random() < 0.3:
            frames = np.round(frames).astype(np.float32)
        sl, sh = (float(x) for x in rng.choice([0.5, 1.0, 2.0, 2.75, 4.0, -1.0], 2))
        mode, weighted = cases[int(rng.integers(0, len(cases)))]
        w = (rng.random(n).astype(np.float32) + np.float32(0.05)) if weighted else None
        ref = float(rng.choice([0.0, 7.5]))
        want = O.stack(frames, mode, sl, sh, weights=w, ref_loc=ref)
        got = emul_stack(hostemul, frames, mode, sl, sh, w, ref)
        assert bits_equal(got[0], want[0]), (it, n, p, mode, weighted, sl, sh, first_mismatch(got[0], want[0]))
        assert got[1:] == want[1:], (it, mode, got[1:], want[1:])


def test_infinities_and_huge_values_every_mode(hostemul):
    """IEEE corner values (the GPU test's data): +-inf samples, denormals, magnitudes whose sums and squares overflow --
    the winsorized clamps must not assume an ordered median there"""
    rng = np.random.default_rng(21)
    n, p = 40, 600
    frames = (rng.standard_normal((n, p)) * 10 + 100).astype(np.float32)
    frames[3, 0:50] = np.inf
    frames[7, 25:80] = -np.inf
    frames[:, 100:150] = (rng.standard_normal((n, 50)) * 1e-41).astype(np.float32)
    frames[:, 150:200] = (rng.standard_normal((n, 50)) * 1e30).astype(np.float32)
    frames[5, 200:220] = np.float32(3e38)
    frames[:, 220:240] = -0.0
    frames[::2, 230:240] = 0.0
    frames[:, 240:260] = np.float32(3.2e38)           # lower + upper of the median overflow
    frames[::3, 240:260] = np.float32(-3.2e38)
    for mode, weighted in mode_cases():
        w = weights_for(n) if weighted else None
        for sl, sh in ((2.75, 2.75), (1.0, 0.5)):
            a = O.stack(frames, mode, sl, sh, weights=w)
            b = emul_stack(hostemul, frames, mode, sl, sh, w)
            assert bits_equal(a[0], b[0]), (mode, weighted, sl, first_mismatch(a[0], b[0]))
            assert a[1:] == b[1:], (mode, weighted)
