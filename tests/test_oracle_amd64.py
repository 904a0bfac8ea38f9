"""The amd64-numerics part of the oracle (oracle/nl_oracle_amd64.c: the reference's AVX2 kernels restated
lane by lane) against (1) the same instruction sequences executed with real AVX2/FMA instructions
(oracle/nl_oracle_simd.c; skipped on hosts without AVX2) and (2) independent numpy statements of the
lane layout.  The reference itself has no test for these kernels."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import oracle as O  # noqa: E402

fp = C.POINTER(C.c_float)


def _p(a):
    return a.ctypes.data_as(fp)


def frame(w, h, seed, specials=False):
    rng = np.random.default_rng(seed)
    img = (rng.standard_normal((h, w)) * 37.0 + 900.0).astype(np.float32)
    img[rng.random((h, w)) < 0.002] += np.float32(5000)
    if specials:
        m = rng.random((h, w))
        img[m < 0.02] = np.float32(0.0)
        img[(m >= 0.02) & (m < 0.04)] = np.float32(-0.0)
        img[(m >= 0.04) & (m < 0.05)] = np.nan
        img[(m >= 0.05) & (m < 0.055)] = np.inf
    return img


needs_simd = pytest.mark.skipif(O.simd() is None, reason="host CPU without AVX2+FMA")


@needs_simd
@pytest.mark.parametrize("specials", [False, True])
def test_min_mean_max_variance_against_real_avx2(specials):
    for seed, n in enumerate([4, 8, 12, 4096, 100000, 1 << 20]):
        data = frame(n, 1, seed, specials).ravel()
        want = [C.c_float(), C.c_float(), C.c_float()]
        O.simd().nlo_simd_min_mean_max(_p(data), n, *[C.byref(x) for x in want])
        got = [C.c_float(), C.c_float(), C.c_float()]
        O.lib().nlo_calc_min_mean_max_avx2(_p(data), n, *[C.byref(x) for x in got])
        a = np.array([x.value for x in want], np.float32).view(np.uint32)
        b = np.array([x.value for x in got], np.float32).view(np.uint32)
        assert np.array_equal(a, b), (n, a, b)
        mean = np.float32(got[1].value) if not specials else np.float32(900.25)
        v1 = O.simd().nlo_simd_variance(_p(data), n, mean)
        v2 = O.lib().nlo_calc_variance_avx2(_p(data), n, mean)
        assert np.array([v1]).view(np.uint64)[0] == np.array([v2]).view(np.uint64)[0], n


@needs_simd
@pytest.mark.parametrize("specials", [False, True])
def test_noise_line_against_real_avx2(specials):
    for seed, w in enumerate(list(range(8, 40)) + [100, 101, 1023, 1024, 4099]):
        rows = frame(w, 3, 100 + seed, specials)
        want = np.float32(O.simd().nlo_simd_noise_line(_p(rows), w))
        got = np.float32(O.lib().nlo_estimate_noise_line_avx2(_p(rows), w))
        assert want.view(np.uint32) == got.view(np.uint32) or (np.isnan(want) and np.isnan(got)), (w, want, got)


@needs_simd
@pytest.mark.parametrize("specials", [False, True])
def test_median_filter_against_real_avx2(specials):
    for seed, (w, h) in enumerate([(8, 3), (9, 5), (13, 7), (14, 4), (37, 11), (256, 33)]):
        img = frame(w, h, 200 + seed, specials)
        got = O.median_filter3x3(img, w, amd64=True).reshape(h, w)
        want = img.copy()
        for y in range(h - 2):
            rows = np.ascontiguousarray(img[y:y + 3])
            dest = np.ascontiguousarray(want[y:y + 3])
            dest[1, 0], dest[1, -1] = img[y + 1, 0], img[y + 1, -1]
            O.simd().nlo_simd_median_line(_p(dest), _p(rows), w)
            want[y + 1] = dest[1]
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (w, h)


def test_stats_lane_layout_numpy():
    """element i -> lane i%4, float64 sums per lane, fold (0+1)+(2+3)"""
    data = frame(4096, 3, 7).ravel()
    lanes = data.reshape(-1, 4).astype(np.float64)
    s = np.zeros(4)
    for row in lanes:
        s += row
    mean = np.float32(((s[2] + s[3]) + (s[0] + s[1])) / float(data.size))
    d = (data - mean).astype(np.float32).reshape(-1, 4).astype(np.float64)
    v = np.zeros(4)
    for row in d:
        v += row * row
    std = np.float32(np.sqrt(((v[2] + v[3]) + (v[0] + v[1])) / float(data.size)))
    got = O.stats(data, amd64=True)
    assert got[0] == data.min() and got[2] == data.max()
    assert got[1].view(np.uint32) == mean.view(np.uint32)
    assert got[3].view(np.uint32) == std.view(np.uint32)
    # the pure-Go definition: one sequential float64 chain
    acc = 0.0
    for x in data.astype(np.float64):
        acc += x
    pg = O.stats(data, amd64=False)
    assert pg[1].view(np.uint32) == np.float32(acc / data.size).view(np.uint32)


def test_median_filter_is_the_median():
    """NaN-free data without signed-zero ties: both variants equal the plain 3x3 median, borders copied"""
    w, h = 61, 23
    img = frame(w, h, 3)
    stack = np.stack([img[dy:h - 2 + dy, dx:w - 2 + dx] for dy in range(3) for dx in range(3)])
    want = img.copy()
    want[1:-1, 1:-1] = np.sort(stack, axis=0)[4]
    for amd64 in (True, False):
        assert np.array_equal(O.median_filter3x3(img, w, amd64).reshape(h, w), want)


def test_noise_variants_agree_to_rounding_and_differ_in_order():
    w, h = 250, 40
    img = frame(w, h, 11)
    a, b = O.estimate_noise(img, w, amd64=True), O.estimate_noise(img, w, amd64=False)
    assert abs(float(a) - float(b)) <= 1e-5 * float(b)
    # the whole-image value is the sequential fp32 sum of the line values (noise_amd64.go:36-42)
    s = np.float32(0)
    for y in range(h - 2):
        s = np.float32(s + np.float32(O.lib().nlo_estimate_noise_line_avx2(_p(np.ascontiguousarray(img[y:y + 3])), w)))
    factor = np.float32(np.float32(np.sqrt(0.5 * np.pi)) / (np.float32(6) * np.float32(w - 2) * np.float32(h - 2)))
    assert np.float32(s * factor).view(np.uint32) == a.view(np.uint32)


def test_bad_pixel_map_definition():
    w, h = 64, 48
    img = frame(w, h, 5)
    bpm, st, diff = O.bad_pixel_map(img, w, 3.0, 5.0, amd64=True)
    med = O.median_filter3x3(img, w, True)
    assert np.array_equal(diff, img.ravel() - med)
    assert np.array_equal(st.view(np.uint32), O.stats(diff, True).view(np.uint32))
    want = np.nonzero((diff < -st[3] * np.float32(3.0)) | (diff > st[3] * np.float32(5.0)))[0]
    assert np.array_equal(bpm, want.astype(np.int32)) and len(bpm) > 0
