"""SURVEY.md 8f N3 on the device: MedianFilter3x3, Stats.Min/Mean/Max/StdDev and pre.BadPixelMap, bit-exact
against the oracle in both numerics of the reference (AVX2 assembly of amd64 builds / pure-Go loops)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import nightlight_b200 as nl  # noqa: E402
from oracle import oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    with nl.Context(0) as c:
        yield c


@pytest.fixture(params=[True, False], ids=["amd64", "purego"])
def numerics(request, ctx):
    ctx.set_numerics(nl.NUMERICS_AMD64 if request.param else nl.NUMERICS_PUREGO)
    yield request.param
    ctx.set_numerics(nl.NUMERICS_AMD64)


def frame(w, h, seed, specials=False, scale=37.0, offset=900.0):
    rng = np.random.default_rng(seed)
    img = (rng.standard_normal((h, w)) * scale + offset).astype(np.float32)
    img[rng.random((h, w)) < 0.002] += np.float32(5000)
    if specials:
        m = rng.random((h, w))
        img[m < 0.02] = np.float32(0.0)
        img[(m >= 0.02) & (m < 0.04)] = np.float32(-0.0)
        img[(m >= 0.04) & (m < 0.05)] = np.nan
        img[(m >= 0.05) & (m < 0.055)] = np.inf
    return img


def same_bits(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.array_equal(a.view(np.uint32)[~(np.isnan(a) & np.isnan(b))], b.view(np.uint32)[~(np.isnan(a) & np.isnan(b))])


@pytest.mark.parametrize("w,h", [(1, 1), (2, 5), (3, 3), (7, 4), (8, 3), (9, 9), (61, 23), (64, 64), (257, 130), (1500, 700)])
@pytest.mark.parametrize("specials", [False, True])
def test_median_filter3x3(ctx, numerics, w, h, specials):
    img = frame(w, h, w * 31 + h, specials)
    got = nl.median_filter3x3(ctx, img, w)
    want = O.median_filter3x3(img, w, amd64=numerics)
    assert same_bits(got, want), np.nonzero(got.view(np.uint32) != want.view(np.uint32))[0][:8]


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8, 127, 128, 4095, 4096, 4097, 5000, 1 << 16, (1 << 20) + 4, 3000 * 2000])
def test_stats_min_mean_max_stddev(ctx, numerics, n):
    data = frame(n, 1, n % 1000 + 5).ravel()
    got = nl.stats(ctx, data)
    want = O.stats(data, amd64=numerics)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (n, got, want)


def test_stats_signed_zero_and_nan_extremes(ctx, numerics):
    """min/max follow the operand roles of VMINPS/VMAXPS (amd64) or the Go comparisons: the sign of a zero
    extreme and what a NaN does to the running extreme are part of the result"""
    rng = np.random.default_rng(3)
    for trial in range(40):
        n = int(rng.integers(1, 300)) * 4
        data = np.abs(rng.standard_normal(n)).astype(np.float32) * (1 if trial % 2 else -1)
        k = rng.integers(0, n, size=int(rng.integers(1, 12)))
        data[k] = np.where(rng.random(k.size) < 0.5, np.float32(0.0), np.float32(-0.0))
        if trial % 3 == 0:
            data[rng.integers(0, n, size=3)] = np.nan
        got, want = nl.stats(ctx, data), O.stats(data, amd64=numerics)
        assert same_bits(got[[0, 2]], want[[0, 2]]), (trial, got, want)
        assert np.isnan(got[[0, 2]]).tolist() == np.isnan(want[[0, 2]]).tolist()


def _steered_to_a_mean_boundary(rng, n, tiny):
    """data whose float64 sum sits (within the rounding of one element) on the midpoint between two float32 means"""
    data = (rng.standard_normal(n) * 3 + 50).astype(np.float32)
    if tiny:
        k = rng.integers(1, n, size=n // 50)
        data[k] = (rng.standard_normal(k.size) * 1e-9).astype(np.float32)      # terms far below the accumulator's ulp
    s = float(np.sum(data.astype(np.float64)))
    m = np.float32(s / n)
    mid = (float(m) + float(np.nextafter(m, np.float32(np.inf)))) / 2
    data[0] = np.float32(float(data[0]) + (mid * n - s))
    return data


def test_stats_on_a_rounding_boundary_proven_without_replay(ctx):
    """The first-level interval cannot decide a sum that sits on a float32 rounding boundary of the mean; when no term
    can round at all (every term a multiple of the accumulator's ulp) the second-level proof settles it exactly."""
    ctx.set_numerics(nl.NUMERICS_AMD64)
    before = ctx.exact_replays()
    for seed in range(6):
        data = _steered_to_a_mean_boundary(np.random.default_rng(100 + seed), 1 << 16, tiny=False)
        got, want = nl.stats(ctx, data), O.stats(data, amd64=True)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (seed, got, want)
    assert ctx.exact_replays() == before


def test_stats_near_rounding_boundaries_take_the_exact_replay(ctx):
    """...and when many terms do round, the chains are replayed in order; the result still equals the oracle's bit for bit"""
    ctx.set_numerics(nl.NUMERICS_AMD64)
    before = ctx.exact_replays()
    for seed in range(6):
        data = _steered_to_a_mean_boundary(np.random.default_rng(200 + seed), 1 << 16, tiny=True)
        got, want = nl.stats(ctx, data), O.stats(data, amd64=True)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (seed, got, want)
    assert ctx.exact_replays() > before


@pytest.mark.parametrize("w,h", [(64, 48), (333, 257), (1500, 1000)])
def test_bad_pixel_map(ctx, numerics, w, h):
    img = frame(w, h, 9 + w)
    bpm, st = nl.bad_pixel_map(ctx, img, w, 3.0, 5.0)
    want_bpm, want_st, _ = O.bad_pixel_map(img, w, 3.0, 5.0, amd64=numerics)
    assert np.array_equal(st.view(np.uint32), want_st.view(np.uint32)), (st, want_st)
    assert np.array_equal(bpm, want_bpm) and len(bpm) > 0


def test_bad_pixel_map_truncated_list_still_counts(ctx):
    w, h = 200, 100
    img = frame(w, h, 77)
    full, _ = nl.bad_pixel_map(ctx, img, w, 1.0, 1.0)
    part, _ = nl.bad_pixel_map(ctx, img, w, 1.0, 1.0, cap=5)       # the wrapper retries with the reported count
    assert np.array_equal(full, part) and len(full) > 5


def test_op_bad_pixel_repairs_in_list_order(ctx, numerics):
    """OpBadPixel.Apply: hot pixel clusters make the in-place, in-order repair observable"""
    w, h = 320, 200
    img = frame(w, h, 21)
    rng = np.random.default_rng(4)
    for _ in range(30):                       # 2x2 hot clusters: a repair sees the repair before it
        y, x = int(rng.integers(1, h - 2)), int(rng.integers(1, w - 2))
        img[y:y + 2, x:x + 2] += np.float32(4000)
    got, n, st = nl.op_bad_pixel(ctx, img, w, 3.0, 5.0)
    want, wn, wst = O.op_bad_pixel(img, w, 3.0, 5.0, amd64=numerics)
    assert n == wn and n >= 100
    assert np.array_equal(st.view(np.uint32), wst.view(np.uint32))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    same, n0, _ = nl.op_bad_pixel(ctx, img, w, 0.0, 5.0)          # a zero sigma switches the operator off
    assert n0 == 0 and np.array_equal(same, img.ravel())


def test_star_detection_threshold_from_device_stats(ctx):
    """medianDiffStats.StdDev() of BadPixelMap is what FindStars takes as its bad-pixel scale (findstars.go:134-169):
    the whole detection input now comes from the device"""
    from test_gpu_project_stars import star_field
    w, h = 400, 300
    img = star_field(w, h, 25, seed=8)
    _, st = nl.bad_pixel_map(ctx, img, w, 3.0, 5.0)
    _, want_st, _ = O.bad_pixel_map(img, w, 3.0, 5.0, amd64=True)
    assert st[3].view(np.uint32) == want_st[3].view(np.uint32)
    loc, scale = np.float32(np.median(img)), np.float32(3.0)
    stars, shifts, hfr = nl.find_stars(ctx, img, w, loc, scale, 10.0, 5.0, 1.4, 12, float(st[3]))
    want = O.find_stars(img, w, loc, scale, 10.0, 5.0, 1.4, 12, float(want_st[3]))
    assert len(stars) == len(want[0]) and len(stars) > 5


def test_resident_frame_pipeline_one_upload(ctx):
    """bad-pixel repair -> noise estimate -> star detection on ONE upload of the frame: the `_dev` entry points keep the
    device copy in step with the host copy (the repaired pixels are scattered into it), results equal the host-call
    sequence and the oracle"""
    import ctypes as C
    from nightlight_b200.binding import check, load_library
    from test_gpu_project_stars import star_field
    lib = load_library()
    w, h = 480, 360
    img = star_field(w, h, 30, seed=31, hot=80).astype(np.float32).ravel()
    host = img.copy()
    dev = ctx.dev_alloc(4 * host.size)
    try:
        ctx.h2d(dev, host)
        st = np.zeros(4, np.float32)
        removed = C.c_int64()
        fp = C.POINTER(C.c_float)
        check(lib.nl_op_bad_pixel_dev(ctx.handle, C.c_void_p(dev), host.ctypes.data_as(C.c_void_p), host.size, w, 3.0, 5.0,
                                      C.byref(removed), st.ctypes.data_as(fp)))
        want_img, want_removed, want_st = O.op_bad_pixel(img, w, 3.0, 5.0, amd64=True)
        assert removed.value == want_removed and removed.value > 20
        assert np.array_equal(host.view(np.uint32), want_img.view(np.uint32))
        back = np.empty_like(host)
        ctx.d2h(back, dev)
        assert np.array_equal(back.view(np.uint32), host.view(np.uint32))          # device copy repaired too
        noise = (C.c_float * 1)()
        check(lib.nl_estimate_noise_dev(ctx.handle, C.c_void_p(dev), 1, host.size, w, h, noise))
        assert np.float32(noise[0]).view(np.uint32) == O.estimate_noise(want_img, w, amd64=True).view(np.uint32)
        loc, scale = np.float32(np.median(want_img)), np.float32(noise[0])
        cap = 4096
        out = np.zeros(cap, dtype=nl.STAR_DTYPE)
        n, sos, hfr = C.c_int32(), C.c_float(), C.c_float()
        check(lib.nl_find_stars_dev(ctx.handle, C.c_void_p(dev), host.ctypes.data_as(C.c_void_p), host.size, w, float(loc), float(scale),
                                    8.0, 5.0, 1.4, 12, float(st[3]), out.ctypes.data_as(C.c_void_p), cap, C.byref(n), C.byref(sos),
                                    C.byref(hfr)))
        want = O.find_stars(want_img, w, loc, scale, 8.0, 5.0, 1.4, 12, float(want_st[3]))
        assert n.value == len(want[0]) and n.value > 5
        for f in ("index", "x", "y", "mass", "hfr"):
            assert np.array_equal(out[:n.value][f], want[0][f]), f
        assert np.float32(hfr.value).view(np.uint32) == want[2].view(np.uint32)
    finally:
        ctx.dev_free(dev)


def test_stats_dev_of_a_frame_that_is_not_16_byte_aligned(ctx, numerics):
    """nl_stats_dev on frame k of a stack job whose pixel count is not a multiple of four (the frame then starts on
    an odd float): the float4 passes must not fault; nl_bad_pixel_map_dev rejects an unaligned scratch image"""
    import ctypes as C
    lib = nl.load_library()
    w, h = 17, 9                                       # 153 pixels per frame
    imgs = [frame(w, h, 500 + k).ravel() for k in range(3)]
    with nl.StackJob(ctx, 3, w * h) as job:
        for k in range(3):
            job.put_frame(k, imgs[k])
        base, stride = job.frames_dev
        for k in range(3):
            st = np.zeros(4, np.float32)
            nl.binding.check(lib.nl_stats_dev(ctx.handle, C.c_void_p(base + 4 * k * stride), w * h, st.ctypes.data_as(C.POINTER(C.c_float))))
            want = O.stats(imgs[k], amd64=numerics)
            assert np.array_equal(st.view(np.uint32), want.view(np.uint32)), k
        tmp = ctx.dev_alloc(4 * (w * h + 8))
        try:
            cnt, st = C.c_int64(), np.zeros(4, np.float32)
            rc = lib.nl_bad_pixel_map_dev(ctx.handle, C.c_void_p(base + 4 * stride), w * h, w, 3.0, 5.0, C.c_void_p(tmp + 4), None, 0,
                                          C.byref(cnt), st.ctypes.data_as(C.POINTER(C.c_float)))
            assert rc == nl.binding.NL_E_INVALID and b"16-byte aligned" in lib.nl_last_error()
            nl.binding.check(lib.nl_bad_pixel_map_dev(ctx.handle, C.c_void_p(base + 4 * stride), w * h, w, 3.0, 5.0, C.c_void_p(tmp), None, 0,
                                                      C.byref(cnt), st.ctypes.data_as(C.POINTER(C.c_float))))
            bpm, want, _ = O.bad_pixel_map(imgs[1], w, 3.0, 5.0, amd64=numerics)
            assert cnt.value == bpm.size and np.array_equal(st.view(np.uint32), want.view(np.uint32))
        finally:
            ctx.dev_free(tmp)
