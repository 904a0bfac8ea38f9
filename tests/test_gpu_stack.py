"""Parity tests proper: the CUDA stacking path, called through the C ABI, against the oracle on the
same seeded inputs (bit-exact results, identical clip counts), the committed golden arrays, the
reference's edge cases, and full-size (BASELINE config) runs checked on sampled tiles."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import nightlight_b200 as nl  # noqa: E402
from oracle import oracle as O  # noqa: E402
from util import GOLDEN, MODE_ID, bits_equal, first_mismatch, from_hex, hx, kats, mode_cases, synth_frames_threaded, weights_for  # noqa: E402

pytestmark = pytest.mark.gpu


def gpu_stack(ctx, frames, mode, sl=2.75, sh=2.75, w=None, ref_loc=0.0):
    frames = np.ascontiguousarray(frames, dtype=np.float32)
    n, p = frames.shape
    with nl.StackJob(ctx, n, p) as job:
        for i in range(n):
            job.put_frame(i, frames[i])
        return job.run(MODE_ID[mode] if isinstance(mode, str) else mode, w, sl, sh, ref_loc)


def check_against_oracle(ctx, frames, mode, weighted, sl=2.75, sh=2.75, ref_loc=0.0, w=None):
    if weighted and w is None:
        w = weights_for(frames.shape[0])
    want = O.stack(frames, mode, sl, sh, weights=w if weighted else None, ref_loc=ref_loc)
    got = gpu_stack(ctx, frames, mode, sl, sh, w if weighted else None, ref_loc)
    assert bits_equal(got[0], want[0]), (mode, weighted, first_mismatch(got[0], want[0]))
    assert got[1:] == want[1:], (mode, weighted, got[1:], want[1:])


def test_column_kats(ctx):
    for k in kats()["columns"]:
        col = np.array([[np.nan if v == "nan" else v for v in k["col"]]], dtype=np.float32).T
        res, cl, ch = gpu_stack(ctx, col, k["mode"], k["sig"], k["sig"])
        assert hx(res[0]) == k["hex"], k
        assert [cl, ch] == k["clip"], k


def test_golden_arrays(ctx):
    g = np.load(os.path.join(GOLDEN, "stack_golden.npz"))
    for n, p0, count in ((16, 0, 4096), (37, 777, 1031), (256, 4096 * 4096 - 512, 512)):
        frames = O.synth_frames(n, p0, count)
        w = weights_for(n)
        for mode, weighted in mode_cases():
            key = "n%d_p%d_c%d_%s%s" % (n, p0, count, mode, "_w" if weighted else "")
            res, cl, ch = gpu_stack(ctx, frames, mode, w=w if weighted else None)
            assert bits_equal(res, g[key]), (key, first_mismatch(res, g[key]))
            assert [cl, ch] == list(g[key + "_clip"]), key


@pytest.mark.parametrize("n", [1, 2, 3, 5, 6, 15, 16, 25, 64, 100, 256])
def test_all_modes_match_oracle(ctx, n):
    pixels = 5000 if n <= 64 else 2100           # ragged: not a multiple of 32 or of 4
    frames = O.synth_frames(n, 31 * n, pixels)
    for mode, weighted in mode_cases():
        check_against_oracle(ctx, frames, mode, weighted)


def test_config1_shape_sigma(ctx):
    """BASELINE config 1: 16 frames of 1024x1024, sigma-clip 2.75/2.75"""
    frames = O.synth_frames(16, 0, 1024 * 1024)
    check_against_oracle(ctx, frames, "sigma", False)
    check_against_oracle(ctx, frames, "auto", False)      # 16 frames -> winsorized sigma


def test_random_data_sigmas_and_edge_pixels(ctx):
    rng = np.random.default_rng(11)
    for n in (6, 30, 100):
        frames = (rng.standard_normal((n, 777)) * 50 + 1000).astype(np.float32)
        frames[rng.random(frames.shape) < 0.02] = np.nan
        frames[rng.random(frames.shape) < 0.03] += 2000
        frames[:, 5] = np.nan                       # all-NaN pixel -> RefFrameLoc
        frames[:, 6] = 7.0                          # constant pixel, sigma 0
        frames[1:, 8] = np.nan                      # a single valid sample
        w = rng.random(n).astype(np.float32) + np.float32(0.1)
        for sl, sh in ((2.75, 2.75), (1.0, 3.0), (-1.0, -1.0), (0.5, 0.5)):
            for mode, weighted in mode_cases():
                check_against_oracle(ctx, frames, mode, weighted, sl, sh, ref_loc=0.25, w=w)


def test_ties_and_integers(ctx):
    rng = np.random.default_rng(12)
    frames = rng.integers(0, 5, (64, 1000)).astype(np.float32)
    for mode, weighted in mode_cases():
        check_against_oracle(ctx, frames, mode, weighted)


@pytest.mark.parametrize("n", [7, 16, 40, 256])
def test_signed_zero_ties_every_mode(ctx, n):
    """columns made of -1, -0.0, +0.0, 1 with both zero signs straddling the median: the sign of a zero result
    depends on the reference's permutation and on `>` (not fmax) keeping the first of equal values"""
    rng = np.random.default_rng(100 + n)
    frames = np.round(rng.standard_normal((n, 900)) * 0.6).astype(np.float32)
    assert (np.signbit(frames) & (frames == 0)).any() and ((~np.signbit(frames)) & (frames == 0)).any()
    frames[rng.random(frames.shape) < 0.02] = np.nan
    for mode, weighted in mode_cases():
        for sl, sh in ((2.75, 2.75), (0.5, 1.0)):
            check_against_oracle(ctx, frames, mode, weighted, sl, sh)


def test_many_frames_small_tile_path(ctx):
    """n_frames beyond what fits 32 pixel columns per warp in shared memory (narrower tile kernels)"""
    frames = O.synth_frames(2000, 99, 70)
    for mode, weighted in (("sigma", False), ("median", False), ("mean", True), ("winsor", True), ("linfit", False)):
        check_against_oracle(ctx, frames, mode, weighted)


def test_512_and_1024_frames_medium_tile_paths(ctx):
    """512 frames -> 16-pixel tiles, 1024 frames (BASELINE config 4's depth) -> 8-pixel tiles"""
    for n, pixels in ((512, 200), (1024, 90)):
        frames = O.synth_frames(n, 7 * n, pixels)
        for mode, weighted in (("sigma", False), ("linfit", False), ("winsor", True), ("median", False)):
            check_against_oracle(ctx, frames, mode, weighted)


def test_empty_and_errors(ctx):
    with nl.StackJob(ctx, 3, 0) as job:
        res, cl, ch = job.run(nl.ST_SIGMA)
        assert res.size == 0 and (cl, ch) == (0, 0)
    frames = O.synth_frames(4, 0, 64)
    with pytest.raises(nl.NightlightError) as e:
        gpu_stack(ctx, frames, 7)
    assert "invalid stacking mode" in str(e.value)
    with pytest.raises(nl.NightlightError) as e:
        gpu_stack(ctx, frames, -1)
    assert "invalid stacking mode" in str(e.value)
    with pytest.raises(nl.NightlightError) as e:
        gpu_stack(ctx, frames, "mad", w=weights_for(4))
    assert "MADSigma stacking with weights" in str(e.value)
    with pytest.raises(nl.NightlightError):
        nl.StackJob(ctx, 0, 10)


def test_opstack_operator(ctx):
    """OpStack.Apply semantics: auto mode, weighting from per-frame scalars, exposure sum, clip totals"""
    n, w, h = 20, 96, 50
    data = O.synth_frames(n, 4242, w * h)
    frames = [nl.ops.Image(data=data[i], naxisn=(w, h), exposure=30.0 + i, noise=1.0 + 0.1 * (i % 5), hfr=2.0 + 0.05 * i, id=i)
              for i in range(n)]
    for weighting in (nl.W_NONE, nl.W_EXPOSURE, nl.W_INVERSE_NOISE, nl.W_INVERSE_HFR):
        op = nl.OpStack(mode=nl.ST_AUTO, weighting=weighting, sigmaLow=2.5, sigmaHigh=3.0)
        out = op.apply(frames, ctx)
        wts = nl.get_weights(frames, weighting)
        want, cl, ch = O.stack(data, "auto", 2.5, 3.0, weights=wts)
        assert bits_equal(out.data, want), (weighting, first_mismatch(out.data, want))
        assert (out.clip_low, out.clip_high) == (cl, ch)
        assert out.exposure == float(np.sum([f.exposure for f in frames], dtype=np.float32))


def test_stack_batches(ctx):
    """stack of stacks (stackbatches.go:84-116): frame-count weighted mean of per-batch stacks"""
    n, pixels = 30, 3000
    data = O.synth_frames(n, 17, pixels)
    batches = [list(range(0, 12)), list(range(12, 24)), list(range(24, 30))]
    imgs = [[nl.ops.Image(data=data[i], naxisn=(pixels, 1), exposure=10.0) for i in b] for b in batches]
    out = nl.OpStackBatches(perBatch=nl.OpStack(mode=nl.ST_SIGMA)).apply(imgs, ctx)
    acc = None
    for b in batches:
        res, _, _ = O.stack(data[b], "sigma")
        wgt = np.float32(len(b))
        acc = res * wgt if acc is None else acc + res * wgt
    want = acc * (np.float32(1.0) / np.float32(n))
    assert bits_equal(out.data, want), first_mismatch(out.data, want)


def test_stripes_equal_whole(ctx):
    """row-stripe sharding: stacking each stripe separately and concatenating == stacking the image"""
    from nightlight_b200.stripes import all_stripes
    n, w, h = 24, 130, 37
    data = O.synth_frames(n, 0, w * h)
    whole = gpu_stack(ctx, data, "sigma")
    parts, cl, ch = [], 0, 0
    for row0, rows in all_stripes(h, 8):
        r = gpu_stack(ctx, data[:, row0 * w:(row0 + rows) * w], "sigma")
        parts.append(r[0]); cl += r[1]; ch += r[2]
    assert bits_equal(np.concatenate(parts), whole[0]) and (cl, ch) == whole[1:]


@pytest.mark.parametrize("mode,weighted", [("sigma", False), ("winsor", True)])
def test_full_image_every_pixel_c2(ctx, mode, weighted):
    """BASELINE configs[1] at size: 256 x 4096x4096 fp32 (16 GiB resident), the metric's sigma-clip mode and the
    config's own winsorized + weighted mode.  The frames are generated twice -- on the device (nl_synth_fill_dev)
    and on the host (the oracle's generator) -- and EVERY one of the 16.7 M stacked pixels and the exact clip totals
    must equal the CPU restatement of the reference (all host threads, the reference's 8 MiB work packages)."""
    n, pixels = 256, 4096 * 4096
    w = weights_for(n) if weighted else None
    with nl.StackJob(ctx, n, pixels) as job:
        job.synth_fill(0)
        res, cl, ch = job.run(MODE_ID[mode], w)
    frames = synth_frames_threaded(n, 0, pixels)
    want, wl, wh = O.stack(frames, mode, weights=w, threads=os.cpu_count() or 1)
    del frames
    assert bits_equal(res, want), first_mismatch(res, want)
    assert (cl, ch) == (wl, wh)
    assert not np.isnan(res).any()


@pytest.mark.parametrize("mode,weighted,n,pixels", [
    ("linfit", False, 64, 6000 * 500),
    ("median", False, 256, 4096 * 512),
    ("linfit", False, 1024, 8192 * 64),        # configs[3] depth and row width: 1024 frames, 8192-pixel rows (8-pixel tiles)
    ("sigma", False, 1024, 8192 * 64),
])
def test_full_size_sampled_tiles(ctx, mode, weighted, n, pixels):
    """Frames generated on the device at full size; the oracle regenerates sampled tiles from the same
    position-addressable hash and must agree bit for bit, and the clip totals must be plausible."""
    w = weights_for(n) if weighted else None
    with nl.StackJob(ctx, n, pixels) as job:
        job.synth_fill(0)
        res, cl, ch = job.run(MODE_ID[mode], w)
    rng = np.random.default_rng(2024)
    span = 4096 if n <= 256 else 512
    starts = [0, pixels - span] + [int(s) for s in rng.integers(0, pixels - span, 6)]
    for s in starts:
        frames = O.synth_frames(n, s, span)
        want, _, _ = O.stack(frames, mode, weights=w)
        assert bits_equal(res[s:s + span], want), (s, first_mismatch(res[s:s + span], want))
    assert not np.isnan(res).any()
    if mode != "median":
        frac = (cl + ch) / (n * pixels)
        assert 0.005 < frac < 0.5, frac
    # size-independent property: every stacked value lies within the range of the synthetic samples
    assert res.min() >= 1024 - 256 - 512 and res.max() <= 1024 + 256 + 4096


def test_goal_seek_matches_oracle_driven_search(ctx):
    """a21: the restated sigma goal-seek around a resident StackJob follows the same trajectory as the
    same host logic driven by the oracle (identical clip counts at every step)"""
    frames = O.synth_frames(40, 5000, 4000)
    n, p = frames.shape
    with nl.StackJob(ctx, n, p) as job:
        for i in range(n):
            job.put_frame(i, frames[i])
        for mode, name, lo, hi in ((nl.ST_SIGMA, "sigma", 1.0, 1.5), (nl.ST_WINSOR_SIGMA, "winsor", 0.5, 2.0),
                                   (nl.ST_AUTO, "linfit", 2.0, 2.0)):
            got = nl.find_sigmas_and_stack(lambda sl, sh: job.run(mode, None, sl, sh), mode, n, p, lo, hi)
            want = nl.find_sigmas_and_stack(lambda sl, sh: O.stack(frames, name, sl, sh), mode, n, p, lo, hi)
            assert got[1:] == want[1:], (name, got[1:], want[1:])
            assert bits_equal(got[0], want[0]), (name, first_mismatch(got[0], want[0]))


def test_bcast_epilogue_stores_every_copy(ctx):
    """nl_stack_run_dev_bcast: the result also lands in every extra (peer) buffer -- the fused
    reassembly epilogue, here with two more buffers on the same device"""
    frames = O.synth_frames(20, 77, 4100)           # 4100 % 4 == 0 -> float4 path of the mean kernel too
    n, p = frames.shape
    bufs = [ctx.dev_alloc(4 * p) for _ in range(3)]
    try:
        with nl.StackJob(ctx, n, p) as job:
            for i in range(n):
                job.put_frame(i, frames[i])
            for mode, name in ((nl.ST_SIGMA, "sigma"), (nl.ST_MEAN, "mean"), (nl.ST_LINEAR_FIT, "linfit")):
                job.run_dev_bcast(mode, bufs[0], bufs[1:])
                ctx.sync()
                want = O.stack(frames, name)
                for b in bufs:
                    got = np.empty(p, np.float32)
                    ctx.d2h(got, b)
                    assert bits_equal(got, want[0]), name
                assert job.clip_counts() == want[1:]
    finally:
        for b in bufs:
            ctx.dev_free(b)


def test_infinities_denormals_and_huge_values(ctx):
    """IEEE corner values go through the same comparisons and fp32 sums as in the reference: +-inf
    samples, denormals (no flush-to-zero), magnitudes whose squares overflow"""
    rng = np.random.default_rng(21)
    n, p = 40, 600
    frames = (rng.standard_normal((n, p)) * 10 + 100).astype(np.float32)
    frames[3, 0:50] = np.inf
    frames[7, 25:80] = -np.inf
    frames[:, 100:150] = (rng.standard_normal((n, 50)) * 1e-41).astype(np.float32)      # denormals
    frames[:, 150:200] = (rng.standard_normal((n, 50)) * 1e30).astype(np.float32)       # squares overflow to inf
    frames[5, 200:220] = np.float32(3e38)
    frames[:, 220:240] = -0.0
    frames[::2, 230:240] = 0.0
    for mode, weighted in mode_cases():
        check_against_oracle(ctx, frames, mode, weighted)


def test_stack_apply_one_call_striped(ctx):
    """nl_stack_apply: all host frame pointers in one call, stripes pipelined inside the library"""
    import ctypes as C
    lib = nl.load_library()
    w, h, n = 64, 37, 20                           # 37 rows: ragged stripes
    frames = O.synth_frames(n, 4242, w * h)
    ptrs = (C.c_void_p * n)(*[frames[i].ctypes.data for i in range(n)])
    wts = weights_for(n)
    for stripes in (1, 3, 8, 100):
        for mode, name, wv in ((nl.ST_SIGMA, "sigma", None), (nl.ST_WINSOR_SIGMA, "winsor", wts), (nl.ST_AUTO, "winsor", None)):
            out = np.empty(w * h, np.float32)
            cl, ch = C.c_int64(), C.c_int64()
            wp = wv.ctypes.data_as(C.POINTER(C.c_float)) if wv is not None else None
            nl.binding.check(lib.nl_stack_apply(ctx.handle, ptrs, n, w * h, w, stripes, mode, wp, 2.75, 2.75, 0.0,
                                                out.ctypes.data_as(C.c_void_p), C.byref(cl), C.byref(ch)))
            want = O.stack(frames, name, weights=wv)
            assert bits_equal(out, want[0]), (stripes, name)
            assert (cl.value, ch.value) == want[1:]
    out = np.empty(w * h, np.float32)
    rc = lib.nl_stack_apply(ctx.handle, ptrs, n, w * h, w, 2, 9, None, 2.75, 2.75, 0.0, out.ctypes.data_as(C.c_void_p), None, None)
    assert rc == nl.binding.NL_E_INVALID and b"invalid stacking mode" in lib.nl_last_error()


@pytest.mark.parametrize("seed", [20261017, 1, 2, 3, 112, 124])     # 112, 124: a linear fit that rejects every sample at a deferral boundary
def test_random_shapes_and_modes_fuzz(ctx, seed):
    """seeded fuzz over frame counts, pixel counts (ragged tiles, unaligned rows -> both staging paths),
    NaN densities, outliers, sigmas and modes: every result bit-identical to the oracle"""
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
            frames = np.round(frames).astype(np.float32)          # ties
        sl, sh = (float(x) for x in rng.choice([0.5, 1.0, 2.0, 2.75, 4.0, -1.0], 2))
        mode, weighted = cases[int(rng.integers(0, len(cases)))]
        w = (rng.random(n).astype(np.float32) + np.float32(0.05)) if weighted else None
        check_against_oracle(ctx, frames, mode, weighted, sl, sh, ref_loc=float(rng.choice([0.0, 7.5])), w=w)


@pytest.mark.parametrize("defer", ["1", "2", "3", "0", "1,2,4", "2,3,5,8,13"])
@pytest.mark.parametrize("mode,weighted", [("sigma", False), ("sigma", True), ("winsor", False), ("winsor", True), ("linfit", False)])
def test_deferred_clipping_passes_pool_and_overflow(ctx, tuning, mode, weighted, defer):
    """Late clipping passes are deferred to a pool of unfinished columns and finished by a second launch; the pool
    holds a quarter of the pixels, the overflow finishes in place.  Heavy-tailed samples make most pixels need
    many passes, so with an early limit the pool overflows; every variant must equal the oracle bit for bit."""
    tuning("defer_passes", defer)
    rng = np.random.default_rng(len(defer) * 7 + len(mode) + weighted)
    n, p = 96, 32 * 61 + 5                              # 32-pixel tiles and a ragged last tile
    frames = (rng.standard_t(2.0, size=(n, p)) * 25 + 500).astype(np.float32)
    frames[rng.random((n, p)) < 0.01] = np.nan
    frames[:, 7] = np.nan                               # an empty column
    frames[1:, 11] = np.nan                             # a single sample
    check_against_oracle(ctx, frames, mode, weighted, 2.0, 2.5, ref_loc=123.0)
    check_against_oracle(ctx, frames[:40], mode, weighted)


@pytest.mark.parametrize("defer", ["1", "3", "2,4,7"])
@pytest.mark.parametrize("n", [520, 1030])
def test_deferred_passes_on_narrow_tiles(ctx, tuning, n, defer):
    """more than 256 frames: 16- and 8-pixel tiles; the pools then hold tiles of that width"""
    tuning("defer_passes", defer)
    rng = np.random.default_rng(n + len(defer))
    p = 16 * 7 + 5
    frames = (rng.standard_t(2.0, size=(n, p)) * 25 + 500).astype(np.float32)
    frames[rng.random((n, p)) < 0.01] = np.nan
    frames[:, 3] = np.nan
    for mode, weighted in (("sigma", False), ("winsor", True), ("linfit", False)):
        check_against_oracle(ctx, frames, mode, weighted, 2.0, 2.5, ref_loc=5.0)


def test_default_deferral_schedules_on_the_synthetic_workload(ctx):
    """the schedules the library uses by default (sigma {3}, winsor {2}, linear fit {8,...,30}), on generator data"""
    frames = O.synth_frames(128, 4096 * 100, 32 * 300)
    for mode, weighted in (("sigma", False), ("winsor", True), ("linfit", False)):
        check_against_oracle(ctx, frames, mode, weighted)


@pytest.mark.parametrize("pixels,offset", [(1, 0), (3, 0), (4, 0), (1000003, 0), (4099, 1), (65536, 3)])
def test_incremental_kernels_vector_and_scalar_paths(ctx, pixels, offset):
    """StackIncremental / StackIncrementalFinalize (stack.go:924-944): acc = light*w, acc += light*w (mul then add),
    acc *= 1/weightSum -- float4 path for 16-byte aligned buffers, scalar path otherwise, ragged tails"""
    import ctypes as C
    from nightlight_b200.binding import check, load_library
    lib = load_library()
    rng = np.random.default_rng(pixels + offset)
    a = (rng.standard_normal(pixels) * 50 + 10).astype(np.float32)
    b = (rng.standard_normal(pixels) * 50 + 10).astype(np.float32)
    dev = ctx.dev_alloc(4 * (2 * pixels + 64))
    try:
        acc, light = dev + 4 * offset, dev + 4 * (pixels + 16 + offset) + (0 if offset == 0 else 4 * ((4 - (pixels + offset) % 4) % 4))
        ctx.h2d(light, a)
        check(lib.nl_stack_incremental_dev(ctx.handle, C.c_void_p(acc), C.c_void_p(light), pixels, 3.0, 1))
        ctx.h2d(light, b)
        check(lib.nl_stack_incremental_dev(ctx.handle, C.c_void_p(acc), C.c_void_p(light), pixels, 5.0, 0))
        check(lib.nl_stack_incremental_finalize_dev(ctx.handle, C.c_void_p(acc), pixels, 8.0))
        got = np.empty(pixels, np.float32)
        ctx.d2h(got, acc)
        want = (a * np.float32(3.0)).astype(np.float32)
        want = (want + (b * np.float32(5.0)).astype(np.float32)).astype(np.float32)
        want = (want * (np.float32(1.0) / np.float32(8.0))).astype(np.float32)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    finally:
        ctx.dev_free(dev)


def test_stack_apply_multi_contexts_one_image(ctx):
    """nl_stack_apply_multi: all host frame pointers in, ONE host image out, the rows dealt to several contexts
    (here on the devices this box has, several contexts per device when it has only one) -- ragged row blocks,
    ragged stripes inside a block, more contexts than rows"""
    import ctypes as C
    lib = nl.load_library()
    cnt = C.c_int()
    nl.binding.check(lib.nl_device_count(C.byref(cnt)))
    w, h, n = 64, 37, 20
    frames = O.synth_frames(n, 777, w * h)
    ptrs = (C.c_void_p * n)(*[frames[i].ctypes.data for i in range(n)])
    wts = weights_for(n)
    for n_ctx in (1, 2, 3, 5):
        ctxs = [nl.Context(g % cnt.value) for g in range(n_ctx)]
        try:
            arr = (C.c_void_p * n_ctx)(*[c.handle for c in ctxs])
            for stripes in (1, 3, 50):
                for mode, name, wv in ((nl.ST_SIGMA, "sigma", None), (nl.ST_WINSOR_SIGMA, "winsor", wts), (nl.ST_MEAN, "mean", None),
                                       (nl.ST_LINEAR_FIT, "linfit", None)):
                    out = np.full(w * h, -1.0, np.float32)
                    cl, ch = C.c_int64(), C.c_int64()
                    wp = wv.ctypes.data_as(C.POINTER(C.c_float)) if wv is not None else None
                    nl.binding.check(lib.nl_stack_apply_multi(arr, n_ctx, ptrs, n, w * h, w, stripes, mode, wp, 2.75, 2.75, 0.0,
                                                              out.ctypes.data_as(C.c_void_p), C.byref(cl), C.byref(ch)))
                    want = O.stack(frames, name, weights=wv)
                    assert bits_equal(out, want[0]), (n_ctx, stripes, name, first_mismatch(out, want[0]))
                    assert (cl.value, ch.value) == want[1:]
        finally:
            for c in ctxs:
                c.close()
    # h = 2 rows over 5 contexts: three contexts get nothing
    ctxs = [nl.Context(0) for _ in range(5)]
    try:
        arr = (C.c_void_p * 5)(*[c.handle for c in ctxs])
        out = np.empty(2 * w, np.float32)
        nl.binding.check(lib.nl_stack_apply_multi(arr, 5, ptrs, n, 2 * w, w, 8, nl.ST_SIGMA, None, 2.75, 2.75, 0.0,
                                                  out.ctypes.data_as(C.c_void_p), None, None))
        assert bits_equal(out, O.stack(frames[:, :2 * w], "sigma")[0])
        rc = lib.nl_stack_apply_multi(arr, 5, ptrs, n, 2 * w, w, 8, 11, None, 2.75, 2.75, 0.0, out.ctypes.data_as(C.c_void_p), None, None)
        assert rc == nl.binding.NL_E_INVALID and b"invalid stacking mode" in lib.nl_last_error()
    finally:
        for c in ctxs:
            c.close()


def test_stack_apply_lanes_are_sized_once(ctx):
    """ragged stripes (rows % n_stripes != 0) run in lanes sized for the largest stripe: repeated calls allocate nothing"""
    import ctypes as C
    lib = nl.load_library()
    w, h, n = 128, 251, 12                         # 251 rows in 2 stripes: 126 + 125
    frames = O.synth_frames(n, 99, w * h)
    ptrs = (C.c_void_p * n)(*[frames[i].ctypes.data for i in range(n)])
    want = O.stack(frames, "sigma")
    with nl.Context(0) as c:
        free = []
        for rep in range(4):
            out = np.empty(w * h, np.float32)
            cl, ch = C.c_int64(), C.c_int64()
            nl.binding.check(lib.nl_stack_apply(c.handle, ptrs, n, w * h, w, 2, nl.ST_SIGMA, None, 2.75, 2.75, 0.0,
                                                out.ctypes.data_as(C.c_void_p), C.byref(cl), C.byref(ch)))
            assert bits_equal(out, want[0]) and (cl.value, ch.value) == want[1:]
            free.append(c.mem_info()[0])
        assert free[1] == free[2] == free[3], free


def test_count_only_runs_and_native_goal_seek(ctx):
    """a21 in the C ABI: nl_stack_clip_counts_only gives the clip totals of a full run without writing an image, and
    nl_find_sigmas_and_stack (count-only trials + one stack) ends where the float32 restatement of the reference's
    commented-out FindSigmasAndStack ends when it is driven by the oracle"""
    frames = O.synth_frames(40, 5000, 4000)
    n, p = frames.shape
    wts = weights_for(n)
    with nl.StackJob(ctx, n, p) as job:
        for i in range(n):
            job.put_frame(i, frames[i])
        for mode, name, wv in ((nl.ST_SIGMA, "sigma", None), (nl.ST_WINSOR_SIGMA, "winsor", wts), (nl.ST_LINEAR_FIT, "linfit", None),
                               (nl.ST_MAD_SIGMA, "mad", None), (nl.ST_MEAN, "mean", wts), (nl.ST_MEDIAN, "median", None)):
            want = O.stack(frames, name, 2.0, 2.5, weights=wv)
            assert job.clip_counts_only(mode, wv, 2.0, 2.5) == want[1:], name
        for mode, name, lo, hi in ((nl.ST_SIGMA, "sigma", 1.0, 1.5), (nl.ST_WINSOR_SIGMA, "winsor", 0.5, 2.0),
                                   (nl.ST_AUTO, "linfit", 2.0, 2.0), (nl.ST_MEDIAN, "median", 1.0, 1.0)):
            got = job.find_sigmas_and_stack(mode, lo, hi)
            calls = []

            def oracle_stack(sl, sh):
                calls.append((sl, sh))
                return O.stack(frames, name, sl, sh)

            want = nl.find_sigmas_and_stack(oracle_stack, mode, n, p, lo, hi)
            assert (np.float32(got[3]), np.float32(got[4])) == (np.float32(want[3]), np.float32(want[4])), (name, got[3:], want[3:])
            assert got[1:3] == want[1:3], name
            assert bits_equal(got[0], want[0]), (name, first_mismatch(got[0], want[0]))
            if name != "median":
                assert got[5] == len(calls)


@pytest.mark.parametrize("mode,weighted", [("mean", False), ("mean", True), ("sigma", False)])
def test_result_pointers_that_are_not_16_byte_aligned(ctx, mode, weighted):
    """a job of pixels % 4 == 0 whose result (or a peer copy) starts at an odd float: the float4 stores of the mean
    kernel would fault there, so the launcher must take the scalar path"""
    frames = O.synth_frames(12, 31, 4096)
    n, p = frames.shape
    w = weights_for(n) if weighted else None
    want = O.stack(frames, mode, weights=w)
    buf = ctx.dev_alloc(4 * (3 * p + 64))
    try:
        with nl.StackJob(ctx, n, p) as job:
            for i in range(n):
                job.put_frame(i, frames[i])
            for off_out, off_peer in ((1, 0), (0, 3), (2, 1)):
                out, peer = buf + 4 * off_out, buf + 4 * (p + 16 + off_peer)
                job.run_dev_bcast(MODE_ID[mode], out, [peer], w)
                ctx.sync()
                for ptr in (out, peer):
                    got = np.empty(p, np.float32)
                    ctx.d2h(got, ptr)
                    assert bits_equal(got, want[0]), (off_out, off_peer)
    finally:
        ctx.dev_free(buf)


@pytest.mark.parametrize("stream", ["1", "0"])
@pytest.mark.parametrize("n,p", [(300, 8 * 40 + 3), (1024, 8 * 70 + 1), (2100, 37)])
def test_linear_fit_long_columns_streaming_rounds_and_in_place(ctx, tuning, stream, n, p):
    """more than 256 frames: the linear fit sorts every column in the column kernel and streams the rejection rounds
    from a pool (linfit_rounds_kernel, survivors as bit masks) -- or, with the tuning switched off, runs them in
    place; both equal the oracle bit for bit, with heavy tails (many rounds), NaNs, an empty and a one-sample column"""
    tuning("linfit_stream", stream)
    rng = np.random.default_rng(n + p)
    frames = (rng.standard_t(2.5, size=(n, p)) * 30 + 700).astype(np.float32)
    frames[rng.random((n, p)) < 0.01] = np.nan
    frames[:, 2] = np.nan
    frames[1:, 5] = np.nan
    frames[:, 9] = 3.25                                 # constant column: sigma 0
    check_against_oracle(ctx, frames, "linfit", False, 2.0, 2.5, ref_loc=9.0)
    check_against_oracle(ctx, frames, "linfit", False, 0.8, 0.8)
    frames2 = O.synth_frames(n, 31 * n, p)
    check_against_oracle(ctx, frames2, "linfit", False)


@pytest.mark.parametrize("n", [16, 48, 64, 96, 97])
def test_short_columns_built_in_regrouping_on_jobs_large_enough_to_use_it(ctx, n):
    """columns of at most 96 frames are regrouped after every pass (sigma) / earlier and more often (linear fit) -- but only
    on jobs of at least 32 x 148 x 8 tiles, more than the other tests stack: every pixel and the clip totals of a
    1.2 M pixel job against the oracle, with the built-in schedules"""
    p = 37888 * 32 + 4096 + 17
    frames = O.synth_frames(n, 12345 * n, p)
    for mode, weighted in (("sigma", False), ("sigma", True), ("winsor", False), ("linfit", False)):
        check_against_oracle(ctx, frames, mode, weighted)
        check_against_oracle(ctx, frames, mode, weighted, 1.0, 2.0)
