"""Parity of the resample kernel and of star detection against the oracle, through the C ABI."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import nightlight_b200 as nl  # noqa: E402
from oracle import oracle as O  # noqa: E402
from util import bits_equal, first_mismatch, from_hex, kats  # noqa: E402

pytestmark = pytest.mark.gpu


def star_field(w, h, nstars, seed, noise=3.0, hot=20):
    rng = np.random.default_rng(seed)
    img = (rng.standard_normal((h, w)) * noise + 100).astype(np.float32)
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(nstars):
        x, y = rng.uniform(0, w), rng.uniform(0, h)
        amp, s = rng.uniform(50, 3000), rng.uniform(1.0, 3.0)
        x0, x1, y0, y1 = int(max(0, x - 15)), int(min(w, x + 16)), int(max(0, y - 15)), int(min(h, y + 16))
        sub = (slice(y0, y1), slice(x0, x1))
        img[sub] += (amp * np.exp(-((xx[sub] - x) ** 2 + (yy[sub] - y) ** 2) / (2 * s * s))).astype(np.float32)
    for _ in range(hot):                                   # single hot pixels
        img[rng.integers(0, h), rng.integers(0, w)] += 5000
    return img.reshape(-1)


def test_project_kat(ctx):
    k = kats()["project"]
    trans = np.array([from_hex(h) for h in k["trans_hex"]], dtype=np.float32)
    src = np.array([10 * r + c for r in range(4) for c in range(4)], dtype=np.float32)
    out = nl.project(ctx, src, 4, 4, 4, 4, trans).reshape(4, 4)
    want = np.array([[from_hex(h) for h in row] for row in k["rows_hex"]], dtype=np.float32)
    assert bits_equal(out, want), first_mismatch(out, want)


@pytest.mark.parametrize("sw,sh,dw,dh", [(640, 480, 640, 480), (333, 257, 400, 300), (1, 1, 5, 5), (2, 2, 3, 3), (6000, 4000, 6000, 4000)])
def test_project_matches_oracle(ctx, sw, sh, dw, dh):
    rng = np.random.default_rng(sw * 7 + dh)
    src = (rng.standard_normal(sw * sh) * 100 + 1000).astype(np.float32)
    th = np.deg2rad(rng.uniform(-3, 3))
    for trans in ([1, 0, 0, 0, 1, 0],
                  [np.cos(th), -np.sin(th), rng.uniform(-20, 20), np.sin(th), np.cos(th), rng.uniform(-20, 20)],
                  [1.01, 0.02, -5.5, -0.015, 0.99, 7.25],
                  [0.5, 0, 0.25, 0, 0.5, 0.75]):
        trans = np.array(trans, dtype=np.float32)
        for oob in (np.float32(np.nan), np.float32(0.0)):
            got = nl.project(ctx, src, sw, sh, dw, dh, trans, oob)
            want = O.project(src, sw, sh, dw, dh, trans, oob)
            assert bits_equal(got, want), (list(trans), first_mismatch(got, want))


def test_project_singular(ctx):
    with pytest.raises(nl.NightlightError) as e:
        nl.project(ctx, np.zeros(16, np.float32), 4, 4, 4, 4, [1, 2, 0, 2, 4, 0])
    assert "Matrix has no inverse" in str(e.value)


@pytest.mark.parametrize("w,h,radius", [(640, 480, 16), (1000, 37, 4), (257, 300, 0), (64, 3, 100), (6000, 4000, 16)])
def test_find_bright_matches_oracle(ctx, w, h, radius):
    img = star_field(w, h, max(5, w * h // 20000), seed=w + h)
    for thr in (110.0, 150.0, 1e9, -1e9 if w * h < 100000 else 105.0):
        got = nl.find_bright_pixels(ctx, img, w, thr, radius)
        want = O.find_bright_pixels(img, w, thr, radius)
        assert len(got) == len(want), (thr, len(got), len(want))
        assert got.tobytes() == want.tobytes(), thr


def test_find_bright_plateaus_and_ties(ctx):
    """equal values keep the older candidate; a brighter one replaces it and moves the window"""
    w, h = 300, 4
    img = np.zeros(w * h, np.float32)
    img[10:40] = 5.0                  # plateau
    img[100:130] = np.arange(30)      # rising ramp: every pixel replaces the previous one
    img[w + 50: w + 80] = np.arange(30, 0, -1)   # falling ramp
    img[2 * w: 3 * w] = 9.0           # a whole bright row
    for radius in (0, 1, 8, 16, 400):
        got = nl.find_bright_pixels(ctx, img, w, 0.5, radius)
        want = O.find_bright_pixels(img, w, 0.5, radius)
        assert got.tobytes() == want.tobytes(), radius


@pytest.mark.parametrize("w,h", [(640, 480), (2048, 1500)])
def test_find_stars_matches_oracle(ctx, w, h):
    img = star_field(w, h, w * h // 5000, seed=3 * w)
    loc, scale = 100.0, 3.0
    for bp_sigma, md_sd in ((0.0, 0.0), (5.0, 4.0)):
        got = nl.find_stars(ctx, img, w, loc, scale, 15.0, bp_sigma, 1.4, 16, md_sd)
        want = O.find_stars(img, w, loc, scale, 15.0, bp_sigma, 1.4, 16, md_sd)
        assert len(got[0]) == len(want[0]) and len(got[0]) > 10
        assert got[0].tobytes() == want[0].tobytes()
        assert bits_equal([got[1], got[2]], [want[1], want[2]])


def test_config3_pipeline_detect_project_stack(ctx):
    """BASELINE config 3 in miniature: star-detect every frame, resample it with its (given) alignment
    transform, stack the aligned frames -- CUDA path against the oracle at every stage.  The triangle
    matcher that produces the transform stays on the host in Go (gonum Nelder-Mead, parity unpinned), so
    the transforms are inputs here."""
    from util import MODE_ID
    w, h, n = 320, 240, 8
    base = star_field(w, h, 40, seed=5, noise=2.0, hot=0).reshape(h, w)
    rng = np.random.default_rng(9)
    aligned_gpu, aligned_cpu = [], []
    for k in range(n):
        th = np.deg2rad(rng.uniform(-1, 1))
        trans = np.array([np.cos(th), -np.sin(th), rng.uniform(-6, 6), np.sin(th), np.cos(th), rng.uniform(-6, 6)], np.float32)
        frame = (base + rng.standard_normal((h, w)).astype(np.float32) * 2).astype(np.float32).reshape(-1)
        sg = nl.find_stars(ctx, frame, w, 100.0, 2.0, 15.0, 0.0, 1.4, 8, 0.0)
        sc = O.find_stars(frame, w, 100.0, 2.0, 15.0, 0.0, 1.4, 8, 0.0)
        assert sg[0].tobytes() == sc[0].tobytes() and len(sg[0]) > 5
        aligned_gpu.append(nl.project(ctx, frame, w, h, w, h, trans))          # NaN outside the source
        aligned_cpu.append(O.project(frame, w, h, w, h, trans, np.float32(np.nan)))
        assert bits_equal(aligned_gpu[-1], aligned_cpu[-1])
    fg, fc = np.stack(aligned_gpu), np.stack(aligned_cpu)
    assert np.isnan(fg).any()
    for mode in ("sigma", "winsor", "median"):
        with nl.StackJob(ctx, n, w * h) as job:
            for i in range(n):
                job.put_frame(i, fg[i])
            got = job.run(MODE_ID[mode])
        want = O.stack(fc, mode)
        assert bits_equal(got[0], want[0]), (mode, first_mismatch(got[0], want[0]))
        assert got[1:] == want[1:]


@pytest.mark.parametrize("amd64", [True, False])
@pytest.mark.parametrize("w,h", [(64, 48), (333, 257), (3, 3), (5, 4), (7, 9), (8, 3), (9, 5), (13, 6), (14, 14), (1024, 1024), (1021, 37)])
def test_estimate_noise_matches_oracle(ctx, w, h, amd64):
    """stats.EstimateNoise on the device, bit-exact against the oracle in both of the reference's
    summation orders: the AVX2 lanes with fused multiply-adds (amd64 builds) and the pure-Go loop"""
    rng = np.random.default_rng(w + h)
    img = (rng.standard_normal(w * h) * 11 + 500).astype(np.float32)
    want = O.estimate_noise(img, w, amd64=amd64)
    ctx.set_numerics(nl.NUMERICS_AMD64 if amd64 else nl.NUMERICS_PUREGO)
    try:
        got = nl.estimate_noise(ctx, img, w)
    finally:
        ctx.set_numerics(nl.NUMERICS_AMD64)
    assert got.view(np.uint32) == want.view(np.uint32) or (np.isnan(got) and np.isnan(want)), (got, want)


def test_inverse_noise_weights_from_resident_frames(ctx):
    """BASELINE configs[1] mode: winsorized sigma-clip + noise-weighted mean; the per-frame noise is
    estimated on the device for all resident frames in one launch"""
    import ctypes as C
    from nightlight_b200.ops import Image, OpStack
    w, h, n = 96, 64, 18
    rng = np.random.default_rng(2)
    frames = [(rng.standard_normal(w * h) * (5 + k % 4) + 300).astype(np.float32) for k in range(n)]
    noise = np.array([O.estimate_noise(f, w, amd64=True) for f in frames], np.float32)
    with nl.StackJob(ctx, n, w * h) as job:
        for i, f in enumerate(frames):
            job.put_frame(i, f)
        got = job.frame_noise(w)
    assert np.array_equal(got.view(np.uint32), noise.view(np.uint32))
    weights = (np.float32(1) / (np.float32(1) + np.float32(4) * (noise - noise.min()) / (noise.max() - noise.min()))).astype(np.float32)
    res = OpStack(mode=nl.ST_WINSOR_SIGMA, weighting=nl.W_INVERSE_NOISE).apply([Image(data=f, naxisn=(w, h)) for f in frames], ctx)
    want = O.stack(np.stack(frames), "winsor", weights=weights)
    assert bits_equal(res.data, want[0]) and (res.clip_low, res.clip_high) == want[1:]


def test_project_scaled_equals_match_histogram_then_project(ctx):
    """N1: the histogram match d*mult + offset (pixelops.go:601-612) fused into the resample"""
    rng = np.random.default_rng(31)
    w, h = 300, 200
    src = (rng.standard_normal(w * h) * 40 + 700).astype(np.float32)
    mult, off = np.float32(1.0371), np.float32(-12.625)
    matched = (src * mult).astype(np.float32) + off            # mul, then add, both rounded to fp32
    th = np.deg2rad(0.7)
    trans = np.array([np.cos(th), -np.sin(th), 3.25, np.sin(th), np.cos(th), -4.5], np.float32)
    got = nl.project_scaled(ctx, src, w, h, w, h, trans, mult, off)
    want = O.project(matched.astype(np.float32), w, h, w, h, trans, np.float32(np.nan))
    assert bits_equal(got, want), first_mismatch(got, want)


@pytest.mark.parametrize("bitpix", [8, 16, 32, 64, -32, -64])
def test_fits_payload_decode_encode(ctx, bitpix):
    """N1: the reader's / writer's conversion loops on the device (read.go:176-443, write.go:182-215)"""
    rng = np.random.default_rng(140 + bitpix)
    n = 10007
    dt = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}[bitpix]
    if bitpix == 8:
        disk = rng.integers(0, 256, n).astype(dt)
    elif bitpix > 0:
        lim = 2 ** (bitpix - 1) - 1
        disk = rng.integers(-lim, lim, n).astype(dt)
    else:
        disk = (rng.standard_normal(n) * 1e3).astype(dt)
    bscale, bzero = (np.float32(1), np.float32(32768)) if bitpix == 16 else (np.float32(1.5), np.float32(-0.25))
    got = nl.fits_decode(ctx, disk.tobytes(), bitpix, float(bscale), float(bzero))
    val = disk.astype(dt[1:]).astype(np.float32)
    want = (val * bscale).astype(np.float32) + bzero
    assert bits_equal(got, want.astype(np.float32)), first_mismatch(got, want)
    # encode: NaN -> 0, network byte order
    data = want.astype(np.float32).copy()
    data[::97] = np.nan
    raw = nl.fits_encode(ctx, data)
    back = np.frombuffer(raw, dtype=">f4").astype(np.float32)
    expect = data.copy()
    expect[np.isnan(expect)] = 0
    assert np.array_equal(back.view(np.uint32), expect.view(np.uint32))


def test_stack_from_raw_int16_payloads(ctx):
    """frames uploaded as raw 16-bit FITS payloads (half the PCIe bytes), decoded on the device, stacked"""
    rng = np.random.default_rng(50)
    n, p = 12, 5000
    adu = rng.integers(-32768, 32767, (n, p)).astype(">i2")
    frames = (adu.astype(np.int16).astype(np.float32) * np.float32(1)).astype(np.float32) + np.float32(32768)
    with nl.StackJob(ctx, n, p) as job:
        for i in range(n):
            job.put_frame_raw(i, adu[i].tobytes(), 16, 1.0, 32768.0)
        got = job.run(nl.ST_SIGMA)
    want = O.stack(frames, "sigma")
    assert bits_equal(got[0], want[0]) and got[1:] == want[1:]


@pytest.mark.parametrize("seed", [11, 12])
def test_project_and_bright_scan_fuzz(ctx, seed):
    """seeded fuzz: image sizes, affine transforms (rotation, scale, shear, large shifts), fill values; star
    scan thresholds and radii on images with plateaus -- bit-identical to the oracle"""
    rng = np.random.default_rng(seed)
    for it in range(25):
        sw, sh = int(rng.integers(1, 200)), int(rng.integers(1, 150))
        dw, dh = int(rng.integers(1, 200)), int(rng.integers(1, 150))
        src = (rng.standard_normal(sw * sh) * 30 + 500).astype(np.float32)
        th = np.deg2rad(rng.uniform(-180, 180))
        sc = rng.uniform(0.3, 3.0)
        trans = np.array([sc * np.cos(th), -sc * np.sin(th) + rng.uniform(-0.2, 0.2), rng.uniform(-sw, sw),
                          sc * np.sin(th), sc * np.cos(th), rng.uniform(-sh, sh)], np.float32)
        oob = np.float32(rng.choice([np.nan, 0.0, -1.5]))
        try:
            want = O.project(src, sw, sh, dw, dh, trans, oob)
        except Exception:
            with pytest.raises(nl.NightlightError):
                nl.project(ctx, src, sw, sh, dw, dh, trans, oob)
            continue
        got = nl.project(ctx, src, sw, sh, dw, dh, trans, oob)
        assert bits_equal(got, want), (it, sw, sh, dw, dh, list(trans), first_mismatch(got, want))
    for it in range(25):
        w, h = int(rng.integers(1, 400)), int(rng.integers(1, 60))
        img = np.round(rng.standard_normal(w * h) * 2 + 10).astype(np.float32)       # many equal values: plateaus
        thr = float(rng.choice([9.5, 11.5, 13.5, 100.0]))
        radius = int(rng.choice([0, 1, 2, 5, 16, 500]))
        got = nl.find_bright_pixels(ctx, img, w, thr, radius)
        want = O.find_bright_pixels(img, w, thr, radius)
        assert got.tobytes() == want.tobytes(), (it, w, h, thr, radius)


def test_find_stars_fuzz(ctx):
    """seeded fuzz of the whole FindStars pipeline: field sizes, star densities, detection sigmas, radii,
    in/out ratios, with and without the bad-pixel rejection"""
    rng = np.random.default_rng(77)
    for it in range(8):
        w, h = int(rng.integers(200, 900)), int(rng.integers(150, 600))
        img = star_field(w, h, int(rng.integers(5, w * h // 3000 + 6)), seed=int(rng.integers(0, 10**6)),
                         noise=float(rng.uniform(1, 5)), hot=int(rng.integers(0, 30)))
        loc, scale = 100.0, float(rng.uniform(1.0, 5.0))
        star_sig = float(rng.choice([5.0, 10.0, 15.0, 25.0]))
        radius = int(rng.choice([4, 8, 16, 24]))
        in_out = float(rng.choice([1.0, 1.4, 2.0]))
        bp_sigma, md_sd = (0.0, 0.0) if rng.random() < 0.5 else (float(rng.choice([3.0, 5.0])), float(rng.uniform(1, 6)))
        got = nl.find_stars(ctx, img, w, loc, scale, star_sig, bp_sigma, in_out, radius, md_sd)
        want = O.find_stars(img, w, loc, scale, star_sig, bp_sigma, in_out, radius, md_sd)
        assert len(got[0]) == len(want[0]), (it, len(got[0]), len(want[0]))
        assert got[0].tobytes() == want[0].tobytes(), it
        assert bits_equal([got[1], got[2]], [want[1], want[2]]), it


def test_project_scatter_into_stripe_jobs(ctx):
    """N4: the resample stores each destination row into the stack job that owns its row stripe (here three
    ragged stripes on one device); every job then holds exactly its rows of every resampled frame"""
    from nightlight_b200.stripes import all_stripes, scatter_project
    from util import MODE_ID
    rng = np.random.default_rng(77)
    sw, sh, dw, dh, n = 150, 110, 131, 97, 5
    stripes = all_stripes(dh, 3)
    jobs = [nl.StackJob(ctx, n, rows * dw) for _, rows in stripes]
    src_dev = ctx.dev_alloc(4 * sw * sh)
    try:
        frames, want = [], []
        for k in range(n):
            src = (rng.standard_normal(sw * sh) * 20 + 400).astype(np.float32)
            th = np.deg2rad(rng.uniform(-3, 3))
            trans = np.array([np.cos(th), -np.sin(th), rng.uniform(-6, 6), np.sin(th), np.cos(th), rng.uniform(-6, 6)], np.float32)
            mult, off = (1.0, 0.0) if k % 2 == 0 else (1.0 + 0.01 * k, -3.0)
            ctx.h2d(src_dev, src)
            scatter_project(ctx, src_dev, sw, sh, dw, dh, trans, k, [j.frames_dev[0] for j in jobs],
                            [s[0] for s in stripes] + [dh], float("nan"), mult, off)
            ctx.sync()
            want.append(nl.project(ctx, src, sw, sh, dw, dh, trans) if k % 2 == 0 else
                        nl.project_scaled(ctx, src, sw, sh, dw, dh, trans, mult, off))
        want = np.stack(want)
        for (row0, rows), job in zip(stripes, jobs):
            got = np.empty((n, rows * dw), np.float32)
            ctx.d2h(got, job.frames_dev[0])
            assert bits_equal(got, want[:, row0 * dw:(row0 + rows) * dw]), (row0, rows)
        # and the stripes stack to the whole-image stack of the resampled frames
        whole = O.stack(want, "median")
        for (row0, rows), job in zip(stripes, jobs):
            res = job.run(MODE_ID["median"])
            assert bits_equal(res[0], whole[0][row0 * dw:(row0 + rows) * dw])
    finally:
        ctx.dev_free(src_dev)
        for j in jobs:
            j.close()


def _resident(ctx, frames):
    """frames [n, px] -> a stack job holding them (device), plus the device base pointer and stride"""
    n, px = frames.shape
    job = nl.StackJob(ctx, n, px)
    for i in range(n):
        job.put_frame(i, frames[i])
    ctx.sync()
    base, stride = job.frames_dev
    return job, base, stride


def test_batched_resample_of_a_resident_stack(ctx):
    """nl_project_batch_dev: all frames in one launch, each with its own transform (and histogram match), written
    straight into the slots of a second stack job -- equal to the per-frame oracle, then stacked"""
    import ctypes as C
    lib = nl.load_library()
    w, h, n = 333, 257, 7
    rng = np.random.default_rng(42)
    frames = (rng.standard_normal((n, w * h)) * 50 + 500).astype(np.float32)
    trans = np.zeros((n, 6), np.float32)
    for k in range(n):
        th = np.deg2rad(rng.uniform(-2, 2))
        trans[k] = [np.cos(th), -np.sin(th), rng.uniform(-9, 9), np.sin(th), np.cos(th), rng.uniform(-9, 9)]
    mult = rng.uniform(0.9, 1.1, n).astype(np.float32)
    off = rng.uniform(-5, 5, n).astype(np.float32)
    mult[2], off[2] = 1.0, 0.0                          # "no match" for one frame
    fp = C.POINTER(C.c_float)
    src_job, sbase, sstride = _resident(ctx, frames)
    with src_job, nl.StackJob(ctx, n, w * h) as dst_job:
        dbase, dstride = dst_job.frames_dev
        for use_match in (False, True):
            nl.binding.check(lib.nl_project_batch_dev(ctx.handle, C.c_void_p(sbase), sstride, w, h, C.c_void_p(dbase), dstride, w, h, n,
                                                      trans.ctypes.data_as(fp), float("nan"),
                                                      mult.ctypes.data_as(fp) if use_match else None,
                                                      off.ctypes.data_as(fp) if use_match else None))
            ctx.sync()
            aligned = []
            for k in range(n):
                got = np.empty(w * h, np.float32)
                ctx.d2h(got, dbase + 4 * k * dstride)
                src = frames[k]
                if use_match and not (mult[k] == 1.0 and off[k] == 0.0):
                    src = (src * mult[k]).astype(np.float32) + off[k]            # MatchHistogram, then Project
                want = O.project(src.astype(np.float32), w, h, w, h, trans[k], np.float32(np.nan))
                assert bits_equal(got, want), (use_match, k, first_mismatch(got, want))
                aligned.append(want)
            res, cl, ch = dst_job.run(nl.ST_SIGMA)
            want = O.stack(np.stack(aligned), "sigma")
            assert bits_equal(res, want[0]) and (cl, ch) == want[1:]
    bad = trans.copy()
    bad[3] = [1, 2, 0, 2, 4, 0]
    rc = lib.nl_project_batch_dev(ctx.handle, C.c_void_p(sbase), sstride, w, h, C.c_void_p(sbase), sstride, w, h, n,
                                  bad.ctypes.data_as(fp), 0.0, None, None)
    assert rc == nl.binding.NL_E_SINGULAR


def test_batched_star_scan_one_read_per_frame(ctx):
    """nl_find_bright_batch_dev / nl_find_stars_batch_dev over the frames of a resident stack: per-row slots + scan +
    compaction equal the oracle's raster-order lists; a frame with a row of more than 32 candidates takes the
    two-pass scan; the sparse steps on host threads give the oracle's stars"""
    import ctypes as C
    lib = nl.load_library()
    w, h, n = 640, 300, 6
    frames = np.stack([star_field(w, h, 60, seed=100 + k) for k in range(n)])
    frames[3].reshape(h, w)[17, ::9] += 4000.0          # 72 isolated hits in one row (radius 4): beyond the 32 slots
    # candidates in the first and last row and at both ends of the frame: rejectBadPixels' gather buffer keeps entries of
    # the candidate before them there (the device tests the interior candidates, the host replays these)
    frames[1].reshape(h, w)[0, 5::40] += 3000.0
    frames[1].reshape(h, w)[h - 1, 7::37] += 3000.0
    frames[2].ravel()[0] += 5000.0
    frames[2].ravel()[-1] += 5000.0
    frames[2].reshape(h, w)[0, 100:300:11] += 900.0
    frames[4].reshape(h, w)[h - 1, 3::23] += 2500.0
    frames[4].reshape(h, w)[h - 2, 9::29] += 2500.0
    radius = 4
    thr = np.array([130.0 + k for k in range(n)], np.float32)
    job, base, stride = _resident(ctx, frames)
    with job:
        cap = 20000
        out = np.zeros((n, cap), dtype=nl.STAR_DTYPE)
        counts = np.zeros(n, np.int32)
        nl.binding.check(lib.nl_find_bright_batch_dev(ctx.handle, C.c_void_p(base), n, stride, w * h, w, thr.ctypes.data_as(C.POINTER(C.c_float)),
                                                      radius, out.ctypes.data_as(C.c_void_p), cap, counts.ctypes.data_as(C.POINTER(C.c_int32))))
        for k in range(n):
            want = O.find_bright_pixels(frames[k], w, float(thr[k]), radius)
            assert counts[k] == len(want) and len(want) > 20, (k, counts[k], len(want))
            assert out[k, :counts[k]].tobytes() == want.tobytes(), k
        # truncated lists still count
        small = np.zeros((n, 5), dtype=nl.STAR_DTYPE)
        nl.binding.check(lib.nl_find_bright_batch_dev(ctx.handle, C.c_void_p(base), n, stride, w * h, w, thr.ctypes.data_as(C.POINTER(C.c_float)),
                                                      radius, small.ctypes.data_as(C.c_void_p), 5, counts.ctypes.data_as(C.POINTER(C.c_int32))))
        for k in range(n):
            want = O.find_bright_pixels(frames[k], w, float(thr[k]), radius)
            assert counts[k] == len(want) and small[k].tobytes() == want[:5].tobytes()
        # the whole FindStars
        loc = np.full(n, 100.0, np.float32)
        scale = np.array([3.0 + 0.1 * k for k in range(n)], np.float32)
        mds = np.full(n, 4.0, np.float32)
        stars = np.zeros((n, 4000), dtype=nl.STAR_DTYPE)
        sos, hfr = np.zeros(n, np.float32), np.zeros(n, np.float32)
        ptrs = (C.c_void_p * n)(*[frames[k].ctypes.data for k in range(n)])
        td, thost = C.c_double(), C.c_double()
        fp = C.POINTER(C.c_float)
        for bp in (0.0, 5.0):
            nl.binding.check(lib.nl_find_stars_batch_dev(ctx.handle, C.c_void_p(base), n, stride, ptrs, w * h, w, loc.ctypes.data_as(fp),
                                                         scale.ctypes.data_as(fp), 15.0, bp, 1.4, 16, mds.ctypes.data_as(fp),
                                                         stars.ctypes.data_as(C.c_void_p), 4000, counts.ctypes.data_as(C.POINTER(C.c_int32)),
                                                         sos.ctypes.data_as(fp), hfr.ctypes.data_as(fp), C.byref(td), C.byref(thost)))
            for k in range(n):
                want = O.find_stars(frames[k], w, 100.0, float(scale[k]), 15.0, bp, 1.4, 16, 4.0)
                assert counts[k] == len(want[0]) and counts[k] > 5, (bp, k)
                assert stars[k, :counts[k]].tobytes() == want[0].tobytes(), (bp, k)
                assert bits_equal([sos[k], hfr[k]], [want[1], want[2]])
            assert td.value > 0 and thost.value > 0


def test_batched_bad_pixel_map(ctx):
    import ctypes as C
    lib = nl.load_library()
    w, h, n = 200, 96, 4
    frames = np.stack([star_field(w, h, 10, seed=7 + k, hot=30) for k in range(n)])
    job, base, stride = _resident(ctx, frames)
    with job:
        cap = 500
        bpm = np.zeros((n, cap), np.int32)
        counts = np.zeros(n, np.int64)
        st = np.zeros((n, 4), np.float32)
        nl.binding.check(lib.nl_bad_pixel_map_batch_dev(ctx.handle, C.c_void_p(base), n, stride, w * h, w, 3.0, 5.0,
                                                        bpm.ctypes.data_as(C.POINTER(C.c_int32)), cap, counts.ctypes.data_as(C.POINTER(C.c_int64)),
                                                        st.ctypes.data_as(C.POINTER(C.c_float))))
        for k in range(n):
            want_bpm, want_st, _ = O.bad_pixel_map(frames[k], w, 3.0, 5.0)
            assert counts[k] == want_bpm.size and np.array_equal(bpm[k, :counts[k]], want_bpm)
            assert np.array_equal(st[k].view(np.uint32), want_st.view(np.uint32))
