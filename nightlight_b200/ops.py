"""Host-side mirrors of the reference's operators for the stacking hot path.

Same field names (the reference's JSON tags), argument meaning and error behaviour as
  OpStack         internal/ops/stack/stack.go:66-227
  OpStackBatches  internal/ops/stack/stackbatches.go:56-119 (stack of stacks)
  Image.Project   internal/fits/project.go:26-76
  FindStars       internal/star/findstars.go:59-100
All arithmetic happens in libnightlight_cuda.so; nothing here computes pixels.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import binding as B
from .binding import NightlightError, check, load_library


@dataclass
class Image:
    """The fields of fits.Image (internal/fits/fits.go:30-54) this path reads or writes."""
    data: np.ndarray                      # float32, Pixels samples
    naxisn: Sequence[int] = ()            # (width, height)
    exposure: float = 0.0
    noise: Optional[float] = None         # Stats.Noise()
    hfr: float = 0.0
    id: int = 0
    clip_low: int = 0                     # outputs of a stack (printed by the reference, stack.go:214-218)
    clip_high: int = 0


def get_weights(frames: Sequence[Image], weighting: int):
    """getWeights, stack.go:231-270 -> float32[n] or None"""
    n = len(frames)
    if weighting == B.W_NONE:
        return None
    if weighting not in (B.W_EXPOSURE, B.W_INVERSE_NOISE, B.W_INVERSE_HFR):
        raise NightlightError(B.NL_E_WEIGHTS, "Invalid weighting mode %d\n" % weighting)
    if weighting == B.W_INVERSE_NOISE:
        for f in frames:
            if f.noise is None:
                raise NightlightError(B.NL_E_WEIGHTS, "%d: Missing stats information for noise-weighted stacking" % f.id)
    exposure = np.array([f.exposure for f in frames], dtype=np.float32)
    noise = np.array([f.noise if f.noise is not None else 0.0 for f in frames], dtype=np.float32)
    hfr = np.array([f.hfr for f in frames], dtype=np.float32)
    w = np.empty(n, dtype=np.float32)
    fp = C.POINTER(C.c_float)
    check(load_library().nl_get_weights(int(weighting), exposure.ctypes.data_as(fp), noise.ctypes.data_as(fp),
                                        hfr.ctypes.data_as(fp), n, w.ctypes.data_as(fp)))
    return w


@dataclass
class OpStack:
    """stack.OpStack (stack.go:66-73); JSON tags mode / weighting / sigmaLow / sigmaHigh."""
    mode: int = B.ST_AUTO
    weighting: int = B.W_NONE
    sigmaLow: float = 2.75
    sigmaHigh: float = 2.75
    refFrameLoc: float = 0.0              # json:"-" in the reference and never assigned -> 0

    def apply(self, frames: Sequence[Image], ctx: B.Context) -> Image:
        """OpStack.Apply, stack.go:115-227"""
        if self.mode < B.ST_MEDIAN or self.mode > B.ST_AUTO:
            raise NightlightError(B.NL_E_INVALID, "invalid stacking mode")
        pixels = int(np.asarray(frames[0].data).size)
        with B.StackJob(ctx, len(frames), pixels) as job:
            for i, f in enumerate(frames):
                job.put_frame(i, f.data)
            if self.weighting == B.W_INVERSE_NOISE and any(f.noise is None for f in frames) and len(frames[0].naxisn) >= 1:
                # the reference's Stats.Noise() is computed lazily from the frame (stats.go, noise.go:24-55);
                # here for all resident frames in one launch
                for f, nz in zip(frames, job.frame_noise(int(frames[0].naxisn[0]))):
                    if f.noise is None:
                        f.noise = float(nz)
            weights = get_weights(frames, self.weighting)
            data, cl, ch = job.run(self.mode, weights, self.sigmaLow, self.sigmaHigh, self.refFrameLoc)
        exposure = np.float32(0)
        for f in frames:                  # stack.go:220-221, sequential fp32 sum
            exposure = np.float32(exposure + np.float32(f.exposure))
        return Image(data=data, naxisn=tuple(frames[0].naxisn), exposure=float(exposure), clip_low=cl, clip_high=ch)


def partition(num_frames, width, height, stack_memory_mb, max_threads=1, has_dark=False, has_flat=False, perm=None):
    """OpStackBatches.partition, stackbatches.go:121-210: frames -> equal batches from a memory budget.
    -> (order, numBatches, batchSize, maxThreads).  `stack_memory_mb` is the reference's StackMemoryMB; on the
    GPU pass the device budget (Context.mem_info).  The reference shuffles with math/rand.Perm when there is
    more than one batch; the permutation is an input here (`perm`, default identity) and each batch's slice
    of it is sorted ascending like stackbatches.go:196-203."""
    if num_frames <= 0:
        raise NightlightError(B.NL_E_INVALID, "No input files to prepare batches")
    nbytes = int(width) * int(height) * 4
    available = (int(stack_memory_mb) * 1024 * 1024) // nbytes
    mt, bs, nb = int(max_threads), 0, 0
    while mt >= 1:
        bs = available - mt - (1 if has_dark else 0) - (1 if has_flat else 0)
        if bs >= 2:
            nb = (num_frames + bs - 1) // bs
            if nb > 1:
                bs -= 2                       # reference frame from batch 0, and the stack of stacks
            if bs >= 2 and bs >= mt:
                break
        mt -= 1
    if mt < 1 or bs < 2:
        raise NightlightError(B.NL_E_NOMEM, "Cannot find a stacking execution path within the given memory constraints.")
    while (bs - 1) * nb >= num_frames:       # even out the size of the last batch
        bs -= 1
    order = list(range(num_frames))
    if nb > 1:
        order = list(perm) if perm is not None else order
        if sorted(order) != list(range(num_frames)):
            raise NightlightError(B.NL_E_INVALID, "perm is not a permutation of the frames")
        for i in range(nb):
            order[i * bs:(i + 1) * bs] = sorted(order[i * bs:(i + 1) * bs])
    return order, nb, bs, mt


@dataclass
class OpStackBatches:
    """The stack-of-stacks arithmetic of OpStackBatches.Apply (stackbatches.go:84-116): every batch is
    stacked on its own, the batch results are averaged weighted by their frame counts
    (StackIncremental / StackIncrementalFinalize, stack.go:924-944) on the device."""
    perBatch: OpStack = field(default_factory=OpStack)

    def apply(self, batches: Sequence[Sequence[Image]], ctx: B.Context) -> Image:
        lib = load_library()
        if len(batches) == 1:
            return self.perBatch.apply(batches[0], ctx)
        pixels = int(np.asarray(batches[0][0].data).size)
        acc = ctx.dev_alloc(4 * pixels)
        tmp = ctx.dev_alloc(4 * pixels)
        try:
            frames_total = 0
            exposure = np.float32(0)
            for b, batch in enumerate(batches):
                weights = get_weights(batch, self.perBatch.weighting)
                with B.StackJob(ctx, len(batch), pixels) as job:
                    for i, f in enumerate(batch):
                        job.put_frame(i, f.data)
                    job.run_dev(self.perBatch.mode, tmp, weights, self.perBatch.sigmaLow, self.perBatch.sigmaHigh,
                                self.perBatch.refFrameLoc)
                    check(lib.nl_stack_incremental_dev(ctx.handle, C.c_void_p(acc), C.c_void_p(tmp), pixels,
                                                       float(len(batch)), 1 if b == 0 else 0))
                    ctx.sync()
                frames_total += len(batch)
                bexp = np.float32(0)
                for f in batch:
                    bexp = np.float32(bexp + np.float32(f.exposure))
                exposure = bexp if b == 0 else np.float32(exposure + bexp)     # stack.go:926-931
            check(lib.nl_stack_incremental_finalize_dev(ctx.handle, C.c_void_p(acc), pixels, float(frames_total)))
            out = np.empty(pixels, dtype=np.float32)
            ctx.d2h(out, acc)
        finally:
            ctx.dev_free(acc)
            ctx.dev_free(tmp)
        return Image(data=out, naxisn=tuple(batches[0][0].naxisn), exposure=float(exposure))


def median_filter3x3(ctx: B.Context, data, width):
    """median.MedianFilter3x3 (median3x3_amd64.go:24-48 / median3x3.go:26-38) -> float32[len]"""
    data = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
    out = np.empty_like(data)
    check(load_library().nl_median_filter3x3(ctx.handle, data.ctypes.data_as(C.c_void_p), data.size, int(width),
                                             out.ctypes.data_as(C.c_void_p)))
    return out


def stats(ctx: B.Context, data):
    """Stats.Min/Mean/Max/StdDev (stats.go:102-153) -> float32[4] = min, mean, max, stddev"""
    data = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
    out = np.zeros(4, dtype=np.float32)
    check(load_library().nl_stats(ctx.handle, data.ctypes.data_as(C.c_void_p), data.size, out.ctypes.data_as(C.POINTER(C.c_float))))
    return out


def bad_pixel_map(ctx: B.Context, data, width, sigma_low, sigma_high, cap=None):
    """pre.BadPixelMap (badpixels.go:32-51) -> (bpm int32[], medianDiffStats float32[4] = min, mean, max, stddev)"""
    data = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
    cap = data.size // 100 + 1024 if cap is None else int(cap)
    st = np.zeros(4, dtype=np.float32)
    lib = load_library()
    while True:
        bpm = np.empty(cap, dtype=np.int32)
        n = C.c_int64()
        check(lib.nl_bad_pixel_map(ctx.handle, data.ctypes.data_as(C.c_void_p), data.size, int(width), float(sigma_low),
                                   float(sigma_high), bpm.ctypes.data_as(C.POINTER(C.c_int32)), cap, C.byref(n),
                                   st.ctypes.data_as(C.POINTER(C.c_float))))
        if n.value <= cap:
            return bpm[:n.value].copy(), st
        cap = n.value


def op_bad_pixel(ctx: B.Context, data, width, sigma_low, sigma_high):
    """OpBadPixel.Apply, monochrome (preprocess.go:180-191) -> (repaired data, number removed, medianDiffStats[4])"""
    data = np.array(data, dtype=np.float32).reshape(-1)          # a copy: the operator repairs in place
    st = np.zeros(4, dtype=np.float32)
    n = C.c_int64()
    check(load_library().nl_op_bad_pixel(ctx.handle, data.ctypes.data_as(C.c_void_p), data.size, int(width), float(sigma_low),
                                         float(sigma_high), C.byref(n), st.ctypes.data_as(C.POINTER(C.c_float))))
    return data, n.value, st


def estimate_noise(ctx: B.Context, data, width):
    """stats.EstimateNoise (noise_amd64.go:25-43, or noise.go:24-55 in pure-Go numerics) of one frame -> float32"""
    data = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
    out = C.c_float()
    check(load_library().nl_estimate_noise(ctx.handle, data.ctypes.data_as(C.c_void_p), data.size, int(width), C.byref(out)))
    return np.float32(out.value)


def transform_invert(trans):
    """Transform2D.Invert, coord.go:159-201"""
    t = np.ascontiguousarray(trans, dtype=np.float32)
    inv = np.empty(6, dtype=np.float32)
    fp = C.POINTER(C.c_float)
    check(load_library().nl_transform_invert(t.ctypes.data_as(fp), inv.ctypes.data_as(fp)))
    return inv


def project(ctx: B.Context, src, src_w, src_h, dst_w, dst_h, trans, out_of_bounds=float("nan")):
    """(*Image).Project, project.go:26-76: src float32[src_h*src_w] -> float32[dst_h*dst_w]"""
    src = np.ascontiguousarray(src, dtype=np.float32).reshape(-1)
    if src.size != src_w * src_h:
        raise NightlightError(B.NL_E_INVALID, "source size does not match its dimensions")
    dst = np.empty(dst_w * dst_h, dtype=np.float32)
    t = np.ascontiguousarray(trans, dtype=np.float32)
    check(load_library().nl_project(ctx.handle, src.ctypes.data_as(C.c_void_p), src_w, src_h,
                                    dst.ctypes.data_as(C.c_void_p), dst_w, dst_h,
                                    t.ctypes.data_as(C.POINTER(C.c_float)), out_of_bounds))
    return dst


def project_scaled(ctx: B.Context, src, src_w, src_h, dst_w, dst_h, trans, multiplier, offset, out_of_bounds=float("nan")):
    """Image.MatchHistogram (pixelops.go:601-612: d*multiplier + offset) fused into (*Image).Project"""
    src = np.ascontiguousarray(src, dtype=np.float32).reshape(-1)
    if src.size != src_w * src_h:
        raise NightlightError(B.NL_E_INVALID, "source size does not match its dimensions")
    dst = np.empty(dst_w * dst_h, dtype=np.float32)
    t = np.ascontiguousarray(trans, dtype=np.float32)
    check(load_library().nl_project_scaled(ctx.handle, src.ctypes.data_as(C.c_void_p), src_w, src_h,
                                           dst.ctypes.data_as(C.c_void_p), dst_w, dst_h,
                                           t.ctypes.data_as(C.POINTER(C.c_float)), out_of_bounds, multiplier, offset))
    return dst


def fits_decode(ctx: B.Context, raw, bitpix, bscale=1.0, bzero=0.0):
    """the reader's conversion loop (read.go:176-443): big-endian payload -> float32(val)*Bscale + Bzero"""
    a = np.frombuffer(raw, dtype=np.uint8) if isinstance(raw, (bytes, bytearray, memoryview)) else np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
    count = a.size // (abs(int(bitpix)) // 8)
    out = np.empty(count, dtype=np.float32)
    check(load_library().nl_fits_decode(ctx.handle, a.ctypes.data_as(C.c_void_p), int(bitpix), count, bscale, bzero,
                                        out.ctypes.data_as(C.c_void_p)))
    return out


def fits_encode(ctx: B.Context, data):
    """the writer's conversion loop (write.go:182-215): float32 -> big-endian bytes, NaN -> 0"""
    a = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
    out = np.empty(4 * a.size, dtype=np.uint8)
    check(load_library().nl_fits_encode(ctx.handle, a.ctypes.data_as(C.c_void_p), a.size, out.ctypes.data_as(C.c_void_p)))
    return out.tobytes()


def find_bright_pixels(ctx: B.Context, data, width, threshold, radius):
    """findBrightPixels, findstars.go:105-129 -> structured array of star.Star in raster order"""
    data = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
    n = C.c_int32()
    lib = load_library()
    check(lib.nl_find_bright(ctx.handle, data.ctypes.data_as(C.c_void_p), data.size, width, threshold, radius,
                             None, 0, C.byref(n)))
    out = np.zeros(max(n.value, 1), dtype=B.STAR_DTYPE)
    if n.value:
        check(lib.nl_find_bright(ctx.handle, data.ctypes.data_as(C.c_void_p), data.size, width, threshold, radius,
                                 out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
    return out[:n.value]


def find_stars(ctx: B.Context, data, width, location, scale, starSig, bpSigma, starInOut, radius, medianDiffStdDev=0.0):
    """star.FindStars, findstars.go:59-100 -> (stars, sumOfShifts, avgHFR)"""
    data = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
    cap = max(data.size // 100, 1024)
    out = np.zeros(cap, dtype=B.STAR_DTYPE)
    n, sos, hfr = C.c_int32(), C.c_float(), C.c_float()
    check(load_library().nl_find_stars(ctx.handle, data.ctypes.data_as(C.c_void_p), data.size, width, location, scale,
                                       starSig, bpSigma, starInOut, radius, medianDiffStdDev,
                                       out.ctypes.data_as(C.c_void_p), cap, C.byref(n), C.byref(sos), C.byref(hfr)))
    return out[:min(n.value, cap)], np.float32(sos.value), np.float32(hfr.value)


def find_sigmas_and_stack(stack_fn, mode, n_frames, pixels, clip_perc_low, clip_perc_high, log=None):
    """Goal-seek of the clipping sigmas for target clip percentages, restated from the reference's
    DEAD code (internal/ops/stack/stackfindsigma.go:27-170 is inside a comment block; nothing calls it,
    so parity is unpinned by the reference).  `stack_fn(sigma_low, sigma_high)` runs one stack and returns
    (result, clipLow, clipHigh): StackJob.run keeps the frames resident in HBM, so the up to 20 trial
    stacks cost no re-upload.  Host arithmetic is float32 like the Go code.
    -> (result, clipLow, clipHigh, sigmaLow, sigmaHigh)"""
    f32 = np.float32
    say = log or (lambda *_: None)
    if mode == B.ST_AUTO:                                   # :29-32
        mode = load_library().nl_auto_select_mode(int(n_frames))
    total = f32(pixels * n_frames)
    lo_t, hi_t = f32(clip_perc_low), f32(clip_perc_high)

    def perc(c):
        return f32(f32(f32(c) * f32(100.0)) / total)

    if mode in (B.ST_WINSOR_SIGMA, B.ST_SIGMA):             # binarySearchAndStack, :48-98
        low_l, low_r, high_l, high_r = f32(1.0), f32(11.0), f32(1.0), f32(11.0)
        low_m, high_m = f32(0.5) * f32(low_l + low_r), f32(0.5) * f32(high_l + high_r)
        i = 0
        while True:
            say("Step %d: stSigLow %.2f stSigHigh %.2f" % (i, low_m, high_m))
            res, cl, ch = stack_fn(float(low_m), float(high_m))
            dl = int(f32(f32(100) * perc(cl)) + f32(0.5)) - int(f32(100) * lo_t)
            dh = int(f32(f32(100) * perc(ch)) + f32(0.5)) - int(f32(100) * hi_t)
            if (dl == 0 and dh == 0) or i >= 20:
                return res, cl, ch, float(low_m), float(high_m)
            if dl > 0:
                low_l = low_m
            elif dl < 0:
                low_r = low_m
            if dl != 0:
                low_m = f32(0.5) * f32(low_l + low_r)
            if dh > 0:
                high_l = high_m
            elif dh < 0:
                high_r = high_m
            if dh != 0:
                high_m = f32(0.5) * f32(high_l + high_r)
            i += 1
    if mode == B.ST_LINEAR_FIT:                             # newtonMethodAndStack, :101-170
        sig_lo, sig_hi, eps = f32(6.0), f32(6.0), f32(0.005)
        i = 0
        while True:
            say("Step %d: stSigLow %.2f stSigHigh %.2f" % (i, sig_lo, sig_hi))
            res, cl, ch = stack_fn(float(sig_lo), float(sig_hi))
            d_l = f32(perc(cl) - lo_t)
            d_h = f32(perc(ch) - lo_t)                      # the reference uses the LOW target here too (:114)
            if (int(f32(100) * d_l + f32(0.5)) == 0 and int(f32(100) * d_h + f32(0.5)) == 0) or i >= 20:
                return res, cl, ch, float(sig_lo), float(sig_hi)
            i += 1
            _, cl2, _ = stack_fn(float(f32(sig_lo + eps)), float(sig_hi))
            diff_l = f32(f32(f32(perc(cl2) - lo_t) - d_l) / eps)
            if diff_l == 0:
                return res, cl, ch, float(sig_lo), float(sig_hi)
            new_lo = min(max(f32(sig_lo - f32(d_l / diff_l)), f32(0.1)), f32(20))
            i += 1
            _, _, ch3 = stack_fn(float(sig_lo), float(f32(sig_hi + eps)))
            diff_h = f32(f32(f32(perc(ch3) - lo_t) - d_h) / eps)   # :155 again the low target
            if diff_h == 0:
                return res, cl, ch, float(sig_lo), float(sig_hi)
            new_hi = min(max(f32(sig_hi - f32(d_h / diff_h)), f32(0.1)), f32(20))
            sig_lo, sig_hi = new_lo, new_hi
            i += 1
    res, cl, ch = stack_fn(0.0, 0.0)                        # :41-44: mode has no sigmas
    return res, cl, ch, 0.0, 0.0
