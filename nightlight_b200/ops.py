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
        weights = get_weights(frames, self.weighting)
        pixels = int(np.asarray(frames[0].data).size)
        with B.StackJob(ctx, len(frames), pixels) as job:
            for i, f in enumerate(frames):
                job.put_frame(i, f.data)
            data, cl, ch = job.run(self.mode, weights, self.sigmaLow, self.sigmaHigh, self.refFrameLoc)
        exposure = np.float32(0)
        for f in frames:                  # stack.go:220-221, sequential fp32 sum
            exposure = np.float32(exposure + np.float32(f.exposure))
        return Image(data=data, naxisn=tuple(frames[0].naxisn), exposure=float(exposure), clip_low=cl, clip_high=ch)


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
