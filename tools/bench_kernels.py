#!/usr/bin/env python
"""Device-resident timing of every kernel on the hot path against the measured HBM roofline.

    python tools/bench_kernels.py [--quick]

One JSON line per kernel: achieved = ALGORITHMIC bytes per launch / CUDA-event time (DESIGN.md section 3
states the per-unit figures), peak = MEASURED_PEAKS.json hbm_gbs.  Inputs are larger than L2 (126 MB): the stack
jobs by themselves, the per-frame kernels by rotating through distinct frames (steady state, no flush).  bench.py remains the headline benchmark; this is the per-kernel view
the ncu captures under profiles/ are taken from.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="smaller inputs (for ncu)")
    ap.add_argument("--only", default="", help="comma list: mean,mean_w,median,sigma,sigma_w,winsor,winsor_w,mad,linfit,project,fits,bright,prestats,incremental")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--tune", default="", help="k=v,k=v: nl_ctx_set_tuning knobs (A/B measurements)")
    ap.add_argument("--tile-width", default="0", help="force the column kernel's tile width (nl_ctx_set_tuning)")
    ap.add_argument("--defer-passes", default=None, help="deferral schedule, e.g. 3 or 8,12,16 or 0 (nl_ctx_set_tuning)")
    args = ap.parse_args()
    import torch
    import nightlight_b200 as nl
    from bench import peaks

    peak, peak_src = peaks()
    lib = nl.load_library()
    ctx = nl.Context(0)
    ctx.set_tuning("tile_width", args.tile_width)
    if args.defer_passes is not None:
        ctx.set_tuning("defer_passes", args.defer_passes)
    for kv in [x for x in args.tune.split(",") if x]:
        k, v = kv.split("=", 1)
        ctx.set_tuning(k, v.replace(":", ","))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    only = set(x for x in args.only.split(",") if x)

    def timed(fn, reps=args.reps, warm=2, flush_l2=True):
        for _ in range(warm):
            fn()
        ctx.sync()
        ms = []
        for _ in range(reps):
            if flush_l2:
                with torch.cuda.stream(ext):
                    flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(ext):
                e0.record()
            fn()
            with torch.cuda.stream(ext):
                e1.record()
            ctx.sync()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return float(np.median(ms))

    def timed_rot(fns, rounds=3):
        """Steady state over a rotation of distinct frames whose total size exceeds L2 (no flush: a flush by writing
        leaves 126 MB of dirty lines whose write-back would be charged to a 30-microsecond kernel).  Average per call."""
        for f in fns:
            f()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record()
        for _ in range(rounds):
            for f in fns:
                f()
        with torch.cuda.stream(ext):
            e1.record()
        ctx.sync()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (rounds * len(fns))

    ROT = 6            # frames in rotation: 6 x (96 MB in + 96 MB out) at 6000x4000

    def report(kernel, workload, algo_bytes, ms, extra=None):
        ach = algo_bytes / (ms * 1e-3) / 1e9
        line = {"kernel": kernel, "workload": workload, "ms": ms, "algorithmic_bytes": algo_bytes,
                "achieved_gbs": ach, "peak_gbs": peak, "frac": ach / peak, "peak_source": peak_src}
        if extra:
            line.update(extra)
        print(json.dumps(line), flush=True)

    # ---- stacking modes: 256 frames (sigma family) / 64 frames (linear fit is the heaviest) ------------
    rows = 128 if args.quick else 512
    width = 4096
    w256 = (np.float32(1) / (np.float32(1) + np.float32(4) * ((np.arange(256) % 7).astype(np.float32) / np.float32(6)))).astype(np.float32)
    stack_cases = [("mean", nl.ST_MEAN, None), ("mean_w", nl.ST_MEAN, w256), ("median", nl.ST_MEDIAN, None),
                   ("sigma", nl.ST_SIGMA, None), ("sigma_w", nl.ST_SIGMA, w256), ("winsor", nl.ST_WINSOR_SIGMA, None),
                   ("winsor_w", nl.ST_WINSOR_SIGMA, w256), ("mad", nl.ST_MAD_SIGMA, None), ("linfit", nl.ST_LINEAR_FIT, None)]
    if not only or only & set(c[0] for c in stack_cases):
        n, pixels = 256, width * rows
        job = nl.StackJob(ctx, n, pixels)
        job.synth_fill()
        out = torch.empty(pixels, dtype=torch.float32, device=dev)
        for name, mode, w in stack_cases:
            if only and name not in only:
                continue
            ms = timed(lambda: job.run_dev(mode, out.data_ptr(), w, 2.75, 2.75, 0.0), flush_l2=False)
            cl, ch = job.clip_counts()
            report("stack<%s>" % name, "%d x %dx%d fp32, sigma 2.75/2.75 (inputs %.1f GiB > L2)" % (n, width, rows, 4.0 * n * pixels / 2**30),
                   4.0 * (n + 1) * pixels, ms, {"mpx_in_per_s": n * pixels / ms / 1e3, "clipped": [cl, ch]})
        job.close()
        del out

    # ---- BASELINE config 4 depth: 1024 frames, linear fit (a row-stripe sample of the 8192-wide image) ----
    for n512 in ((512,) if "n512" in only else ()):
        n, pixels = 512, 4096 * 32
        job = nl.StackJob(ctx, n, pixels)
        job.synth_fill()
        out = torch.empty(pixels, dtype=torch.float32, device=dev)
        ms = timed(lambda: job.run_dev(nl.ST_SIGMA, out.data_ptr(), None, 2.75, 2.75, 0.0), flush_l2=False, reps=3)
        report("stack<sigma> n=512 tile=%s" % args.tile_width, "%d x 4096x32" % n, 4.0 * (n + 1) * pixels, ms)
        job.close()
        del out
    if "linfit1024big" in only:
        # BASELINE configs[3] depth on a 128-row stripe (1 M pixels, 4 GiB): enough columns to fill the GPU
        n, pixels = 1024, 8192 * 128
        job = nl.StackJob(ctx, n, pixels)
        job.synth_fill()
        out = torch.empty(pixels, dtype=torch.float32, device=dev)
        ms = timed(lambda: job.run_dev(nl.ST_LINEAR_FIT, out.data_ptr(), None, 2.75, 2.75, 0.0), flush_l2=False, reps=args.reps, warm=1)
        report("stack<linfit> n=1024, 1M px", "%d x 8192x128 fp32 (inputs %.2f GiB > L2)" % (n, 4.0 * n * pixels / 2**30),
               4.0 * (n + 1) * pixels, ms, {"mpx_in_per_s": n * pixels / ms / 1e3})
        job.close()
        del out
    if "linfit1024" in only or (not only and not args.quick):
        n, pixels = 1024, 8192 * 8
        job = nl.StackJob(ctx, n, pixels)
        job.synth_fill()
        out = torch.empty(pixels, dtype=torch.float32, device=dev)
        for name, mode in (("linfit", nl.ST_LINEAR_FIT), ("sigma", nl.ST_SIGMA)):
            ms = timed(lambda: job.run_dev(mode, out.data_ptr(), None, 2.75, 2.75, 0.0), flush_l2=False, reps=3)
            report("stack<%s> n=1024 tile=%s" % (name, args.tile_width), "%d x 8192x8 fp32 (inputs %.2f GiB > L2)" % (n, 4.0 * n * pixels / 2**30),
                   4.0 * (n + 1) * pixels, ms, {"mpx_in_per_s": n * pixels / ms / 1e3})
        job.close()
        del out

    # ---- resample: 6000x4000 -> 6000x4000, rotation 0.5 deg + shift ---------------------------------
    if not only or "project" in only or "fits" in only:
        w, h = (3000, 2000) if args.quick else (6000, 4000)
        srcs = [torch.empty(w * h, dtype=torch.float32, device=dev) for _ in range(ROT)]
        dsts = [torch.empty(w * h, dtype=torch.float32, device=dev) for _ in range(ROT)]
        for k, t in enumerate(srcs):
            ctx.synth_fill(t.data_ptr(), 0, w * h, k)
        th = np.deg2rad(0.5)
        trans = (C.c_float * 6)(np.cos(th), -np.sin(th), 7.25, np.sin(th), np.cos(th), -3.5)
        rot = "%d frames of %dx%d in rotation (%.1f GB > L2), no flush" % (ROT, w, h, ROT * 8.0 * w * h / 1e9)
        if not only or "project" in only:
            ms = timed_rot([lambda a=a, b=b: nl.binding.check(lib.nl_project_dev(ctx.handle, C.c_void_p(a.data_ptr()), w, h, C.c_void_p(b.data_ptr()),
                                                                                 w, h, trans, float("nan"))) for a, b in zip(srcs, dsts)])
            report("project_kernel", "rot 0.5 deg + shift; " + rot, 8.0 * w * h, ms, {"mpx_per_s": w * h / ms / 1e3})
        # ---- N1: resample fused with the histogram match, FITS payload decode (16-bit) and encode ---------
        if not only or "fits" in only:
            ms = timed_rot([lambda a=a, b=b: nl.binding.check(lib.nl_project_scaled_dev(ctx.handle, C.c_void_p(a.data_ptr()), w, h,
                                                                                        C.c_void_p(b.data_ptr()), w, h, trans, float("nan"), 1.03, -5.0))
                            for a, b in zip(srcs, dsts)])
            report("project_kernel<scaled> (match histogram fused)", rot, 8.0 * w * h, ms)
            raws = [torch.randint(-32768, 32767, (w * h,), dtype=torch.int16, device=dev) for _ in range(2 * ROT)]
            ms = timed_rot([lambda r=r, b=dsts[i % ROT]: nl.binding.check(lib.nl_fits_decode_dev(ctx.handle, C.c_void_p(r.data_ptr()), 16, w * h, 1.0, 32768.0,
                                                                                                  C.c_void_p(b.data_ptr()))) for i, r in enumerate(raws)])
            report("fits_decode_kernel<16>", "%d frames of %d samples in rotation, no flush" % (2 * ROT, w * h), 6.0 * w * h, ms)
            ms = timed_rot([lambda a=a, b=b: nl.binding.check(lib.nl_fits_encode_dev(ctx.handle, C.c_void_p(a.data_ptr()), w * h, C.c_void_p(b.data_ptr())))
                            for a, b in zip(srcs, dsts)])
            report("fits_encode_kernel", rot, 8.0 * w * h, ms)
            del raws
        del srcs, dsts

    # ---- star candidate scan and frame statistics: 6000x4000 sky noise + 0.02 % bright pixels, frames in rotation ----
    if not only or "bright" in only or "prestats" in only:
        w, h = (3000, 2000) if args.quick else (6000, 4000)
        g = torch.Generator(device=dev).manual_seed(7)
        imgs = []
        for k in range(2 * ROT):
            img = torch.randn(w * h, dtype=torch.float32, device=dev, generator=g) * 30 + 1000
            img += (torch.rand(w * h, device=dev, generator=g) < 2e-4).float() * 5000
            imgs.append(img)
        tmps = [torch.empty(w * h, dtype=torch.float32, device=dev) for _ in range(ROT)]
        rot = "%d frames of %dx%d in rotation (%.1f GB > L2), no flush" % (2 * ROT, w, h, 2 * ROT * 4.0 * w * h / 1e9)
        if not only or "bright" in only:
            cap = w * h // 50
            out = np.zeros(cap, dtype=nl.STAR_DTYPE)
            cnt = C.c_int32()
            ms = timed_rot([lambda im=im: nl.binding.check(lib.nl_find_bright_dev(ctx.handle, C.c_void_p(im.data_ptr()), w * h, w, 3000.0, 16,
                                                                                  out.ctypes.data_as(C.c_void_p), cap, C.byref(cnt))) for im in imgs])
            report("find_bright (2 scans + offsets, one host round trip)", "radius 16, %d candidates; %s" % (cnt.value, rot),
                   4.0 * w * h, ms, {"mpx_per_s": w * h / ms / 1e3, "note": "whole call incl. host sync; the image is read twice (count, write)"})
        if not only or "prestats" in only:
            st = (C.c_float * 4)()
            one = (C.c_float * 1)()
            for numerics, tag in ((nl.NUMERICS_AMD64, "amd64"), (nl.NUMERICS_PUREGO, "purego")):
                ctx.set_numerics(numerics)
                ms = timed_rot([lambda im=im: nl.binding.check(lib.nl_estimate_noise_dev(ctx.handle, C.c_void_p(im.data_ptr()), 1, w * h, w, h, one))
                                for im in imgs])
                report("estimate_noise (%s order; rows kernel + finalize + D2H)" % tag, rot, 4.0 * w * h, ms)
                ms = timed_rot([lambda im=im, t=tmps[i % ROT]: nl.binding.check(lib.nl_median_filter3x3_dev(ctx.handle, C.c_void_p(im.data_ptr()), w, h,
                                                                                                             C.c_void_p(t.data_ptr()))) for i, im in enumerate(imgs)])
                report("median3x3_kernel (%s)" % tag, rot, 8.0 * w * h, ms)
                r0 = ctx.exact_replays()
                ms = timed_rot([lambda im=im: nl.binding.check(lib.nl_stats_dev(ctx.handle, C.c_void_p(im.data_ptr()), w * h, st)) for im in imgs])
                report("stats min/mean/max/stddev (%s; 2 fused passes, one read-back)" % tag, rot, 8.0 * w * h, ms,
                       {"exact_replays": ctx.exact_replays() - r0})
                cap = w * h // 50
                bpm = np.empty(cap, dtype=np.int32)
                cnt = C.c_int64()
                r0 = ctx.exact_replays()
                ms = timed_rot([lambda im=im, t=tmps[i % ROT]: nl.binding.check(lib.nl_bad_pixel_map_dev(
                    ctx.handle, C.c_void_p(im.data_ptr()), w * h, w, 3.0, 5.0, C.c_void_p(t.data_ptr()), bpm.ctypes.data_as(C.POINTER(C.c_int32)),
                    cap, C.byref(cnt), st)) for i, im in enumerate(imgs)])
                report("bad_pixel_map whole call (%s; median-diff, stats, 2 scans)" % tag, "%d bad pixels; %s" % (cnt.value, rot),
                       24.0 * w * h, ms, {"exact_replays": ctx.exact_replays() - r0})
            ctx.set_numerics(nl.NUMERICS_AMD64)
            # the in-order replay of the float64 chains (taken when neither proof decides), forced for the measurement
            img2 = imgs[0]
            ctx.set_tuning("stats_force_replay", "1")
            r0 = ctx.exact_replays()
            ms = timed(lambda: nl.binding.check(lib.nl_stats_dev(ctx.handle, C.c_void_p(img2.data_ptr()), w * h, st)), reps=3, warm=1)
            ctx.set_tuning("stats_force_replay", "0")
            report("stats with both chains replayed in order (worst case)", "%dx%d" % (w, h), 8.0 * w * h, ms,
                   {"exact_replays": ctx.exact_replays() - r0})
            del img2
        del imgs, tmps

    # ---- stack-of-stacks accumulate -----------------------------------------------------------------
    if not only or "incremental" in only:
        px = 4096 * (1024 if args.quick else 4096)
        acc = torch.zeros(px, dtype=torch.float32, device=dev)
        light = torch.ones(px, dtype=torch.float32, device=dev)
        ms = timed(lambda: nl.binding.check(lib.nl_stack_incremental_dev(ctx.handle, C.c_void_p(acc.data_ptr()), C.c_void_p(light.data_ptr()),
                                                                         px, 3.0, 0)))
        report("incremental_kernel", "%d px acc += light*w, L2 flushed" % px, 12.0 * px, ms)
    ctx.close()


if __name__ == "__main__":
    main()
