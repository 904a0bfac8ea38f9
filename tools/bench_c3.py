"""bench.py --config c3: BASELINE.json configs[2] as one measured pipeline on one GPU.

    star detection (OpStarDetect / FindStars) -> [alignment: host, out of scope, the transforms are inputs] ->
    bilinear resample (OpAlign / Image.Project) -> stack (OpStack, StAuto -> linear fit for 64 frames)

over 64 frames of 6000x4000 fp32 that are resident in HBM, through the batched C-ABI entry points
(nl_bad_pixel_map_batch_dev, nl_find_stars_batch_dev, nl_project_batch_dev, nl_stack_run_dev).  Reference:
internal/ops/pre/preprocess.go:440-465, internal/star/findstars.go:59-100, internal/ops/post/postprocess.go:142-191,
internal/fits/project.go:26-76, internal/ops/stack/stack.go:115-227.

The frames are a synthetic star field (Gaussian blobs at hashed positions, flat background, Gaussian noise, a few hot
pixels) rendered on the host once per frame at the frame's own affine pose; the poses are what the alignment would find.
"""
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H, N = 6000, 4000, 64
BACKGROUND, NOISE = 1000.0, 12.0
STAR_SIG, BP_SIGMA, IN_OUT, RADIUS = 15.0, 5.0, 1.4, 16
METRIC = "Mpixels/s through star-detect + resample + stack (input samples N*P/t; 64 x 6000x4000 fp32, linear-fit stack)"


def poses(n, seed=7):
    """Transform2D of every frame (frame -> reference), small rotations and shifts like a night of tracking drift"""
    rng = np.random.default_rng(seed)
    t = np.zeros((n, 6), np.float32)
    for k in range(n):
        th = np.deg2rad(rng.uniform(-0.4, 0.4)) if k else 0.0
        dx, dy = (rng.uniform(-25, 25), rng.uniform(-25, 25)) if k else (0.0, 0.0)
        t[k] = [np.cos(th), -np.sin(th), dx, np.sin(th), np.cos(th), dy]
    return t


def render_frame(k, trans, stars, w=W, h=H):
    """frame k: the star list seen through the inverse of its pose, plus background, noise and hot pixels"""
    rng = np.random.default_rng(1000 + k)
    img = rng.standard_normal((h, w), dtype=np.float32)
    img *= np.float32(NOISE)
    img += np.float32(BACKGROUND)
    a, b, c, d, e, f = [float(x) for x in trans]
    det = a * e - b * d
    for (x, y, amp, s) in stars:
        # reference position (x, y) = T(frame position): frame position = T^-1
        fx = (e * (x - c) - b * (y - f)) / det
        fy = (-d * (x - c) + a * (y - f)) / det
        x0, x1, y0, y1 = int(fx) - 12, int(fx) + 13, int(fy) - 12, int(fy) + 13
        if x0 < 0 or y0 < 0 or x1 > w or y1 > h:
            continue
        yy, xx = np.mgrid[y0:y1, x0:x1]
        img[y0:y1, x0:x1] += (amp * np.exp(-((xx - fx) ** 2 + (yy - fy) ** 2) / (2 * s * s))).astype(np.float32)
    hot = rng.integers(0, w * h, 300)
    img.reshape(-1)[hot] += np.float32(8000.0)
    return img.reshape(-1)


def star_list(n_stars, seed=3, w=W, h=H):
    rng = np.random.default_rng(seed)
    return [(rng.uniform(40, w - 40), rng.uniform(40, h - 40), float(np.exp(rng.uniform(np.log(150), np.log(30000)))),
             rng.uniform(1.2, 2.8)) for _ in range(n_stars)]


def render_all(n, trans, stars, out, w=W, h=H):
    cores = os.cpu_count() or 1

    def work(k0):
        for k in range(k0, n, cores):
            out[k] = render_frame(k, trans[k], stars, w, h)

    th = [threading.Thread(target=work, args=(k0,)) for k0 in range(min(cores, n))]
    [t.start() for t in th]
    [t.join() for t in th]


def run_c3(args):
    import torch
    import nightlight_b200 as nl
    from bench import ClockSampler, peaks, source_hash, UNIT

    w, h, n = W, H, N
    if args.rows:                       # smaller frames for quick checks: --rows = frame height
        h = args.rows
    px = w * h
    lib = nl.load_library()
    torch.cuda.set_device(0)
    ctx = nl.Context(0)
    for kv in [x for x in (getattr(args, "tune", "") or "").split(",") if x]:      # nl_ctx_set_tuning knobs (A/B measurements)
        k, v = kv.split("=", 1)
        ctx.set_tuning(k, v.replace(":", ","))
    ext = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
    fp = C.POINTER(C.c_float)

    # ---- the frame set in pinned host memory
    t0 = time.perf_counter()
    host = C.c_void_p()
    nl.binding.check(lib.nl_host_alloc_pinned(4 * n * px, C.byref(host)))
    frames = np.ctypeslib.as_array(C.cast(host, fp), shape=(n, px))
    trans = poses(n)
    stars = star_list(2500 * h // H + 50, w=w, h=h)
    render_all(n, trans, stars, frames, w, h)
    gen_s = time.perf_counter() - t0

    raw = ctx.dev_alloc(4 * n * px)                      # the frames as loaded (after calibration), resident
    job = nl.StackJob(ctx, n, px)                        # the aligned frames, resident
    jbase, jstride = job.frames_dev
    out_dev = ctx.dev_alloc(4 * px)
    host_out = np.empty(px, np.float32)
    loc = np.full(n, BACKGROUND, np.float32)             # Stats.Location / Scale: randomized estimators in the reference, inputs here
    scale = np.full(n, NOISE, np.float32)
    cap = 20000
    found = np.zeros((n, cap), dtype=nl.STAR_DTYPE)
    counts = np.zeros(n, np.int32)
    sos, hfr = np.zeros(n, np.float32), np.zeros(n, np.float32)
    bcounts = np.zeros(n, np.int64)
    bstats = np.zeros((n, 4), np.float32)
    ptrs = (C.c_void_p * n)(*[host.value + 4 * k * px for k in range(n)])
    mode = lib.nl_auto_select_mode(n)                    # StAuto: linear fit from 25 frames up
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def rec(e):
        with torch.cuda.stream(ext):
            e.record()

    def upload():
        nl.binding.check(lib.nl_memcpy_h2d(ctx.handle, C.c_void_p(raw), host, 4 * n * px))

    def pipeline(timers):
        """device-resident pipeline; timers: dict of lists (ms)"""
        ea = ev()
        rec(ea)                                          # device timestamp before the first stage (the library's stream)
        t = time.perf_counter()
        nl.binding.check(lib.nl_bad_pixel_map_batch_dev(ctx.handle, C.c_void_p(raw), n, px, px, w, 3.0, 5.0, None, 0,
                                                        bcounts.ctypes.data_as(C.POINTER(C.c_int64)), bstats.ctypes.data_as(fp)))
        t1 = time.perf_counter()
        mds = np.ascontiguousarray(bstats[:, 3])
        td, thost = C.c_double(), C.c_double()
        nl.binding.check(lib.nl_find_stars_batch_dev(ctx.handle, C.c_void_p(raw), n, px, ptrs, px, w, loc.ctypes.data_as(fp),
                                                     scale.ctypes.data_as(fp), STAR_SIG, BP_SIGMA, IN_OUT, RADIUS, mds.ctypes.data_as(fp),
                                                     found.ctypes.data_as(C.c_void_p), cap, counts.ctypes.data_as(C.POINTER(C.c_int32)),
                                                     sos.ctypes.data_as(fp), hfr.ctypes.data_as(fp), C.byref(td), C.byref(thost)))
        t2 = time.perf_counter()
        e0, e1, e2 = ev(), ev(), ev()
        rec(e0)
        nl.binding.check(lib.nl_project_batch_dev(ctx.handle, C.c_void_p(raw), px, w, h, C.c_void_p(jbase), jstride, w, h, n,
                                                  trans.ctypes.data_as(fp), float("nan"), None, None))
        rec(e1)
        job.run_dev(mode, out_dev, None, 2.75, 2.75, 0.0)
        rec(e2)
        ctx.sync()
        t3 = time.perf_counter()
        timers["badpixel_map_ms"].append((t1 - t) * 1e3)
        timers["detect_device_ms"].append(td.value * 1e3)
        timers["detect_host_ms"].append(thost.value * 1e3)
        timers["detect_ms"].append((t2 - t1) * 1e3)
        timers["resample_ms"].append(e0.elapsed_time(e1))
        timers["stack_ms"].append(e1.elapsed_time(e2))
        timers["total_ms"].append((t3 - t) * 1e3)
        timers["total_device_ms"].append(ea.elapsed_time(e2))      # the same pass between two events on the device

    upload()
    ctx.sync()
    names = ["badpixel_map_ms", "detect_device_ms", "detect_host_ms", "detect_ms", "resample_ms", "stack_ms", "total_ms", "total_device_ms"]
    warm = {k: [] for k in names}
    for _ in range(max(3, args.warmup) if not args.rows else 1):
        pipeline(warm)
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = ctx.launch_count
    timers = {k: [] for k in names}
    steps = args.steps
    for _ in range(steps):
        pipeline(timers)
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    med = {k: float(np.median(v)) for k, v in timers.items()}
    clip = job.clip_counts()
    # the step is timed on the device: two events on the library's stream bracket the pass, host stages included (the
    # wall clock around the same calls is total_ms in stages_ms)
    value = n * px / (med["total_device_ms"] * 1e-3) / 1e6

    # ---- end to end: upload of the frame set, the pipeline, download of the stacked image.  The per-frame stages
    # (bad-pixel statistics, star detection) of a chunk of frames run while the next chunk uploads on a second
    # context's stream; the resample and the stack need every frame.
    up = nl.Context(0)
    chunk = max(1, n // 8)                               # (4 chunks: 161 ms, 8: 154 ms, 16: 155 ms)
    found2 = np.zeros((n, cap), dtype=nl.STAR_DTYPE)
    counts2 = np.zeros(n, np.int32)
    bstats2 = np.zeros((n, 4), np.float32)

    def upload_chunk(c0, c1):
        nl.binding.check(lib.nl_memcpy_h2d(up.handle, C.c_void_p(raw + 4 * c0 * px), C.c_void_p(host.value + 4 * c0 * px), 4 * (c1 - c0) * px))

    def detect_chunk(c0, c1):
        m = c1 - c0
        nl.binding.check(lib.nl_bad_pixel_map_batch_dev(ctx.handle, C.c_void_p(raw + 4 * c0 * px), m, px, px, w, 3.0, 5.0, None, 0,
                                                        bcounts[c0:c1].ctypes.data_as(C.POINTER(C.c_int64)), bstats2[c0:c1].ctypes.data_as(fp)))
        mds = np.ascontiguousarray(bstats2[c0:c1, 3])
        td, thost = C.c_double(), C.c_double()
        cptrs = (C.c_void_p * m)(*[host.value + 4 * k * px for k in range(c0, c1)])
        nl.binding.check(lib.nl_find_stars_batch_dev(ctx.handle, C.c_void_p(raw + 4 * c0 * px), m, px, cptrs, px, w, loc[c0:c1].ctypes.data_as(fp),
                                                     scale[c0:c1].ctypes.data_as(fp), STAR_SIG, BP_SIGMA, IN_OUT, RADIUS, mds.ctypes.data_as(fp),
                                                     found2[c0:c1].ctypes.data_as(C.c_void_p), cap, counts2[c0:c1].ctypes.data_as(C.POINTER(C.c_int32)),
                                                     sos[c0:c1].ctypes.data_as(fp), hfr[c0:c1].ctypes.data_as(fp), C.byref(td), C.byref(thost)))

    e2e_ms = []
    for _ in range(max(1, min(steps, args.e2e_steps))):
        found2[:] = 0
        t = time.perf_counter()
        bounds = [(c0, min(n, c0 + chunk)) for c0 in range(0, n, chunk)]
        upload_chunk(*bounds[0])
        up.sync()
        for i, (c0, c1) in enumerate(bounds):
            if i + 1 < len(bounds):
                upload_chunk(*bounds[i + 1])
            detect_chunk(c0, c1)
            up.sync()
        nl.binding.check(lib.nl_project_batch_dev(ctx.handle, C.c_void_p(raw), px, w, h, C.c_void_p(jbase), jstride, w, h, n,
                                                  trans.ctypes.data_as(fp), float("nan"), None, None))
        job.run_dev(mode, out_dev, None, 2.75, 2.75, 0.0)
        ctx.d2h(host_out, out_dev)
        e2e_ms.append((time.perf_counter() - t) * 1e3)
    # the chunked run found what the whole-set run found
    if not (np.array_equal(counts2, counts) and found2.tobytes() == found.tobytes() and bstats2.tobytes() == bstats.tobytes()):
        raise SystemExit("c3: the chunked end-to-end run differs from the resident run")
    up.close()
    e2e = {"value": n * px / (float(np.median(e2e_ms)) * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": 4 * n * px,
           "d2h_bytes_per_step": 4 * px + n * (24 * int(counts.max()) + 16), "ms_per_step": float(np.median(e2e_ms)),
           "steps": len(e2e_ms), "host_memory": "pinned", "chunks": len(bounds),
           "api": "frames uploaded in %d chunks (nl_memcpy_h2d on a second context); per chunk nl_bad_pixel_map_batch_dev + "
                  "nl_find_stars_batch_dev while the next chunk uploads; then nl_project_batch_dev, nl_stack_run_dev, download of the "
                  "stacked image" % len(bounds)}

    # ---- parity and CPU baseline on a bounded sample (the oracle = restatement of the Go code)
    parity, cpu = {}, None
    if not args.no_cpu:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        k = 1
        t = time.perf_counter()
        want = O.find_stars(frames[k], w, BACKGROUND, NOISE, STAR_SIG, BP_SIGMA, IN_OUT, RADIUS, float(bstats[k, 3]))
        cpu_detect = time.perf_counter() - t
        stars_ok = counts[k] == len(want[0]) and found[k, :counts[k]].tobytes() == want[0].tobytes()
        t = time.perf_counter()
        wantp = O.project(frames[k], w, h, w, h, trans[k], np.float32(np.nan))
        cpu_project = time.perf_counter() - t
        gotp = np.empty(px, np.float32)
        ctx.d2h(gotp, jbase + 4 * k * jstride)
        gn, wn = np.isnan(gotp), np.isnan(wantp)
        proj_ok = bool(np.array_equal(gn, wn) and np.array_equal(gotp.view(np.uint32)[~gn], wantp.view(np.uint32)[~wn]))
        t = time.perf_counter()
        _, st_want, _ = O.bad_pixel_map(frames[k], w, 3.0, 5.0)
        cpu_bpm = time.perf_counter() - t
        bpm_ok = bool(np.array_equal(st_want.view(np.uint32), bstats[k].view(np.uint32)))
        # the stack: the first rows of the aligned frames as the GPU produced them, through the CPU linear fit
        srows = min(64, h)
        spx = srows * w
        al = np.empty((n, spx), np.float32)
        for i in range(n):
            ctx.d2h(al[i], jbase + 4 * i * jstride)
        t = time.perf_counter()
        sw_, cl_, ch_ = O.stack(al, mode, 2.75, 2.75, threads=cores)
        cpu_stack = time.perf_counter() - t
        got = np.empty(spx, np.float32)
        ctx.d2h(got, out_dev)
        gn, wn = np.isnan(got), np.isnan(sw_)
        stack_ok = bool(np.array_equal(gn, wn) and np.array_equal(got.view(np.uint32)[~gn], sw_.view(np.uint32)[~wn]))
        parity = {"frame_checked": k, "stars_bit_exact": bool(stars_ok), "stars": int(counts[k]), "resample_bit_exact": proj_ok,
                  "median_diff_stats_bit_exact": bpm_ok, "stack_rows_checked": srows, "stack_bit_exact": stack_ok}
        if not (stars_ok and proj_ok and bpm_ok and stack_ok):
            raise SystemExit("c3 parity failure: %s" % json.dumps(parity))
        # extrapolation: the per-frame stages run one frame per core (the reference's goroutine per frame), the stack on all cores
        rounds = (n + cores - 1) // cores
        cpu_total = (cpu_bpm + cpu_detect + cpu_project) * rounds + cpu_stack * (h / srows)
        cpu = {"value": n * px / cpu_total / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "bad-pixel map %.2f s, FindStars %.2f s, Project %.2f s on one frame (one thread each; x %d rounds of %d "
                         "frames in parallel), linear-fit stack of %d rows %.2f s on %d threads (x %d for the image)" % (
                             cpu_bpm, cpu_detect, cpu_project, rounds, cores, srows, cpu_stack, cores, h // srows)}

    peak, peak_src = peaks()
    algo = 4.0 * (n + 1) * px
    roofline = {"bound": "hbm", "achieved": algo / (med["stack_ms"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": algo / (med["stack_ms"] * 1e-3) / 1e9 / peak, "traffic": None,
                "kernel": "stack_column_kernel<linfit> (sort + rejection rounds, unfinished columns regrouped between launches)",
                "kernel_ms": med["stack_ms"], "algorithmic_bytes_per_launch": algo, "peak_source": peak_src,
                "resample": {"achieved": 8.0 * n * px / (med["resample_ms"] * 1e-3) / 1e9, "frac": 8.0 * n * px / (med["resample_ms"] * 1e-3) / 1e9 / peak,
                             "algorithmic_bytes": 8.0 * n * px, "kernel": "project_batch_kernel (one launch, 64 frames)"},
                "detect_scan": {"achieved": 4.0 * n * px / (med["detect_device_ms"] * 1e-3) / 1e9,
                                "frac": 4.0 * n * px / (med["detect_device_ms"] * 1e-3) / 1e9 / peak, "algorithmic_bytes": 4.0 * n * px,
                                "kernel": "bright_rows_slots_kernel + row_offsets_batch_kernel + bright_compact_kernel (whole call incl. two host round trips)"}}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": max(3, args.warmup), "ms_per_step": med["total_device_ms"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "star detection (bad-pixel statistics + FindStars) + bilinear resample + StAuto (linear fit) stack of %d x %dx%d fp32 "
                               "frames resident in HBM; alignment transforms are inputs (host step, out of scope)" % (n, w, h),
                   "config": "c3", "n_frames": n, "width": w, "height": h, "stages_ms": med, "stars_per_frame": [int(counts.min()), int(counts.max())],
                   "bad_pixels_per_frame": [int(bcounts.min()), int(bcounts.max())], "clipped": list(clip), "mode": int(mode),
                   "l2": "64 frames of %.0f MB each, every stage larger than L2" % (4 * px / 1e6), "host_generation_s": gen_s,
                   "parity": parity, "source_hash": source_hash()},
        "clocks": clocks, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "gpu_launches": launches,
    }
    print(json.dumps(line))
    ctx.dev_free(raw)
    ctx.dev_free(out_dev)
    job.close()
    lib.nl_host_free_pinned(host)
    ctx.close()
    return 0


def run_c3_multi(args):
    """configs[2] on N GPUs of one box: detection and resampling shard over FRAMES (rank r owns frames
    r, r+N, ...), the stack over ROW STRIPES; the exchange between the two is fused into the resample's stores
    (nl_project_scatter_dev writes every destination row into the peer-mapped stack job of the rank that owns it).
    One step = bad-pixel statistics + FindStars of the local frames, scatter-resample, barrier, linear-fit stack of the
    local stripe, all-gather of the stripes.  Launched by torchrun; rank 0 prints the line."""
    import torch
    import torch.distributed as dist
    import nightlight_b200 as nl
    from nightlight_b200 import stripes
    from bench import ClockSampler, peaks, source_hash, UNIT
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib = nl.load_library()
    ctx = nl.Context(local)
    fp = C.POINTER(C.c_float)
    n, w, h = N, W, H
    if args.rows:
        h = args.rows
    px = w * h
    ids = stripes.frame_shard(n, world, rank)
    m = len(ids)
    trans = poses(n)
    stars = star_list(2500 * h // H + 50, w=w, h=h)

    # ---- this rank's frames in pinned host memory, then resident
    t0 = time.perf_counter()
    host = C.c_void_p()
    nl.binding.check(lib.nl_host_alloc_pinned(4 * m * px, C.byref(host)))
    frames = np.ctypeslib.as_array(C.cast(host, fp), shape=(m, px))
    cores = max(1, (os.cpu_count() or 1) // world)

    def work(i0):
        for i in range(i0, m, cores):
            frames[i] = render_frame(ids[i], trans[ids[i]], stars, w, h)

    th = [threading.Thread(target=work, args=(i0,)) for i0 in range(min(cores, m))]
    [t.start() for t in th]
    [t.join() for t in th]
    gen_s = time.perf_counter() - t0
    raw = ctx.dev_alloc(4 * m * px)

    def upload():
        nl.binding.check(lib.nl_memcpy_h2d(ctx.handle, C.c_void_p(raw), host, 4 * m * px))

    sc = stripes.PeerScatter(ctx, nl.StackJob, n, w, h)
    row0, rows = sc.row0, sc.rows
    out_stripe = torch.empty(rows * w, dtype=torch.float32, device=dev)
    loc = np.full(m, BACKGROUND, np.float32)
    scale = np.full(m, NOISE, np.float32)
    cap = 20000
    found = np.zeros((m, cap), dtype=nl.STAR_DTYPE)
    counts = np.zeros(m, np.int32)
    sos, hfr = np.zeros(m, np.float32), np.zeros(m, np.float32)
    bcounts = np.zeros(m, np.int64)
    bstats = np.zeros((m, 4), np.float32)
    ptrs = (C.c_void_p * m)(*[host.value + 4 * i * px for i in range(m)])
    mode = lib.nl_auto_select_mode(n)
    state = {}

    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def step(seg):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record()                               # device timestamp before the first stage (the library's stream)
        t = time.perf_counter()
        nl.binding.check(lib.nl_bad_pixel_map_batch_dev(ctx.handle, C.c_void_p(raw), m, px, px, w, 3.0, 5.0, None, 0,
                                                        bcounts.ctypes.data_as(C.POINTER(C.c_int64)), bstats.ctypes.data_as(fp)))
        mds = np.ascontiguousarray(bstats[:, 3])
        td, thost = C.c_double(), C.c_double()
        nl.binding.check(lib.nl_find_stars_batch_dev(ctx.handle, C.c_void_p(raw), m, px, ptrs, px, w, loc.ctypes.data_as(fp),
                                                     scale.ctypes.data_as(fp), STAR_SIG, BP_SIGMA, IN_OUT, RADIUS, mds.ctypes.data_as(fp),
                                                     found.ctypes.data_as(C.c_void_p), cap, counts.ctypes.data_as(C.POINTER(C.c_int32)),
                                                     sos.ctypes.data_as(fp), hfr.ctypes.data_as(fp), C.byref(td), C.byref(thost)))
        t1 = time.perf_counter()
        for i, k in enumerate(ids):
            sc.project(raw + 4 * i * px, w, h, k, trans[k])
        sc.finish()                                   # every rank's rows have landed in every job
        t2 = time.perf_counter()
        sc.job.run_dev(mode, out_stripe.data_ptr(), None, 2.75, 2.75, 0.0)
        ctx.sync()
        t3 = time.perf_counter()
        state["full"] = stripes.allgather_image(out_stripe, w, h)
        e1.record()                                   # ... and after the all-gather (torch's stream)
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        state["device_ms"] = e0.elapsed_time(e1)
        seg["detect_ms"].append((t1 - t) * 1e3)
        seg["scatter_resample_ms"].append((t2 - t1) * 1e3)
        seg["stack_ms"].append((t3 - t2) * 1e3)
        seg["gather_ms"].append((t4 - t3) * 1e3)
        seg["total_ms"].append((t4 - t) * 1e3)

    names = ["detect_ms", "scatter_resample_ms", "stack_ms", "gather_ms", "total_ms"]

    def timed(fn, reps, on_device):
        """median over reps of the slowest rank's time for fn (barrier before): between two events on the device when
        fn is the step itself, wall clock when it also uploads and downloads"""
        out = []
        for _ in range(reps):
            dist.barrier()
            torch.cuda.synchronize()
            t = time.perf_counter()
            fn()
            mine = state["device_ms"] if on_device else (time.perf_counter() - t) * 1e3
            ms = torch.tensor([mine], device=dev, dtype=torch.float64)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            out.append(float(ms.item()))
        return float(np.median(out)), out

    upload()
    ctx.sync()
    warm = {k: [] for k in names}
    for _ in range(max(3, args.warmup) if not args.rows else 1):
        dist.barrier()
        step(warm)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launch_count
    seg = {k: [] for k in names}
    steps = args.steps
    ms_step, _ = timed(lambda: step(seg), steps, True)
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    med = {k: float(np.median(v)) for k, v in seg.items()}
    host_full = np.empty(px, np.float32) if rank == 0 else None

    def e2e_step():
        upload()
        step({k: [] for k in names})
        if rank == 0:
            ctx.sync()
            host_full[:] = state["full"].cpu().numpy()

    ms_e2e, _ = timed(e2e_step, max(1, min(steps, args.e2e_steps)), False)

    # ---- parity on rank 0 (bounded CPU samples), and every rank's stripe arrived in the gathered image
    mine = state["full"][row0 * w:(row0 + rows) * w]
    ok_t = torch.tensor([1 if torch.equal(mine.view(torch.int32), out_stripe.view(torch.int32)) else 0], device=dev)
    dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
    parity = {}
    if rank == 0 and not args.no_cpu:
        from oracle import oracle as O
        i = 1 if m > 1 else 0
        want = O.find_stars(frames[i], w, BACKGROUND, NOISE, STAR_SIG, BP_SIGMA, IN_OUT, RADIUS, float(bstats[i, 3]))
        stars_ok = counts[i] == len(want[0]) and found[i, :counts[i]].tobytes() == want[0].tobytes()
        jbase, jstride = sc.job.frames_dev

        def stripe_of_frame(k, nrows):
            a = np.empty(nrows * w, np.float32)
            ctx.d2h(a, jbase + 4 * k * jstride)
            return a

        def same(a, b):
            an, bn = np.isnan(a), np.isnan(b)
            return bool(np.array_equal(an, bn) and np.array_equal(a.view(np.uint32)[~an], b.view(np.uint32)[~bn]))

        # a frame this rank resampled itself and one that arrived from a peer (rendered again here for the CPU side)
        k_own, k_peer = ids[i], (1 if world > 1 else ids[0])
        proj_ok = True
        for k in (k_own, k_peer):
            src = frames[ids.index(k)] if k in ids else render_frame(k, trans[k], stars, w, h)
            wantp = O.project(np.ascontiguousarray(src), w, h, w, h, trans[k], np.float32(np.nan))
            proj_ok = proj_ok and same(stripe_of_frame(k, rows), wantp[row0 * w:(row0 + rows) * w])
        srows = min(32, rows)
        al = np.stack([stripe_of_frame(k, srows) for k in range(n)])
        sw_, _, _ = O.stack(al, mode, 2.75, 2.75, threads=os.cpu_count() or 1)
        stack_ok = same(out_stripe[:srows * w].cpu().numpy(), sw_)
        parity = {"frame_checked": int(k_own), "stars_bit_exact": bool(stars_ok), "stars": int(counts[i]),
                  "resample_bit_exact_own_and_peer_frame": [int(k_own), int(k_peer), bool(proj_ok)],
                  "stack_rows_checked": srows, "stack_bit_exact": stack_ok, "stripes_in_gathered_image": bool(ok_t.item())}
        if not (stars_ok and proj_ok and stack_ok and ok_t.item()):
            raise SystemExit("c3 multi-GPU parity failure: %s" % json.dumps(parity))
    if rank == 0:
        peak, peak_src = peaks()
        algo = 4.0 * (n + 1) * rows * w
        line = {
            "metric": METRIC, "value": n * px / (ms_step * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "star detection + resample sharded over frames (%d per GPU), linear-fit stack sharded over row stripes "
                                   "(%d rows per GPU), %d x %dx%d fp32; the frame->stripe exchange is fused into the resample's stores "
                                   "(peer-mapped stack jobs), stripes all-gathered with NCCL" % (m, rows, n, w, h),
                       "config": "c3", "n_frames": n, "width": w, "height": h, "segments_ms_rank0": med, "mode": int(mode),
                       "timing": "two events on the device around the step (host stages included), barrier before, max over ranks", "host_generation_s": gen_s,
                       "parity": parity, "source_hash": source_hash()},
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": algo / (med["stack_ms"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": algo / (med["stack_ms"] * 1e-3) / 1e9 / peak, "traffic": None, "kernel_ms": med["stack_ms"],
                         "kernel": "stack_column_kernel<linfit> on this rank's stripe", "algorithmic_bytes_per_launch": algo,
                         "peak_source": peak_src},
            "e2e": {"value": n * px / (ms_e2e * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": 4 * n * px,
                    "d2h_bytes_per_step": 4 * px, "host_memory": "pinned",
                    "api": "every rank uploads its frame shard (nl_memcpy_h2d), the step, rank 0 downloads the gathered image"},
            "cpu_baseline": None, "gpu_launches": int(launches) * world,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    ctx.dev_free(raw)
    sc.close()
    lib.nl_host_free_pinned(host)
    ctx.close()
    dist.destroy_process_group()
    return 0
