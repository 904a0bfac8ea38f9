#!/usr/bin/env python
"""bench.py -- the stacking hot path on BASELINE.json's headline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one sigma-clip stacking pass (sigma 2.75/2.75, unweighted) over 256 synthetic 4096x4096 fp32
frames (BASELINE.json configs[1] geometry with the metric's sigma-clip mode).  `value` is measured with
the frames already resident in HBM (CUDA events on the library's stream); `e2e` is the same pass through
the C ABI with HOST buffers: 256 frame uploads from pinned memory, the kernel, and the download of the
stacked image, all inside the timed region.  Multi-GPU: one process per GPU, every rank owns a row stripe
of all frames; weak scaling (each rank stacks a full 256x4096x4096 stripe of a 4096 x 4096*N image), no
data-path collective, one NCCL all-gather to reassemble the stacked image inside the timed step.
PyTorch is used for torch.distributed, CUDA events and the NCCL all-gather only.

`--impl reference` times the CPU restatement of the reference (oracle/, all host threads, the reference's
own 8 MiB work packages) on a bounded sample of the same workload; the Go reference itself cannot be
built here (no Go toolchain, un-vendored modules).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FRAMES, WIDTH, HEIGHT = 256, 4096, 4096
SIG_LO = SIG_HI = 2.75
METRIC = "Mpixels/s stacked (input samples N*P/t; sigma-clip, 256x4096^2 fp32)"
UNIT = "Mpx/s"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle (restatement of the reference's Go code) on a bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_stack_sample(rows, reps_budget_s, max_reps):
    """-> (Mpx/s, cores, sample description, seconds per rep)"""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    pixels = WIDTH * rows
    frames = np.empty((N_FRAMES, pixels), dtype=np.float32)
    fp = C.POINTER(C.c_float)
    lib = O.lib()
    # generate the stripe with all cores (the generator is not part of the measurement)
    def gen(k0, k1):
        for k in range(k0, k1):
            lib.nlo_synth_frame(frames[k].ctypes.data_as(fp), 0, pixels, k, 12345)
    th = [threading.Thread(target=gen, args=(k0, min(N_FRAMES, k0 + (N_FRAMES + cores - 1) // cores)))
          for k0 in range(0, N_FRAMES, (N_FRAMES + cores - 1) // cores)]
    [t.start() for t in th]; [t.join() for t in th]
    ptrs = (fp * N_FRAMES)(*[frames[k].ctypes.data_as(fp) for k in range(N_FRAMES)])
    res = np.empty(pixels, dtype=np.float32)
    cl, ch = C.c_int64(), C.c_int64()
    times = []
    t_all = time.perf_counter()
    while len(times) < max_reps and (not times or time.perf_counter() - t_all < reps_budget_s):
        t0 = time.perf_counter()
        rc = lib.nlo_stack_apply(2, ptrs, N_FRAMES, pixels, None, 0.0, SIG_LO, SIG_HI, res.ctypes.data_as(fp),
                                 C.byref(cl), C.byref(ch), cores)
        assert rc == 0
        times.append(time.perf_counter() - t0)
    best = float(np.median(times))
    sample = "%d frames x %dx%d rows (1/%d of the workload), median of %d passes, %d threads" % (
        N_FRAMES, WIDTH, rows, HEIGHT // rows, len(times), cores)
    return N_FRAMES * pixels / best / 1e6, cores, sample, best, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    rows = args.cpu_rows
    # warm-up passes are the first W reps; K timed passes
    from oracle import oracle as O  # noqa: F401  (build the oracle before timing)
    mpx, cores, sample, sec, times = cpu_stack_sample(rows, 1e9, args.warmup + args.steps)
    timed = times[args.warmup:] or times
    sec = float(np.mean(timed))
    value = N_FRAMES * WIDTH * rows / sec / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(timed),
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "sigma-clip stack 2.75/2.75, 256 x 4096x4096 fp32 (each step: a %d-row stripe)" % rows,
                   "n_frames": N_FRAMES, "width": WIDTH, "height": HEIGHT, "sample_rows": rows},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d frames x %dx%d rows per step, %d threads, C restatement of the Go reference "
                                   "(Go toolchain absent)" % (N_FRAMES, WIDTH, rows, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import nightlight_b200 as nl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch --gpus %d under torch.distributed.run (one process per GPU)" % args.gpus)
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    rows = args.rows
    pixels = WIDTH * rows
    ctx = nl.Context(local_rank)
    ext = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    job = nl.StackJob(ctx, N_FRAMES, pixels)
    # weak scaling: rank r owns rows [r*rows, (r+1)*rows) of a WIDTH x rows*world image
    job.synth_fill(p0=rank * pixels)
    ctx.sync()
    out = torch.empty(pixels, dtype=torch.float32, device="cuda")
    gathered = torch.empty(pixels * world, dtype=torch.float32, device="cuda") if world > 1 else None
    # Multi-GPU reassembly of the stacked image: fused into the stack kernel's epilogue (every result is
    # also stored into the peer-mapped gathered image of every other rank over NVLink), with one tiny
    # all-reduce as the "everybody's stores have landed" signal.  --gather nccl keeps the plain
    # all_gather_into_tensor of the stripes; the fused path is verified against it once before timing.
    peer = None
    flag = torch.zeros(1, dtype=torch.int32, device="cuda") if world > 1 else None
    if world > 1 and args.gather == "peer":
        from nightlight_b200.stripes import PeerGather
        try:
            peer = PeerGather(ctx, pixels)
        except Exception as e:                       # no peer access on this box: say so and fall back
            if rank == 0:
                print("peer mapping unavailable (%s), using the NCCL all-gather" % e, file=sys.stderr)
            peer = None
        ok = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0 and peer is not None:
            peer.close()
            peer = None

    def step():
        if peer is not None:
            job.run_dev_bcast(nl.ST_SIGMA, peer.local_out, peer.peer_outs, None, SIG_LO, SIG_HI, 0.0)
            with torch.cuda.stream(ext):
                dist.all_reduce(flag)
            return
        job.run_dev(nl.ST_SIGMA, out.data_ptr(), None, SIG_LO, SIG_HI, 0.0)
        if world > 1:
            with torch.cuda.stream(ext):
                dist.all_gather_into_tensor(gathered, out)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    gather_desc = "none (1 GPU)"
    if world > 1:
        gather_desc = "NCCL all_gather_into_tensor of the stripes"
    if peer is not None:
        # verify the fused reassembly against NCCL once
        job.run_dev(nl.ST_SIGMA, out.data_ptr(), None, SIG_LO, SIG_HI, 0.0)
        with torch.cuda.stream(ext):
            dist.all_gather_into_tensor(gathered, out)
        barrier()
        same = bool(np.array_equal(peer.to_host().view(np.uint32), gathered.cpu().numpy().view(np.uint32)))
        agree = torch.tensor([1 if same else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        if int(agree.item()) != 1:
            raise SystemExit("fused peer-store reassembly differs from the NCCL all-gather")
        gather_desc = "fused: stack kernel epilogue stores every stripe into all peers' images (CUDA IPC over NVLink), verified against NCCL all_gather"
        barrier()

    # ---- timed region: K steps, device-resident frames (16 GiB per GPU >> 126 MB L2: no flush needed)
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 2)]
    with torch.cuda.stream(ext):
        ev[0].record()
    for i in range(args.steps):
        with torch.cuda.stream(ext):
            ev[2 + 2 * i].record()
        if peer is not None:
            job.run_dev_bcast(nl.ST_SIGMA, peer.local_out, peer.peer_outs, None, SIG_LO, SIG_HI, 0.0)
        else:
            job.run_dev(nl.ST_SIGMA, out.data_ptr(), None, SIG_LO, SIG_HI, 0.0)
        with torch.cuda.stream(ext):
            ev[3 + 2 * i].record()
            if peer is not None:
                dist.all_reduce(flag)
            elif world > 1:
                dist.all_gather_into_tensor(gathered, out)
    with torch.cuda.stream(ext):
        ev[1].record()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    total_ms = ev[0].elapsed_time(ev[1])
    kernel_ms = [ev[2 + 2 * i].elapsed_time(ev[3 + 2 * i]) for i in range(args.steps)]
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * N_FRAMES * pixels / (ms_per_step * 1e-3) / 1e6
    clip_low, clip_high = job.clip_counts()

    # ---- roofline of the dominant kernel (this rank): algorithmic bytes = 4*(N+1) per output pixel
    peak, peak_src = peaks()
    k_ms = float(np.mean(kernel_ms))
    algo_bytes = 4.0 * (N_FRAMES + 1) * pixels
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_stack_sigma.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("rows") == rows:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "stack_column_kernel<sigma> (per step: one launch over the frame stack + one over the pool of columns whose late clipping passes were deferred)", "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src}

    # ---- end to end through the C ABI with host buffers (rank-local; max over ranks)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, nl, ctx, job, pixels, world, dist, torch)
        clipped = e2e.pop("clipped")
        # one GPU: the end-to-end pass sees exactly the resident frames, so its clip totals must agree
        e2e["matches_resident_run"] = (clipped == [clip_low, clip_high]) if world == 1 else None

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        mpx, cores, sample, sec, _ = cpu_stack_sample(args.cpu_rows, 12.0, 6)
        cpu = {"value": mpx, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "sigma-clip stack 2.75/2.75 of 256 x 4096x%d fp32 frames per GPU (row stripe of a "
                                   "4096x%d image), frames resident in HBM" % (rows, rows * world),
                       "n_frames": N_FRAMES, "width": WIDTH, "rows_per_gpu": rows, "mode": "sigma", "sigma": [SIG_LO, SIG_HI],
                       "parallelism": "row stripes x%d" % world, "gather": gather_desc,
                       "l2": "inputs (%.1f GiB per GPU) larger than L2, no flush" % (4.0 * N_FRAMES * pixels / 2**30),
                       "mpx_out_per_s": world * pixels / (ms_per_step * 1e-3) / 1e6,
                       "clipped": [clip_low, clip_high]},
            "clocks": clocks, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "gpu_launches": launches,
        }
        print(json.dumps(line))
    if peer is not None:
        peer.close()
    job.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_e2e(args, nl, ctx, job, pixels, world, dist, torch):
    """One stacking pass through the public C ABI with HOST buffers: nl_stack_apply takes the 256 host frame
    pointers, cuts the image into row stripes and alternates them on two streams, so the upload of one
    stripe (pinned memory) overlaps the stacking of the previous one; the stacked image comes back to host
    memory.  All copies are inside the timed region."""
    lib = nl.load_library()
    # One GPU: all 256 frames in pinned host memory (16 GiB), result checked against the HBM-resident pass.
    # Several ranks on one host: 64 distinct frames per rank (4 GiB), every pointer k -> frame k % 64: the same
    # bytes cross PCIe per step with a quarter of the pinned memory.
    distinct = N_FRAMES if world == 1 else min(64, N_FRAMES)
    nbytes = 4 * distinct * pixels
    host = C.c_void_p()
    pinned = lib.nl_host_alloc_pinned(nbytes, C.byref(host)) == 0
    if not pinned:
        arr = np.empty(distinct * pixels, dtype=np.float32)
        host = C.c_void_p(arr.ctypes.data)
    host_out = C.c_void_p()
    out_pinned = lib.nl_host_alloc_pinned(4 * pixels, C.byref(host_out)) == 0
    if not out_pinned:
        oarr = np.empty(pixels, dtype=np.float32)
        host_out = C.c_void_p(oarr.ctypes.data)
    base, stride = job.frames_dev                # fill the host frames once from the device-generated ones
    nl.binding.check(lib.nl_memcpy_d2h(ctx.handle, host, C.c_void_p(base), nbytes))
    ctx.sync()
    ptrs = (C.c_void_p * N_FRAMES)(*[host.value + 4 * (k % distinct) * pixels for k in range(N_FRAMES)])
    cl, ch = C.c_int64(), C.c_int64()

    def e2e_step():
        nl.binding.check(lib.nl_stack_apply(ctx.handle, ptrs, N_FRAMES, pixels, WIDTH, args.e2e_stripes, nl.ST_SIGMA, None,
                                            SIG_LO, SIG_HI, 0.0, host_out, C.byref(cl), C.byref(ch)))

    e2e_step()                                   # warm-up
    if world > 1:
        dist.barrier()
    steps = max(1, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    sec = (time.perf_counter() - t0) / steps
    if world > 1:
        t = torch.tensor([sec], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    if pinned:
        lib.nl_host_free_pinned(host)
    if out_pinned:
        lib.nl_host_free_pinned(host_out)
    return {"value": world * N_FRAMES * pixels / sec / 1e6, "unit": UNIT, "h2d_bytes_per_step": 4 * N_FRAMES * pixels,
            "d2h_bytes_per_step": 4 * pixels + 16, "ms_per_step": sec * 1e3, "steps": steps,
            "host_memory": "pinned" if pinned else "pageable",
            "api": "nl_stack_apply: %d host frame pointers in, host image out; %d row stripes alternating on two streams "
                   "inside the library" % (N_FRAMES, args.e2e_stripes), "clipped": [cl.value, ch.value]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=HEIGHT, help="rows per GPU (default: the full 4096)")
    ap.add_argument("--cpu-rows", type=int, default=256, help="rows of the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-stripes", type=int, default=8, help="row stripes of the pipelined end-to-end pass")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"], help="multi-GPU reassembly of the stacked image")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
