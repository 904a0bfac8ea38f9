#!/usr/bin/env python
"""Frame-sharded resample -> row-sharded stack on N GPUs: checks the fused path (the resample
kernel stores into the peer-mapped stack jobs, stripes.PeerScatter) bit for bit against the plain one (resample into
a local image, NCCL all-to-all, copy into the job) and times both on the device.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/check_scatter.py [--frames 32] [--width 6000] [--height 4000]

Rank 0 prints one JSON line.  Every rank resamples its frame shard (frames r, r+N, ...) of synthetic frames with
per-frame affine transforms; afterwards every rank holds its row stripe of ALL frames and stacks it."""
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
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--width", type=int, default=6000)
    ap.add_argument("--height", type=int, default=4000)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import nightlight_b200 as nl
    from nightlight_b200 import stripes

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib = nl.load_library()
    ctx = nl.Context(local)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    w, h, n = args.width, args.height, args.frames
    ids = stripes.frame_shard(n, world, rank)
    row0, rows = stripes.stripe_rows(h, world, rank)

    def trans_of(k):
        th = np.deg2rad(0.05 * (k % 11) - 0.25)
        return np.array([np.cos(th), -np.sin(th), 1.5 * (k % 7) - 4, np.sin(th), np.cos(th), 2.25 - 0.75 * (k % 5)], np.float32)

    # this rank's source frames (synthetic, resident)
    src = torch.empty(len(ids), w * h, dtype=torch.float32, device=dev)
    for i, k in enumerate(ids):
        ctx.synth_fill(src[i].data_ptr(), 0, w * h, k)
    ctx.sync()

    sc = stripes.PeerScatter(ctx, nl.StackJob, n, w, h)

    def fused():
        for i, k in enumerate(ids):
            sc.project(src[i].data_ptr(), w, h, k, trans_of(k))

    # baseline: full local images, NCCL all-to-all, copy into a second job's frame buffer
    full = torch.empty(len(ids), w * h, dtype=torch.float32, device=dev)
    base_job = nl.StackJob(ctx, n, rows * w)
    base_frames = base_job.frames_dev[0]

    def baseline():
        for i, k in enumerate(ids):
            t = trans_of(k)
            nl.binding.check(lib.nl_project_dev(ctx.handle, C.c_void_p(src[i].data_ptr()), w, h, C.c_void_p(full[i].data_ptr()), w, h,
                                                t.ctypes.data_as(C.POINTER(C.c_float)), float("nan")))
        ctx.sync()
        mine = stripes.alltoall_frames_to_stripes(full, ids, n, w, h)
        torch.cuda.synchronize()
        nl.binding.check(lib.nl_memcpy_d2d(ctx.handle, C.c_void_p(base_frames), C.c_void_p(mine.data_ptr()), mine.numel() * 4))
        ctx.sync()

    def timed(fn, sync_all):
        best = []
        for _ in range(args.reps):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(ext):
                e0.record()
            fn()
            sync_all()
            with torch.cuda.stream(ext):
                e1.record()
            ctx.sync()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best.append(float(t.item()))
        return float(np.median(best))

    fused(); sc.finish()
    baseline()
    a = torch.empty(n * rows * w, dtype=torch.float32, device=dev)
    b = torch.empty_like(a)
    nl.binding.check(lib.nl_memcpy_d2d(ctx.handle, C.c_void_p(a.data_ptr()), C.c_void_p(sc.job.frames_dev[0]), a.numel() * 4))
    nl.binding.check(lib.nl_memcpy_d2d(ctx.handle, C.c_void_p(b.data_ptr()), C.c_void_p(base_frames), b.numel() * 4))
    ctx.sync()
    same = bool(torch.equal(a.view(torch.int32), b.view(torch.int32)))
    ok = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    del a, b
    # the stripes stack identically (sigma clip), clip counters summed over the ranks
    r1 = sc.job.run(nl.ST_SIGMA)
    r2 = base_job.run(nl.ST_SIGMA)
    stack_same = bool(np.array_equal(r1[0].view(np.uint32), r2[0].view(np.uint32)) and r1[1:] == r2[1:])
    ms_fused = timed(fused, sc.finish)
    ms_base = timed(baseline, lambda: None)
    if rank == 0:
        px = float(n) * w * h
        print(json.dumps({"check": "frame-sharded resample -> row-sharded stack", "n_gpus": world, "frames": n, "width": w, "height": h,
                          "bit_identical_job_buffers": bool(ok.item()), "stack_identical": stack_same,
                          "fused_scatter_ms": ms_fused, "project_then_nccl_alltoall_ms": ms_base,
                          "fused_gpx_per_s": px / ms_fused / 1e6, "baseline_gpx_per_s": px / ms_base / 1e6}), flush=True)
    base_job.close()
    sc.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
