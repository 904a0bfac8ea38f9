#!/usr/bin/env python
"""bench.py -- the stacking hot path on BASELINE.json's configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c2|c2w|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Default (`--config c2`, the configuration BASELINE.json's metric is quoted on): a "step" is one sigma-clip stacking
pass (sigma 2.75/2.75, unweighted) over 256 synthetic 4096x4096 fp32 frames.  `value` is measured with the frames
already resident in HBM (CUDA events on the library's stream); `e2e` is the same pass through ONE C-ABI call with HOST
buffers -- nl_stack_apply (1 GPU) / nl_stack_apply_multi (N GPUs: one process, N devices inside the call): all host
frame pointers in, one host image out, uploads from pinned memory and the download inside the timed region.

Multi-GPU: one process per GPU; the metric's FIXED image is cut into row stripes (rank g owns rows g*H/G ..
(g+1)*H/G of all frames): strong scaling.  No data-path collective; the reassembly of the stacked image
is fused into the stack kernel's epilogue (peer stores over NVLink) or, with --gather nccl, one all-gather.
`--weak` keeps a full 4096-row stripe per GPU (a 4096 x 4096*N image) instead.

Other configurations (BASELINE.json configs[1..4]), same JSON contract:
  c2w  256 x 4096^2, winsorized sigma clip + inverse-noise weights (noise estimated on the device)
  c4   1024 x 8192^2, linear-fit stacking, 1024-row stripes per GPU (8 GPUs = the whole image)
  c5   4096 x 4096^2, batches sized by OpStackBatches.partition from the device memory (capped at 256 frames per batch),
       sigma goal-seek on the first batch (count-only trial stacks), stack of stacks per stripe, one reassembly
  c3   star detection + resample + stack over 64 x 6000x4000 resident frames (one GPU; under torchrun: frame-sharded
       detection and resample, fused scatter into row-stripe stack jobs, stack per stripe, all-gather)

`--impl reference` times the CPU restatement of the reference (oracle/, all host threads, the reference's own 8 MiB
work packages) on a bounded sample of the same workload; the Go reference itself cannot be built here (no Go
toolchain, un-vendored modules).  `--traffic` measures the DRAM bytes of the stack kernels with ncu
(one step) and stores them under profiles/, keyed by a hash of the kernel sources; normal runs report that number
as roofline.traffic only while the hash still matches.
PyTorch is used for torch.distributed, CUDA events and the NCCL collectives only.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "Mpx/s"
SEED = 12345

WORKLOADS = {
    # the configuration the metric is quoted on
    "c2": dict(n_frames=256, width=4096, height=4096, mode="sigma", weighting="none", sig=(2.75, 2.75), cpu_rows=256,
               metric="Mpixels/s stacked (input samples N*P/t; sigma-clip, 256x4096^2 fp32)",
               name="sigma-clip stack 2.75/2.75, 256 x 4096x4096 fp32"),
    "c2w": dict(n_frames=256, width=4096, height=4096, mode="winsor", weighting="inverse_noise", sig=(2.75, 2.75), cpu_rows=128,
                metric="Mpixels/s stacked (input samples N*P/t; winsorized sigma-clip + inverse-noise weights, 256x4096^2 fp32)",
                name="winsorized sigma-clip 2.75/2.75 + noise-weighted mean, 256 x 4096x4096 fp32"),
    "c4": dict(n_frames=1024, width=8192, height=8192, mode="linfit", weighting="none", sig=(2.75, 2.75), cpu_rows=8,
               max_rows_per_gpu=1024,
               metric="Mpixels/s stacked (input samples N*P/t; linear-fit clip, 1024x8192^2 fp32, 1024-row stripes per GPU)",
               name="linear-fit stack 2.75/2.75, 1024 x 8192x8192 fp32, row-striped"),
    "c5": dict(n_frames=4096, width=4096, height=4096, mode="sigma", weighting="none", sig=None, cpu_rows=16,
               metric="Mpixels/s stacked (input samples N*P/t; batched sigma-clip with sigma goal-seek, 4096x4096^2 fp32)",
               name="batched stack of stacks, sigma goal-seek on the first batch, 4096 x 4096x4096 fp32"),
}
MODE_ID = {"median": 0, "mean": 1, "sigma": 2, "winsor": 3, "mad": 4, "linfit": 5}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def source_hash():
    """identifies the kernel sources a DRAM-traffic measurement belongs to (there is no .git on the GPU box)"""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "nightlight_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(d, name), "rb") as f:
                h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def stored_traffic(config, rows, n_gpus):
    """DRAM bytes per step of the stack kernels as measured by `bench.py --traffic` for THESE sources, else None"""
    # traffic_<config>.json: the whole image on one GPU; traffic_<config>_r<rows>.json: a row stripe (one rank of N)
    for name in ("traffic_%s.json" % config, "traffic_%s_r%d.json" % (config, rows)):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)
            if t.get("source_hash") == source_hash() and t.get("rows") == rows:
                return t.get("dram_bytes_per_step"), t
        except Exception:
            pass
    return None, None


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
# CPU side: the oracle (restatement of the reference's Go code).  Used as the checker of sampled rows and as the
# CPU baseline on a bounded sample -- never as the thing measured in the B200 arm.
# ------------------------------------------------------------------------------------------------
def host_synth_frames(n_frames, p0, length, frame_ids=None, threads=None):
    """[n, length] float32 of the synthetic workload from the oracle's generator, all host threads"""
    from oracle import oracle as O
    lib = O.lib()
    fp = C.POINTER(C.c_float)
    ids = list(range(n_frames)) if frame_ids is None else list(frame_ids)
    frames = np.empty((len(ids), length), dtype=np.float32)
    cores = threads or os.cpu_count() or 1
    per = (len(ids) + cores - 1) // cores

    def gen(i0, i1):
        for i in range(i0, i1):
            lib.nlo_synth_frame(frames[i].ctypes.data_as(fp), p0, length, ids[i], SEED)

    th = [threading.Thread(target=gen, args=(i0, min(len(ids), i0 + per))) for i0 in range(0, len(ids), per)]
    [t.start() for t in th]
    [t.join() for t in th]
    return frames


def cpu_stack(frames, mode, weights, sig, threads):
    from oracle import oracle as O
    return O.stack(frames, mode, sig[0], sig[1], weights=weights, threads=threads)


def standin_weights(n):
    """SURVEY.md 8d stand-in for inverse-noise weights: w[k] = 1/(1+4*((k%7)/6)) in float32"""
    k = np.arange(n) % 7
    return (np.float32(1) / (np.float32(1) + np.float32(4) * (k.astype(np.float32) / np.float32(6)))).astype(np.float32)


def cpu_sample(cfg, rows, weights, sig, reps_budget_s, max_reps, threads=None):
    """times the oracle on the first `rows` rows of the workload -> dict(mpx, cores, sample, sec, times, res, clip)"""
    cores = threads or os.cpu_count() or 1
    n, width = cfg["n_frames"], cfg["width"]
    pixels = width * rows
    frames = host_synth_frames(n, 0, pixels)
    times, res = [], None
    t_all = time.perf_counter()
    while len(times) < max_reps and (not times or time.perf_counter() - t_all < reps_budget_s):
        t0 = time.perf_counter()
        res = cpu_stack(frames, cfg["mode"], weights, sig, cores)
        times.append(time.perf_counter() - t0)
    sec = float(np.median(times))
    sample = "%d frames x %dx%d rows (1/%d of the image), median of %d passes, %d threads" % (
        n, width, rows, max(1, cfg["height"] // rows), len(times), cores)
    return dict(mpx=n * pixels / sec / 1e6, cores=cores, sample=sample, sec=sec, times=times, res=res[0], clip=list(res[1:]))


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O  # noqa: F401  (build the oracle before timing)
    rows = args.cpu_rows or cfg["cpu_rows"]
    sig = cfg["sig"] or (2.75, 2.75)
    weights = standin_weights(cfg["n_frames"]) if cfg["weighting"] != "none" else None
    s = cpu_sample(cfg, rows, weights, sig, 1e9, args.warmup + args.steps)
    timed = s["times"][args.warmup:] or s["times"]
    sec = float(np.mean(timed))
    value = cfg["n_frames"] * cfg["width"] * rows / sec / 1e6
    note = "each step stacks a %d-row stripe = 1/%d of the image (a rate; the B200 arm stacks the whole image per step)" % (
        rows, max(1, cfg["height"] // rows))
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(timed),
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s (each step: a %d-row stripe)" % (cfg["name"], rows), "config": args.config,
                   "n_frames": cfg["n_frames"], "width": cfg["width"], "height": cfg["height"], "sample_rows": rows,
                   "reference_sample": "1/%d rows" % max(1, cfg["height"] // rows), "note": note,
                   "weights": "stand-in w[k]=1/(1+4*((k%7)/6))" if weights is not None else "none"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": s["cores"], "kind": "port",
                         "sample": "%d frames x %dx%d rows per step, %d threads, C restatement of the Go reference "
                                   "(Go toolchain absent)" % (cfg["n_frames"], cfg["width"], rows, s["cores"])},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
class Env:
    """rank / device / NCCL plumbing shared by the configurations"""

    def __init__(self, args):
        import torch
        import nightlight_b200 as nl
        self.torch, self.nl, self.args = torch, nl, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch --gpus %d under torch.distributed.run (one process per GPU)" % args.gpus)
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        self.ctx = nl.Context(self.local_rank)
        for kv in [x for x in (args.tune or "").split(",") if x]:      # nl_ctx_set_tuning knobs (A/B measurements)
            k, v = kv.split("=", 1)
            self.ctx.set_tuning(k, v.replace(":", ","))
        self.ext = torch.cuda.ExternalStream(self.ctx.stream, device=torch.device("cuda", self.local_rank))
        self.lib = nl.load_library()

    def barrier(self):
        self.ctx.sync()
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def allmax(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum_int(self, xs):
        if self.world == 1:
            return [int(x) for x in xs]
        t = self.torch.tensor(list(xs), dtype=self.torch.int64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [int(x) for x in t.tolist()]

    def allmin_flag(self, ok):
        if self.world == 1:
            return bool(ok)
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return int(t.item()) == 1

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)

    def record(self, ev):
        with self.torch.cuda.stream(self.ext):
            ev.record()

    def close(self):
        self.ctx.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def stripe_of(cfg, env, args):
    """(row0, rows, image_rows): this rank's rows of the image the run stacks"""
    from nightlight_b200.stripes import stripe_rows
    h = cfg["height"]
    cap = cfg.get("max_rows_per_gpu")
    if args.rows:                                       # explicit rows per GPU (measurements on parts of the image)
        return env.rank * args.rows, args.rows, args.rows * env.world
    if args.weak:
        rows = cap or h
        return env.rank * rows, rows, rows * env.world
    if cap and h // env.world > cap:                    # the image does not fit this few GPUs: a part of it, stated in the line
        return env.rank * cap, cap, cap * env.world
    row0, rows = stripe_rows(h, env.world, env.rank)
    return row0, rows, h


def frame_weights(cfg, env, job, whole_frames):
    """inverse-noise weights (getWeights, stack.go:247-259) from stats.EstimateNoise of every frame, on the device.
    One GPU holding whole frames: straight from the resident job.  Row stripes: every rank generates the whole frames
    k = rank (mod world) into a scratch image and estimates their noise; one all-gather makes the weights identical."""
    nl, ctx, torch = env.nl, env.ctx, env.torch
    n, width, height = cfg["n_frames"], cfg["width"], cfg["height"]
    t0 = time.perf_counter()
    launches0 = ctx.launch_count
    if whole_frames:
        noise = job.frame_noise(width)
    else:
        noise_t = torch.zeros(n, dtype=torch.float32, device="cuda")
        img = ctx.dev_alloc(4 * width * height)
        one = np.zeros(1, np.float32)
        for k in range(env.rank, n, env.world):
            ctx.synth_fill(img, 0, width * height, k, SEED)
            nl.binding.check(env.lib.nl_estimate_noise_dev(ctx.handle, C.c_void_p(img), 1, width * height, width, height,
                                                           one.ctypes.data_as(C.POINTER(C.c_float))))
            noise_t[k] = float(one[0])
        ctx.dev_free(img)
        if env.world > 1:
            env.dist.all_reduce(noise_t)                # every frame was filled in by exactly one rank
        noise = noise_t.cpu().numpy()
    frames = [nl.ops.Image(data=np.zeros(1, np.float32), noise=float(x), id=i) for i, x in enumerate(noise)]
    w = nl.get_weights(frames, nl.W_INVERSE_NOISE)
    return w, (time.perf_counter() - t0) * 1e3, ctx.launch_count - launches0


def verify_rows(cfg, env, row0, n_rows, dev_result, weights, sig, threads):
    """the first n_rows rows of this rank's stripe against the CPU restatement of the reference, bit for bit"""
    width = cfg["width"]
    px = n_rows * width
    got = np.empty(px, np.float32)
    env.ctx.d2h(got, dev_result)
    frames = host_synth_frames(cfg["n_frames"], row0 * width, px, threads=threads)
    want = cpu_stack(frames, cfg["mode"], weights, sig, threads)[0]
    gn, wn = np.isnan(got), np.isnan(want)
    same = bool(np.array_equal(gn, wn) and np.array_equal(got.view(np.uint32)[~gn], want.view(np.uint32)[~wn]))
    return same


def run_stack_config(args, cfg):
    """c2 / c2w / c4: one resident stack job per rank, one stacking mode"""
    env = Env(args)
    nl, ctx, torch, dist, lib = env.nl, env.ctx, env.torch, env.dist, env.lib
    world, rank = env.world, env.rank
    n, width = cfg["n_frames"], cfg["width"]
    row0, rows, image_rows = stripe_of(cfg, env, args)
    pixels = width * rows
    mode = MODE_ID[cfg["mode"]]
    sig = cfg["sig"]
    strong = not args.weak and not args.rows and image_rows == cfg["height"]

    job = nl.StackJob(ctx, n, pixels)
    job.synth_fill(p0=row0 * width)
    ctx.sync()
    weights, weights_ms, weights_launches = None, None, 0
    if cfg["weighting"] == "inverse_noise":
        whole = world == 1 and rows == cfg["height"]
        if whole or strong:
            weights, weights_ms, weights_launches = frame_weights(cfg, env, job, whole)
            weights_desc = "1/(1+4*(noise-min)/(max-min)) from stats.EstimateNoise of every frame on the device (%.1f ms, outside the step, as in the reference where Stats are computed at load time)" % weights_ms
        else:
            weights = standin_weights(n)
            weights_desc = "stand-in w[k]=1/(1+4*((k%7)/6)) (part of the image only: no whole frames to estimate noise from)"
    else:
        weights_desc = "none"

    out = torch.empty(pixels, dtype=torch.float32, device="cuda")
    # Multi-GPU reassembly of the stacked image: fused into the stack kernel's epilogue (every result is also stored
    # into the peer-mapped gathered image of every other rank over NVLink), with one tiny all-reduce as the
    # "everybody's stores have landed" signal.  --gather nccl keeps the plain all_gather_into_tensor of the stripes;
    # the fused path is verified against it once before timing.
    equal_stripes = image_rows == rows * world
    gathered = torch.empty(pixels * world, dtype=torch.float32, device="cuda") if world > 1 and equal_stripes else None
    peer = None
    flag = torch.zeros(1, dtype=torch.int32, device="cuda") if world > 1 else None
    if world > 1 and args.gather == "peer" and equal_stripes:
        from nightlight_b200.stripes import PeerGather
        try:
            peer = PeerGather(ctx, pixels)
        except Exception as e:                       # no peer access on this box: say so and fall back
            if rank == 0:
                print("peer mapping unavailable (%s), using the NCCL all-gather" % e, file=sys.stderr)
            peer = None
        if not env.allmin_flag(peer is not None) and peer is not None:
            peer.close()
            peer = None

    def gather_nccl():
        with torch.cuda.stream(env.ext):
            if equal_stripes:
                dist.all_gather_into_tensor(gathered, out)
            else:
                from nightlight_b200.stripes import allgather_image
                allgather_image(out, width, image_rows)

    def step():
        if peer is not None:
            job.run_dev_bcast(mode, peer.local_out, peer.peer_outs, weights, sig[0], sig[1], 0.0)
            with torch.cuda.stream(env.ext):
                dist.all_reduce(flag)
            return
        job.run_dev(mode, out.data_ptr(), weights, sig[0], sig[1], 0.0)
        if world > 1:
            gather_nccl()

    for _ in range(args.warmup):
        step()
    env.barrier()
    gather_desc = "none (1 GPU)"
    if world > 1:
        gather_desc = "NCCL all_gather_into_tensor of the stripes"
    if peer is not None:
        # verify the fused reassembly against NCCL once
        job.run_dev(mode, out.data_ptr(), weights, sig[0], sig[1], 0.0)
        gather_nccl()
        env.barrier()
        same = bool(np.array_equal(peer.to_host().view(np.uint32), gathered.cpu().numpy().view(np.uint32)))
        if not env.allmin_flag(same):
            raise SystemExit("fused peer-store reassembly differs from the NCCL all-gather")
        gather_desc = ("fused: stack kernel epilogue stores every stripe into all peers' images (CUDA IPC over NVLink), "
                       "verified against NCCL all_gather")
        env.barrier()

    # ---- timed region: K steps, device-resident frames (inputs >> 126 MB L2: no flush needed)
    sampler = ClockSampler(env.local_rank)
    sampler.start()
    launches0 = ctx.launch_count
    ev = [env.event() for _ in range(2 * args.steps + 2)]
    env.record(ev[0])
    for i in range(args.steps):
        env.record(ev[2 + 2 * i])
        if peer is not None:
            job.run_dev_bcast(mode, peer.local_out, peer.peer_outs, weights, sig[0], sig[1], 0.0)
        else:
            job.run_dev(mode, out.data_ptr(), weights, sig[0], sig[1], 0.0)
        env.record(ev[3 + 2 * i])
        if peer is not None:
            with torch.cuda.stream(env.ext):
                dist.all_reduce(flag)
        elif world > 1:
            gather_nccl()
    env.record(ev[1])
    env.barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    total_ms = env.allmax(ev[0].elapsed_time(ev[1]))
    kernel_ms = [ev[2 + 2 * i].elapsed_time(ev[3 + 2 * i]) for i in range(args.steps)]
    ms_per_step = total_ms / args.steps
    total_pixels = width * image_rows
    value = n * total_pixels / (ms_per_step * 1e-3) / 1e6
    clip_low, clip_high = env.allsum_int(job.clip_counts())

    # ---- parity of sampled rows on every rank (the oracle as the checker), and the CPU baseline beside the number
    threads = max(1, (os.cpu_count() or 1) // world)
    cpu, parity = None, {}
    result_dev = peer.local_out if peer is not None else out.data_ptr()
    if rank == 0 and world == 1 and not args.no_cpu:
        # the bounded CPU sample is the first rows of the image: a second, small job stacks exactly that stripe on the
        # GPU, and result and clip totals must agree with the CPU run
        crow = min(args.cpu_rows or cfg["cpu_rows"], rows)
        s = cpu_sample(cfg, crow, weights, sig, 12.0, 6)
        with nl.StackJob(ctx, n, width * crow) as sjob:
            sjob.synth_fill(p0=row0 * width)
            got, gl, gh = sjob.run(mode, weights, sig[0], sig[1], 0.0)
        gn, wn = np.isnan(got), np.isnan(s["res"])
        same = bool(np.array_equal(gn, wn) and np.array_equal(got.view(np.uint32)[~gn], s["res"].view(np.uint32)[~wn]))
        parity = {"rows_checked": crow, "bit_exact": same, "clip_totals_gpu": [gl, gh], "clip_totals_cpu": s["clip"],
                  "clip_totals_equal": [gl, gh] == s["clip"]}
        if not same or [gl, gh] != s["clip"]:
            raise SystemExit("parity failure against the CPU restatement on the sampled stripe: %s" % json.dumps(parity))
        cpu = {"value": s["mpx"], "unit": UNIT, "cores": s["cores"], "kind": "port", "sample": s["sample"],
               "reference_sample": "1/%d rows" % max(1, cfg["height"] // crow)}
    elif args.verify_rows > 0:
        vrows = min(args.verify_rows, rows)
        same = verify_rows(cfg, env, row0, vrows, result_dev, weights, sig, threads)
        ok = env.allmin_flag(same)
        parity = {"rows_checked_per_rank": vrows, "bit_exact_on_every_rank": ok}
        if not ok:
            raise SystemExit("parity failure against the CPU restatement on rank %d's sampled rows" % rank)

    if args.dump_rows > 0:
        # more rows than the run itself can afford to check on the host while it holds GPUs: saved for
        # tools/verify_rows.py, which replays them through the CPU restatement anywhere (no GPU needed)
        drows = min(args.dump_rows, rows)
        got = np.empty(drows * width, np.float32)
        ctx.d2h(got, result_dev)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.save(os.path.join(ROOT, "gpurun_out", "rows_%s_rank%d.npy" % (args.config, rank)), got)
        with open(os.path.join(ROOT, "gpurun_out", "rows_%s_rank%d.json" % (args.config, rank)), "w") as f:
            json.dump({"config": args.config, "rank": rank, "world": world, "row0": row0, "rows": drows, "width": width,
                       "n_frames": n, "mode": cfg["mode"], "sigma": list(sig), "seed": SEED,
                       "weights": [float(x) for x in weights] if weights is not None else None}, f)
        parity["rows_dumped_per_rank"] = drows

    # ---- roofline of the dominant kernel (this rank): algorithmic bytes = 4*(N+1) per output pixel
    peak, peak_src = peaks()
    k_ms = float(np.mean(kernel_ms))
    algo_bytes = 4.0 * (n + 1) * pixels
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9
    traffic, tinfo = stored_traffic(args.config, rows, world)
    kernel_name = {"sigma": "stack_column_kernel<sigma> (per step: one launch over the frame stack + one over the pool of columns whose late clipping passes were deferred)",
                   "winsor": "stack_column_kernel<winsor, weighted> (per step: one launch over the frame stack + one over the pool of deferred columns)",
                   "linfit": "stack_column_kernel<linfit> (per step: sort + first rejection rounds over the frame stack, then regrouping launches over the pools of unfinished columns)"}[cfg["mode"]]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_over_algorithmic": (traffic / algo_bytes) if traffic else None,
                "traffic_source": ("ncu dram__bytes_read.sum + dram__bytes_write.sum over one step, %s" % tinfo.get("when", "")) if tinfo
                else "not measured for these kernel sources (python bench.py --traffic)",
                "kernel": kernel_name, "kernel_ms": k_ms, "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src}

    # ---- end to end through the C ABI with host buffers: ONE call, all devices inside it
    clip_resident = [clip_low, clip_high]
    if peer is not None:
        peer.close()
        peer = None
    job.close()
    del out, gathered
    torch.cuda.empty_cache()
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, cfg, env, image_rows, mode, weights, sig, clip_resident if strong or world == 1 else None)

    if rank == 0:
        part = "" if image_rows == cfg["height"] else " -- a %d-row part of the %d-row image (%d GPUs hold no more)" % (image_rows, cfg["height"], world)
        line = {
            "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: %d x %dx%d fp32 frames, row stripes of %d rows per GPU%s, frames resident in HBM" % (
                           cfg["name"], n, width, image_rows, rows, part),
                       "config": args.config, "n_frames": n, "width": width, "height": image_rows, "rows_per_gpu": rows,
                       "mode": cfg["mode"], "sigma": list(sig), "weights": weights_desc,
                       "parallelism": "row stripes x%d" % world, "gather": gather_desc,
                       "l2": "inputs (%.1f GiB per GPU) larger than L2, no flush" % (4.0 * n * pixels / 2**30),
                       "mpx_out_per_s": total_pixels / (ms_per_step * 1e-3) / 1e6,
                       "clipped": clip_resident, "parity": parity, "source_hash": source_hash()},
            "clocks": clocks, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu,
            "gpu_launches": launches + weights_launches,
        }
        print(json.dumps(line))
    env.close()
    return 0


def run_e2e(args, cfg, env, image_rows, mode, weights, sig, clip_expected):
    """One stacking pass through the public C ABI with HOST buffers, ONE call: nl_stack_apply_multi takes all host
    frame pointers and the contexts of all N devices (nl_stack_apply on one GPU), deals the image's rows to the
    devices, pipelines row stripes on two streams per device, and fills one host image.  Runs in rank 0's process --
    this is what the Go OpStack.Apply (stack.go:115) would bind; the other ranks wait.  All copies are inside the
    timed region."""
    nl, lib, torch = env.nl, env.lib, env.torch
    world, rank = env.world, env.rank
    n, width = cfg["n_frames"], cfg["width"]
    pixels = width * image_rows
    result = None
    if rank == 0:
        frame_bytes = 4 * pixels
        distinct = max(1, min(n, int(args.e2e_host_gib * 2**30) // frame_bytes))
        host = C.c_void_p()
        pinned = lib.nl_host_alloc_pinned(distinct * frame_bytes, C.byref(host)) == 0
        if not pinned:
            arr = np.empty(distinct * pixels, dtype=np.float32)
            host = C.c_void_p(arr.ctypes.data)
        host_out = C.c_void_p()
        out_pinned = lib.nl_host_alloc_pinned(frame_bytes, C.byref(host_out)) == 0
        if not out_pinned:
            oarr = np.empty(pixels, dtype=np.float32)
            host_out = C.c_void_p(oarr.ctypes.data)
        # the host frames: generated on the device once, copied down (outside the measurement)
        scratch = env.ctx.dev_alloc(frame_bytes)
        for k in range(distinct):
            env.ctx.synth_fill(scratch, 0, pixels, k, SEED)
            nl.binding.check(lib.nl_memcpy_d2h(env.ctx.handle, C.c_void_p(host.value + k * frame_bytes), C.c_void_p(scratch), frame_bytes))
        env.ctx.sync()
        env.ctx.dev_free(scratch)
        ptrs = (C.c_void_p * n)(*[host.value + (k % distinct) * frame_bytes for k in range(n)])
        ctxs = [env.ctx] + [nl.Context(g) for g in range(1, world)]
        arr_ctx = (C.c_void_p * world)(*[c.handle for c in ctxs])
        cl, ch = C.c_int64(), C.c_int64()
        wp = weights.ctypes.data_as(C.POINTER(C.c_float)) if weights is not None else None

        def e2e_step():
            if world == 1:
                nl.binding.check(lib.nl_stack_apply(env.ctx.handle, ptrs, n, pixels, width, args.e2e_stripes, mode, wp,
                                                    sig[0], sig[1], 0.0, host_out, C.byref(cl), C.byref(ch)))
            else:
                nl.binding.check(lib.nl_stack_apply_multi(arr_ctx, world, ptrs, n, pixels, width, args.e2e_stripes, mode, wp,
                                                          sig[0], sig[1], 0.0, host_out, C.byref(cl), C.byref(ch)))

        e2e_step()                                   # warm-up (allocates the stripe lanes on every device)
        steps = max(1, min(args.steps, args.e2e_steps))
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        sec = (time.perf_counter() - t0) / steps
        for c in ctxs[1:]:
            c.close()
        lib.nl_stack_apply_release(env.ctx.handle)
        if pinned:
            lib.nl_host_free_pinned(host)
        if out_pinned:
            lib.nl_host_free_pinned(host_out)
        h2d = 4 * n * pixels
        result = {"value": n * pixels / sec / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * pixels + 16,
                  "ms_per_step": sec * 1e3, "steps": steps, "host_memory": "pinned" if pinned else "pageable",
                  "h2d_gb_per_s_aggregate": h2d / sec / 1e9, "h2d_gb_per_s_per_gpu": h2d / sec / 1e9 / world,
                  "e2e_distinct_host_frames": distinct,
                  "api": ("nl_stack_apply" if world == 1 else "nl_stack_apply_multi (one process, %d devices inside the call)" % world) +
                         ": %d host frame pointers in, one host image out; %d row stripes per device alternating on two streams "
                         "inside the library" % (n, args.e2e_stripes),
                  "clipped": [cl.value, ch.value]}
        if distinct == n and clip_expected is not None:
            result["matches_resident_run"] = [cl.value, ch.value] == list(clip_expected)
        else:
            result["matches_resident_run"] = None
            if distinct < n:
                result["note"] = "host memory bounded to %.0f GiB: frame k is read from host buffer k %% %d (same bytes over PCIe)" % (args.e2e_host_gib, distinct)
    if world > 1:
        env.dist.barrier()
    return result


# ------------------------------------------------------------------------------------------------
# c5: 4096 frames, batches sized from the device memory, sigma goal-seek, stack of stacks per stripe
# ------------------------------------------------------------------------------------------------
def run_c5(args, cfg):
    env = Env(args)
    nl, ctx, torch, dist, lib = env.nl, env.ctx, env.torch, env.dist, env.lib
    world, rank = env.world, env.rank
    n, width = cfg["n_frames"], cfg["width"]
    row0, rows, image_rows = stripe_of(cfg, env, args)
    pixels = width * rows
    mode = MODE_ID[cfg["mode"]]
    free_b, total_b = ctx.mem_info()
    # OpStackBatches.partition (stackbatches.go:121-210) with the device's free memory in the place of StackMemoryMB:
    # 60 % of it for the frames of a batch (the rest: the pool of deferred columns, accumulator, result)
    # ... capped so that a batch holds at most 256 frames per column: beyond that the exact emulation of the reference's
    # quick-select runs on narrower tiles and the stack slows down by an order of magnitude;
    # --stack-memory-mb sets the budget explicitly (e.g. the whole free memory: one batch of 4096 frames on 8 GPUs)
    frame_mb = 4.0 * pixels / 2**20
    mem_mb = args.stack_memory_mb or min(int(free_b * 0.6) >> 20, int((256 + 3) * frame_mb) + 1)
    if world > 1:                                       # every rank must cut the same batches
        t = torch.tensor([mem_mb], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        mem_mb = int(t.item())
    perm = [int(x) for x in np.random.default_rng(SEED).permutation(n)]      # math/rand.Perm stand-in: an input
    order, nb, bs, _ = nl.partition(n, width, rows, mem_mb, 1, perm=perm)
    batches = [order[b * bs:(b + 1) * bs] for b in range(nb)]
    batches = [b for b in batches if b]
    acc = ctx.dev_alloc(4 * pixels)
    tmp = ctx.dev_alloc(4 * pixels)
    jobs = {}

    def job_for(size):
        if size not in jobs:
            for j in jobs.values():
                j.close()
            jobs.clear()
            jobs[size] = nl.StackJob(ctx, size, pixels)
        return jobs[size]

    def fill(batch):
        job = job_for(len(batch))
        base, stride = job.frames_dev
        for i, k in enumerate(batch):
            ctx.synth_fill(base + 4 * i * stride, row0 * width, pixels, k, SEED)
        return job

    # ---- pass 1: goal-seek of the sigmas on the first batch (count-only trial stacks; totals summed over the stripes)
    sampler = ClockSampler(env.local_rank)
    sampler.start()
    launches0 = ctx.launch_count
    job = fill(batches[0])
    for _ in range(max(3, args.warmup)):
        job.run_dev(mode, 0, None, 2.75, 2.75, 0.0)
    env.barrier()
    seek = nl.binding.SigmaSeek()
    nl.binding.check(lib.nl_sigma_seek_begin(C.byref(seek), mode, len(batches[0]), width * image_rows, args.clip_perc_low, args.clip_perc_high))
    ev0, ev1 = env.event(), env.event()
    env.record(ev0)
    trial_log = []
    while not seek.done:
        job.run_dev(mode, 0, None, seek.trial_low, seek.trial_high, 0.0)
        ctx.sync()
        cl, ch = env.allsum_int(job.clip_counts())
        trial_log.append([round(float(seek.trial_low), 4), round(float(seek.trial_high), 4), cl, ch])
        assert lib.nl_sigma_seek_step(C.byref(seek), cl, ch) >= 0
    env.record(ev1)
    env.barrier()
    seek_ms = env.allmax(ev0.elapsed_time(ev1))
    sig = (float(seek.result_low), float(seek.result_high))

    # ---- pass 2: every batch stacked with the sigmas found, stack of stacks on the device, one reassembly
    gathered = torch.empty(pixels * world, dtype=torch.float32, device="cuda") if world > 1 else None
    acc_t = None
    if world > 1:
        # NCCL builds its all-gather channels on first use (>100 ms): outside the timed segments
        acc_t = torch.zeros(pixels, dtype=torch.float32, device="cuda")
        with torch.cuda.stream(env.ext):
            dist.all_gather_into_tensor(gathered, acc_t)
        env.barrier()
    stack_ms, k_ms_batches = 0.0, []
    clip_tot = [0, 0]
    steps = max(1, args.steps if args.steps_given else 1)
    for s in range(steps):
        stack_ms = 0.0
        clip_tot = [0, 0]
        k_ms_batches = []
        for b, batch in enumerate(batches):
            job = fill(batch)
            ctx.sync()
            e0, e1, e2 = env.event(), env.event(), env.event()
            env.record(e0)
            job.run_dev(mode, tmp, None, sig[0], sig[1], 0.0)
            env.record(e1)
            nl.binding.check(lib.nl_stack_incremental_dev(ctx.handle, C.c_void_p(acc), C.c_void_p(tmp), pixels, float(len(batch)), 1 if b == 0 else 0))
            if b == len(batches) - 1:
                nl.binding.check(lib.nl_stack_incremental_finalize_dev(ctx.handle, C.c_void_p(acc), pixels, float(n)))
                if world > 1:
                    # (torch view of the accumulator for the one all-gather at the end)
                    if acc_t is None:
                        acc_t = torch.empty(pixels, dtype=torch.float32, device="cuda")
                    nl.binding.check(lib.nl_memcpy_d2d(ctx.handle, C.c_void_p(acc_t.data_ptr()), C.c_void_p(acc), 4 * pixels))
                    with torch.cuda.stream(env.ext):
                        dist.all_gather_into_tensor(gathered, acc_t)
            env.record(e2)
            env.barrier()
            stack_ms += env.allmax(e0.elapsed_time(e2))
            k_ms_batches.append(e0.elapsed_time(e1))
            c = job.clip_counts()
            clip_tot[0] += c[0]; clip_tot[1] += c[1]
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    clip_tot = env.allsum_int(clip_tot)
    total_pixels = width * image_rows
    job_ms = seek_ms + stack_ms
    value = n * total_pixels / (job_ms * 1e-3) / 1e6

    # ---- parity: the first rows of every rank's stripe against the oracle driven through the same batches
    threads = max(1, (os.cpu_count() or 1) // world)
    parity = {}
    if args.verify_rows > 0:
        from oracle import oracle as O
        vrows = min(args.verify_rows, rows)
        px = vrows * width
        got = np.empty(px, np.float32)
        ctx.d2h(got, acc)
        want = np.empty(px, np.float32)
        fp = C.POINTER(C.c_float)
        for b, batch in enumerate(batches):
            frames = host_synth_frames(n, row0 * width, px, frame_ids=batch, threads=threads)
            res = O.stack(frames, cfg["mode"], sig[0], sig[1], threads=threads)[0]
            O.lib().nlo_stack_incremental(want.ctypes.data_as(fp), res.ctypes.data_as(fp), px, float(len(batch)), 1 if b == 0 else 0)
        O.lib().nlo_stack_incremental_finalize(want.ctypes.data_as(fp), px, float(n))
        same = bool(np.array_equal(got.view(np.uint32), want.view(np.uint32)))
        ok = env.allmin_flag(same)
        parity = {"rows_checked_per_rank": vrows, "bit_exact_on_every_rank": ok}
        if not ok:
            raise SystemExit("c5 parity failure against the CPU restatement on rank %d" % rank)

    peak, peak_src = peaks()
    algo_bytes = sum(4.0 * (len(b) + 1) * pixels for b in batches)
    k_ms = float(sum(k_ms_batches))
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": "stack_column_kernel<sigma> over the %d batches of pass 2 (%d frames per column)" % (len(batches), bs),
                "kernel_ms": k_ms, "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src}

    for j in jobs.values():
        j.close()
    jobs.clear()
    e2e = None
    if not args.no_e2e:
        e2e = run_c5_e2e(args, cfg, env, batches, mode, sig, pixels, row0, acc, tmp)
    ctx.dev_free(acc)
    ctx.dev_free(tmp)
    if rank == 0:
        line = {
            "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(3, args.warmup),
            "ms_per_step": job_ms, "higher_is_better": True, "scaling": "strong" if image_rows == cfg["height"] else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: %d x %dx%d fp32 frames in %d batches of <= %d (OpStackBatches.partition with %d MiB of device "
                                   "memory per GPU), row stripes of %d rows per GPU; a step = goal-seek (%d count-only trial stacks of "
                                   "batch 0) + every batch stacked and accumulated + one reassembly; batch frames generated on the "
                                   "device between the timed segments" % (cfg["name"], n, width, image_rows, len(batches), bs, mem_mb, rows, seek.trials),
                       "config": "c5", "n_frames": n, "width": width, "height": image_rows, "rows_per_gpu": rows, "mode": cfg["mode"],
                       "batches": len(batches), "batch_size": bs, "stack_memory_mb": mem_mb,
                       "clip_targets_percent": [args.clip_perc_low, args.clip_perc_high], "sigma_found": list(sig),
                       "goal_seek": {"trials": int(seek.trials), "converged": bool(seek.converged), "ms": seek_ms, "log": trial_log},
                       "pass2_ms": stack_ms, "clipped": clip_tot,
                       "clipped_percent": [100.0 * c / (n * total_pixels) for c in clip_tot],
                       "parallelism": "row stripes x%d" % world, "gather": "one NCCL all_gather_into_tensor at the end" if world > 1 else "none (1 GPU)",
                       "l2": "inputs (%.1f GiB per batch and GPU) larger than L2, no flush" % (4.0 * bs * pixels / 2**30),
                       "parity": parity, "source_hash": source_hash()},
            "clocks": clocks, "roofline": roofline, "e2e": e2e, "cpu_baseline": None, "gpu_launches": launches,
        }
        print(json.dumps(line))
    env.close()
    return 0


def run_c5_e2e(args, cfg, env, batches, mode, sig, pixels, row0, acc, tmp):
    """pass 2 end to end with HOST frames: every batch uploaded from pinned host memory (this rank's row stripe of each
    frame), stacked, accumulated on the device; the stripe of the stack of stacks comes back to the host.  One process
    per GPU (OpStackBatches has no one-call C entry point: its frames are lazy promises in the reference,
    stackbatches.go:84-116).  The host holds a bounded ring of distinct frames."""
    nl, ctx, lib = env.nl, env.ctx, env.lib
    n = cfg["n_frames"]
    frame_bytes = 4 * pixels
    distinct = max(1, min(n, int(args.e2e_host_gib * 2**30 / env.world) // frame_bytes))
    host = C.c_void_p()
    nl.binding.check(lib.nl_host_alloc_pinned(distinct * frame_bytes, C.byref(host)))
    host_out = C.c_void_p()
    nl.binding.check(lib.nl_host_alloc_pinned(frame_bytes, C.byref(host_out)))
    for k in range(distinct):
        ctx.synth_fill(tmp, row0 * cfg["width"], pixels, k, SEED)
        nl.binding.check(lib.nl_memcpy_d2h(ctx.handle, C.c_void_p(host.value + k * frame_bytes), C.c_void_p(tmp), frame_bytes))
    ctx.sync()
    env.barrier()
    t0 = time.perf_counter()
    job = None
    for b, batch in enumerate(batches):
        if job is None or job.n_frames != len(batch):
            if job is not None:
                job.close()
            job = nl.StackJob(ctx, len(batch), pixels)
        for i, k in enumerate(batch):
            job.put_frame(i, (host.value + (k % distinct) * frame_bytes, pixels))
        job.run_dev(mode, tmp, None, sig[0], sig[1], 0.0)
        nl.binding.check(lib.nl_stack_incremental_dev(ctx.handle, C.c_void_p(acc), C.c_void_p(tmp), pixels, float(len(batch)), 1 if b == 0 else 0))
    nl.binding.check(lib.nl_stack_incremental_finalize_dev(ctx.handle, C.c_void_p(acc), pixels, float(n)))
    nl.binding.check(lib.nl_memcpy_d2h(ctx.handle, host_out, C.c_void_p(acc), frame_bytes))
    ctx.sync()
    sec = env.allmax(time.perf_counter() - t0)
    job.close()
    lib.nl_host_free_pinned(host)
    lib.nl_host_free_pinned(host_out)
    h2d = 4 * n * pixels
    return {"value": env.world * n * pixels / sec / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": frame_bytes,
            "ms_per_step": sec * 1e3, "steps": 1, "host_memory": "pinned", "h2d_gb_per_s_per_gpu": h2d / sec / 1e9,
            "e2e_distinct_host_frames": distinct,
            "api": "per GPU: nl_stack_put_frame x batch, nl_stack_run_dev, nl_stack_incremental_dev, ..., finalize, download "
                   "(pass 2 with the sigmas of the goal-seek; the jobs are allocated inside the timed region)",
            "note": "bytes per step are per GPU; frame k is read from host buffer k %% %d" % distinct}


# ------------------------------------------------------------------------------------------------
# --traffic: DRAM bytes of the stack kernels of one step, measured with ncu
# ------------------------------------------------------------------------------------------------
def run_traffic(args, cfg):
    """runs `bench.py --traffic-child` under ncu (device-resident steps only), sums dram bytes over the stack kernels
    of ONE step and writes profiles/traffic_<config>.json keyed by the hash of the kernel sources"""
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = os.path.join(ROOT, "gpurun_out", "traffic_%s%s.csv" % (args.config, "_r%d" % args.rows if args.rows else ""))
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
           "-k", "regex:stack_column_kernel|stack_mean_kernel", "--csv", "--log-file", log,
           sys.executable, os.path.abspath(__file__), "--config", args.config, "--traffic-child", "--no-e2e", "--no-cpu",
           "--verify-rows", "0", "--steps", "1", "--warmup", "3"] + (["--rows", str(args.rows)] if args.rows else [])
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    import csv
    rows_csv = []
    with open(log) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        rows_csv.append(r)
    per_launch = {}
    for r in rows_csv:
        key = int(r["ID"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
        d = per_launch.setdefault(key, {"kernel": r["Kernel Name"][:60]})
        if r["Metric Name"].startswith("dram__bytes"):
            d[r["Metric Name"]] = v * scale
        else:
            d["ns"] = v
    ids = sorted(per_launch)
    # the child ran 3 warm-up steps + 1 timed step, all alike: the last quarter of the launches is one step
    per_step = len(ids) // 4
    last = ids[-per_step:] if per_step else ids
    rd = sum(per_launch[i].get("dram__bytes_read.sum", 0.0) for i in last)
    wr = sum(per_launch[i].get("dram__bytes_write.sum", 0.0) for i in last)
    cfg_rows = args.rows or (cfg.get("max_rows_per_gpu") or cfg["height"])
    out = {"config": args.config, "rows": cfg_rows, "source_hash": source_hash(), "launches_per_step": per_step,
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_step": rd + wr,
           "algorithmic_bytes": 4.0 * (cfg["n_frames"] + 1) * cfg["width"] * cfg_rows,
           "launch_ns_under_ncu": [per_launch[i].get("ns") for i in last], "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
           "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over `bench.py --traffic-child` (1 GPU)"}
    out["traffic_over_algorithmic"] = out["dram_bytes_per_step"] / out["algorithmic_bytes"]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    name = "traffic_%s_r%d.json" % (args.config, args.rows) if args.rows else "traffic_%s.json" % args.config
    with open(os.path.join(ROOT, "profiles", name), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(WORKLOADS) + ["c3"])
    ap.add_argument("--weak", action="store_true", help="weak scaling: a full stripe per GPU instead of the fixed image cut into N stripes")
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (default: the image's rows / N)")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the bounded CPU sample (default per configuration)")
    ap.add_argument("--verify-rows", type=int, default=None, help="rows per rank checked against the CPU restatement")
    ap.add_argument("--dump-rows", type=int, default=0, help="save the first rows of every rank's result under gpurun_out/ for tools/verify_rows.py")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-stripes", type=int, default=8, help="row stripes per device of the pipelined end-to-end pass")
    ap.add_argument("--e2e-host-gib", type=float, default=16.0, help="pinned host memory for the end-to-end frames")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"], help="multi-GPU reassembly of the stacked image")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--stack-memory-mb", type=int, default=0, help="c5: memory budget of a batch per GPU (default: 60 %% of the free device memory, capped at 256-frame batches)")
    ap.add_argument("--clip-perc-low", type=float, default=2.0, help="c5: target percentage of samples clipped on the low side")
    ap.add_argument("--clip-perc-high", type=float, default=2.0)
    ap.add_argument("--tune", default="", help="k=v,k=v: nl_ctx_set_tuning knobs of the library (A/B measurements; lists with ':')")
    ap.add_argument("--traffic", action="store_true", help="measure the DRAM traffic of one step with ncu -> profiles/traffic_<config>.json")
    ap.add_argument("--traffic-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    args.steps_given = args.steps is not None
    if args.steps is None:
        args.steps = 5
    if args.warmup < 3:
        args.warmup = 3
    if args.config == "c3":
        from tools.bench_c3 import run_c3, run_c3_multi
        multi = int(os.environ.get("WORLD_SIZE", "1")) > 1 or os.environ.get("NL_C3_MULTI") == "1"   # (the N-GPU path on one rank: a check)
        return run_c3_multi(args) if multi else run_c3(args)
    cfg = WORKLOADS[args.config]
    if args.verify_rows is None:
        args.verify_rows = {"c4": 8, "c5": 32}.get(args.config, 16)
    if args.impl == "reference":
        return run_reference(args, cfg)
    if args.traffic:
        return run_traffic(args, cfg)
    if args.config == "c5":
        return run_c5(args, cfg)
    return run_stack_config(args, cfg)


if __name__ == "__main__":
    sys.exit(main())
