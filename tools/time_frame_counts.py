"""Sigma and linear-fit stack time for the same sample count at 32 .. 256 frames (occupancy vs frame count)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nightlight_b200 as nl
ctx = nl.Context(0)
dev = torch.device("cuda", 0)
for n in (32, 64, 96, 128, 160, 192, 224, 256):
    px = 4096 * 512 * 256 // n // 32 * 32
    out = torch.empty(px, dtype=torch.float32, device=dev)
    with nl.StackJob(ctx, n, px) as job:
        job.synth_fill(); ctx.sync()
        for mode, name in ((nl.ST_SIGMA, "sigma"), (nl.ST_LINEAR_FIT, "linfit")):
            ts = []
            for rep in range(3):
                ctx.sync(); t0 = time.perf_counter()
                job.run_dev(mode, out.data_ptr())
                ctx.sync(); ts.append((time.perf_counter() - t0) * 1e3)
            gb = 4.0 * (n + 1) * px / 1e9
            print("n=%3d px=%8d %-7s ms %6.2f -> %5.0f GB/s" % (n, px, name, min(ts), gb / (min(ts) * 1e-3)), flush=True)
