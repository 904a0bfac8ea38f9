"""Wall-clock timing of the stack modes at BASELINE configs[2] size (64 frames of 6000x4000), with and without the deferral of late passes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nightlight_b200 as nl
ctx = nl.Context(0)
n, px = 64, 6000 * 4000
dev = torch.device("cuda", 0)
out = torch.empty(px, dtype=torch.float32, device=dev)
with nl.StackJob(ctx, n, px) as job:
    job.synth_fill()
    ctx.sync()
    for mode, name in ((nl.ST_SIGMA, "sigma"), (nl.ST_WINSOR_SIGMA, "winsor"), (nl.ST_LINEAR_FIT, "linfit (StAuto for 64 frames)"), (nl.ST_MEDIAN, "median"), (nl.ST_MEAN, "mean")):
        for sched in ("0", None):
            ctx.set_tuning("defer_passes", sched or "")
            ts = []
            for rep in range(3):
                ctx.sync(); t0 = time.perf_counter()
                job.run_dev(mode, out.data_ptr())
                ctx.sync(); ts.append((time.perf_counter() - t0) * 1e3)
            gb = 4.0 * (n + 1) * px / 1e9
            print("C3 64 x 6000x4000 %-32s defer %-4s ms %s  -> %.0f GB/s" % (name, sched, ["%.2f" % t for t in ts], gb / (min(ts) * 1e-3)), flush=True)
