"""Extended fuzz run against the oracle (GPU box; not part of pytest: minutes, thousands of cases).

    python tools/fuzz/fuzz_stack.py
"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import nightlight_b200 as nl
import test_gpu_stack as T
ctx = nl.Context(0)
bad = 0
for seed in range(300, 380):
    try:
        T.test_random_shapes_and_modes_fuzz(ctx, seed)
    except AssertionError as e:
        bad += 1; print("FAIL seed", seed, str(e)[:300])
# heavy-tailed, bigger, default schedules and a few odd ones
from oracle import oracle as O
for sched in (None, "1", "2,4,6", "5"):
    ctx.set_tuning("defer_passes", sched or "")
    rng = np.random.default_rng(7)
    for n, p in ((256, 32 * 97 + 3), (200, 4096), (33, 5000)):
        fr = (rng.standard_t(1.5, size=(n, p)) * 10 + 100).astype(np.float32)
        fr[rng.random((n, p)) < 0.05] = np.nan
        for mode, w in (("sigma", False), ("sigma", True), ("winsor", False), ("winsor", True), ("linfit", False)):
            try:
                T.check_against_oracle(ctx, fr, mode, w, 1.5, 2.0)
            except AssertionError as e:
                bad += 1; print("FAIL", sched, n, p, mode, w, str(e)[:200])
print("done, failures:", bad)
