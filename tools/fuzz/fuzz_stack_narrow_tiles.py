"""Extended fuzz run against the oracle (GPU box; not part of pytest: minutes, thousands of cases).

    python tools/fuzz/fuzz_stack_narrow_tiles.py
"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import nightlight_b200 as nl
import test_gpu_stack as T
from util import mode_cases
ctx = nl.Context(0)
bad = 0
cases = mode_cases()
for seed in range(40):
    rng = np.random.default_rng(1000 + seed)
    sched = rng.choice(["", "1", "2", "3", "1,2,3", "2,5,9", "0"])
    ctx.set_tuning("defer_passes", str(sched))
    n = int(rng.choice([257, 300, 480, 513, 520, 777, 1024, 1100, 2100]))
    p = int(rng.integers(1, 140))
    scale = float(rng.choice([1e-3, 1.0, 50.0]))
    fr = (rng.standard_t(float(rng.choice([1.5, 3.0, 30.0])), size=(n, p)) * scale + float(rng.choice([0.0, 1000.0]))).astype(np.float32)
    if rng.random() < 0.7: fr[rng.random(fr.shape) < float(rng.choice([0.001, 0.05, 0.4]))] = np.nan
    if rng.random() < 0.3: fr = np.round(fr).astype(np.float32)
    sl, sh = (float(x) for x in rng.choice([0.5, 1.0, 2.0, 2.75, -1.0], 2))
    mode, weighted = cases[int(rng.integers(0, len(cases)))]
    w = (rng.random(n).astype(np.float32) + np.float32(0.05)) if weighted else None
    try:
        T.check_against_oracle(ctx, fr, mode, weighted, sl, sh, ref_loc=3.5, w=w)
    except AssertionError as e:
        bad += 1; print("FAIL", seed, sched, n, p, mode, weighted, str(e)[:200])
print("narrow fuzz done, failures:", bad)
