"""Extended fuzz run against the oracle (GPU box; not part of pytest: minutes, thousands of cases).

    python tools/fuzz/fuzz_prestats.py
"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import nightlight_b200 as nl
from oracle import oracle as O
ctx = nl.Context(0)
bad = 0
def same(a, b):
    a, b = np.asarray(a, np.float32).ravel(), np.asarray(b, np.float32).ravel()
    m = ~(np.isnan(a) & np.isnan(b))
    return np.array_equal(a.view(np.uint32)[m], b.view(np.uint32)[m])
rng = np.random.default_rng(2026)
for it in range(400):
    amd = bool(rng.integers(0, 2))
    ctx.set_numerics(nl.NUMERICS_AMD64 if amd else nl.NUMERICS_PUREGO)
    w = int(rng.choice([1, 2, 3, 5, 7, 8, 9, 13, 14, 15, 16, 17, 31, 32, 33, 100, 257, 640, 1023]))
    h = int(rng.choice([1, 2, 3, 4, 5, 9, 16, 33, 64, 200]))
    scale = float(rng.choice([1e-3, 1.0, 300.0, 6e4]))
    img = (rng.standard_normal((h, w)) * scale + float(rng.choice([0.0, 1000.0, -5.0]))).astype(np.float32)
    if rng.random() < 0.4:
        m = rng.random((h, w)); img[m < 0.03] = 0.0; img[(m > 0.03) & (m < 0.06)] = -0.0
    if rng.random() < 0.3:
        img = np.round(img).astype(np.float32)
    nanny = rng.random() < 0.25
    if nanny:
        img[rng.random((h, w)) < 0.02] = np.nan
    try:
        assert same(nl.median_filter3x3(ctx, img, w), O.median_filter3x3(img, w, amd)), "median"
        if not nanny:
            g, o = nl.stats(ctx, img), O.stats(img, amd)
            assert np.array_equal(g.view(np.uint32), o.view(np.uint32)), ("stats", g, o)
            if h >= 3 and w >= 3:
                g, o = nl.estimate_noise(ctx, img, w), O.estimate_noise(img, w, amd)
                assert g.view(np.uint32) == o.view(np.uint32), ("noise", g, o)
            sl, sh = float(rng.choice([1.0, 3.0])), float(rng.choice([2.0, 5.0]))
            gb, gs = nl.bad_pixel_map(ctx, img, w, sl, sh)
            ob, os_, _ = O.bad_pixel_map(img, w, sl, sh, amd)
            assert np.array_equal(gs.view(np.uint32), os_.view(np.uint32)) and np.array_equal(gb, ob), "bpm"
            gd, gn, _ = nl.op_bad_pixel(ctx, img, w, sl, sh)
            od, on, _ = O.op_bad_pixel(img, w, sl, sh, amd)
            assert gn == on and same(gd, od), "op_bad_pixel"
        else:
            g, o = nl.stats(ctx, img), O.stats(img, amd)
            assert same(g[[0, 2]], o[[0, 2]]) and np.isnan(g).tolist() == np.isnan(o).tolist(), ("stats nan", g, o)
    except AssertionError as e:
        bad += 1; print("FAIL", it, amd, w, h, str(e)[:200])
print("prestats fuzz done, failures:", bad, "replays", ctx.exact_replays())
