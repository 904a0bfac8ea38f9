import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, ctypes as C
from oracle import oracle as O
from util import bits_equal, first_mismatch, MODE_ID
import subprocess
so = "/root/repo/tests/hostemul/libnl_hostemul.so"
L = C.CDLL(so); fp = C.POINTER(C.c_float)
L.emul_stack.restype = C.c_int
L.emul_stack.argtypes = [C.c_int, C.POINTER(fp), C.c_int, C.c_size_t, fp, C.c_float, C.c_float, C.c_float, fp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
def emul(frames, mode, sl, sh, w, ref):
    frames = [np.ascontiguousarray(f, dtype=np.float32) for f in frames]
    n, p = len(frames), frames[0].size
    ptrs = (fp * n)(*[f.ctypes.data_as(fp) for f in frames]); res = np.empty(p, np.float32)
    cl, ch = C.c_longlong(), C.c_longlong()
    wp = None
    if w is not None: w = np.ascontiguousarray(w, dtype=np.float32); wp = w.ctypes.data_as(fp)
    assert L.emul_stack(MODE_ID[mode], ptrs, n, p, wp, ref, sl, sh, res.ctypes.data_as(fp), C.byref(cl), C.byref(ch)) == 0
    return res, cl.value, ch.value
bad = 0
for seed in range(400):
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.choice([2, 3, 5, 8, 15, 16, 17, 31, 32, 33, 64, 100, 255, 256, 300]))
    p = int(rng.integers(1, 400))
    scale = float(rng.choice([1e-30, 1e-3, 1.0, 50.0, 4e4, 1e30, 3e37]))
    dist = rng.choice(["normal", "t", "ties", "bimodal"])
    if dist == "normal": fr = rng.standard_normal((n, p))
    elif dist == "t": fr = rng.standard_t(1.5, size=(n, p))
    elif dist == "ties": fr = np.round(rng.standard_normal((n, p)) * 2)
    else: fr = np.where(rng.random((n, p)) < 0.5, -1.0, 1.0) * (1 + 0.01 * rng.standard_normal((n, p)))
    fr = (fr * scale + float(rng.choice([0.0, 1000.0, -3.0])) * (scale if scale > 1e20 else 1)).astype(np.float32)
    if rng.random() < 0.6: fr[rng.random(fr.shape) < float(rng.choice([0.001, 0.05, 0.4]))] = np.nan
    if rng.random() < 0.2: fr[rng.random(fr.shape) < 0.02] = np.inf
    if rng.random() < 0.2: fr[rng.random(fr.shape) < 0.02] = -np.inf
    sl, sh = (float(x) for x in rng.choice([0.0, 0.5, 1.0, 2.0, 2.75, 4.0, -1.0], 2))
    weighted = bool(rng.integers(0, 2))
    w = (rng.random(n).astype(np.float32) + np.float32(0.05)) if weighted else None
    with np.errstate(all="ignore"):
        want = O.stack(fr, "winsor", sl, sh, weights=w, ref_loc=1.5)
        got = emul(fr, "winsor", sl, sh, w, 1.5)
    if not (bits_equal(got[0], want[0]) and got[1:] == want[1:]):
        bad += 1; print("FAIL seed", seed, n, p, scale, dist, sl, sh, weighted, first_mismatch(got[0], want[0]), got[1:], want[1:])
print("winsor host-emulation fuzz: 400 cases,", bad, "failures")
