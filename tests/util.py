"""Shared helpers of the test suite (test infrastructure)."""
import json
import os
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def hx(v):
    return "%08x" % struct.unpack("<I", struct.pack("<f", float(v)))[0]


def from_hex(h):
    if h == "nan":
        return np.float32(np.nan)
    return np.frombuffer(struct.pack("<I", int(h, 16)), dtype=np.float32)[0]


def bits_equal(a, b):
    """bit-exact equality of float32 arrays, except that any NaN equals any NaN"""
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    b = np.ascontiguousarray(b, dtype=np.float32).reshape(-1)
    if a.shape != b.shape:
        return False
    an, bn = np.isnan(a), np.isnan(b)
    if not np.array_equal(an, bn):
        return False
    return np.array_equal(a.view(np.uint32)[~an], b.view(np.uint32)[~bn])


def first_mismatch(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    b = np.ascontiguousarray(b, dtype=np.float32).reshape(-1)
    bad = np.nonzero((a.view(np.uint32) != b.view(np.uint32)) & ~(np.isnan(a) & np.isnan(b)))[0]
    if bad.size == 0:
        return "none"
    i = int(bad[0])
    return "%d mismatches, first at %d: %s (%r) vs %s (%r)" % (bad.size, i, hx(a[i]), a[i], hx(b[i]), b[i])


def kats():
    with open(os.path.join(GOLDEN, "kats.json")) as f:
        return json.load(f)


def weights_for(n):
    """SURVEY.md section 8c/8d stand-in weights: w[k] = 1/(1+4*((k%7)/6)) in float32"""
    k = np.arange(n) % 7
    return (np.float32(1) / (np.float32(1) + np.float32(4) * (k.astype(np.float32) / np.float32(6)))).astype(np.float32)


MODES = ["median", "mean", "sigma", "winsor", "mad", "linfit"]
MODE_ID = {"median": 0, "mean": 1, "sigma": 2, "winsor": 3, "mad": 4, "linfit": 5, "auto": 6}


def mode_cases():
    """(mode, weighted) pairs the reference implements (MAD+weights panics; median/linfit ignore weights)"""
    out = []
    for m in MODES:
        out.append((m, False))
        if m in ("mean", "sigma", "winsor"):
            out.append((m, True))
    return out


def synth_frames_threaded(n, p0, length, seed=12345):
    """[n, length] float32 of the synthetic workload (oracle generator), generated with all host threads
    (ctypes calls release the GIL) -- for the full-size parity tests"""
    import ctypes as C
    import threading
    from oracle import oracle as O
    lib = O.lib()
    fp = C.POINTER(C.c_float)
    frames = np.empty((n, length), dtype=np.float32)
    cores = os.cpu_count() or 1
    per = (n + cores - 1) // cores

    def gen(k0, k1):
        for k in range(k0, k1):
            lib.nlo_synth_frame(frames[k].ctypes.data_as(fp), p0, length, k, seed)

    th = [threading.Thread(target=gen, args=(k0, min(n, k0 + per))) for k0 in range(0, n, per)]
    [t.start() for t in th]
    [t.join() for t in th]
    return frames
