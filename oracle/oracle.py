"""ctypes front end of the parity oracle (oracle/nl_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package nightlight_b200 never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnl_oracle.so")

MODES = {"median": 0, "mean": 1, "sigma": 2, "winsor": 3, "mad": 4, "linfit": 5, "auto": 6}


class Star(C.Structure):
    _fields_ = [("index", C.c_int32), ("value", C.c_float), ("x", C.c_float), ("y", C.c_float),
                ("mass", C.c_float), ("hfr", C.c_float)]


STAR_DTYPE = np.dtype([("index", "<i4"), ("value", "<f4"), ("x", "<f4"), ("y", "<f4"),
                       ("mass", "<f4"), ("hfr", "<f4")])


class Transform(C.Structure):
    _fields_ = [(n, C.c_float) for n in "abcdef"]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("nl_oracle.c", "nl_oracle_amd64.c", "nl_oracle_simd.c", "nl_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        fp = C.POINTER(C.c_float)
        L.nlo_qselect_median_f32.restype = C.c_float
        L.nlo_qselect_median_f32.argtypes = [fp, C.c_int]
        L.nlo_qselect_f32.restype = C.c_float
        L.nlo_qselect_f32.argtypes = [fp, C.c_int, C.c_int]
        L.nlo_qsort_f32.argtypes = [fp, C.c_int]
        L.nlo_qpartition_f32.restype = C.c_int
        L.nlo_qpartition_f32.argtypes = [fp, C.c_int]
        L.nlo_mean_stddev.argtypes = [fp, C.c_int, fp, fp]
        L.nlo_linear_regression.argtypes = [fp, fp, C.c_int] + [fp] * 6
        L.nlo_estimate_noise.restype = C.c_float
        L.nlo_estimate_noise.argtypes = [fp, C.c_int32, C.c_int32]
        L.nlo_stack_apply.restype = C.c_int
        L.nlo_stack_apply.argtypes = [C.c_int, C.POINTER(fp), C.c_int, C.c_size_t, fp, C.c_float, C.c_float,
                                      C.c_float, fp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int]
        L.nlo_get_weights.restype = C.c_int
        L.nlo_get_weights.argtypes = [C.c_int, fp, fp, fp, C.c_int, fp]
        L.nlo_stack_incremental.argtypes = [fp, fp, C.c_size_t, C.c_float, C.c_int]
        L.nlo_stack_incremental_finalize.argtypes = [fp, C.c_size_t, C.c_float]
        L.nlo_partition.restype = C.c_int
        L.nlo_partition.argtypes = [C.c_int64] * 5 + [C.c_int, C.c_int] + [C.POINTER(C.c_int64)] * 3
        L.nlo_transform_invert.restype = C.c_int
        L.nlo_transform_invert.argtypes = [C.POINTER(Transform), C.POINTER(Transform)]
        L.nlo_new_transform2d.restype = C.c_int
        L.nlo_new_transform2d.argtypes = [fp, C.POINTER(Transform)]
        L.nlo_project.restype = C.c_int
        L.nlo_project.argtypes = [fp, C.c_int32, C.c_int32, fp, C.c_int32, C.c_int32, C.POINTER(Transform), C.c_float]
        L.nlo_median9.restype = C.c_float
        L.nlo_median9.argtypes = [fp]
        sp = C.POINTER(Star)
        L.nlo_find_bright_pixels.restype = C.c_int
        L.nlo_find_bright_pixels.argtypes = [fp, C.c_int32, C.c_int32, C.c_float, C.c_int32, sp, C.c_int]
        L.nlo_find_stars.restype = C.c_int
        L.nlo_find_stars.argtypes = [fp, C.c_int32, C.c_int32] + [C.c_float] * 5 + [C.c_int32, C.c_float, sp, C.c_int, fp, fp]
        L.nlo_synth_sample.restype = C.c_float
        L.nlo_synth_sample.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.nlo_synth_frame.argtypes = [fp, C.c_uint64, C.c_size_t, C.c_uint32, C.c_uint32]
        L.nlo_lowbias32.restype = C.c_uint32
        L.nlo_lowbias32.argtypes = [C.c_uint32]
        L.nlo_stats.argtypes = [fp, C.c_int64, C.c_int, fp]
        L.nlo_estimate_noise_amd64.restype = C.c_float
        L.nlo_estimate_noise_amd64.argtypes = [fp, C.c_int32, C.c_int32]
        L.nlo_estimate_noise_line_avx2.restype = C.c_float
        L.nlo_estimate_noise_line_avx2.argtypes = [fp, C.c_int64]
        L.nlo_median_filter3x3.argtypes = [fp, fp, C.c_int32, C.c_int32, C.c_int]
        L.nlo_bad_pixel_map.restype = C.c_int64
        L.nlo_bad_pixel_map.argtypes = [fp, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_int, fp,
                                        C.POINTER(C.c_int32), C.c_int64, fp]
        L.nlo_op_bad_pixel.restype = C.c_int64
        L.nlo_op_bad_pixel.argtypes = [fp, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_int, fp, C.POINTER(C.c_int32), fp]
        L.nlo_calc_variance_avx2.restype = C.c_double
        L.nlo_calc_variance_avx2.argtypes = [fp, C.c_int64, C.c_float]
        L.nlo_calc_min_mean_max_avx2.argtypes = [fp, C.c_int64, fp, fp, fp]
        _lib = L
    return _lib


_simd = None


def simd():
    """libnl_oracle_simd.so (the AVX2 kernels replayed with real instructions), or None when this host
    has no AVX2+FMA or the file did not build.  Only a cross-check of nl_oracle_amd64.c."""
    global _simd
    if _simd is None:
        build()
        so = os.path.join(_HERE, "libnl_oracle_simd.so")
        try:
            flags = open("/proc/cpuinfo").read()
        except OSError:
            flags = ""
        if not os.path.exists(so) or " avx2" not in flags or " fma" not in flags:
            _simd = False
        else:
            L = C.CDLL(so)
            fp = C.POINTER(C.c_float)
            L.nlo_simd_min_mean_max.argtypes = [fp, C.c_int64, fp, fp, fp]
            L.nlo_simd_variance.restype = C.c_double
            L.nlo_simd_variance.argtypes = [fp, C.c_int64, C.c_float]
            L.nlo_simd_noise_line.restype = C.c_float
            L.nlo_simd_noise_line.argtypes = [fp, C.c_int64]
            L.nlo_simd_median_line.argtypes = [fp, fp, C.c_int64]
            _simd = L
    return _simd or None


def stats(data, amd64=True):
    """Stats.Min/Mean/Max/StdDev -> float32[4]"""
    data = np.ascontiguousarray(data, dtype=np.float32).ravel()
    out = np.zeros(4, np.float32)
    lib().nlo_stats(_fp(data), data.size, int(amd64 and data.size % 4 == 0), _fp(out))
    return out


def estimate_noise(data, width, amd64=True):
    data = np.ascontiguousarray(data, dtype=np.float32).ravel()
    f = lib().nlo_estimate_noise_amd64 if amd64 else lib().nlo_estimate_noise
    return np.float32(f(_fp(data), int(width), data.size // int(width)))


def median_filter3x3(data, width, amd64=True):
    data = np.ascontiguousarray(data, dtype=np.float32).ravel()
    out = np.empty_like(data)
    lib().nlo_median_filter3x3(_fp(out), _fp(data), int(width), data.size // int(width), int(amd64))
    return out


def bad_pixel_map(data, width, sigma_low, sigma_high, amd64=True):
    """pre.BadPixelMap -> (bpm int32[], stats float32[4] of data - median3x3, diff image)"""
    data = np.ascontiguousarray(data, dtype=np.float32).ravel()
    tmp = np.empty_like(data)
    bpm = np.empty(data.size, np.int32)
    st = np.zeros(4, np.float32)
    n = lib().nlo_bad_pixel_map(_fp(data), data.size, int(width), float(sigma_low), float(sigma_high), int(amd64),
                                _fp(tmp), bpm.ctypes.data_as(C.POINTER(C.c_int32)), bpm.size, _fp(st))
    return bpm[:n].copy(), st, tmp


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def qselect_median(a):
    a = np.ascontiguousarray(a, dtype=np.float32).copy()
    v = lib().nlo_qselect_median_f32(_fp(a), len(a))
    return np.float32(v), a


def qsort(a):
    a = np.ascontiguousarray(a, dtype=np.float32).copy()
    lib().nlo_qsort_f32(_fp(a), len(a))
    return a


def mean_stddev(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    m, s = C.c_float(), C.c_float()
    lib().nlo_mean_stddev(_fp(a), len(a), C.byref(m), C.byref(s))
    return np.float32(m.value), np.float32(s.value)


def linear_regression(xs, ys):
    xs = np.ascontiguousarray(xs, dtype=np.float32)
    ys = np.ascontiguousarray(ys, dtype=np.float32)
    out = [C.c_float() for _ in range(6)]
    lib().nlo_linear_regression(_fp(xs), _fp(ys), len(xs), *[C.byref(o) for o in out])
    return tuple(np.float32(o.value) for o in out)


def stack(frames, mode, sig_lo=2.75, sig_hi=2.75, weights=None, ref_loc=0.0, threads=0):
    """frames: [N, P] float32 (or list of 1-D arrays). Returns (res[P], clipLow, clipHigh)."""
    if isinstance(mode, str):
        mode = MODES[mode]
    frames = [np.ascontiguousarray(f, dtype=np.float32).reshape(-1) for f in frames]
    n, p = len(frames), frames[0].size
    ptrs = (C.POINTER(C.c_float) * n)(*[_fp(f) for f in frames])
    res = np.empty(p, dtype=np.float32)
    w = None
    if weights is not None:
        w = np.ascontiguousarray(weights, dtype=np.float32)
    cl, ch = C.c_int64(), C.c_int64()
    rc = lib().nlo_stack_apply(mode, ptrs, n, p, _fp(w) if w is not None else None, ref_loc, sig_lo, sig_hi,
                               _fp(res), C.byref(cl), C.byref(ch), threads)
    if rc == -1:
        raise ValueError("invalid stacking mode")
    if rc == -2:
        raise RuntimeError("MADSigma stacking with weights is still unimplemented")
    return res, cl.value, ch.value


def project(src, sw, sh, dw, dh, trans, oob):
    src = np.ascontiguousarray(src, dtype=np.float32).reshape(-1)
    dst = np.empty(dw * dh, dtype=np.float32)
    t = Transform(*[float(x) for x in trans])
    rc = lib().nlo_project(_fp(src), sw, sh, _fp(dst), dw, dh, C.byref(t), oob)
    if rc != 0:
        raise ZeroDivisionError("Matrix has no inverse")
    return dst


def transform_invert(trans):
    t, inv = Transform(*[float(x) for x in trans]), Transform()
    if lib().nlo_transform_invert(C.byref(t), C.byref(inv)) != 0:
        raise ZeroDivisionError("Matrix has no inverse")
    return np.array([inv.a, inv.b, inv.c, inv.d, inv.e, inv.f], dtype=np.float32)


def find_bright_pixels(data, width, threshold, radius):
    data = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
    n = lib().nlo_find_bright_pixels(_fp(data), data.size, width, threshold, radius, None, 0)
    out = np.zeros(max(n, 1), dtype=STAR_DTYPE)
    lib().nlo_find_bright_pixels(_fp(data), data.size, width, threshold, radius,
                                 out.ctypes.data_as(C.POINTER(Star)), n)
    return out[:n]


def find_stars(data, width, location, scale, star_sig, bp_sigma, star_in_out, radius, median_diff_stddev=0.0):
    data = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
    cap = max(data.size // 100, 1024)
    out = np.zeros(cap, dtype=STAR_DTYPE)
    sos, hfr = C.c_float(), C.c_float()
    n = lib().nlo_find_stars(_fp(data), data.size, width, location, scale, star_sig, bp_sigma, star_in_out, radius,
                             median_diff_stddev, out.ctypes.data_as(C.POINTER(Star)), cap, C.byref(sos), C.byref(hfr))
    return out[:min(n, cap)], np.float32(sos.value), np.float32(hfr.value)


def synth_frame(p0, length, k, seed=12345):
    dst = np.empty(length, dtype=np.float32)
    lib().nlo_synth_frame(_fp(dst), p0, length, k, seed)
    return dst


def synth_frames(n, p0, length, seed=12345):
    return np.stack([synth_frame(p0, length, k, seed) for k in range(n)])


def op_bad_pixel(data, width, sigma_low, sigma_high, amd64=True):
    """OpBadPixel.Apply (monochrome) -> (repaired data, number removed, medianDiffStats)"""
    data = np.array(data, dtype=np.float32).ravel()
    tmp = np.empty_like(data)
    bpm = np.empty(data.size, np.int32)
    st = np.zeros(4, np.float32)
    n = lib().nlo_op_bad_pixel(_fp(data), data.size, int(width), float(sigma_low), float(sigma_high), int(amd64),
                               _fp(tmp), bpm.ctypes.data_as(C.POINTER(C.c_int32)), _fp(st))
    return data, int(n), st
