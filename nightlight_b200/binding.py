"""ctypes binding of include/nightlight_cuda.h.  No torch, no numpy tricks: plain pointers and sizes."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NL_CUDA_LIB: development override to A/B-test another build of the same library
_SO = os.environ.get("NL_CUDA_LIB") or os.path.join(_HERE, "libnightlight_cuda.so")

ST_MEDIAN, ST_MEAN, ST_SIGMA, ST_WINSOR_SIGMA, ST_MAD_SIGMA, ST_LINEAR_FIT, ST_AUTO = range(7)
W_NONE, W_EXPOSURE, W_INVERSE_NOISE, W_INVERSE_HFR = range(4)
NUMERICS_AMD64, NUMERICS_PUREGO = 0, 1

NL_E_INVALID, NL_E_CUDA, NL_E_UNSUPPORTED, NL_E_SINGULAR, NL_E_NOMEM, NL_E_WEIGHTS = -1, -2, -3, -4, -5, -6


class NightlightError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


class SigmaSeek(C.Structure):
    """nl_sigma_seek (include/nightlight_cuda.h): the state of the sigma goal-seek"""
    _fields_ = [("mode", C.c_int32), ("done", C.c_int32), ("converged", C.c_int32), ("trials", C.c_int32),
                ("trial_low", C.c_float), ("trial_high", C.c_float), ("result_low", C.c_float), ("result_high", C.c_float),
                ("step", C.c_int32), ("phase", C.c_int32), ("perc_low", C.c_float), ("perc_high", C.c_float), ("total", C.c_float),
                ("low_l", C.c_float), ("low_r", C.c_float), ("low_m", C.c_float), ("high_l", C.c_float), ("high_r", C.c_float),
                ("high_m", C.c_float), ("sig_lo", C.c_float), ("sig_hi", C.c_float), ("d_l", C.c_float), ("d_h", C.c_float),
                ("new_lo", C.c_float)]


class Star(C.Structure):
    """star.Star, internal/star/findstars.go:30-37"""
    _fields_ = [("index", C.c_int32), ("value", C.c_float), ("x", C.c_float), ("y", C.c_float),
                ("mass", C.c_float), ("hfr", C.c_float)]


STAR_DTYPE = np.dtype([("index", "<i4"), ("value", "<f4"), ("x", "<f4"), ("y", "<f4"), ("mass", "<f4"), ("hfr", "<f4")])

_fp = C.POINTER(C.c_float)
_vp = C.c_void_p
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)

# every symbol include/nightlight_cuda.h declares: name -> (restype, argtypes)
DECLARED_SYMBOLS = {
    "nl_last_error": (C.c_char_p, []),
    "nl_version": (C.c_int, []),
    "nl_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "nl_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "nl_ctx_destroy": (C.c_int, [_vp]),
    "nl_ctx_sync": (C.c_int, [_vp]),
    "nl_ctx_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
    "nl_ctx_device": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "nl_ctx_mem_info": (C.c_int, [_vp, _i64p, _i64p]),
    "nl_ctx_launch_count": (C.c_int, [_vp, _i64p]),
    "nl_ctx_set_tuning": (C.c_int, [_vp, C.c_char_p, C.c_char_p]),
    "nl_stack_job_shape": (C.c_int, [_vp, _i32p, _i64p]),
    "nl_stack_apply_multi": (C.c_int, [C.POINTER(_vp), C.c_int32, C.POINTER(_vp), C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _fp,
                                       C.c_float, C.c_float, C.c_float, _vp, _i64p, _i64p]),
    "nl_stack_clip_counts_only": (C.c_int, [_vp, C.c_int32, _fp, C.c_float, C.c_float, _i64p, _i64p]),
    "nl_sigma_seek_begin": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int64, C.c_float, C.c_float]),
    "nl_sigma_seek_step": (C.c_int, [_vp, C.c_int64, C.c_int64]),
    "nl_find_sigmas_and_stack": (C.c_int, [_vp, C.c_int32, _fp, C.c_float, C.c_float, C.c_float, _vp, _i64p, _i64p, _fp, _fp, _i32p]),
    "nl_stack_begin": (C.c_int, [_vp, C.c_int32, C.c_int64, C.POINTER(_vp)]),
    "nl_stack_put_frame": (C.c_int, [_vp, C.c_int32, _vp, C.c_int64]),
    "nl_stack_frames_dev": (C.c_int, [_vp, C.POINTER(_vp), _i64p]),
    "nl_stack_run": (C.c_int, [_vp, C.c_int32, _fp, C.c_float, C.c_float, C.c_float, _vp, _i64p, _i64p]),
    "nl_stack_apply": (C.c_int, [_vp, C.POINTER(_vp), C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _fp, C.c_float, C.c_float,
                                 C.c_float, _vp, _i64p, _i64p]),
    "nl_stack_apply_release": (C.c_int, [_vp]),
    "nl_stack_run_dev": (C.c_int, [_vp, C.c_int32, _fp, C.c_float, C.c_float, C.c_float, _vp]),
    "nl_stack_run_dev_bcast": (C.c_int, [_vp, C.c_int32, _fp, C.c_float, C.c_float, C.c_float, _vp, C.POINTER(_vp), C.c_int32]),
    "nl_stack_clip_counts": (C.c_int, [_vp, _i64p, _i64p]),
    "nl_stack_end": (C.c_int, [_vp]),
    "nl_auto_select_mode": (C.c_int, [C.c_int32]),
    "nl_get_weights": (C.c_int, [C.c_int32, _fp, _fp, _fp, C.c_int32, _fp]),
    "nl_estimate_noise_dev": (C.c_int, [_vp, _vp, C.c_int32, C.c_int64, C.c_int32, C.c_int32, _fp]),
    "nl_estimate_noise": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, _fp]),
    "nl_ctx_set_numerics": (C.c_int, [_vp, C.c_int32]),
    "nl_ctx_exact_replays": (C.c_int, [_vp, _i64p]),
    "nl_median_filter3x3": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, _vp]),
    "nl_median_filter3x3_dev": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, _vp]),
    "nl_stats": (C.c_int, [_vp, _vp, C.c_int64, _fp]),
    "nl_stats_dev": (C.c_int, [_vp, _vp, C.c_int64, _fp]),
    "nl_bad_pixel_map": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, C.c_float, C.c_float, _i32p, C.c_int64, _i64p, _fp]),
    "nl_bad_pixel_map_dev": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, C.c_float, C.c_float, _vp, _i32p, C.c_int64, _i64p, _fp]),
    "nl_op_bad_pixel": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, C.c_float, C.c_float, _i64p, _fp]),
    "nl_op_bad_pixel_dev": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int32, C.c_float, C.c_float, _i64p, _fp]),
    "nl_find_stars_dev": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32] + [C.c_float] * 5 + [C.c_int32, C.c_float, _vp, C.c_int32,
                                    _i32p, _fp, _fp]),
    "nl_stack_incremental_dev": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_float, C.c_int]),
    "nl_stack_incremental_finalize_dev": (C.c_int, [_vp, _vp, C.c_int64, C.c_float]),
    "nl_transform_invert": (C.c_int, [_fp, _fp]),
    "nl_project": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, _vp, C.c_int32, C.c_int32, _fp, C.c_float]),
    "nl_project_dev": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, _vp, C.c_int32, C.c_int32, _fp, C.c_float]),
    "nl_project_scaled": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, _vp, C.c_int32, C.c_int32, _fp, C.c_float, C.c_float, C.c_float]),
    "nl_project_scaled_dev": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, _vp, C.c_int32, C.c_int32, _fp, C.c_float, C.c_float, C.c_float]),
    "nl_project_scatter_dev": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _fp, C.c_float, C.c_float, C.c_float,
                                         C.c_int32, C.POINTER(_vp), _i32p, C.c_int32]),
    "nl_fits_decode": (C.c_int, [_vp, _vp, C.c_int32, C.c_int64, C.c_float, C.c_float, _vp]),
    "nl_fits_decode_dev": (C.c_int, [_vp, _vp, C.c_int32, C.c_int64, C.c_float, C.c_float, _vp]),
    "nl_fits_encode": (C.c_int, [_vp, _vp, C.c_int64, _vp]),
    "nl_fits_encode_dev": (C.c_int, [_vp, _vp, C.c_int64, _vp]),
    "nl_stack_put_frame_raw": (C.c_int, [_vp, C.c_int32, _vp, C.c_int32, C.c_int64, C.c_float, C.c_float]),
    "nl_find_bright": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_float, C.c_int32, _vp, C.c_int32, _i32p]),
    "nl_find_bright_dev": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_float, C.c_int32, _vp, C.c_int32, _i32p]),
    "nl_find_stars": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                C.c_int32, C.c_float, _vp, C.c_int32, _i32p, _fp, _fp]),
    "nl_bad_pixel_map_batch_dev": (C.c_int, [_vp, _vp, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_float, C.c_float, _i32p, C.c_int64,
                                            _i64p, _fp]),
    "nl_project_batch_dev": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, C.c_int32, _vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _fp,
                                      C.c_float, _fp, _fp]),
    "nl_find_bright_batch_dev": (C.c_int, [_vp, _vp, C.c_int32, C.c_int64, C.c_int32, C.c_int32, _fp, C.c_int32, _vp, C.c_int32, _i32p]),
    "nl_find_stars_batch_dev": (C.c_int, [_vp, _vp, C.c_int32, C.c_int64, C.POINTER(_vp), C.c_int32, C.c_int32, _fp, _fp, C.c_float,
                                         C.c_float, C.c_float, C.c_int32, _fp, _vp, C.c_int32, _i32p, _fp, _fp,
                                         C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "nl_star_reject_bad_pixels_host": (C.c_int, [_vp, C.c_int32, _fp, C.c_int32, C.c_int32, C.c_float, C.c_float, _i32p]),
    "nl_star_sort_desc_host": (C.c_int, [_vp, C.c_int32]),
    "nl_star_filter_overlaps_host": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _i32p]),
    "nl_synth_fill_dev": (C.c_int, [_vp, _vp, C.c_uint64, C.c_int64, C.c_uint32, C.c_uint32]),
    "nl_dev_alloc": (C.c_int, [_vp, C.c_int64, C.POINTER(_vp)]),
    "nl_dev_free": (C.c_int, [_vp, _vp]),
    "nl_ipc_get_handle": (C.c_int, [_vp, _vp, C.c_char_p]),
    "nl_ipc_open_handle": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "nl_ipc_close_handle": (C.c_int, [_vp, _vp]),
    "nl_host_alloc_pinned": (C.c_int, [C.c_int64, C.POINTER(_vp)]),
    "nl_host_free_pinned": (C.c_int, [_vp]),
    "nl_host_register": (C.c_int, [_vp, C.c_int64]),
    "nl_host_unregister": (C.c_int, [_vp]),
    "nl_memcpy_h2d": (C.c_int, [_vp, _vp, _vp, C.c_int64]),
    "nl_memcpy_d2h": (C.c_int, [_vp, _vp, _vp, C.c_int64]),
    "nl_memcpy_d2d": (C.c_int, [_vp, _vp, _vp, C.c_int64]),
}

_lib = None


def library_path():
    return _SO


def load_library():
    """Loads libnightlight_cuda.so.  Fails loudly when it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise NightlightError(NL_E_CUDA, "libnightlight_cuda.so is missing (%s): build it with "
                                  "`make -C nightlight_b200/csrc` or __graft_entry__.build(); there is no CPU fallback" % _SO)
        lib = C.CDLL(_SO)
        for name, (res, args) in DECLARED_SYMBOLS.items():
            fn = getattr(lib, name)     # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load_library().nl_last_error()
        raise NightlightError(rc, (msg or b"").decode("utf-8", "replace") or "nightlight_cuda error %d" % rc)


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


class Context:
    """One CUDA device + one stream (nl_ctx)."""

    def __init__(self, device=0):
        self._h = _vp()
        check(load_library().nl_ctx_create(int(device), C.byref(self._h)))
        self.device = int(device)

    def close(self):
        if self._h:
            load_library().nl_ctx_destroy(self._h)
            self._h = _vp()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def sync(self):
        check(load_library().nl_ctx_sync(self._h))

    @property
    def stream(self):
        s = _vp()
        check(load_library().nl_ctx_stream(self._h, C.byref(s)))
        return s.value or 0

    def set_numerics(self, numerics):
        """NUMERICS_AMD64 (default: the reference's AVX2 kernels) or NUMERICS_PUREGO for the frame statistics"""
        check(load_library().nl_ctx_set_numerics(self._h, int(numerics)))

    def exact_replays(self):
        n = C.c_int64()
        check(load_library().nl_ctx_exact_replays(self._h, C.byref(n)))
        return n.value

    def set_tuning(self, key, value):
        """nl_ctx_set_tuning: "defer_passes", "tile_width", "stats_debug", "stats_force_replay" (A/B measurements, tests)"""
        check(load_library().nl_ctx_set_tuning(self._h, str(key).encode(), str(value).encode()))

    def mem_info(self):
        """(free, total) bytes of device memory"""
        f, t = C.c_int64(), C.c_int64()
        check(load_library().nl_ctx_mem_info(self._h, C.byref(f), C.byref(t)))
        return f.value, t.value

    @property
    def launch_count(self):
        n = C.c_int64()
        check(load_library().nl_ctx_launch_count(self._h, C.byref(n)))
        return n.value

    # raw device memory (bench / tests)
    def dev_alloc(self, nbytes):
        p = _vp()
        check(load_library().nl_dev_alloc(self._h, int(nbytes), C.byref(p)))
        return p.value

    def dev_free(self, ptr):
        check(load_library().nl_dev_free(self._h, _vp(ptr)))

    def h2d(self, dev, host_array):
        a = np.ascontiguousarray(host_array)
        check(load_library().nl_memcpy_h2d(self._h, _vp(dev), a.ctypes.data_as(_vp), a.nbytes))
        self.sync()

    def d2h(self, host_array, dev):
        assert host_array.flags["C_CONTIGUOUS"]
        check(load_library().nl_memcpy_d2h(self._h, host_array.ctypes.data_as(_vp), _vp(dev), host_array.nbytes))
        self.sync()

    # cross-process peer mapping (CUDA IPC)
    def ipc_handle(self, dev):
        buf = C.create_string_buffer(64)
        check(load_library().nl_ipc_get_handle(self._h, _vp(dev), buf))
        return buf.raw

    def ipc_open(self, handle):
        p = _vp()
        check(load_library().nl_ipc_open_handle(self._h, handle, C.byref(p)))
        return p.value

    def ipc_close(self, dev):
        check(load_library().nl_ipc_close_handle(self._h, _vp(dev)))

    def synth_fill(self, dev, p0, count, frame, seed=12345):
        check(load_library().nl_synth_fill_dev(self._h, _vp(dev), int(p0), int(count), int(frame), int(seed)))


class StackJob:
    """nl_stack_job: N frames (or one row stripe of them) resident in device memory, frame-major."""

    def __init__(self, ctx, n_frames, pixels):
        self.ctx, self.n_frames, self.pixels = ctx, int(n_frames), int(pixels)
        self._h = _vp()
        check(load_library().nl_stack_begin(ctx.handle, self.n_frames, self.pixels, C.byref(self._h)))

    def close(self):
        if self._h:
            load_library().nl_stack_end(self._h)
            self._h = _vp()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def put_frame(self, i, host):
        """host: float32 array of `pixels` samples, or a raw (pointer, count) pair."""
        if isinstance(host, tuple):
            ptr, count = host
            check(load_library().nl_stack_put_frame(self._h, int(i), _vp(ptr), int(count)))
            return
        # pageable memory is staged by the CUDA runtime before the call returns; a pinned buffer must
        # stay valid until the next sync
        a = _f32(host).reshape(-1)
        check(load_library().nl_stack_put_frame(self._h, int(i), a.ctypes.data_as(_vp), a.size))

    def put_frame_raw(self, i, raw, bitpix, bscale=1.0, bzero=0.0):
        """raw: the frame's big-endian FITS payload (bytes / uint8 array); decoded on the device"""
        a = np.frombuffer(raw, dtype=np.uint8) if isinstance(raw, (bytes, bytearray, memoryview)) else np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
        count = a.size // (abs(int(bitpix)) // 8)
        check(load_library().nl_stack_put_frame_raw(self._h, int(i), a.ctypes.data_as(_vp), int(bitpix), count, bscale, bzero))
        self.ctx.sync()                     # `a` may be a temporary

    @property
    def frames_dev(self):
        p, stride = _vp(), C.c_int64()
        check(load_library().nl_stack_frames_dev(self._h, C.byref(p), C.byref(stride)))
        return p.value, stride.value

    def synth_fill(self, p0=0, seed=12345):
        base, stride = self.frames_dev
        for k in range(self.n_frames):
            self.ctx.synth_fill(base + 4 * k * stride, p0, self.pixels, k, seed)

    @staticmethod
    def _weights(weights, n):
        if weights is None:
            return None, None
        w = _f32(weights).reshape(-1)
        if w.size != n:
            raise NightlightError(NL_E_INVALID, "weights must have one entry per frame")
        return w, w.ctypes.data_as(_fp)

    def run(self, mode, weights=None, sigma_low=2.75, sigma_high=2.75, ref_frame_loc=0.0, out=None):
        """-> (result float32[pixels], clipLow, clipHigh); blocks until the result is on the host."""
        w, wp = self._weights(weights, self.n_frames)
        if out is None:
            out = np.empty(self.pixels, dtype=np.float32)
        cl, ch = C.c_int64(), C.c_int64()
        check(load_library().nl_stack_run(self._h, int(mode), wp, sigma_low, sigma_high, ref_frame_loc,
                                          out.ctypes.data_as(_vp), C.byref(cl), C.byref(ch)))
        return out, cl.value, ch.value

    def run_dev(self, mode, dev_out, weights=None, sigma_low=2.75, sigma_high=2.75, ref_frame_loc=0.0):
        """asynchronous; result stays in device memory at dev_out"""
        w, wp = self._weights(weights, self.n_frames)
        check(load_library().nl_stack_run_dev(self._h, int(mode), wp, sigma_low, sigma_high, ref_frame_loc, _vp(dev_out)))
        if w is not None:
            self.ctx.sync()

    def run_dev_bcast(self, mode, dev_out, peer_outs, weights=None, sigma_low=2.75, sigma_high=2.75, ref_frame_loc=0.0):
        """like run_dev, and the result is also stored to every device pointer of peer_outs (peer stripes)"""
        w, wp = self._weights(weights, self.n_frames)
        arr = (_vp * max(len(peer_outs), 1))(*[_vp(p) for p in peer_outs])
        check(load_library().nl_stack_run_dev_bcast(self._h, int(mode), wp, sigma_low, sigma_high, ref_frame_loc, _vp(dev_out),
                                                    arr, len(peer_outs)))
        if w is not None:
            self.ctx.sync()

    def frame_noise(self, width):
        """stats.EstimateNoise of every resident frame (whole frames of `width` columns) -> float32[n_frames]"""
        base, stride = self.frames_dev
        out = np.empty(self.n_frames, dtype=np.float32)
        check(load_library().nl_estimate_noise_dev(self.ctx.handle, _vp(base), self.n_frames, stride, int(width),
                                                   self.pixels // int(width), out.ctypes.data_as(_fp)))
        return out

    def clip_counts(self):
        cl, ch = C.c_int64(), C.c_int64()
        check(load_library().nl_stack_clip_counts(self._h, C.byref(cl), C.byref(ch)))
        return cl.value, ch.value

    def clip_counts_only(self, mode, weights=None, sigma_low=2.75, sigma_high=2.75):
        """one stacking pass that only counts what it clips (no image written) -> (clipLow, clipHigh)"""
        w, wp = self._weights(weights, self.n_frames)
        cl, ch = C.c_int64(), C.c_int64()
        check(load_library().nl_stack_clip_counts_only(self._h, int(mode), wp, sigma_low, sigma_high, C.byref(cl), C.byref(ch)))
        return cl.value, ch.value

    def find_sigmas_and_stack(self, mode, clip_perc_low, clip_perc_high, weights=None, ref_frame_loc=0.0):
        """nl_find_sigmas_and_stack -> (result, clipLow, clipHigh, sigmaLow, sigmaHigh, trials)"""
        w, wp = self._weights(weights, self.n_frames)
        out = np.empty(self.pixels, dtype=np.float32)
        cl, ch, sl, sh, tr = C.c_int64(), C.c_int64(), C.c_float(), C.c_float(), C.c_int32()
        check(load_library().nl_find_sigmas_and_stack(self._h, int(mode), wp, ref_frame_loc, clip_perc_low, clip_perc_high,
                                                      out.ctypes.data_as(_vp), C.byref(cl), C.byref(ch), C.byref(sl), C.byref(sh),
                                                      C.byref(tr)))
        return out, cl.value, ch.value, sl.value, sh.value, tr.value
