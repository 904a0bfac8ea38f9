"""The C++ host layer's FITS in/out (host/nightlight_host.cpp, mirroring internal/fits/read.go and write.go)
against an independent numpy FITS codec.  No GPU needed."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fitsutil import read_fits, write_fits  # noqa: E402
from oracle import oracle as O  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim():
    if not os.path.exists(os.path.join(ROOT, "nightlight_b200", "libnightlight_cuda.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "nightlight_b200", "csrc"), "-j8"])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host")])
    L = C.CDLL(os.path.join(ROOT, "host", "libnl_host_test.so"))
    L.nlh_last_error.restype = C.c_char_p
    L.nlh_fits_write.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float]
    L.nlh_fits_read.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]
    return L


def _read(L, path, n):
    out = np.empty(n, np.float32)
    mmm = np.empty(3, np.float32)
    rc = L.nlh_fits_read(path.encode(), out.ctypes.data_as(C.c_void_p), n, mmm.ctypes.data_as(C.c_void_p))
    assert rc == 0, L.nlh_last_error()
    return out, mmm


def test_write_is_bitpix_minus32_big_endian_nan_to_zero(shim, tmp_path):
    rng = np.random.default_rng(1)
    img = rng.standard_normal((37, 53)).astype(np.float32) * 1000
    img[3, 4] = np.nan
    img[0, 0] = -0.0
    p = str(tmp_path / "w.fits")
    naxisn = np.array([53, 37], np.int32)
    assert shim.nlh_fits_write(p.encode(), img.ctypes.data_as(C.c_void_p), naxisn.ctypes.data_as(C.c_void_p), 2, 120.5) == 0
    hdr, data = read_fits(p)
    assert hdr["SIMPLE"] == "T" and hdr["NAXIS"] == "2" and hdr["EXPOSURE"] == "120.5" and hdr["BZERO"] == "0" and hdr["BSCALE"] == "1"
    want = img.copy()
    want[3, 4] = 0.0                                          # write.go:192: NaN -> 0
    assert np.array_equal(data.view(np.uint32), want.view(np.uint32))
    assert os.path.getsize(p) % 2880 == 0


@pytest.mark.parametrize("bitpix", [8, 16, 32, 64, -32, -64])
def test_read_every_bitpix_with_bzero_bscale(shim, tmp_path, bitpix):
    rng = np.random.default_rng(bitpix + 100)
    h, w = 19, 31
    if bitpix == 8:
        disk = rng.integers(0, 256, (h, w))
    elif bitpix > 0:
        lim = min(2 ** (bitpix - 1) - 1, 2 ** 40)
        disk = rng.integers(-lim, lim, (h, w))
    else:
        disk = rng.standard_normal((h, w)) * 1e3
    bzero, bscale = (32768, 1) if bitpix == 16 else ((0.5, 2.0) if bitpix < 0 else (None, None))
    p = str(tmp_path / ("r%d.fits" % bitpix))
    write_fits(p, disk, bitpix, bzero, bscale, exposure=30)
    got, mmm = _read(shim, p, h * w)
    dt = {8: np.uint8, 16: np.int16, 32: np.int32, 64: np.int64, -32: np.float32, -64: np.float64}[bitpix]
    val = disk.astype(dt).astype(np.float32).reshape(-1)      # float32(val)
    want = (val * np.float32(1 if bscale is None else bscale)).astype(np.float32) + np.float32(0 if bzero is None else bzero)
    assert np.array_equal(got.view(np.uint32), want.astype(np.float32).view(np.uint32))
    assert mmm[0] == want.min() and mmm[2] == want.max()
    assert mmm[1] == np.float32(want.astype(np.float64).sum() / want.size)


def test_header_errors(shim, tmp_path):
    p = str(tmp_path / "bad.fits")
    with open(p, "wb") as f:
        f.write(("%-8s= %20s" % ("SIMPLE", "F")).ljust(80).encode() + "END".ljust(80).encode() + b" " * (2880 - 160))
    out = np.empty(4, np.float32)
    assert shim.nlh_fits_read(p.encode(), out.ctypes.data_as(C.c_void_p), 4, out.ctypes.data_as(C.c_void_p)) == -1
    assert b"SIMPLE=T missing" in shim.nlh_last_error()


def test_header_floats_are_the_shortest_round_trip_like_go(shim, tmp_path):
    """write.go formats EXPOSURE / BZERO / BSCALE with Go's %g of a float32: the shortest decimal that parses back to
    the same float32, exponent form from 1e6 up and below 1e-4 (C's %g would cut 12345.67 to 12345.7)"""
    img = np.zeros((2, 3), np.float32)
    nax = np.array([3, 2], np.int32)
    for exposure, want in ((12345.67, "12345.67"), (300.0, "300"), (0.1, "0.1"), (1234567.0, "1.234567e+06"),
                           (1e-5, "1e-05"), (2.5e10, "2.5e+10"), (16777216.0, "1.6777216e+07"), (0.33333334, "0.33333334")):
        path = str(tmp_path / "e.fits")
        rc = shim.nlh_fits_write(path.encode(), img.ctypes.data_as(C.c_void_p), nax.ctypes.data_as(C.c_void_p), 2, exposure)
        assert rc == 0, shim.nlh_last_error()
        head = open(path, "rb").read(2880).decode("ascii")
        cards = {head[i:i + 8].strip(): head[i + 10:i + 30].strip() for i in range(0, 2880, 80)}
        assert cards["EXPOSURE"] == want, (exposure, cards["EXPOSURE"])
        assert np.float32(float(cards["EXPOSURE"])) == np.float32(exposure)
        assert cards["BZERO"] == "0" and cards["BSCALE"] == "1"
