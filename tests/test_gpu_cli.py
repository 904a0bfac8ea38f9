"""BASELINE config 1 through the C++ host layer: synthetic fp32 FITS frames -> `nlstack stack` (C ABI ->
CUDA) -> stacked FITS, compared bit for bit with the oracle's stack of the same frames (NaN written as 0
like write.go:192); log lines like the reference's.  Also row stripes over several contexts and the
stack-of-stacks path."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fitsutil import read_fits, write_fits  # noqa: E402
from oracle import oracle as O  # noqa: E402
from util import bits_equal, first_mismatch  # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NLSTACK = os.path.join(ROOT, "host", "nlstack")


@pytest.fixture(scope="module")
def frames_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("c1")
    n, w, h = 16, 1024, 1024
    frames = O.synth_frames(n, 0, w * h)
    for k in range(n):
        write_fits(str(d / ("c1_%02d.fits" % k)), frames[k].reshape(h, w), -32, exposure=30 + k)
    return d, frames, w, h


def run(args):
    r = subprocess.run([NLSTACK] + args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def expect(frames, mode, **kw):
    res, cl, ch = O.stack(frames, mode, **kw)
    res = res.copy()
    res[np.isnan(res)] = 0.0
    return res, cl, ch


def test_config1_sigma_clip_through_fits(frames_dir):
    d, frames, w, h = frames_dir
    files = sorted(str(p) for p in d.glob("c1_*.fits"))
    out = str(d / "out_sigma.fits")
    log = run(["stack", "-stMode", "2", "-stSigLow", "2.75", "-stSigHigh", "2.75", "-out", out] + files)
    want, cl, ch = expect(frames, "sigma")
    hdr, got = read_fits(out)
    assert bits_equal(got, want), first_mismatch(got, want)
    assert "Stacking 16 frames with stacking mode 2 and sigma low 2.75 high 2.75:" in log
    m = re.search(r"Clipped low (\d+) \(([\d.]+)%\) high (\d+) \(([\d.]+)%\)", log)
    assert m and (int(m.group(1)), int(m.group(3))) == (cl, ch)
    assert float(hdr["EXPOSURE"]) == sum(30 + k for k in range(16))       # stack.go:220-221


def test_auto_mode_weights_and_stripes(frames_dir):
    d, frames, w, h = frames_dir
    files = sorted(str(p) for p in d.glob("c1_*.fits"))
    out = str(d / "out_auto.fits")
    # 16 frames -> winsorized sigma (stack.go:45-55), exposure weights, two contexts = two row stripes
    log = run(["stack", "-stWeight", "1", "-gpus", "0,0", "-out", out] + files)
    weights = np.array([30 + k for k in range(16)], np.float32)
    want, cl, ch = expect(frames, "winsor", weights=weights)
    _, got = read_fits(out)
    assert bits_equal(got, want), first_mismatch(got, want)
    assert "stacking mode 3" in log and ("Clipped low %d " % cl) in log


def test_stack_of_stacks_and_errors(frames_dir):
    d, frames, w, h = frames_dir
    files = sorted(str(p) for p in d.glob("c1_*.fits"))
    out = str(d / "out_batches.fits")
    run(["stack", "-stMode", "2", "-stBatch", "8", "-out", out] + files)
    a, _, _ = O.stack(frames[:8], "sigma")
    b, _, _ = O.stack(frames[8:], "sigma")
    acc = (a * np.float32(8)).astype(np.float32)
    acc = (acc + (b * np.float32(8)).astype(np.float32)).astype(np.float32)     # stack.go:924-937
    want = (acc * (np.float32(1) / np.float32(16))).astype(np.float32)          # stack.go:940-944
    want[np.isnan(want)] = 0
    _, got = read_fits(out)
    assert bits_equal(got, want), first_mismatch(got, want)
    r = subprocess.run([NLSTACK, "stack", "-stMode", "9", "-out", out] + files[:2], capture_output=True, text=True)
    assert r.returncode == 1 and "invalid stacking mode" in r.stdout
    r = subprocess.run([NLSTACK, "stack", "-stMode", "4", "-stWeight", "1", "-out", out] + files[:6], capture_output=True, text=True)
    assert r.returncode == 1 and "MADSigma stacking with weights" in r.stdout


def test_stars_verb_with_bad_pixel_repair(tmp_path):
    """`nlstack stars -bpSigLow 3 -bpSigHigh 5 -starBpSig 5`: OpBadPixel (BadPixelMap on the device + in-order sparse
    repair) leaves the frame's MedianDiffStats, whose StdDev is star detection's bad-pixel scale"""
    from test_gpu_project_stars import star_field
    w, h = 640, 480
    img = star_field(w, h, 30, seed=12, hot=60)
    path = str(tmp_path / "field.fits")
    write_fits(path, img.reshape(h, w), -32)
    loc, scale = float(np.float32(np.median(img))), 3.0
    log = run(["stars", "-bpSigLow", "3", "-bpSigHigh", "5", "-starBpSig", "5", "-starSig", "10", "-starRadius", "12",
               "-loc", repr(loc), "-scale", repr(scale), path])
    fixed, removed, st = O.op_bad_pixel(img, w, 3.0, 5.0, amd64=True)
    stars, _, hfr = O.find_stars(fixed, w, np.float32(loc), np.float32(scale), 10.0, 5.0, 1.4, 12, float(st[3]))
    m = re.search(r"Removed (\d+) bad pixels \(([\d.]+)%\) with sigma low=3.00 high=5.00", log)
    assert m and int(m.group(1)) == removed and removed > 20
    m = re.search(r"Stars (\d+) HFR ([\d.]+)", log)
    assert m and int(m.group(1)) == len(stars) and len(stars) > 5
    assert m.group(2) == "%.2f" % float(hfr)
