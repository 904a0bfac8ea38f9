"""The C-ABI library loads and exports every symbol include/nightlight_cuda.h declares; the host-only
entry points (no device work) agree with the oracle; device entry points fail loudly without a GPU."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import nightlight_b200 as nl  # noqa: E402
from nightlight_b200 import binding  # noqa: E402
from oracle import oracle as O  # noqa: E402
from util import ROOT, bits_equal  # noqa: E402


def declared_in_header():
    text = open(os.path.join(ROOT, "include", "nightlight_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_in_header()
    assert len(names) >= 30
    lib = C.CDLL(nl.library_path())
    for n in names:
        assert hasattr(lib, n), "libnightlight_cuda.so lacks %s" % n
    # the ctypes binding covers the header exactly
    assert sorted(binding.DECLARED_SYMBOLS) == names


def test_version_and_auto_mode():
    lib = nl.load_library()
    assert lib.nl_version() >= 100
    # autoSelectStackingMode, stack.go:45-55
    for n, want in ((1, 1), (5, 1), (6, 2), (14, 2), (15, 3), (24, 3), (25, 5), (1000, 5)):
        assert lib.nl_auto_select_mode(n) == want


def test_no_gpu_fails_loudly():
    cnt = C.c_int(-1)
    rc = nl.load_library().nl_device_count(C.byref(cnt))
    if rc == 0 and cnt.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(nl.NightlightError) as e:
        nl.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_transform_invert_matches_oracle():
    rng = np.random.default_rng(5)
    for _ in range(200):
        t = (rng.standard_normal(6) * np.array([1, 0.1, 50, 0.1, 1, 50])).astype(np.float32)
        assert bits_equal(nl.transform_invert(t), O.transform_invert(t))
    with pytest.raises(nl.NightlightError) as e:
        nl.transform_invert([1, 2, 0, 2, 4, 0])
    assert e.value.code == binding.NL_E_SINGULAR and "Matrix has no inverse" in str(e.value)


def test_get_weights_matches_oracle():
    rng = np.random.default_rng(6)
    n = 23
    frames = [nl.ops.Image(data=np.zeros(1, np.float32), exposure=float(rng.integers(1, 300)),
                           noise=float(rng.random() + 0.5), hfr=float(rng.random() * 3 + 1), id=i) for i in range(n)]
    exposure = np.array([f.exposure for f in frames], np.float32)
    noise = np.array([f.noise for f in frames], np.float32)
    hfr = np.array([f.hfr for f in frames], np.float32)
    fp = C.POINTER(C.c_float)
    for weighting in (1, 2, 3):
        want = np.empty(n, np.float32)
        assert O.lib().nlo_get_weights(weighting, exposure.ctypes.data_as(fp), noise.ctypes.data_as(fp),
                                       hfr.ctypes.data_as(fp), n, want.ctypes.data_as(fp)) == 0
        assert bits_equal(nl.get_weights(frames, weighting), want)
    assert nl.get_weights(frames, 0) is None
    frames[3].exposure = 0.0
    with pytest.raises(nl.NightlightError) as e:
        nl.get_weights(frames, 1)
    assert "Missing exposure information" in str(e.value)
    with pytest.raises(nl.NightlightError):
        nl.get_weights(frames, 9)


def test_stripe_rows_cover_image():
    from nightlight_b200.stripes import all_stripes
    for h in (1, 7, 8, 4096, 4097, 6000):
        for g in (1, 2, 3, 4, 8):
            s = all_stripes(h, g)
            assert s[0][0] == 0 and sum(r for _, r in s) == h
            for (a0, ar), (b0, _) in zip(s, s[1:]):
                assert a0 + ar == b0
            assert max(r for _, r in s) - min(r for _, r in s) <= 1


def test_partition_matches_oracle():
    """OpStackBatches.partition (stackbatches.go:121-210) host mirror against the oracle's restatement"""
    import ctypes as C
    import numpy as np
    import nightlight_b200 as nl
    from oracle import oracle as O
    L = O.lib()
    L.nlo_partition.restype = C.c_int
    L.nlo_partition.argtypes = [C.c_int64] * 5 + [C.c_int, C.c_int] + [C.POINTER(C.c_int64)] * 3
    rng = np.random.default_rng(0)
    for _ in range(400):
        n = int(rng.integers(1, 5000))
        w, h = int(rng.integers(100, 9000)), int(rng.integers(100, 9000))
        mem = int(rng.integers(64, 200000))
        mt = int(rng.integers(1, 65))
        dark, flat = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        nb, bs, mo = C.c_int64(), C.c_int64(), C.c_int64()
        rc = L.nlo_partition(n, w, h, mem, mt, int(dark), int(flat), C.byref(nb), C.byref(bs), C.byref(mo))
        try:
            order, gnb, gbs, gmt = nl.partition(n, w, h, mem, mt, dark, flat, perm=list(rng.permutation(n)))
        except nl.NightlightError as e:
            assert rc != 0 and "Cannot find a stacking execution path" in str(e)
            continue
        assert rc == 0 and (gnb, gbs, gmt) == (nb.value, bs.value, mo.value)
        assert sorted(order) == list(range(n))
        for i in range(gnb):
            chunk = order[i * gbs:(i + 1) * gbs]
            assert chunk == sorted(chunk)


def _seek_native(count_fn, mode, n, p, lo, hi):
    """drive nl_sigma_seek_* (host code of the C ABI, no device work) with a clip-count function"""
    lib = nl.load_library()
    s = binding.SigmaSeek()
    binding.check(lib.nl_sigma_seek_begin(C.byref(s), mode, n, p, lo, hi))
    traj = []
    while not s.done:
        traj.append((s.trial_low, s.trial_high))
        cl, ch = count_fn(s.trial_low, s.trial_high)
        assert lib.nl_sigma_seek_step(C.byref(s), cl, ch) >= 0
    return s.result_low, s.result_high, s.trials, s.converged, traj


@pytest.mark.parametrize("mode", [2, 3, 5, 6, 0])
def test_sigma_goal_seek_state_machine_matches_the_python_restatement(mode):
    """a21: nl_sigma_seek_begin/step (C ABI) follow, trial by trial, the float32 restatement of the reference's
    commented-out FindSigmasAndStack (ops.find_sigmas_and_stack) on synthetic clip-count curves -- smooth ones
    that converge, flat ones that stop Newton's method, and ones that never reach the target (20-step limit)"""
    n, p = 40, 100000
    rng = np.random.default_rng(mode)
    for case in range(40):
        a, b = float(rng.uniform(0.5, 3.0)), float(rng.uniform(0.5, 3.0))
        kind = case % 4
        lo_t, hi_t = float(rng.uniform(0.05, 3.0)), float(rng.uniform(0.05, 3.0))

        def counts(sl, sh):
            if kind == 3:                                   # flat: derivative 0
                return 1234, 4321
            fl = np.exp(-a * sl) * 0.2 + (0.0 if kind != 2 else 0.05)
            fh = np.exp(-b * sh) * 0.2 + (0.001 * sl if kind == 1 else 0.0)   # kind 1: the sides interact (linear fit)
            return int(fl * n * p), int(fh * n * p)

        trace = []

        def stack_fn(sl, sh):
            trace.append((np.float32(sl), np.float32(sh)))
            cl, ch = counts(sl, sh)
            return None, cl, ch

        want = nl.find_sigmas_and_stack(stack_fn, mode, n, p, lo_t, hi_t)
        got = _seek_native(counts, mode, n, p, lo_t, hi_t)
        assert (np.float32(got[0]), np.float32(got[1])) == (np.float32(want[3]), np.float32(want[4])), (case, got, want)
        # the python version's last call is the stack it returns; the native search counts trials only
        resolved = nl.load_library().nl_auto_select_mode(n) if mode == 6 else mode
        if resolved in (2, 3, 5):
            assert [tuple(np.float32(x) for x in t) for t in got[4]] == trace, case
            assert got[2] == len(trace)
        else:
            assert got[4] == [] and got[:2] == (0.0, 0.0)


def test_go_shims_only_use_declared_symbols_with_the_declared_arity():
    """integration/go cannot be compiled here (no Go toolchain): at least every C.nl_* the cgo files name must be declared
    in include/nightlight_cuda.h, and every call must pass as many arguments as the declaration has parameters"""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "nightlight_cuda.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    decl = {}
    for m in re.finditer(r"\b(nl_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S):
        params = m.group(2).strip()
        decl[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    types = set(re.findall(r"\b(nl_[a-z0-9_]+)\b", header))

    def call_args(src, start):
        """number of top-level arguments of the call whose '(' is at src[start]"""
        depth, n, seen = 0, 0, False
        for ch in src[start:]:
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
                if depth == 0:
                    return n + (1 if seen else 0)
            elif ch == "," and depth == 1:
                n += 1
            elif depth >= 1 and not ch.isspace():
                seen = True
        raise AssertionError("unbalanced call")

    files = glob.glob(os.path.join(root, "integration", "go", "**", "*.go"), recursive=True)
    assert files
    calls = 0
    for f in files:
        src = open(f).read()
        src = re.sub(r"//[^\n]*", "", src)
        for m in re.finditer(r"\bC\.(nl_[a-zA-Z0-9_]+)", src):
            name = m.group(1)
            assert name in types, (os.path.basename(f), name, "not in include/nightlight_cuda.h")
            rest = src[m.end():]
            if name in decl and rest.lstrip().startswith("("):
                got = call_args(src, m.end() + (len(rest) - len(rest.lstrip())))
                assert got == decl[name], (os.path.basename(f), name, "passes %d arguments, declared with %d" % (got, decl[name]))
                calls += 1
    assert calls >= 10
