"""The oracle against everything that pins it: the reference's own qsort test, the known-answer
vectors of SURVEY.md section 8c, an independent float32 transliteration (tests/pyref.py), and the
committed golden arrays."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import oracle as O  # noqa: E402
import pyref  # noqa: E402
from util import GOLDEN, bits_equal, first_mismatch, from_hex, hx, kats, mode_cases, weights_for  # noqa: E402


def test_qselect_median_like_reference_test():
    """internal/qsort/qsort_test.go:25-53: median of a shuffled 1..n is (n+1)/2 or the mean of the two middles"""
    rng = np.random.default_rng(1)
    for n in range(1, 1000):
        a = np.arange(1, n + 1, dtype=np.float32)
        rng.shuffle(a)
        med, _ = O.qselect_median(a)
        want = np.float32((n + 1) // 2) if n & 1 else np.float32(0.5) * (np.float32(n // 2) + np.float32(n // 2 + 1))
        assert med == want, n


def test_qselect_kats():
    for k in kats()["qselect"]:
        med, after = O.qselect_median(k["col"])
        assert med == np.float32(k["median"])
        assert list(after) == k["after"]


def test_column_kats():
    for k in kats()["columns"]:
        col = np.array([[np.nan if v == "nan" else v for v in k["col"]]], dtype=np.float32).T
        res, cl, ch = O.stack(col, k["mode"], k["sig"], k["sig"])
        assert hx(res[0]) == k["hex"], k
        assert [cl, ch] == k["clip"], k


def test_generator_kats():
    g = kats()["generator"]
    first = [from_hex(h) for h in g["first_column_n16_p0"].split()]
    col = O.synth_frames(16, 0, 1)[:, 0]
    assert bits_equal(col, first)
    for row in g["rows"]:
        n, p = row[0], row[1]
        frames = O.synth_frames(n, p, 1)
        w = weights_for(n)
        for name, want in zip(g["order"], row[2:]):
            mode, weighted = (name[:-2], True) if name.endswith("_w") else (name, False)
            res, cl, ch = O.stack(frames, mode, 2.75, 2.75, weights=w if weighted else None)
            got = hx(res[0]) if "/" not in want else "%s/%d/%d" % (hx(res[0]), cl, ch)
            assert got == want, (n, p, name)


def test_project_kat():
    k = kats()["project"]
    trans = np.array([from_hex(h) for h in k["trans_hex"]], dtype=np.float32)
    inv = O.transform_invert(trans)
    assert [hx(v) for v in inv] == k["inverse_hex"]
    src = np.array([10 * r + c for r in range(4) for c in range(4)], dtype=np.float32)
    out = O.project(src, 4, 4, 4, 4, trans, np.float32(np.nan)).reshape(4, 4)
    want = np.array([[from_hex(h) for h in row] for row in k["rows_hex"]], dtype=np.float32)
    assert bits_equal(out, want), first_mismatch(out, want)
    assert bits_equal(pyref.project(src, 4, 4, 4, 4, trans, np.float32(np.nan)), want)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 16, 33, 64])
def test_oracle_vs_independent_python(n):
    """two separately written readings of the Go source agree bit for bit on every mode"""
    frames = O.synth_frames(n, 5000 + n, 24)
    w = weights_for(n)
    for mode, weighted in mode_cases():
        res, cl, ch = O.stack(frames, mode, 2.75, 2.75, weights=w if weighted else None)
        tl = th = 0
        for p in range(frames.shape[1]):
            v, a, b = pyref.stack_column(frames[:, p], mode, 2.75, 2.75, w if weighted else None)
            assert hx(v) == hx(res[p]) or (np.isnan(v) and np.isnan(res[p])), (mode, weighted, p)
            tl += a
            th += b
        assert (tl, th) == (cl, ch), (mode, weighted)


def test_golden_arrays():
    g = np.load(os.path.join(GOLDEN, "stack_golden.npz"))
    for n, p0, count in ((16, 0, 4096), (37, 777, 1031), (256, 4096 * 4096 - 512, 512)):
        frames = O.synth_frames(n, p0, count)
        w = weights_for(n)
        for mode, weighted in mode_cases():
            key = "n%d_p%d_c%d_%s%s" % (n, p0, count, mode, "_w" if weighted else "")
            res, cl, ch = O.stack(frames, mode, 2.75, 2.75, weights=w if weighted else None)
            assert bits_equal(res, g[key]), (key, first_mismatch(res, g[key]))
            assert [cl, ch] == list(g[key + "_clip"]), key


def test_invalid_mode_and_mad_weights():
    frames = O.synth_frames(4, 0, 8)
    with pytest.raises(ValueError):
        O.stack(frames, 7)
    with pytest.raises(RuntimeError):
        O.stack(frames, "mad", weights=weights_for(4))


def test_synth_matches_python_generator():
    got = O.synth_frame(123456, 64, 9)
    want = np.array([pyref.synth_sample(123456 + i, 9) for i in range(64)], dtype=np.float32)
    assert bits_equal(got, want)


def test_goal_seek_host_logic_on_the_oracle():
    """find_sigmas_and_stack (restated dead code, parity unpinned) driven by the oracle: the binary
    search reaches the requested clip percentages, the Newton variant terminates within its 20 steps"""
    import nightlight_b200 as nl
    frames = O.synth_frames(32, 1000, 3000)
    n, p = frames.shape
    steps = []

    def stack_fn(mode):
        def fn(sl, sh):
            steps.append((sl, sh))
            return O.stack(frames, mode, sl, sh)
        return fn
    res, cl, ch, sl, sh = nl.find_sigmas_and_stack(stack_fn("sigma"), nl.ST_SIGMA, n, p, 1.0, 1.5)
    assert int(100 * (cl * 100.0 / (n * p)) + 0.5) == 100 and int(100 * (ch * 100.0 / (n * p)) + 0.5) == 150
    assert 1.0 <= sl <= 11.0 and 1.0 <= sh <= 11.0 and len(steps) <= 21
    want = O.stack(frames, "sigma", sl, sh)
    assert np.array_equal(res.view(np.uint32), want[0].view(np.uint32)) and (cl, ch) == want[1:]
    steps.clear()
    res, cl, ch, sl, sh = nl.find_sigmas_and_stack(stack_fn("linfit"), nl.ST_AUTO, n, p, 2.0, 2.0)   # 32 frames -> linear fit
    assert len(steps) <= 23 and 0.1 <= sl <= 20 and 0.1 <= sh <= 20
    res2, cl2, ch2, _, _ = nl.find_sigmas_and_stack(stack_fn("median"), nl.ST_MEDIAN, n, p, 1.0, 1.0)
    assert (cl2, ch2) == (0, 0)
