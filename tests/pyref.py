"""Independent pure-Python/NumPy float32 transliteration of the reference's per-pixel stackers.

Test infrastructure: a second, separately written restatement of the Go code used only to
cross-check oracle/nl_oracle.c on small columns (two independent readings of the reference
must agree bit for bit).  Every np.float32 op rounds to fp32 once, like Go's float32 math;
no fused multiply-add exists here.

Reference: internal/qsort/qsort.go:68-126, internal/stats/stats.go:246-261,569-586,
internal/ops/stack/stack.go:274-918, internal/fits/project.go:26-76,
internal/star/coord.go:141-201.
"""
import math

import numpy as np

f32 = np.float32
_ERR = np.seterr(all="ignore")


def qselect(a, k):
    """qsort.go:94-126 (k is 1-based, list a is permuted in place)."""
    left, right = 0, len(a) - 1
    while left < right:
        pivot = a[(left + right) >> 1]
        l, r = left - 1, right + 1
        while True:
            while True:
                l += 1
                if a[l] >= pivot:
                    break
            while True:
                r -= 1
                if a[r] <= pivot:
                    break
            if l >= r:
                break
            a[l], a[r] = a[r], a[l]
        index = r
        offset = index - left + 1
        if k <= offset:
            right = index
        else:
            left = index + 1
            k -= offset
    return a[left]


def qselect_median(a):
    """qsort.go:68-82"""
    n = len(a)
    k = (n >> 1) + 1
    upper = qselect(a, k)
    if n & 1:
        return upper
    lower = a[0]
    for i in range(1, k - 1):
        if a[i] > lower:
            lower = a[i]
    return f32(0.5) * (lower + upper)


def qpartition(a, lo, hi):
    """qsort.go:38-56 on a[lo..hi] inclusive, returns absolute pivot index."""
    pivot = a[(lo + hi) >> 1]
    l, r = lo - 1, hi + 1
    while True:
        while True:
            l += 1
            if a[l] >= pivot:
                break
        while True:
            r -= 1
            if a[r] <= pivot:
                break
        if l >= r:
            return r
        a[l], a[r] = a[r], a[l]


def qsort(a, lo=0, hi=None):
    """qsort.go:26-32"""
    if hi is None:
        hi = len(a) - 1
    if hi - lo + 1 > 1:
        idx = qpartition(a, lo, hi)
        qsort(a, lo, idx)
        qsort(a, idx + 1, hi)


def mean_stddev(xs):
    """stats.go:246-261"""
    n = f32(len(xs))
    m = f32(0)
    for x in xs:
        m = m + x
    m = m / n
    v = f32(0)
    for x in xs:
        d = x - m
        v = v + d * d
    v = v / n
    return m, f32(math.sqrt(float(v)))   # float32(math.Sqrt(float64(xvar)))


def linear_regression(xs, ys):
    """stats.go:569-586"""
    xm, xs_ = mean_stddev(xs)
    ym, ys_ = mean_stddev(ys)
    corr = f32(0)
    for x, y in zip(xs, ys):
        corr = corr + (x - xm) * (y - ym)
    corr = corr / (xs_ * ys_ * (f32(len(xs)) + f32(1)))
    slope = corr * ys_ / xs_
    icpt = ym - slope * xm
    return slope, icpt, xm, xs_, ym, ys_


def _gather(col, weights=None):
    g, w = [], []
    for i, v in enumerate(col):
        v = f32(v)
        if not np.isnan(v):
            g.append(v)
            if weights is not None:
                w.append(f32(weights[i]))
    return g, w


def _clip(g, w, lo, hi, cl, ch):
    j = 0
    while j < len(g):
        v = g[j]
        if v < lo:
            g[j] = g[-1]; g.pop()
            if w: w[j] = w[-1]; w.pop()
            cl += 1
        elif v > hi:
            g[j] = g[-1]; g.pop()
            if w: w[j] = w[-1]; w.pop()
            ch += 1
        else:
            j += 1
    return cl, ch


def _wmean(g, w):
    s, ws = f32(0), f32(0)
    for a, b in zip(g, w):
        s = s + a * b
        ws = ws + b
    return s / ws


def _winsor_sigma(g, median, sd):
    wz = list(g)
    while True:
        lo = median - f32(1.5) * sd
        hi = median + f32(1.5) * sd
        changed = 0
        for i, v in enumerate(wz):
            if v < lo:
                wz[i] = lo; changed += 1
            elif v > hi:
                wz[i] = hi; changed += 1
        old = sd
        _, sd = mean_stddev(wz)
        sd = f32(1.134) * sd
        factor = f32(abs(sd - old)) / old
        if changed == 0 or factor <= f32(0.0005):
            break
    return sd


def stack_column(col, mode, sig_lo=2.75, sig_hi=2.75, weights=None, ref_loc=0.0):
    """One pixel through Stack<Mode>[Weighted]; returns (value, clipLow, clipHigh).
    mode: 'median','mean','sigma','winsor','mad','linfit' (stack.go:274-918)."""
    sig_lo, sig_hi = f32(sig_lo), f32(sig_hi)
    cl = ch = 0
    if mode == "mean":
        s, ws, n = f32(0), f32(0), 0
        for i, v in enumerate(col):
            v = f32(v)
            if np.isnan(v):
                continue
            if weights is None:
                s = s + v
            else:
                s = s + v * f32(weights[i]); ws = ws + f32(weights[i])
            n += 1
        if n == 0:
            return f32(ref_loc), 0, 0
        return (s / f32(n) if weights is None else s / ws), 0, 0
    g, w = _gather(col, weights)
    if not g:
        return f32(ref_loc), 0, 0
    if mode == "median":
        return qselect_median(g), 0, 0
    if mode == "mad":
        med = qselect_median(g)
        ad = [abs(x - med) for x in g]
        mad = qselect_median(ad)
        sd = mad * f32(1.4826)
        cl, ch = _clip(g, None, med - sig_lo * sd, med + sig_hi * sd, cl, ch)
        s = f32(0)
        for x in g:
            s = s + x
        return s / f32(len(g)), cl, ch
    if mode in ("sigma", "winsor"):
        while True:
            med = qselect_median(g)
            mean, sd = mean_stddev(g)
            if mode == "winsor":
                sd = _winsor_sigma(g, med, sd)
            prev = cl + ch
            cl, ch = _clip(g, w if weights is not None else None, med - sig_lo * sd, med + sig_hi * sd, cl, ch)
            if cl + ch == prev or len(g) <= 1:
                return (mean if weights is None else _wmean(g, w)), cl, ch
    if mode == "linfit":
        mean = f32(0)
        while True:
            qsort(g)
            xs = [f32(i) for i in range(len(g))]
            slope, icpt, _, _, mean, _ = linear_regression(xs, g)
            sigma = f32(0)
            for i, y in enumerate(g):
                sigma = sigma + abs(y - (f32(i) * slope + icpt))
            sigma = sigma / f32(len(g))
            left = 0
            lob, hib = sig_lo * sigma, sig_hi * sigma
            for i in range(len(g)):
                y = g[i]
                lin = f32(i) * slope + icpt
                if lin - y > lob:
                    g[i] = g[left]; left += 1; cl += 1
                elif y - lin > hib:
                    g[i] = g[left]; left += 1; ch += 1
            if left == 0 or len(g) < 3:
                break
            g = g[left:]
        return mean, cl, ch
    raise ValueError(mode)


def transform_invert(t):
    """coord.go:159-201; t = (a,b,c,d,e,f) float32"""
    a, b, c, d, e, f = (f32(x) for x in t)
    eps = b * d - a * e
    if eps < f32(1e-8) and -eps < f32(1e-8):
        raise ZeroDivisionError("Matrix has no inverse")
    return (-e / (b * d - a * e), b / (b * d - a * e), (c * e - b * f) / (b * d - a * e),
            -d / (a * e - b * d), a / (a * e - b * d), (c * d - a * f) / (a * e - b * d))


def project(src, sw, sh, dw, dh, trans, oob):
    """project.go:26-76"""
    ia, ib, ic, id_, ie, if_ = transform_invert(trans)
    out = np.empty(dw * dh, dtype=np.float32)
    for row in range(dh):
        for col in range(dw):
            x, y = f32(col), f32(row)
            px = ia * x + ib * y + ic
            py = id_ * x + ie * y + if_
            xl, yl = math.floor(float(px)), math.floor(float(py))
            xh, yh = xl + 1, yl + 1
            xr, yr = px - f32(xl), py - f32(yl)
            if xl < 0 or xh >= sw or yl < 0 or yh >= sh:
                out[col + row * dw] = oob
                continue
            i00 = xl + yl * sw
            one = f32(1)
            vyl = f32(src[i00]) * (one - xr) + f32(src[i00 + 1]) * xr
            vyh = f32(src[i00 + sw]) * (one - xr) + f32(src[i00 + sw + 1]) * xr
            out[col + row * dw] = vyl * (one - yr) + vyh * yr
    return out


def lowbias32(x):
    x &= 0xFFFFFFFF
    x ^= x >> 16; x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15; x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def synth_sample(p, k, seed=12345):
    """SURVEY.md section 8d generator."""
    h = lowbias32(lowbias32((p + 0x9E3779B9 * k) & 0xFFFFFFFF) ^ seed)
    v = f32(1024.0) + f32((h & 0xFFFF) + (h >> 16) - 65535) * f32(1.0 / 256.0)
    h2 = lowbias32(h ^ 0xA5A5A5A5)
    if h2 % 61 == 0:
        v = v + f32(4096)
    elif h2 % 61 == 1:
        v = v - f32(512)
    elif h2 % 251 == 2:
        v = f32(np.nan)
    return v
