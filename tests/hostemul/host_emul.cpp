// host_emul.cpp -- compiles nightlight_b200/csrc/nl_column.cuh (the product's per-pixel device
// routines, __host__ __device__) for the CPU with element stride S = 1, so the "not gpu" tests can
// compare the exact code the kernels run against the oracle without a GPU.  Test infrastructure.
#include "../../nightlight_b200/csrc/nl_column.cuh"
#include <cmath>
#include <vector>
#include <cstddef>

using namespace nl;

extern "C" int emul_stack(int mode, const float *const *lights, int n, size_t len, const float *weights,
                          float ref_loc, float sig_lo, float sig_hi, float *res, long long *clip_lo, long long *clip_hi) {
    if (mode == ST_AUTO) mode = auto_select_mode(n);
    // clip_pass reads whole 32-slot blocks; the quick-select windows read QW-1 slots outside a column
    std::vector<float> g_(n + 40), wz_(n + 40), ramp(2 * (n + 2));
    std::vector<unsigned short> gw_(n + 40);
    float *g = g_.data() + 4, *wz = wz_.data() + 4;
    unsigned short *gw = gw_.data() + 4;
    for (int c = 1; c <= n; c++) ramp_mean_stddev(c, ramp[2 * c], ramp[2 * c + 1]);
    long long tl = 0, th = 0;
    bool W = weights != nullptr;
    for (size_t p = 0; p < len; p++) {
        int cur = 0;
        int ncl = 0, nch = 0;
        float out;
        if (mode == ST_MEAN) {
            float s = 0.0f, ws = 0.0f;
            for (int k = 0; k < n; k++) {
                float v = lights[k][p];
                if (v == v) {
                    if (W) { s = s + v * weights[k]; ws = ws + weights[k]; } else s = s + v;
                    cur++;
                }
            }
            res[p] = cur == 0 ? ref_loc : (W ? s / ws : s / (float)cur);
            continue;
        }
        for (int k = 0; k < n; k++) {
            float v = lights[k][p];
            g[cur] = v;
            if (W) gw[cur] = (unsigned short)k;
            cur += (v == v) ? 1 : 0;
        }
        if (cur == 0) { res[p] = ref_loc; continue; }
        switch (mode) {
        case ST_MEDIAN: {          // like the kernel: a -0.0 sample sends the column to the emulated quick-select
            bool negzero = false;
            for (int i = 0; i < cur; i++) negzero |= (g[i] == 0.0f && std::signbit(g[i]));
            out = negzero ? qselect_median<1, true>(g, cur) : median_by_value<1, true>(g, cur);
            break;
        }
        case ST_SIGMA:
            out = W ? reduce_sigma<1, true, unsigned short>(g, gw, weights, cur, sig_lo, sig_hi, ncl, nch)
                    : reduce_sigma<1, false, unsigned short>(g, nullptr, nullptr, cur, sig_lo, sig_hi, ncl, nch);
            break;
        case ST_WINSOR:
            out = W ? reduce_winsor<1, true, unsigned short>(g, gw, weights, cur, sig_lo, sig_hi, ncl, nch)
                    : reduce_winsor<1, false, unsigned short>(g, nullptr, nullptr, cur, sig_lo, sig_hi, ncl, nch);
            break;
        case ST_MAD: out = reduce_mad<1>(g, wz, cur, sig_lo, sig_hi, ncl, nch); break;
        case ST_LINFIT: out = reduce_linfit<1>(g, cur, cur, ramp.data(), sig_lo, sig_hi, ncl, nch); break;
        default: return -1;
        }
        res[p] = out;
        tl += ncl; th += nch;
    }
    *clip_lo = tl; *clip_hi = th;
    return 0;
}

// permutation check: run the flattened quick-select and hand back the permuted buffer
extern "C" float emul_qselect_median(float *a, int n) {
    std::vector<float> b(n + 8, 0.0f);
    for (int i = 0; i < n; i++) b[4 + i] = a[i];
    float m = qselect_median<1, true>(b.data() + 4, n);
    for (int i = 0; i < n; i++) a[i] = b[4 + i];
    return m;
}
extern "C" void emul_sort(float *a, int n, int nmax) { sort_column<1>(a, n, nmax > n ? nmax : n); }
