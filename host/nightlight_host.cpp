// nightlight_host.cpp -- see nightlight_host.hpp.  Host logic only: FITS in/out, operator plumbing, logs,
// error texts; every pixel of the hot path is computed by libnightlight_cuda.so.
// Build with -ffp-contract=off: the few fp32 expressions below (BSCALE/BZERO, exposure sum, noise) must
// not be contracted, like Go/amd64.
#include "nightlight_host.hpp"

#include <cerrno>
#include <cmath>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace nightlight {

static void check(int rc) {
    if (rc != NL_OK) throw Error(nl_last_error());
}

// ------------------------------------------------------------------------------------------------
// Context
// ------------------------------------------------------------------------------------------------
Context::Context(const std::vector<int> &devices, FILE *log) : Log(log) {
    for (int d : devices) {
        nl_ctx *c = nullptr;
        int rc = nl_ctx_create(d, &c);
        if (rc != NL_OK) {
            std::string msg = nl_last_error();
            for (nl_ctx *o : ctxs_) nl_ctx_destroy(o);
            throw Error(msg);
        }
        ctxs_.push_back(c);
    }
    if (ctxs_.empty()) throw Error("no CUDA device selected; this library has no CPU fallback");
}

Context::~Context() {
    for (nl_ctx *c : ctxs_) nl_ctx_destroy(c);
}

// ------------------------------------------------------------------------------------------------
// FITS in (read.go) / out (write.go)
// ------------------------------------------------------------------------------------------------
static const size_t kBlock = 2880, kCard = 80;

static std::string trim(const std::string &s) {
    size_t a = s.find_first_not_of(' '), b = s.find_last_not_of(' ');
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

Image NewImageFromNaxisn(const std::vector<int32_t> &naxisn, std::vector<float> data) {
    Image img;
    int64_t px = 1;
    for (int32_t n : naxisn) px *= n;
    img.Naxisn = naxisn;
    img.Pixels = (int32_t)px;
    if (data.empty()) data.assign((size_t)px, 0.0f);
    img.Data = std::move(data);
    return img;
}

void Image::ReadFile(const std::string &fileName, bool readData, FILE *log) {
    FILE *f = fopen(fileName.c_str(), "rb");
    if (!f) throw Error("open " + fileName + ": " + strerror(errno));
    struct Closer { FILE *f; ~Closer() { fclose(f); } } closer{f};
    FileName = fileName;
    std::map<std::string, bool> bools;
    std::map<std::string, int64_t> ints;
    std::map<std::string, double> floats;
    HeaderStrings.clear();
    std::vector<char> block(kBlock);
    bool end = false;
    while (!end) {                                           // read.go:445-466: header units of 2880 bytes
        if (fread(block.data(), 1, kBlock, f) != kBlock) throw Error(std::to_string(ID) + ": unexpected EOF");
        for (size_t ln = 0; ln < kBlock / kCard && !end; ln++) {
            std::string line(block.data() + ln * kCard, kCard);
            std::string key = trim(line.substr(0, 8));
            if (key == "END") { end = true; break; }
            if (key.empty() || key == "HISTORY" || key == "COMMENT") continue;
            size_t eq = line.find('=');
            if (eq == std::string::npos || eq > 9) {
                if (log) fprintf(log, "%d: Warning:Cannot parse '%s', ignoring\n", ID, line.c_str());
                continue;
            }
            std::string rest = trim(line.substr(eq + 1));
            if (!rest.empty() && rest[0] == '\'') {          // string value
                size_t q = rest.find('\'', 1);
                HeaderStrings[key] = rest.substr(1, q == std::string::npos ? std::string::npos : q - 1);
                continue;
            }
            size_t slash = rest.find('/');
            std::string val = trim(slash == std::string::npos ? rest : rest.substr(0, slash));
            if (val == "T" || val == "F") { bools[key] = val == "T"; continue; }
            char *endp = nullptr;
            long long iv = strtoll(val.c_str(), &endp, 10);
            if (endp && *endp == 0 && !val.empty()) { ints[key] = iv; continue; }
            std::string fv = val;
            for (char &ch : fv) if (ch == 'D') ch = 'E';     // FORTRAN double exponent
            double dv = strtod(fv.c_str(), &endp);
            if (endp && *endp == 0 && !val.empty()) { floats[key] = dv; continue; }
            if (log) fprintf(log, "%d: Warning:Cannot parse '%s', ignoring\n", ID, line.c_str());
        }
    }
    if (!bools.count("SIMPLE") || !bools["SIMPLE"])          // read.go:102-105
        throw Error(std::to_string(ID) + ": Not a valid FITS file; SIMPLE=T missing in header");
    auto popInt = [&](const std::string &k) -> int32_t {
        auto it = ints.find(k);
        if (it == ints.end()) throw Error(std::to_string(ID) + ": FITS header does not contain key " + k);
        return (int32_t)it->second;
    };
    auto intOrFloat = [&](const std::string &k, float &out) -> bool {   // PopHeaderInt32OrFloat
        auto i = ints.find(k);
        if (i != ints.end()) { out = (float)(int32_t)i->second; return true; }
        auto fl = floats.find(k);
        if (fl != floats.end()) { out = (float)fl->second; return true; }
        return false;
    };
    Bitpix = popInt("BITPIX");
    int32_t naxis = popInt("NAXIS");
    Naxisn.assign((size_t)naxis, 0);
    Pixels = 1;
    for (int32_t i = 1; i <= naxis; i++) {
        Naxisn[(size_t)i - 1] = popInt("NAXIS" + std::to_string(i));
        Pixels *= Naxisn[(size_t)i - 1];
    }
    if (!intOrFloat("BZERO", Bzero)) Bzero = 0;
    if (!intOrFloat("BSCALE", Bscale)) Bscale = 1;
    if (!intOrFloat("EXPOSURE", Exposure) && !intOrFloat("EXPTIME", Exposure)) Exposure = 0;
    if (!readData) return;

    const int bytesPer = Bitpix < 0 ? -Bitpix / 8 : Bitpix / 8;
    if (Bitpix != 8 && Bitpix != 16 && Bitpix != 32 && Bitpix != 64 && Bitpix != -32 && Bitpix != -64)
        throw Error(std::to_string(ID) + ": Unknown BITPIX value " + std::to_string(Bitpix));
    if (log && (Bitpix == 32 || Bitpix == 64))
        fprintf(log, "%d: Warning: loss of precision converting int%d to float32 values\n", ID, Bitpix);
    if (log && Bitpix == -64)
        fprintf(log, "%d: Warning: loss of precision converting float%d to float32 values\n", ID, -Bitpix);
    Data.assign((size_t)Pixels, 0.0f);
    std::vector<unsigned char> raw((size_t)Pixels * bytesPer);
    if (fread(raw.data(), 1, raw.size(), f) != raw.size()) throw Error(std::to_string(ID) + ": unexpected EOF");
    float mn = 3.40282346638528859811704183484516925440e+38f, mx = -mn;
    double sum = 0;
    const volatile float bscale = Bscale, bzero = Bzero;     // v = float32(val)*Bscale + Bzero, mul then add
    for (size_t i = 0; i < (size_t)Pixels; i++) {
        const unsigned char *p = raw.data() + i * bytesPer;
        float val;
        switch (Bitpix) {
        case 8: val = (float)p[0]; break;
        case 16: val = (float)(int16_t)((uint16_t)(p[0] << 8) | p[1]); break;
        case 32: val = (float)(int32_t)(((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]); break;
        case 64: {
            uint64_t u = 0;
            for (int b = 0; b < 8; b++) u = (u << 8) | p[b];
            val = (float)(int64_t)u;
            break;
        }
        case -32: {
            uint32_t u = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
            memcpy(&val, &u, 4);
            break;
        }
        default: {
            uint64_t u = 0;
            for (int b = 0; b < 8; b++) u = (u << 8) | p[b];
            double d;
            memcpy(&d, &u, 8);
            val = (float)d;
        }
        }
        volatile float prod = val * bscale;
        float v = prod + bzero;
        if (v < mn) mn = v;
        if (v > mx) mx = v;
        sum += (double)v;
        Data[i] = v;
    }
    Bzero = 0; Bscale = 1;                                   // data values incorporate these now
    Min = mn; Max = mx; Mean = (float)(sum / (double)Pixels);
    HasStats = true;
}

static void writeCard(std::string &sb, std::string key, const std::string &value20, std::string comment) {
    if (key.size() > 8) key.resize(8);
    if (comment.size() > 47) comment.resize(47);
    char buf[128];
    snprintf(buf, sizeof buf, "%-8s= %20s / %-47s", key.c_str(), value20.c_str(), comment.c_str());
    sb += buf;
}

// Go's %g of a float32 (write.go:135-139 via fmt): the SHORTEST decimal that parses back to the same float32, in %e
// form when the decimal exponent is < -4 or >= 6 (strconv.FormatFloat(v, 'g', -1, 32): eprec = 6 for shortest)
static std::string gfloat(float v) {
    char b[64];
    if (v != v) return "NaN";
    if (v == INFINITY) return "+Inf";
    if (v == -INFINITY) return "-Inf";
    int prec = 0;
    for (; prec < 9; prec++) {                               // digits after the first: 0 .. 8 (9 significant digits always suffice)
        snprintf(b, sizeof b, "%.*e", prec, (double)v);
        if (strtof(b, nullptr) == v) break;
    }
    std::string e = b;                                       // d.ddddde[+-]XX
    const size_t ep = e.find('e');
    const int exp10 = atoi(e.c_str() + ep + 1);
    if (exp10 < -4 || exp10 >= 6) return e;                  // Go prints at least two exponent digits, like C
    snprintf(b, sizeof b, "%.*f", prec - exp10 > 0 ? prec - exp10 : 0, (double)v);
    return b;
}

void Image::WriteFile(const std::string &fileName) const {
    std::string sb;                                           // write.go:54-80
    writeCard(sb, "SIMPLE", "T", "    FITS standard 4.0");
    writeCard(sb, "BITPIX", "-32", "    32-bit floating point");
    writeCard(sb, "NAXIS", std::to_string(Naxisn.size()), "[1] Number of array dimensions");
    for (size_t i = 0; i < Naxisn.size(); i++)
        writeCard(sb, "NAXIS" + std::to_string(i + 1), std::to_string(Naxisn[i]), "[1] Array dimension");
    writeCard(sb, "BZERO", gfloat(Bzero), "[1] Zero offset");
    writeCard(sb, "BSCALE", gfloat(Bscale), "[1] Data scale");
    if (Exposure != 0) writeCard(sb, "EXPOSURE", gfloat(Exposure), "[s] Exposure duration");
    {
        char buf[128];
        snprintf(buf, sizeof buf, "%-8s= '%s'%s / %-47s", "PROGRAM", "nightlight", "        ", "    https://github.com/mlnoga/nightlight");
        sb += buf;
    }
    sb += "END" + std::string(kCard - 3, ' ');
    if (sb.size() % kBlock) sb.append(kBlock - sb.size() % kBlock, ' ');

    FILE *f = fopen(fileName.c_str(), "wb");
    if (!f) throw Error("open " + fileName + ": " + strerror(errno));
    struct Closer { FILE *f; ~Closer() { fclose(f); } } closer{f};
    if (fwrite(sb.data(), 1, sb.size(), f) != sb.size()) throw Error("write " + fileName + " failed");
    std::vector<unsigned char> raw(Data.size() * 4);          // write.go:182-215: network byte order, NaN -> 0
    for (size_t i = 0; i < Data.size(); i++) {
        float d = Data[i];
        if (d != d) d = 0;
        uint32_t u;
        memcpy(&u, &d, 4);
        raw[4 * i + 0] = (unsigned char)(u >> 24);
        raw[4 * i + 1] = (unsigned char)(u >> 16);
        raw[4 * i + 2] = (unsigned char)(u >> 8);
        raw[4 * i + 3] = (unsigned char)u;
    }
    if (raw.size() % kBlock) raw.insert(raw.end(), kBlock - raw.size() % kBlock, (unsigned char)' ');
    if (fwrite(raw.data(), 1, raw.size(), f) != raw.size()) throw Error("write " + fileName + " failed");
}

// ------------------------------------------------------------------------------------------------
// Resample and star detection
// ------------------------------------------------------------------------------------------------
Image Image::Project(Context &c, const std::vector<int32_t> &destNaxisn, const Transform2D &trans, float outOfBounds) const {
    if (Naxisn.size() < 2 || destNaxisn.size() < 2) throw Error("Project needs two-dimensional images");
    Image res = NewImageFromNaxisn(destNaxisn, {});
    const float t[6] = {trans.A, trans.B, trans.C, trans.D, trans.E, trans.F};
    check(nl_project(c.Device(0), Data.data(), Naxisn[0], Naxisn[1], res.Data.data(), destNaxisn[0], destNaxisn[1], t, outOfBounds));
    res.ID = ID;                                              // project.go:36-44
    res.FileName = FileName;
    res.Exposure = Exposure;
    res.Trans = trans;
    return res;
}

std::vector<Star> FindStars(Context &c, const std::vector<float> &data, int32_t width, float location, float scale,
                            float starSig, float bpSigma, float starInOut, int32_t radius, float medianDiffStdDev,
                            float *sumOfShifts, float *avgHFR) {
    std::vector<Star> stars(data.size() / 100 + 1024);
    int32_t n = 0;
    float sos = 0, hfr = 0;
    check(nl_find_stars(c.Device(0), data.data(), (int32_t)data.size(), width, location, scale, starSig, bpSigma, starInOut,
                        radius, medianDiffStdDev, stars.data(), (int32_t)stars.size(), &n, &sos, &hfr));
    stars.resize((size_t)std::min<int32_t>(n, (int32_t)stars.size()));
    if (sumOfShifts) *sumOfShifts = sos;
    if (avgHFR) *avgHFR = hfr;
    return stars;
}

float EstimateNoise(Context &c, const std::vector<float> &data, int32_t width) {
    float noise = 0;
    check(nl_estimate_noise(c.Device(0), data.data(), (int32_t)data.size(), width, &noise));
    return noise;
}

BasicStats NewStats(Context &c, const std::vector<float> &data) {
    float st[4] = {0, 0, 0, 0};
    check(nl_stats(c.Device(0), data.data(), (int64_t)data.size(), st));
    return BasicStats{st[0], st[1], st[2], st[3]};
}

std::vector<float> MedianFilter3x3(Context &c, const std::vector<float> &data, int32_t width) {
    std::vector<float> out(data.size());
    check(nl_median_filter3x3(c.Device(0), data.data(), (int32_t)data.size(), width, out.data()));
    return out;
}

std::vector<int32_t> BadPixelMap(Context &c, const std::vector<float> &data, int32_t width, float sigmaLow, float sigmaHigh,
                                 BasicStats *medianDiffStats) {
    std::vector<int32_t> bpm(data.size() / 100 + 1024);       // badpixels.go:42
    int64_t count = 0;
    float st[4] = {0, 0, 0, 0};
    check(nl_bad_pixel_map(c.Device(0), data.data(), (int64_t)data.size(), width, sigmaLow, sigmaHigh, bpm.data(), (int64_t)bpm.size(),
                           &count, st));
    if (count > (int64_t)bpm.size()) {
        bpm.resize((size_t)count);
        check(nl_bad_pixel_map(c.Device(0), data.data(), (int64_t)data.size(), width, sigmaLow, sigmaHigh, bpm.data(),
                               (int64_t)bpm.size(), &count, st));
    }
    bpm.resize((size_t)count);
    if (medianDiffStats) *medianDiffStats = BasicStats{st[0], st[1], st[2], st[3]};
    return bpm;
}

void OpBadPixel::Apply(Image &f, Context &c, BasicStats *medianDiffStats) {
    if (SigmaLow == 0 || SigmaHigh == 0) return;              // preprocess.go:181-183
    int64_t removed = 0;
    float st[4] = {0, 0, 0, 0};
    check(nl_op_bad_pixel(c.Device(0), f.Data.data(), (int64_t)f.Data.size(), f.Naxisn[0], SigmaLow, SigmaHigh, &removed, st));
    if (medianDiffStats) *medianDiffStats = BasicStats{st[0], st[1], st[2], st[3]};
    fprintf(c.Log, "%d: Removed %d bad pixels (%.2f%%) with sigma low=%.2f high=%.2f\n", f.ID, (int)removed,
            (double)(100.0f * (float)removed / (float)f.Pixels), (double)SigmaLow, (double)SigmaHigh);     // :190-191
}

// ------------------------------------------------------------------------------------------------
// Stacking
// ------------------------------------------------------------------------------------------------
std::vector<float> getWeights(const std::vector<const Image *> &f, StackWeighting weighting) {
    if (weighting == StWeightNone) return {};
    if (weighting < StWeightNone || weighting > StWeightInverseHFR)
        throw Error("Invalid weighting mode " + std::to_string((int)weighting) + "\n");
    std::vector<float> exposure, noise, hfr;
    for (const Image *img : f) {
        if (weighting == StWeightExposure && img->Exposure == 0)
            throw Error(std::to_string(img->ID) + ": Missing exposure information for exposure-weighted stacking");
        if (weighting == StWeightInverseNoise && !img->HasStats)
            throw Error(std::to_string(img->ID) + ": Missing stats information for noise-weighted stacking");
        exposure.push_back(img->Exposure);
        noise.push_back(img->Noise);
        hfr.push_back(img->HFR);
    }
    std::vector<float> w(f.size());
    check(nl_get_weights((int32_t)weighting, exposure.data(), noise.data(), hfr.data(), (int32_t)f.size(), w.data()));
    return w;
}

Image OpStack::Apply(const std::vector<const Image *> &f, Context &c) {
    int mode = Mode;
    if (mode < StMedian || mode > StAuto) throw Error("invalid stacking mode");        // stack.go:118-120
    if (f.empty()) throw Error("stack operator needs inputs");                        // stack.go:101
    if (mode == StAuto) mode = nl_auto_select_mode((int32_t)f.size());
    if (c.Log)
        fprintf(c.Log, "Stacking %d frames with stacking mode %d and sigma low %g high %g:\n", (int)f.size(), mode,
                (double)SigmaLow, (double)SigmaHigh);
    std::vector<float> weights = getWeights(f, Weighting);
    const size_t pixels = f[0]->Data.size();
    for (const Image *img : f)
        if (img->Data.size() != pixels) throw Error(std::to_string(img->ID) + ": frame size differs from the first frame");
    std::vector<float> data(pixels);

    // Row stripes: device g stacks the pixels [lo_g, hi_g) of all frames; stripes are cut on row
    // boundaries (the reference cuts 8 MiB pixel ranges for its goroutines)
    const size_t ndev = c.NumDevices();
    const size_t width = f[0]->Naxisn.empty() ? pixels : (size_t)f[0]->Naxisn[0];
    const size_t rows = width ? pixels / width : 0;
    std::vector<int64_t> cl(ndev, 0), ch(ndev, 0);
    std::vector<std::string> errs(ndev);
    auto work = [&](size_t g) {
        try {
            size_t lo = rows * g / ndev * width, hi = g + 1 == ndev ? pixels : rows * (g + 1) / ndev * width;
            if (hi <= lo) return;
            // all frame pointers of this device's stripe in one call; the library pipelines sub-stripes
            std::vector<const float *> ptrs(f.size());
            for (size_t i = 0; i < f.size(); i++) ptrs[i] = f[i]->Data.data() + lo;
            check(nl_stack_apply(c.Device(g), ptrs.data(), (int32_t)f.size(), (int64_t)(hi - lo), (int64_t)width, 8, mode,
                                 weights.empty() ? nullptr : weights.data(), SigmaLow, SigmaHigh, RefFrameLoc, data.data() + lo,
                                 &cl[g], &ch[g]));
        } catch (const std::exception &e) {
            errs[g] = e.what();
        }
    };
    if (ndev == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (size_t g = 0; g < ndev; g++) th.emplace_back(work, g);
        for (auto &t : th) t.join();
    }
    for (const std::string &e : errs)
        if (!e.empty()) throw Error(e);
    NumClippedLow = NumClippedHigh = 0;
    for (size_t g = 0; g < ndev; g++) { NumClippedLow += cl[g]; NumClippedHigh += ch[g]; }
    if (mode >= StSigma && c.Log)                                                      // stack.go:213-218
        fprintf(c.Log, "Clipped low %lld (%.2f%%) high %lld (%.2f%%)\n", (long long)NumClippedLow,
                (double)((float)NumClippedLow * 100.0f / (float)(pixels * f.size())), (long long)NumClippedHigh,
                (double)((float)NumClippedHigh * 100.0f / (float)(pixels * f.size())));
    float exposureSum = 0;                                                             // stack.go:220-221
    for (const Image *img : f) exposureSum += img->Exposure;
    Image stack = NewImageFromNaxisn(f[0]->Naxisn, std::move(data));                   // stack.go:224-226
    stack.Exposure = exposureSum;
    return stack;
}

Image OpStackBatches::Apply(const std::vector<std::vector<const Image *>> &batches, Context &c) {
    if (batches.empty()) throw Error("stackBatches operator needs inputs");
    Image stack;
    int64_t stackFrames = 0;
    const size_t nb = batches.size();
    for (size_t b = 0; b < nb; b++) {
        if (c.Log) fprintf(c.Log, "\nStarting batch %d of %d with %d frames...\n", (int)b + 1, (int)nb, (int)batches[b].size());
        Image batch = PerBatch.Apply(batches[b], c);
        if (nb == 1) return batch;
        // StackIncremental (stack.go:924-937) / Finalize (:940-944) on the device
        const int64_t px = (int64_t)batch.Data.size();
        nl_ctx *ctx = c.Device(0);
        void *acc = nullptr, *light = nullptr;
        check(nl_dev_alloc(ctx, 4 * px, &acc));
        struct Free { nl_ctx *c; void *p; ~Free() { nl_dev_free(c, p); } } fa{ctx, acc};
        check(nl_dev_alloc(ctx, 4 * px, &light));
        Free fl{ctx, light};
        if (b > 0) check(nl_memcpy_h2d(ctx, acc, stack.Data.data(), 4 * px));
        check(nl_memcpy_h2d(ctx, light, batch.Data.data(), 4 * px));
        const float weight = (float)batches[b].size();
        check(nl_stack_incremental_dev(ctx, (float *)acc, (const float *)light, px, weight, b == 0));
        stackFrames += (int64_t)batches[b].size();
        if (b + 1 == nb) check(nl_stack_incremental_finalize_dev(ctx, (float *)acc, px, (float)stackFrames));
        if (b == 0) {
            stack = NewImageFromNaxisn(batch.Naxisn, {});
            stack.Exposure = batch.Exposure;
        } else {
            stack.Exposure += batch.Exposure;
        }
        check(nl_memcpy_d2h(ctx, stack.Data.data(), acc, 4 * px));
        check(nl_ctx_sync(ctx));
    }
    return stack;
}

}  // namespace nightlight
