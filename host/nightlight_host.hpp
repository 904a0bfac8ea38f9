// nightlight_host.hpp -- C++ host layer above the C ABI (include/nightlight_cuda.h).
//
// The reference's host side is Go (compiled code); this image has no Go toolchain, so the operators
// of the hot path are mirrored here in C++ with the reference's names, argument meaning, log lines and
// error texts, and they call the CUDA library exactly where the cgo stubs of INTEGRATION.md would:
//   fits::Image, ReadFile / WriteFile    internal/fits/fits.go:30-115, read.go:34-171, write.go:31-215
//   OpStack                              internal/ops/stack/stack.go:66-227
//   OpStackBatches (stack of stacks)     internal/ops/stack/stackbatches.go:56-119
//   Image::Project                       internal/fits/project.go:26-76
//   FindStars                            internal/star/findstars.go:59-100
//   EstimateNoise                        internal/stats/noise.go:24-55 (portable definition)
// No pixel of the hot path is computed on the host: there is no CPU fallback.
#pragma once

#include <cstdint>
#include <cstdio>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/nightlight_cuda.h"

namespace nightlight {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// star.Transform2D, internal/star/coord.go:52-59
struct Transform2D {
    float A = 1, B = 0, C = 0, D = 0, E = 1, F = 0;
};

using Star = nl_star;   // star.Star, internal/star/findstars.go:30-37

// fits.Image, internal/fits/fits.go:30-54: the fields this path reads or writes
struct Image {
    int ID = 0;
    std::string FileName;
    int32_t Bitpix = -32;
    float Bzero = 0, Bscale = 1;
    std::vector<int32_t> Naxisn;
    int32_t Pixels = 0;
    std::vector<float> Data;
    float Exposure = 0;
    bool HasStats = false;       // Stats != nil
    float Min = 0, Max = 0, Mean = 0, Noise = 0;
    std::vector<Star> Stars;
    float HFR = 0;
    Transform2D Trans;
    float Residual = 0;
    std::map<std::string, std::string> HeaderStrings;   // remaining header cards, kept verbatim

    // read.go:47-171 (FITS only; gzip and TIFF input stay with the Go front end)
    void ReadFile(const std::string &fileName, bool readData = true, FILE *log = stderr);
    // write.go:31-89: always BITPIX -32, big-endian, NaN -> 0, padded to 2880-byte blocks
    void WriteFile(const std::string &fileName) const;
    // project.go:26-76
    Image Project(class Context &c, const std::vector<int32_t> &destNaxisn, const Transform2D &trans, float outOfBounds) const;
};

Image NewImageFromNaxisn(const std::vector<int32_t> &naxisn, std::vector<float> data);   // fits.go:83-100

// ops.Context (internal/ops/operator.go:37-56), reduced to what the path needs: the log and the devices
class Context {
public:
    explicit Context(const std::vector<int> &devices = {0}, FILE *log = stdout);
    ~Context();
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    FILE *Log;
    size_t NumDevices() const { return ctxs_.size(); }
    nl_ctx *Device(size_t i) const { return ctxs_[i]; }

private:
    std::vector<nl_ctx *> ctxs_;
};

enum StackMode { StMedian = 0, StMean, StSigma, StWinsorSigma, StMADSigma, StLinearFit, StAuto };          // stack.go:33-42
enum StackWeighting { StWeightNone = 0, StWeightExposure, StWeightInverseNoise, StWeightInverseHFR };        // stack.go:57-63

// stack.OpStack, stack.go:66-73
struct OpStack {
    StackMode Mode = StAuto;
    StackWeighting Weighting = StWeightNone;
    float SigmaLow = 2.75f, SigmaHigh = 2.75f;
    float RefFrameLoc = 0;
    int64_t NumClippedLow = 0, NumClippedHigh = 0;   // outputs of the last Apply (printed like stack.go:214-218)

    // stack.go:115-227.  With several devices in the context every device stacks one row stripe.
    Image Apply(const std::vector<const Image *> &f, Context &c);
};

std::vector<float> getWeights(const std::vector<const Image *> &f, StackWeighting weighting);   // stack.go:231-270

// The stack-of-stacks arithmetic of OpStackBatches.Apply (stackbatches.go:84-116) for given batches
struct OpStackBatches {
    OpStack PerBatch;
    Image Apply(const std::vector<std::vector<const Image *>> &batches, Context &c);
};

// star.FindStars, findstars.go:59-100.  medianDiffStdDev stands for medianDiffStats.StdDev()
std::vector<Star> FindStars(Context &c, const std::vector<float> &data, int32_t width, float location, float scale,
                            float starSig, float bpSigma, float starInOut, int32_t radius, float medianDiffStdDev,
                            float *sumOfShifts, float *avgHFR);

// stats.EstimateNoise (noise_amd64.go:25-43; noise.go:32-55 in pure-Go numerics), computed on the device
float EstimateNoise(Context &c, const std::vector<float> &data, int32_t width);

// The eagerly evaluated part of stats.Stats (stats.go:43-60, 102-153)
struct BasicStats {
    float Min = 0, Mean = 0, Max = 0, StdDev = 0;
};
BasicStats NewStats(Context &c, const std::vector<float> &data);

// median.MedianFilter3x3, median3x3_amd64.go:24-48
std::vector<float> MedianFilter3x3(Context &c, const std::vector<float> &data, int32_t width);

// pre.BadPixelMap, badpixels.go:32-51
std::vector<int32_t> BadPixelMap(Context &c, const std::vector<float> &data, int32_t width, float sigmaLow, float sigmaHigh,
                                 BasicStats *medianDiffStats);

// pre.OpBadPixel for monochrome frames, preprocess.go:160-201: repairs f.Data in place, sets *medianDiffStats
// (f.MedianDiffStats) and logs the reference's line.  A zero sigma returns without touching the frame.
struct OpBadPixel {
    float SigmaLow = 3, SigmaHigh = 5;
    void Apply(Image &f, Context &c, BasicStats *medianDiffStats);
};

}  // namespace nightlight
