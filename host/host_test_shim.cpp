// C entry points over the host layer's FITS code, for the ctypes tests (test infrastructure).
#include <cstring>
#include "nightlight_host.hpp"
using namespace nightlight;
static thread_local std::string g_err;
extern "C" const char *nlh_last_error() { return g_err.c_str(); }
extern "C" int nlh_fits_info(const char *path, int32_t *bitpix, int32_t *naxisn, int32_t *naxis, float *exposure) {
    try {
        Image im;
        im.ReadFile(path, false, nullptr);
        *bitpix = im.Bitpix; *naxis = (int32_t)im.Naxisn.size(); *exposure = im.Exposure;
        for (size_t i = 0; i < im.Naxisn.size() && i < 8; i++) naxisn[i] = im.Naxisn[i];
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}
extern "C" int nlh_fits_read(const char *path, float *data, int64_t cap, float *mmm) {
    try {
        Image im;
        im.ReadFile(path, true, nullptr);
        if ((int64_t)im.Data.size() > cap) { g_err = "buffer too small"; return -2; }
        memcpy(data, im.Data.data(), im.Data.size() * 4);
        mmm[0] = im.Min; mmm[1] = im.Mean; mmm[2] = im.Max;
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}
extern "C" int nlh_fits_write(const char *path, const float *data, const int32_t *naxisn, int32_t naxis, float exposure) {
    try {
        Image im = NewImageFromNaxisn(std::vector<int32_t>(naxisn, naxisn + naxis), std::vector<float>());
        memcpy(im.Data.data(), data, im.Data.size() * 4);
        im.Exposure = exposure;
        im.WriteFile(path);
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}
