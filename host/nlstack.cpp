// nlstack -- command-line front end of the C++ host layer, mirroring the I/O of `nightlight stack`
// (cmd/nightlight/main.go:49-166, 309-336) for the hot path only: FITS frames in -> stacked FITS out plus the
// reference's log lines.  Calibration, reference selection and triangle alignment stay with the Go CLI.
//
//   nlstack stack [-stMode m] [-stWeight w] [-stSigLow x] [-stSigHigh y] [-stBatch n] [-gpus list] [-out file] in.fits...
//   nlstack stars [-bpSigLow l -bpSigHigh h] [-starSig s] [-starBpSig b] [-starInOut r] [-starRadius r] [-loc l] [-scale s] in.fits...
//   nlstack project -trans A,B,C,D,E,F [-oob nan|0] -out out.fits in.fits
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <sstream>

#include "nightlight_host.hpp"

using namespace nightlight;

static std::vector<int> parseDevices(const std::string &s) {
    std::vector<int> d;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) d.push_back(atoi(tok.c_str()));
    return d;
}

static int usage() {
    fprintf(stderr, "usage: nlstack (stack|stars|project) [flags] files...   (see host/nlstack.cpp)\n");
    return 2;
}

int main(int argc, char **argv) {
    if (argc < 2) return usage();
    const std::string cmd = argv[1];
    std::map<std::string, std::string> flag = {
        {"stMode", "6"}, {"stWeight", "0"}, {"stSigLow", "2.75"}, {"stSigHigh", "2.75"}, {"stBatch", "0"}, {"gpus", "0"},
        {"out", "out.fits"}, {"starSig", "15"}, {"starBpSig", "0"}, {"bpSigLow", "0"}, {"bpSigHigh", "0"}, {"starInOut", "1.4"}, {"starRadius", "16"},
        {"loc", "nan"}, {"scale", "nan"}, {"trans", "1,0,0,0,1,0"}, {"oob", "nan"}};
    std::vector<std::string> files;
    for (int i = 2; i < argc; i++) {
        if (argv[i][0] == '-' && flag.count(argv[i] + 1) && i + 1 < argc) { flag[argv[i] + 1] = argv[i + 1]; i++; }
        else if (argv[i][0] == '-' && !isdigit((unsigned char)argv[i][1])) { fprintf(stderr, "unknown flag %s\n", argv[i]); return usage(); }
        else files.push_back(argv[i]);
    }
    try {
        const auto t0 = std::chrono::steady_clock::now();
        Context c(parseDevices(flag["gpus"]), stdout);
        std::vector<std::unique_ptr<Image>> imgs;
        for (size_t i = 0; i < files.size(); i++) {
            imgs.emplace_back(new Image());
            imgs.back()->ID = (int)i;
            imgs.back()->ReadFile(files[i], true, stdout);
            Image &im = *imgs.back();
            fprintf(stdout, "%d: Loaded %s %dx%d exposure %g min %g mean %g max %g\n", im.ID, files[i].c_str(),
                    im.Naxisn.size() > 0 ? im.Naxisn[0] : 0, im.Naxisn.size() > 1 ? im.Naxisn[1] : 1, (double)im.Exposure,
                    (double)im.Min, (double)im.Mean, (double)im.Max);
        }
        if (imgs.empty()) throw Error(cmd + " operator needs inputs");
        if (cmd == "stack") {
            OpStack op;
            op.Mode = (StackMode)atoi(flag["stMode"].c_str());
            op.Weighting = (StackWeighting)atoi(flag["stWeight"].c_str());
            op.SigmaLow = (float)atof(flag["stSigLow"].c_str());
            op.SigmaHigh = (float)atof(flag["stSigHigh"].c_str());
            if (op.Weighting == StWeightInverseNoise)
                for (auto &im : imgs) im->Noise = EstimateNoise(c, im->Data, im->Naxisn[0]);
            std::vector<const Image *> f;
            for (auto &im : imgs) f.push_back(im.get());
            Image res;
            const size_t batch = (size_t)atoi(flag["stBatch"].c_str());
            if (batch > 0 && batch < f.size()) {                // stack of stacks with consecutive batches of `batch` frames
                OpStackBatches ob;
                ob.PerBatch = op;
                std::vector<std::vector<const Image *>> batches;
                for (size_t i = 0; i < f.size(); i += batch) batches.emplace_back(f.begin() + i, f.begin() + std::min(f.size(), i + batch));
                res = ob.Apply(batches, c);
            } else {
                res = op.Apply(f, c);
            }
            res.WriteFile(flag["out"]);
            fprintf(stdout, "Wrote %s\n", flag["out"].c_str());
        } else if (cmd == "stars") {
            for (auto &im : imgs) {
                float loc = (float)atof(flag["loc"].c_str()), scale = (float)atof(flag["scale"].c_str());
                if (std::isnan(loc)) loc = im->Mean;            // the reference's estimators are randomised: inputs here
                if (std::isnan(scale)) scale = EstimateNoise(c, im->Data, im->Naxisn[0]);
                // preprocess.go:88-96: bad-pixel repair runs before star detection and leaves the frame's MedianDiffStats
                OpBadPixel bp;
                bp.SigmaLow = (float)atof(flag["bpSigLow"].c_str());
                bp.SigmaHigh = (float)atof(flag["bpSigHigh"].c_str());
                BasicStats mds;
                bp.Apply(*im, c, &mds);
                float sos = 0, hfr = 0;
                im->Stars = FindStars(c, im->Data, im->Naxisn[0], loc, scale, (float)atof(flag["starSig"].c_str()),
                                      (float)atof(flag["starBpSig"].c_str()), (float)atof(flag["starInOut"].c_str()),
                                      atoi(flag["starRadius"].c_str()), mds.StdDev, &sos, &hfr);
                im->HFR = hfr;
                fprintf(stdout, "%d: Stars %d HFR %.2f\n", im->ID, (int)im->Stars.size(), (double)hfr);   // preprocess.go:455
            }
        } else if (cmd == "project") {
            Transform2D t;
            if (sscanf(flag["trans"].c_str(), "%f,%f,%f,%f,%f,%f", &t.A, &t.B, &t.C, &t.D, &t.E, &t.F) != 6) throw Error("bad -trans");
            const float oob = flag["oob"] == "nan" ? NAN : (float)atof(flag["oob"].c_str());
            Image res = imgs[0]->Project(c, imgs[0]->Naxisn, t, oob);
            res.WriteFile(flag["out"]);
        } else {
            return usage();
        }
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stdout, "\nDone after %.3fs\n", sec);            // main.go:427-429
    } catch (const std::exception &e) {
        fprintf(stdout, "Error: %s\n", e.what());
        return 1;
    }
    return 0;
}
