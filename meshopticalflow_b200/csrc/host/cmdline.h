// Command line of the OpticalFlow host: the reference's flags, names and defaults
// (OpticalFlow/OpticalFlow.cpp:56-63, usage text :65-109; parser include/Misha/CmdLineParser.inl:239-262:
// `--name value`, names case-insensitive, unknown names warn and list the valid ones).
#ifndef MOF_CMDLINE_H
#define MOF_CMDLINE_H

#include <string>

namespace mof {

struct Options {
    // I/O
    bool inSet = false, meshSet = false, outSet = false;
    std::string in[2], mesh, out;
    // processing
    int vfMode = 0, cMode = 0, iterations = 10, threads = 0, pad = 2;
    float sSmooth = 3e-3f, vfSmooth = 0.f, vfSThreshold = 1e-8f, eLength = 0.006f, dogWeight = 1.f, dogSmooth = (float)1e-4, search = 1.f;
    float sMultiply = 0.25f, vMultiply = 1.0f;
    bool vfSmoothSet = false;
    bool divFree = false, verbose = false, showError = false, nearest = false, debug = false, logSpace = false;
    // additions of this build (not in the reference): PCG controls and device choice
    double flowTol = 1e-8, smoothTol = 1e-10;
    int device = 0;
};

void show_usage(const char* exe, const Options& defaults);
// Mirrors cmdLineParse: consumes argv[1..]; warnings go to stderr.
void parse_command_line(int argc, char** argv, Options& opt);

}  // namespace mof

#endif
