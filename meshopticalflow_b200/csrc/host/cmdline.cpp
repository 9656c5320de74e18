// See cmdline.h.
#include "cmdline.h"

#include <strings.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace mof {
namespace {

const char* kNames[] = {"mesh", "in", "out", "vfMode", "cMode", "iterations", "threads", "sSmooth", "vfSmooth", "verbose", "error", "sMultiply", "vMultiply",
                        "vfSThreshold", "dogWeight", "dogSmooth", "eLength", "nearest", "pad", "debug", "log", "search", "divFree", "flowTol", "smoothTol", "device", nullptr};

}  // namespace

void show_usage(const char* exe, const Options& d) {
    printf("Usage %s:\n", exe);
    printf("I/O Parameters: \n");
    printf("\t[--in <input textures (.ply or .png)>\n");
    printf("\t[--mesh <input geometry (.ply)>]\n");
    printf("\t[--out <output file (.ply or .png)>]\n");
    printf("Processing Parameters: \n");
    printf("\t[--eLength <subdivide edges up to this diagonal fraction> = %0.3f]\n", d.eLength);
    printf("\t[--iterations <alignment iterations>=%d]\n", d.iterations);
    printf("Scalar Field Parameters: \n");
    printf("\t[--sSmooth <scalar smoothing weight>=%f]\n", d.sSmooth);
    printf("\t[--sMultiply <scalar weight multiplication factor>=%g]\n", d.sMultiply);
    printf("\t[--dogWeight <difference of Gaussians blending weight>=%g]\n", d.dogWeight);
    printf("\t[--dogSmooth <difference of Gaussians smoothing weight>=%g]\n", d.dogSmooth);
    printf("Vector Field Parameters: \n");
    printf("\t[--vfMode <vector field mode >=%d]\n", d.vfMode);
    printf("\t \t [0] Whitney \n\t \t [1] Conformal \n\t \t [2] Connection \n");
    printf("\t[--cMode <connection mode >=%d]\n", d.cMode);
    printf("\t \t [0] Projected baricentric \n\t \t [1] Baricentric dual \n\t \t [2] Inverse cotangents \n");
    printf("\t[--vfSmooth <vector field smoothing weight>= Whitney -> %g,Conformal -> %g,Connection -> %g]\n", 3e-6, 5e-7, 1e4);
    printf("\t[--vMultiply <vector field weight multiplication factor>=%g]\n", d.vMultiply);
    printf("\t[--vfSThreshold <vector field weight threshold>=%g]\n", d.vfSThreshold);
    printf("Auxiliar Parameters: \n");
    printf("\t[--threads <parallelization threads> (ignored: the solver runs on the GPU)]\n");
    printf("\t[--pad <padding radius>=%d]\n", d.pad);
    printf("\t[--search <golden secition search range multiplier>=%g]\n", d.search);
    printf("\t[--divFree]\n\t[--log]\n\t[--nearest]\n\t[--error]\n\t[--verbose]\n\t[--debug]\n");
    printf("GPU solver parameters (this build): \n");
    printf("\t[--flowTol <PCG relative residual, flow system>=%g]\n", d.flowTol);
    printf("\t[--smoothTol <PCG relative residual, scalar smoothing>=%g]\n", d.smoothTol);
    printf("\t[--device <CUDA device>=%d]\n", d.device);
}

void parse_command_line(int argc, char** argv, Options& o) {
    int i = 1;
    auto value = [&](int k) -> const char* { return i + k < argc ? argv[i + k] : nullptr; };
    while (i < argc) {
        const char* a = argv[i];
        if (a[0] == '-' && a[1] == '-') {
            const char* n = a + 2;
            int used = 0;
            bool known = true;
            auto is = [&](const char* name) { return !strcasecmp(n, name); };
            auto f32 = [&](float& dst, bool* set = nullptr) { if (value(1)) dst = (float)atof(value(1)), used = 1; if (set && value(1)) *set = true; };
            auto i32 = [&](int& dst) { if (value(1)) dst = atoi(value(1)), used = 1; };
            if (is("mesh")) { if (value(1)) o.mesh = value(1), o.meshSet = true, used = 1; }
            else if (is("out")) { if (value(1)) o.out = value(1), o.outSet = true, used = 1; }
            else if (is("in")) { if (value(1) && value(2)) o.in[0] = value(1), o.in[1] = value(2), o.inSet = true, used = 2; }
            else if (is("vfMode")) i32(o.vfMode);
            else if (is("cMode")) i32(o.cMode);
            else if (is("iterations")) i32(o.iterations);
            else if (is("threads")) i32(o.threads);
            else if (is("pad")) i32(o.pad);
            else if (is("device")) i32(o.device);
            else if (is("sSmooth")) f32(o.sSmooth);
            else if (is("vfSmooth")) f32(o.vfSmooth, &o.vfSmoothSet);
            else if (is("vfSThreshold")) f32(o.vfSThreshold);
            else if (is("eLength")) f32(o.eLength);
            else if (is("dogWeight")) f32(o.dogWeight);
            else if (is("dogSmooth")) f32(o.dogSmooth);
            else if (is("search")) f32(o.search);
            else if (is("sMultiply")) f32(o.sMultiply);
            else if (is("vMultiply")) f32(o.vMultiply);
            else if (is("flowTol")) { if (value(1)) o.flowTol = atof(value(1)), used = 1; }
            else if (is("smoothTol")) { if (value(1)) o.smoothTol = atof(value(1)), used = 1; }
            else if (is("divFree")) o.divFree = true;
            else if (is("verbose")) o.verbose = true;
            else if (is("error")) o.showError = true;
            else if (is("nearest")) o.nearest = true;
            else if (is("debug")) o.debug = true;
            else if (is("log")) o.logSpace = true;
            else known = false;
            if (!known) {
                fprintf(stderr, "[WARNING] Invalid option: %s\n", a);
                for (int k = 0; kNames[k]; k++) printf("\t--%s\n", kNames[k]);
            }
            i += used;
        } else
            fprintf(stderr, "[WARNING] Parameter name should be of the form --<name>: %s\n", a);
        i++;
    }
}

}  // namespace mof
