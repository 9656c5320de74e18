// OpticalFlow — drop-in command line for the reference's headless alignment path:
//
//     OpticalFlow --in A.ply B.ply --out result.ply                 (per-vertex colours)
//     OpticalFlow --mesh mesh.ply --in A.png B.png --out result.png (textures over a uv-mapped mesh)
//
// Same flags, defaults, messages and output files as OpticalFlow/OpticalFlow.cpp (main :1096-1116,
// _main :1059-1094, WhitneyFlowViewer::Init :680-917, IterativeOptimization :1036-1056). File I/O and
// the one-time texture-map preparation run here on the host; everything from the metric to the final
// advection runs on the GPU through the C ABI of include/mof_b200.h. The interactive viewer
// (no --out) is not part of this build.
#include <strings.h>
#include <sys/time.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cmdline.h"
#include "mof_b200.h"
#include "ply_io.h"
#include "png_codec.h"
#ifdef MOF_WITH_HOST_TEXPREP
#include "texture_prep.h"  // test builds only (see below)
#else
namespace mof {
struct TexturedMesh {          // the texture configuration's mesh as the files give it (texture_prep.h has the same, with the host restatement)
    std::vector<float> xyz;    // 3 per vertex, float like the reference's PlyVertex<float>
    std::vector<int> tri;      // 3 per triangle
    std::vector<double> uv;    // 6 per triangle: (u,v) of each corner
};
}  // namespace mof
#endif

namespace {

struct Stopwatch {
    double start;
    Stopwatch() { reset(); }
    static double now() {
        timeval t;
        gettimeofday(&t, nullptr);
        return t.tv_sec + t.tv_usec * 1e-6;
    }
    void reset() { start = now(); }
    double elapsed() const { return now() - start; }
};

std::string extension(const std::string& name) {
    size_t dot = name.rfind('.');
    return dot == std::string::npos ? std::string() : name.substr(dot + 1);
}

bool mof_ok(mof_ctx* ctx, int rc) {
    if (rc == MOF_OK) return true;
    fprintf(stderr, "%s\n", mof_last_error(ctx));
    return false;
}

}  // namespace

int main(int argc, char* argv[]) {
    mof::Options opt;
    mof::parse_command_line(argc, argv, opt);
    if (!opt.inSet) {
        mof::show_usage(argv[0], mof::Options());
        return EXIT_FAILURE;
    }
    if (opt.search <= 0) fprintf(stderr, "[WARNING] Search range must be positive: %g<=0\n", opt.search);
    opt.dogWeight = std::min(1.f, std::max(0.f, opt.dogWeight));
    if (opt.vfMode < 0 || opt.vfMode > 2) {
        printf("ERROR: Unsupported vector field! \n");  // OpticalFlow.cpp:867
        return 0;
    }
    if (opt.vfMode == 2 && (opt.cMode < 0 || opt.cMode > 2)) {
        printf("Undefined Connection Mode \n");  // Connection.inl:68
        return EXIT_FAILURE;
    }
    if (!opt.outSet) {
        fprintf(stderr, "[ERROR] the interactive viewer is not part of this build: pass --out <file>\n");
        return EXIT_FAILURE;
    }

    const bool processTexture = opt.meshSet;
    // The texture configuration's one-time preparation (Subdivide, SampleTextureToVertices, GetTextureSource) runs on the GPU
    // (csrc/texprep_kernels.cu). The serial host restatement of it (host/texture_prep.cpp — same outputs, bit for bit) is NOT part of
    // the product: it is compiled in by the tests only (-DMOF_WITH_HOST_TEXPREP, where MOF_GPU_TEXPREP=0 selects it as the cross-check).
#ifdef MOF_WITH_HOST_TEXPREP
    const char* gpuPrepEnv = getenv("MOF_GPU_TEXPREP");
    const bool gpuPrep = processTexture && !(gpuPrepEnv && *gpuPrepEnv == '0');
#else
    const bool gpuPrep = processTexture;
#endif
    mof_ctx* ctx = nullptr;
    auto ensure_context = [&]() {
        if (ctx) return true;
        if (mof_create(opt.device, nullptr, &ctx) != MOF_OK) {
            fprintf(stderr, "[ERROR] no usable CUDA device %d (this build has no CPU solver)\n", opt.device);
            return false;
        }
        return true;
    };
    mof::TexturedMesh tmesh;          // texture configuration
    std::vector<unsigned char> textures[2];
    int tW = 0, tH = 0;
    std::vector<double> vertices;     // 3V, double
    std::vector<int> triangles;       // 3T
    std::vector<double> signal[2];    // 3V
    std::vector<float> outXyz;

    if (processTexture) {
        // OpticalFlow.cpp:686-751
        mof::PlyMesh ply;
        std::string err;
        if (!mof::ply_read(opt.mesh.c_str(), ply, err)) {
            printf("ERROR: Unable to read mesh %s \n", opt.mesh.c_str());
            return 0;
        }
        size_t nf = ply.faceCount();
        size_t ip = 0, up = 0;
        tmesh.xyz = ply.xyz;
        tmesh.tri.resize(3 * nf), tmesh.uv.resize(6 * nf);
        for (size_t i = 0; i < nf; i++) {
            int nvtx = ply.faceSize[i], nuv = i < ply.uvSize.size() ? ply.uvSize[i] : 0;
            if (nvtx != 3 || nuv != 6) {
                fprintf(stderr, "[ERROR] Bad face: %d %d\n", nvtx, nuv);
                return 0;
            }
            for (int j = 0; j < 3; j++) tmesh.tri[3 * i + j] = ply.faceIndex[ip + j];
            for (int j = 0; j < 6; j++) tmesh.uv[6 * i + j] = (double)ply.uv[up + j];
            ip += nvtx, up += nuv;
        }
        double lo[3], hi[3];
        for (int c = 0; c < 3; c++) lo[c] = hi[c] = (double)tmesh.xyz[c];
        for (size_t v = 0; v < tmesh.xyz.size() / 3; v++)
            for (int c = 0; c < 3; c++) lo[c] = std::min(lo[c], (double)tmesh.xyz[3 * v + c]), hi[c] = std::max(hi[c], (double)tmesh.xyz[3 * v + c]);
        double diagonal = std::sqrt((hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1]) + (hi[2] - lo[2]) * (hi[2] - lo[2]));
        float eLength = (float)(opt.eLength * diagonal);  // float parameter times double, stored back as float (:713)
        if (eLength > 0 && gpuPrep) {
            if (!ensure_context()) return EXIT_FAILURE;
            int newV = 0, newT = 0;
            if (!mof_ok(ctx, mof_subdivide(ctx, tmesh.xyz.data(), (int)(tmesh.xyz.size() / 3), tmesh.tri.data(), tmesh.uv.data(), (int)nf, (double)eLength, &newV, &newT)))
                return EXIT_FAILURE;
            tmesh.xyz.resize(3 * (size_t)newV), tmesh.tri.resize(3 * (size_t)newT), tmesh.uv.resize(6 * (size_t)newT);
            if (!mof_ok(ctx, mof_get_subdivision(ctx, tmesh.xyz.data(), tmesh.tri.data(), tmesh.uv.data()))) return EXIT_FAILURE;
        }
#ifdef MOF_WITH_HOST_TEXPREP
        else if (eLength > 0)
            mof::subdivide(tmesh, (double)eLength);
#endif
        printf("Num vertices %d  \n", (int)(tmesh.xyz.size() / 3));

        for (int s = 0; s < 2; s++) {
            if (strcasecmp(extension(opt.in[s]).c_str(), "png")) {
                fprintf(stderr, "[ERROR] Unrecognized image extension: %s\n", extension(opt.in[s]).c_str());
                return EXIT_FAILURE;
            }
            int w, h;
            if (!mof::png_read_rgb8(opt.in[s].c_str(), textures[s], w, h, err)) {
                fprintf(stderr, "[ERROR] %s\n", err.c_str());
                return 0;
            }
            if (s == 0) tW = w, tH = h;
            else if (tW != w || tH != h) {
                fprintf(stderr, "[ERROR] Texture resolutions don't match: %d x %d != %d x %d\n", tW, tH, w, h);
                return EXIT_FAILURE;
            }
        }
#ifdef MOF_WITH_HOST_TEXPREP
        if (!gpuPrep)
            for (int s = 0; s < 2; s++) mof::sample_texture_to_vertices(tmesh, textures[s].data(), tW, tH, !opt.nearest, signal[s]);
#endif
        vertices.assign(tmesh.xyz.begin(), tmesh.xyz.end());
        triangles = tmesh.tri;
    } else {
        // OpticalFlow.cpp:755-779
        mof::PlyMesh ply[2];
        std::string err;
        for (int s = 0; s < 2; s++)
            if (!mof::ply_read(opt.in[s].c_str(), ply[s], err)) {
                fprintf(stderr, "[ERROR] %s\n", err.c_str());
                return EXIT_FAILURE;
            }
        if (ply[0].vertexCount() != ply[1].vertexCount()) {
            fprintf(stderr, "[ERROR] Vertex counts differ: %d != %d\n", (int)ply[0].vertexCount(), (int)ply[1].vertexCount());
            return EXIT_FAILURE;
        }
        if (ply[0].faceCount() != ply[1].faceCount()) {
            fprintf(stderr, "[ERROR] Different number of triangles in meshes: %d != %d\n", (int)ply[0].faceCount(), (int)ply[1].faceCount());
            return EXIT_FAILURE;
        }
        for (int s = 0; s < 2; s++) {
            for (size_t i = 0; i < ply[s].faceCount(); i++)
                if (ply[s].faceSize[i] != 3) {
                    fprintf(stderr, "[ERROR] only triangle meshes are supported (face %d has %d vertices)\n", (int)i, ply[s].faceSize[i]);
                    return EXIT_FAILURE;
                }
            if (ply[s].rgb.empty()) ply[s].rgb.assign(ply[s].xyz.size(), 0.f);  // the reference reads absent colours as 0
        }
        triangles = ply[0].faceIndex;
        for (size_t i = 0; i < triangles.size(); i++)
            if (ply[0].faceIndex[i] != ply[1].faceIndex[i]) {
                fprintf(stderr, "[ERROR] Triangle indices don't match: [%d,%d] %d != %d\n", (int)(i / 3), (int)(i % 3), ply[0].faceIndex[i], ply[1].faceIndex[i]);
                return EXIT_FAILURE;
            }
        size_t n = ply[0].xyz.size();
        vertices.resize(n);
        for (size_t i = 0; i < n; i++) vertices[i] = (double)ply[0].xyz[i] * 0.5 + (double)ply[1].xyz[i] * 0.5;
        for (int s = 0; s < 2; s++) signal[s].assign(ply[s].rgb.begin(), ply[s].rgb.end());
    }
    const int V = (int)(vertices.size() / 3), T = (int)(triangles.size() / 3);
    if (opt.verbose) printf("Vertices / Triangles: %d / %d\n", V, T);
    outXyz.resize(vertices.size());
    for (size_t i = 0; i < vertices.size(); i++) outXyz[i] = (float)vertices[i];

    if (!ensure_context()) return EXIT_FAILURE;
    mof_params params;
    mof_default_params(&params);
    params.iterations = opt.iterations;
    params.sSmooth = (double)opt.sSmooth, params.sMultiply = (double)opt.sMultiply;
    const double vfSmoothDefault[3] = {3e-6, 5e-7, 1e4};  // _main, OpticalFlow.cpp:1064-1069
    params.vfSmooth = opt.vfSmoothSet ? (double)opt.vfSmooth : vfSmoothDefault[opt.vfMode];
    params.vfMode = opt.vfMode, params.cMode = opt.cMode;
    params.logSpace = opt.logSpace ? 1 : 0;  // :821 — the comparison signals only; the colours advected at the end stay raw (:482-489, :1049-1054)
    params.vMultiply = (double)opt.vMultiply, params.vfSThreshold = (double)opt.vfSThreshold;
    params.dogWeight = (double)opt.dogWeight, params.dogSmooth = (double)opt.dogSmooth;
    params.flowTol = opt.flowTol, params.smoothTol = opt.smoothTol;
    if (!mof_ok(ctx, mof_set_params(ctx, &params))) return EXIT_FAILURE;
    // The texture configuration's map follows the file's triangle order (first-writer rule, MeshFlow.inl:281-337) and --debug dumps the
    // library's per-vertex arrays: both keep the caller's numbering; otherwise the library may renumber a badly ordered mesh (mof_set_reorder).
    if (!mof_ok(ctx, mof_set_reorder(ctx, processTexture || opt.debug ? 0 : -1))) return EXIT_FAILURE;

    {
        Stopwatch t;
        if (!mof_ok(ctx, mof_set_mesh(ctx, vertices.data(), V, triangles.data(), T))) return 0;
        if (opt.verbose) printf("Got edge transforms: %.2f (s)\nGot system matrices: %.2f (s)\n", t.elapsed(), 0.);
    }
    if (gpuPrep) {
        // GetTextureSource (:818) and SampleTextureToVertices (:741-745) on the GPU
        int misses = 0;
        if (!mof_ok(ctx, mof_build_texture_map(ctx, tW, tH, opt.pad, tmesh.uv.data(), textures[0].data(), textures[1].data(), &misses))) return misses ? 0 : EXIT_FAILURE;
        for (int s = 0; s < 2; s++) signal[s].resize(3 * (size_t)V);
        if (!mof_ok(ctx, mof_sample_textures_to_vertices(ctx, opt.nearest ? 0 : 1, signal[0].data(), signal[1].data()))) return EXIT_FAILURE;
    }
#ifdef MOF_WITH_HOST_TEXPREP
    else if (processTexture) {
        // GetTextureSource (:818) on the host, from the edge transforms the GPU just built
        std::vector<int> opp(3 * (size_t)T), srcT;
        std::vector<double> lin(12 * (size_t)T), cst(6 * (size_t)T), srcP;
        if (!mof_ok(ctx, mof_get_array(ctx, MOF_ARR_OPPOSITE, opp.data())) || !mof_ok(ctx, mof_get_array(ctx, MOF_ARR_XFORM_LINEAR, lin.data())) ||
            !mof_ok(ctx, mof_get_array(ctx, MOF_ARR_XFORM_CONSTANT, cst.data())))
            return EXIT_FAILURE;
        mof::EdgeTransforms edges{opp.data(), lin.data(), cst.data()};
        int misses = mof::texture_source(tmesh, edges, tW, tH, opt.pad, srcT, srcP);
        if (misses) {
            fprintf(stderr, "[ERROR] FEM::Mesh::exp:\n        Ray does not intersect triangle (%d texels)\n", misses);
            return 0;
        }
        if (!mof_ok(ctx, mof_set_texture_map(ctx, tW, tH, srcT.data(), srcP.data(), tmesh.uv.data(), textures[0].data(), textures[1].data()))) return EXIT_FAILURE;
    }
#endif
    {
        Stopwatch t;
        if (!mof_ok(ctx, mof_set_signals(ctx, signal[0].data(), signal[1].data(), 3))) return EXIT_FAILURE;
        if (opt.verbose && opt.dogWeight > 0) printf("Set comparison values: %.2f (s)\n", t.elapsed());
    }

    // IterativeOptimization (:1036-1056)
    mof_stats before, after;
    for (int i = 0; i < opt.iterations; i++) {
        Stopwatch t;
        mof_get_stats(ctx, &before);
        if (!mof_ok(ctx, mof_iterate(ctx, 1))) return EXIT_FAILURE;
        if (opt.debug) {
            // UpdateFlow's --debug dump (:458-465): the two resampled signals (channels 0-2) as coloured binary meshes
            std::vector<double> res6(6 * (size_t)V);
            const bool blend = opt.dogWeight > 0 && opt.dogWeight < 1;
            if (!mof_ok(ctx, mof_get_array(ctx, blend ? MOF_ARR_RESAMPLED_RAW : MOF_ARR_RESAMPLED, res6.data()))) return EXIT_FAILURE;
            std::vector<float> xyzf(vertices.begin(), vertices.end()), rgb(3 * (size_t)V);
            std::string werr;
            for (int s = 0; s < 2; s++) {
                for (int v = 0; v < V; v++)
                    for (int c = 0; c < 3; c++) rgb[3 * (size_t)v + c] = std::min(255.f, std::max(0.f, (float)res6[6 * (size_t)v + 3 * s + c]));
                char name[64];
                snprintf(name, sizeof(name), "resampled.%c.%d.ply", s ? 'T' : 'S', i);
                if (!mof::ply_write_colored_binary(name, xyzf, rgb, triangles, werr)) fprintf(stderr, "[ERROR] %s\n", werr.c_str());
            }
        }
        if (opt.verbose) {
            mof_get_stats(ctx, &after);
            printf("\t Signal Smoothing: %.4f(s)\n", (after.smoothSolveMs - before.smoothSolveMs) * 1e-3);
            printf("\t Signal advection : %.4f(s)\n", (after.advectMs - before.advectMs) * 1e-3);
            printf("\t PCG solve: %.4f(s) [%lld iterations, relative residual %.3g]\n", (after.flowSolveMs - before.flowSolveMs) * 1e-3,
                   after.flowCgIterations - before.flowCgIterations, after.lastFlowResidual);
            printf("Got flow: %.2f (s)\n", t.elapsed());
        }
    }
    std::string err;
    if (processTexture) {
        size_t n = (size_t)tW * tH;
        std::vector<double> outA(3 * n), outB(3 * n);
        if (!mof_ok(ctx, mof_advect_texels(ctx, 0.5, opt.nearest ? 0 : 1, outA.data(), outB.data()))) return EXIT_FAILURE;
        // average (:1046) and OutputImage with flipY (:126-137): (int) truncation, clamp, rows flipped
        std::vector<unsigned char> pixels(3 * n);
        for (int x = 0; x < tW; x++)
            for (int y = 0; y < tH; y++)
                for (int c = 0; c < 3; c++) {
                    double v = (outA[3 * ((size_t)y * tW + x) + c] + outB[3 * ((size_t)y * tW + x) + c]) / 2.0;
                    pixels[3 * ((size_t)(tH - 1 - y) * tW + x) + c] = (unsigned char)std::max(0, std::min(255, (int)v));
                }
        if (strcasecmp(extension(opt.out).c_str(), "png")) {
            fprintf(stderr, "[ERROR] Unrecognized image extension: %s\n", extension(opt.out).c_str());
            return 0;
        }
        if (!mof::png_write_rgb8(opt.out.c_str(), pixels.data(), tW, tH, err)) {
            fprintf(stderr, "[ERROR] %s\n", err.c_str());
            return 0;
        }
    } else {
        std::vector<double> outA(3 * (size_t)V), outB(3 * (size_t)V);
        if (!mof_ok(ctx, mof_advect_vertices(ctx, 0.5, outA.data(), outB.data()))) return EXIT_FAILURE;
        // average in double, through float (:1053), float again and clamp in OutputMesh (:144-145)
        std::vector<float> rgb(3 * (size_t)V);
        for (size_t i = 0; i < rgb.size(); i++) rgb[i] = std::min(255.f, std::max(0.f, (float)((outA[i] + outB[i]) / 2.0)));
        if (!mof::ply_write_colored_ascii(opt.out.c_str(), outXyz, rgb, triangles, err)) {
            fprintf(stderr, "[ERROR] %s\n", err.c_str());
            return 0;
        }
    }
    if (opt.verbose) {
        mof_get_stats(ctx, &after);
        printf("GPU: %lld kernel launches, %lld flow PCG iterations, %lld smoothing PCG iterations\n", after.kernelLaunches, after.flowCgIterations, after.smoothCgIterations);
    }
    mof_destroy(ctx);
    return EXIT_SUCCESS;
}
