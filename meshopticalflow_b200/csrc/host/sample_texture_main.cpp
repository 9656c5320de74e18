// SampleTextureToVertices — drop-in command line for the reference's sibling tool of the same name
// (SampleTextureToVertices/SampleTextureToVertices.cpp): reads a uv-mapped triangle mesh and a PNG texture,
// optionally splits edges longer than a fraction of the bounding-box diagonal, gives every vertex the texture
// colour of one of its wedges and writes a coloured PLY in the input's file type. It is how the per-vertex inputs
// (A.ply / B.ply) of `OpticalFlow --in A.ply B.ply` are made from the texture configuration's files.
//
//     SampleTextureToVertices --in mesh.ply --texture A.png --out A.ply [--eLength 0.006] [--verbose]
//
// Pure host code, single precision throughout like the reference (`_Execute<float>`, :129): the subdivision of
// Src/Subdivide.inl:81-155 with float edge lengths and float uv midpoints, SampleTexture of Src/Texture.inl:4-23
// in float, last wedge wins (:107-111), colours to uchar by truncation on output.
#include <strings.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "ply_io.h"
#include "png_codec.h"

namespace {

struct Wedges {
    int v[3];
    float uv[3][2];
};

int64_t edge_key(int a, int b) { return a > b ? ((int64_t)a << 32) | (int64_t)b : ((int64_t)b << 32) | (int64_t)a; }

// One sweep of _Subdivide (Src/Subdivide.inl:81-153), Real = float.
int subdivide_once(std::vector<float>& xyz, std::vector<Wedges>& tris, float edgeLength) {
    std::unordered_map<int64_t, int> midpoint;
    std::vector<Wedges> out;
    out.reserve(tris.size() * 2);
    const std::vector<float> old = xyz;
    int added = 0;
    auto emit = [&](int a, int b, int c, const float* ua, const float* ub, const float* uc) {
        Wedges w;
        w.v[0] = a, w.v[1] = b, w.v[2] = c;
        for (int k = 0; k < 2; k++) w.uv[0][k] = ua[k], w.uv[1][k] = ub[k], w.uv[2][k] = uc[k];
        out.push_back(w);
    };
    for (const Wedges& t : tris) {
        int e[3] = {-1, -1, -1}, split = 0;
        float mid[3][2] = {{0, 0}, {0, 0}, {0, 0}};
        for (int j = 0; j < 3; j++) {
            int a = t.v[j], b = t.v[(j + 1) % 3];
            float dx = old[3 * a] - old[3 * b], dy = old[3 * a + 1] - old[3 * b + 1], dz = old[3 * a + 2] - old[3 * b + 2];
            float len2 = dx * dx + dy * dy + dz * dz;
            if (len2 > edgeLength * edgeLength) {
                auto it = midpoint.find(edge_key(a, b));
                if (it == midpoint.end()) {
                    e[j] = (int)(xyz.size() / 3);
                    midpoint.emplace(edge_key(a, b), e[j]);
                    for (int k = 0; k < 3; k++) xyz.push_back((old[3 * a + k] + old[3 * b + k]) / 2);
                    added++;
                } else
                    e[j] = it->second;
                for (int k = 0; k < 2; k++) mid[j][k] = (t.uv[j][k] + t.uv[(j + 1) % 3][k]) / 2;
                split++;
            }
        }
        if (split == 0) out.push_back(t);
        else if (split == 1) {
            for (int j = 0; j < 3; j++)
                if (e[j] != -1) {
                    int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                    emit(t.v[j], e[j], t.v[j2], t.uv[j], mid[j], t.uv[j2]);
                    emit(t.v[j1], t.v[j2], e[j], t.uv[j1], t.uv[j2], mid[j]);
                }
        } else if (split == 2) {
            for (int j = 0; j < 3; j++)
                if (e[j] == -1) {
                    int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                    emit(e[j1], t.v[j2], e[j2], mid[j1], t.uv[j2], mid[j2]);
                    emit(t.v[j], t.v[j1], e[j2], t.uv[j], t.uv[j1], mid[j2]);
                    emit(t.v[j1], e[j1], e[j2], t.uv[j1], mid[j1], mid[j2]);
                }
        } else {
            for (int j = 0; j < 3; j++) emit(t.v[j], e[j], e[(j + 2) % 3], t.uv[j], mid[j], mid[(j + 2) % 3]);
            emit(e[0], e[1], e[2], mid[0], mid[1], mid[2]);
        }
    }
    tris.swap(out);
    return added;
}

// SampleTexture<float> (Src/Texture.inl:4-23): three weights are products in double rounded to float (a `1.` literal is
// involved), dx * dy is a float product; the sum is float.
void sample_texture(const unsigned char* tex, int W, int H, float u, float v, float rgb[3]) {
    v = 1 - v;
    u = std::min<float>(1.f, std::max<float>(0.f, u));
    v = std::min<float>(1.f, std::max<float>(0.f, v));
    u *= W - 1, v *= H - 1;
    int x0 = (int)std::floor(u), y0 = (int)std::floor(v);
    float dx = u - x0, dy = v - y0;
    int x1 = std::min(x0 + 1, W - 1), y1 = std::min(y0 + 1, H - 1);
    float w00 = (float)((1. - dx) * (1. - dy)), w10 = (float)(dx * (1. - dy)), w11 = dx * dy, w01 = (float)((1. - dx) * dy);
    // the shipped build (-O3 -ffast-math, SampleTextureToVertices/Makefile) adds the four products pairwise; the order is kept
    // because the result is truncated to uchar and a last-bit difference shows as 254 for 255
    for (int c = 0; c < 3; c++) {
        float c00 = tex[3 * (W * y0 + x0) + c], c10 = tex[3 * (W * y0 + x1) + c], c11 = tex[3 * (W * y1 + x1) + c], c01 = tex[3 * (W * y1 + x0) + c];
        rgb[c] = (c00 * w00 + c10 * w10) + (c11 * w11 + c01 * w01);
    }
}

std::string extension(const std::string& name) {
    size_t dot = name.rfind('.');
    return dot == std::string::npos ? std::string() : name.substr(dot + 1);
}

void usage(const char* exe) {
    printf("Usage %s:\n", exe);
    printf("\t --in <input mesh>\n\t --texture <input texture file>\n\t --out <output mesh>]\n");
    printf("\t --eLength <diagonal fraction edge length subdivision> [%f]\n\t --verbose\n", 0.006f);
}

}  // namespace

int main(int argc, char* argv[]) {
    std::string in, texture, out;
    bool inSet = false, textureSet = false, outSet = false, eSet = false, verbose = false;
    float eLength = 0.006f;
    const char* names[] = {"in", "texture", "out", "eLength", "verbose", nullptr};
    for (int i = 1; i < argc; i++) {
        const char* a = argv[i];
        if (a[0] == '-' && a[1] == '-') {  // CmdLineParser.inl:238-258: case-insensitive names, unknown ones warn and list
            const char* n = a + 2;
            bool more = i + 1 < argc;
            if (!strcasecmp(n, "in")) { if (more) in = argv[++i], inSet = true; }
            else if (!strcasecmp(n, "texture")) { if (more) texture = argv[++i], textureSet = true; }
            else if (!strcasecmp(n, "out")) { if (more) out = argv[++i], outSet = true; }
            else if (!strcasecmp(n, "eLength")) { if (more) eLength = (float)atof(argv[++i]), eSet = true; }
            else if (!strcasecmp(n, "verbose")) verbose = true;
            else {
                fprintf(stderr, "[WARNING] Invalid option: %s\n", a);
                for (int k = 0; names[k]; k++) printf("\t--%s\n", names[k]);
            }
        } else
            fprintf(stderr, "[WARNING] Parameter name should be of the form --<name>: %s\n", a);
    }
    if (!inSet || !textureSet) {
        usage(argv[0]);
        return EXIT_FAILURE;
    }
    if (strcasecmp(extension(texture).c_str(), "png")) {
        fprintf(stderr, "[ERROR] Unrecognized image extension: %s\n", extension(texture).c_str());
        return EXIT_FAILURE;
    }
    std::string err;
    std::vector<unsigned char> tex;
    int W = 0, H = 0;
    if (!mof::png_read_rgb8(texture.c_str(), tex, W, H, err)) {
        fprintf(stderr, "[ERROR] %s\n", err.c_str());
        return EXIT_FAILURE;
    }
    mof::PlyMesh ply;
    if (!mof::ply_read(in.c_str(), ply, err)) {
        fprintf(stderr, "[ERROR] %s\n", err.c_str());
        return EXIT_FAILURE;
    }
    size_t nf = ply.faceCount(), ip = 0, up = 0;
    std::vector<Wedges> tris(nf);
    for (size_t i = 0; i < nf; i++) {
        int nvtx = ply.faceSize[i], nuv = i < ply.uvSize.size() ? ply.uvSize[i] : 0;
        if (nvtx != 3 || nuv != 6) {
            fprintf(stderr, "[ERROR] Bad face: %d %d\n", nvtx, nuv);
            return EXIT_FAILURE;
        }
        for (int j = 0; j < 3; j++) tris[i].v[j] = ply.faceIndex[ip + j], tris[i].uv[j][0] = ply.uv[up + 2 * j], tris[i].uv[j][1] = ply.uv[up + 2 * j + 1];
        ip += nvtx, up += nuv;
    }
    std::vector<float> xyz = ply.xyz;
    if (eSet) {  // :89-102
        float lo[3], hi[3];
        for (int c = 0; c < 3; c++) lo[c] = hi[c] = xyz[c];
        for (size_t v = 0; v < xyz.size() / 3; v++)
            for (int c = 0; c < 3; c++) lo[c] = std::min(lo[c], xyz[3 * v + c]), hi[c] = std::max(hi[c], xyz[3 * v + c]);
        float d[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
        float diagonal = (float)std::sqrt((double)(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]));
        eLength = eLength * diagonal;
        while (subdivide_once(xyz, tris, eLength)) {}
    }
    size_t nv = xyz.size() / 3;
    std::vector<float> rgb(3 * nv, 0.f);
    std::vector<int> faces(3 * tris.size());
    for (size_t i = 0; i < tris.size(); i++)
        for (int j = 0; j < 3; j++) {
            faces[3 * i + j] = tris[i].v[j];
            // "Assuming texture mapping is seamless so that it doesn't make a difference which wedge we sample from" (:109)
            sample_texture(tex.data(), W, H, tris[i].uv[j][0], tris[i].uv[j][1], &rgb[3 * (size_t)tris[i].v[j]]);
        }
    if (verbose) printf("Vertices / Triangles: %d / %d\n", (int)nv, (int)tris.size());
    if (outSet) {
        bool ok = ply.format == 0 ? mof::ply_write_colored_ascii(out.c_str(), xyz, rgb, faces, err) : mof::ply_write_colored_binary(out.c_str(), xyz, rgb, faces, err);
        if (!ok) {
            fprintf(stderr, "[ERROR] %s\n", err.c_str());
            return EXIT_FAILURE;
        }
    }
    return EXIT_SUCCESS;
}
