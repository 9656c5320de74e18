// See texture_prep.h.
#include "texture_prep.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <unordered_map>

namespace mof {
namespace {

struct Vec2 {
    double x = 0, y = 0;
    Vec2() {}
    Vec2(double x_, double y_) : x(x_), y(y_) {}
    Vec2 operator+(const Vec2& o) const { return Vec2(x + o.x, y + o.y); }
    Vec2 operator-(const Vec2& o) const { return Vec2(x - o.x, y - o.y); }
    Vec2 operator*(double s) const { return Vec2(x * s, y * s); }
    Vec2 operator/(double s) const { return Vec2(x / s, y / s); }
};

struct UvTriangle {
    int v[3];
    Vec2 uv[3];
};

int64_t undirected_key(int a, int b) { return a > b ? ((int64_t)a << 32) | (int64_t)b : ((int64_t)b << 32) | (int64_t)a; }

// One sweep of _Subdivide (MeshFlow.inl:158-220). Midpoint vertices are numbered in the order their
// edges are first met while scanning triangles and corners.
int subdivide_once(std::vector<float>& xyz, std::vector<UvTriangle>& tris, double edgeLength) {
    std::unordered_map<int64_t, int> midpoint;
    std::vector<UvTriangle> out;
    out.reserve(tris.size() * 2);
    const std::vector<float> old = xyz;
    int added = 0;
    for (const UvTriangle& t : tris) {
        int e[3] = {-1, -1, -1}, split = 0;
        Vec2 mid[3];
        for (int j = 0; j < 3; j++) {
            int a = t.v[j], b = t.v[(j + 1) % 3];
            float dx = old[3 * a] - old[3 * b], dy = old[3 * a + 1] - old[3 * b + 1], dz = old[3 * a + 2] - old[3 * b + 2];
            double len2 = (double)dx * dx + (double)dy * dy + (double)dz * dz;
            if (len2 > edgeLength * edgeLength) {
                auto it = midpoint.find(undirected_key(a, b));
                if (it == midpoint.end()) {
                    e[j] = (int)(xyz.size() / 3);
                    midpoint.emplace(undirected_key(a, b), e[j]);
                    for (int k = 0; k < 3; k++) xyz.push_back((old[3 * a + k] + old[3 * b + k]) / 2);
                    added++;
                } else
                    e[j] = it->second;
                mid[j] = (t.uv[j] + t.uv[(j + 1) % 3]) / 2;
                split++;
            }
        }
        auto emit = [&](int a, int b, int c, Vec2 ua, Vec2 ub, Vec2 uc) { out.push_back(UvTriangle{{a, b, c}, {ua, ub, uc}}); };
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

}  // namespace

int subdivide(TexturedMesh& mesh, double edgeLength) {
    size_t nt = mesh.tri.size() / 3;
    std::vector<UvTriangle> tris(nt);
    for (size_t i = 0; i < nt; i++)
        for (int j = 0; j < 3; j++) tris[i].v[j] = mesh.tri[3 * i + j], tris[i].uv[j] = Vec2(mesh.uv[6 * i + 2 * j], mesh.uv[6 * i + 2 * j + 1]);
    int total = 0;
    while (int n = subdivide_once(mesh.xyz, tris, edgeLength)) total += n;
    mesh.tri.resize(3 * tris.size()), mesh.uv.resize(6 * tris.size());
    for (size_t i = 0; i < tris.size(); i++)
        for (int j = 0; j < 3; j++) mesh.tri[3 * i + j] = tris[i].v[j], mesh.uv[6 * i + 2 * j] = tris[i].uv[j].x, mesh.uv[6 * i + 2 * j + 1] = tris[i].uv[j].y;
    return total;
}

void sample_texture(const unsigned char* tex, int W, int H, double u, double v, bool bilinear, double rgb[3]) {
    v = 1 - v;
    u = std::min(1., std::max(0., u)) * (W - 1);
    v = std::min(1., std::max(0., v)) * (H - 1);
    int x0 = (int)std::floor(u), y0 = (int)std::floor(v);
    auto px = [&](int x, int y, int c) { return (double)tex[3 * (W * y + x) + c]; };
    if (!bilinear) {
        for (int c = 0; c < 3; c++) rgb[c] = px(x0, y0, c);
        return;
    }
    double dx = u - x0, dy = v - y0;
    int x1 = std::min(x0 + 1, W - 1), y1 = std::min(y0 + 1, H - 1);
    for (int c = 0; c < 3; c++)
        rgb[c] = px(x0, y0, c) * ((1. - dx) * (1. - dy)) + px(x1, y0, c) * (dx * (1. - dy)) + px(x1, y1, c) * (dx * dy) + px(x0, y1, c) * ((1. - dx) * dy);
}

void sample_texture_to_vertices(const TexturedMesh& mesh, const unsigned char* tex, int W, int H, bool bilinear, std::vector<double>& colors) {
    size_t nv = mesh.xyz.size() / 3, nt = mesh.tri.size() / 3;
    colors.assign(3 * nv, 0.);
    std::vector<int> wedges(nv, 0);
    for (size_t t = 0; t < nt; t++)
        for (int j = 0; j < 3; j++) {
            double rgb[3];
            sample_texture(tex, W, H, mesh.uv[6 * t + 2 * j], mesh.uv[6 * t + 2 * j + 1], bilinear, rgb);
            int v = mesh.tri[3 * t + j];
            for (int c = 0; c < 3; c++) colors[3 * v + c] += rgb[c];
            wedges[v]++;
        }
    for (size_t v = 0; v < nv; v++)
        for (int c = 0; c < 3; c++) colors[3 * v + c] /= wedges[v];
}

namespace {

// BarycentricCoordinate (MeshFlow.inl:268-278): p = c0 + s (c1-c0) + t (c2-c0).
Vec2 barycentric(const Vec2 c[3], Vec2 p) {
    Vec2 a = c[1] - c[0], b = c[2] - c[0], r = p - c[0];
    double inv = 1. / (a.x * b.y - b.x * a.y);
    return Vec2((b.y * r.x - b.x * r.y) * inv, (-a.y * r.x + a.x * r.y) * inv);
}

struct TexelMap {
    int W, H;
    std::vector<int>& tri;
    std::vector<double>& bary;
};

// RasterizeTriangle (MeshFlow.inl:281-337). Corners are ordered by v; scan lines run between the two
// active sides; a texel keeps its first triangle unless the rule at :334 fires.
void rasterize(const Vec2 uvIn[3], int t, TexelMap& map) {
    Vec2 c[3];
    for (int j = 0; j < 3; j++) c[j] = Vec2(uvIn[j].x * (map.W - 1), uvIn[j].y * (map.H - 1));
    int order[3];
    double y0 = uvIn[0].y, y1 = uvIn[1].y, y2 = uvIn[2].y;
    if (y0 <= y1 && y0 <= y2) order[0] = 0, order[1] = y1 <= y2 ? 1 : 2, order[2] = y1 <= y2 ? 2 : 1;
    else if (y1 <= y0 && y1 <= y2) order[0] = 1, order[1] = y0 <= y2 ? 0 : 2, order[2] = y0 <= y2 ? 2 : 0;
    else order[0] = 2, order[1] = y0 <= y1 ? 0 : 1, order[2] = y0 <= y1 ? 1 : 0;
    Vec2 lo = c[order[0]], midc = c[order[1]], hi = c[order[2]];
    auto clampi = [](int v, int n) { return std::max(0, std::min(n - 1, v)); };
    int yBegin = clampi((int)std::ceil(lo.y), map.H), yEnd = clampi((int)std::floor(hi.y), map.H);
    Vec2 apex = lo, sideA = midc - lo, sideB = hi - lo;
    for (int y = yBegin; y <= yEnd; y++) {
        if (y >= midc.y) apex = hi, sideA = midc - hi, sideB = lo - hi;
        if (sideA.y == 0 || sideB.y == 0) continue;
        double xa = apex.x + ((double)y - apex.y) * sideA.x / sideA.y, xb = apex.x + ((double)y - apex.y) * sideB.x / sideB.y;
        int xBegin = clampi((int)std::ceil(std::min(xa, xb)), map.W), xEnd = clampi((int)std::floor(std::max(xa, xb)), map.W);
        Vec2 bBegin = barycentric(c, Vec2((double)xBegin, (double)y)), bEnd = barycentric(c, Vec2((double)xEnd, (double)y));
        for (int x = xBegin; x <= xEnd; x++) {
            double s = xBegin == xEnd ? 0. : (double)(x - xBegin) / (double)(xEnd - xBegin);
            Vec2 b = bBegin * (1. - s) + bEnd * s;
            size_t i = (size_t)y * map.W + x;
            if (map.tri[i] == -1 || (b.x >= 0 && b.y >= 1 && b.x + b.y <= 1)) map.tri[i] = t, map.bary[2 * i] = b.x, map.bary[2 * i + 1] = b.y;
        }
    }
}

// RiemannianMesh::exp (FEM.inl:835-899): follow the straight line from p with velocity v across
// edges until the velocity is used up. Returns false if the ray misses the triangle.
bool exp_map(const EdgeTransforms& e, int& t, Vec2& p, Vec2 v) {
    if (!(v.x * v.x + v.y * v.y)) return true;
    int cameFrom = -1;
    auto cross = [&](int side) {
        int h = 3 * t + side, o = e.opposite[h];
        const double* L = e.linear + 4 * (size_t)h;
        const double* c = e.constant + 2 * (size_t)h;
        p = Vec2(L[0] * p.x + L[1] * p.y + c[0], L[2] * p.x + L[3] * p.y + c[1]);
        v = Vec2(L[0] * v.x + L[1] * v.y, L[2] * v.x + L[3] * v.y);
        t = o / 3, cameFrom = o % 3;
    };
    if (p.x <= 0 && v.x < 0) cross(1);
    else if (p.y <= 0 && v.y < 0) cross(2);
    else if (p.x + p.y >= 1 && v.x + v.y > 0) cross(0);
    for (int count = 0; count < 10000; count++) {
        double best = 0;
        int side = -1;
        double s2 = -p.y / v.y, s1 = -p.x / v.x, s0 = (1. - p.x - p.y) / (v.y + v.x);
        if (cameFrom != 2 && s2 > 0) { double q = p.x + v.x * s2; if (q >= 0 && q <= 1 && s2 > best) side = 2, best = s2; }
        if (cameFrom != 1 && s1 > 0) { double q = p.y + v.y * s1; if (q >= 0 && q <= 1 && s1 > best) side = 1, best = s1; }
        if (cameFrom != 0 && s0 > 0) { double q = p.x + v.x * s0; if (q >= 0 && q <= 1 && s0 > best) side = 0, best = s0; }
        if (side == -1) return false;
        if (best > 1) { p = p + v; return true; }
        p = p + v * best, v = v - v * best;
        cross(side);
    }
    return true;  // "[WARNING] Failed to converge exp" in the reference
}

}  // namespace

int texture_source(const TexturedMesh& mesh, const EdgeTransforms& edges, int W, int H, int padRadius, std::vector<int>& srcT, std::vector<double>& srcP) {
    size_t nt = mesh.tri.size() / 3, n = (size_t)W * H;
    srcT.assign(n, -1), srcP.assign(2 * n, 0.);
    TexelMap map{W, H, srcT, srcP};
    auto corners = [&](size_t t, Vec2 c[3]) { for (int j = 0; j < 3; j++) c[j] = Vec2(mesh.uv[6 * t + 2 * j], mesh.uv[6 * t + 2 * j + 1]); };
    for (size_t t = 0; t < nt; t++) {
        Vec2 c[3];
        corners(t, c);
        rasterize(c, (int)t, map);
    }
    // Grow the map by padRadius rings (MeshFlow.inl:426-455): an empty texel takes the triangle of a
    // covered 4-neighbour (the last one found in the reference's scan order: x-1, x+1, y-1, y+1).
    std::vector<int> grow(n);
    for (int ring = 0; ring < padRadius; ring++) {
        for (int x = 0; x < W; x++)
            for (int y = 0; y < H; y++) {
                size_t i = (size_t)y * W + x;
                grow[i] = -1;
                if (srcT[i] != -1) continue;
                for (int dx = -1; dx <= 1; dx++)
                    if (x + dx >= 0 && x + dx < W && srcT[(size_t)y * W + (x + dx)] != -1) grow[i] = srcT[(size_t)y * W + (x + dx)];
                for (int dy = -1; dy <= 1; dy++)
                    if (y + dy >= 0 && y + dy < H && srcT[(size_t)(y + dy) * W + x] != -1) grow[i] = srcT[(size_t)(y + dy) * W + x];
            }
        for (int x = 0; x < W; x++)
            for (int y = 0; y < H; y++) {
                size_t i = (size_t)y * W + x;
                if (grow[i] == -1) continue;
                Vec2 c[3];
                corners((size_t)grow[i], c);
                Vec2 b = barycentric(c, Vec2((double)x / (W - 1), (double)y / (H - 1)));
                srcT[i] = grow[i], srcP[2 * i] = b.x, srcP[2 * i + 1] = b.y;
            }
    }
    // RemapSamplePoint (MeshFlow.inl:340-350): points outside their triangle are reached from the
    // centroid along the surface.
    int misses = 0;
    for (int x = 0; x < W; x++)
        for (int y = 0; y < H; y++) {
            size_t i = (size_t)y * W + x;
            if (srcT[i] == -1) continue;
            Vec2 p(srcP[2 * i], srcP[2 * i + 1]);
            if (p.x >= 0 && p.y >= 0 && p.x + p.y <= 1) continue;
            Vec2 start(1. / 3, 1. / 3);
            int t = srcT[i];
            Vec2 q = start;
            if (!exp_map(edges, t, q, p - start)) misses++;
            srcT[i] = t, srcP[2 * i] = q.x, srcP[2 * i + 1] = q.y;
        }
    return misses;
}

}  // namespace mof
