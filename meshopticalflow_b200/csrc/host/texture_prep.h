// Host-side preparation for the texture (--mesh m.ply --in A.png B.png) configuration: edge-length
// subdivision, wedge-averaged vertex colours and the texel -> (triangle, barycentric) map
// (include/Src/MeshFlow.inl:66-84, 158-232, 252-266, 281-467). Runs on the CPU, once per alignment,
// like the reference; the per-iteration work and the texel advection run on the GPU.
#ifndef MOF_TEXTURE_PREP_H
#define MOF_TEXTURE_PREP_H

#include <vector>

namespace mof {

struct TexturedMesh {
    std::vector<float> xyz;    // 3 per vertex, float like the reference's PlyVertex<float>
    std::vector<int> tri;      // 3 per triangle
    std::vector<double> uv;    // 6 per triangle: (u,v) of each corner
};

// Subdivide (MeshFlow.inl:223-232): split every edge longer than edgeLength until none is left.
// Returns the number of vertices added.
int subdivide(TexturedMesh& mesh, double edgeLength);

// Sample (MeshFlow.inl:66-84): RGB8 texture, rows top to bottom, (u, v) with v up.
void sample_texture(const unsigned char* tex, int W, int H, double u, double v, bool bilinear, double rgb[3]);

// SampleTextureToVertices (MeshFlow.inl:252-266): colours [3V], each vertex the mean over its wedges.
void sample_texture_to_vertices(const TexturedMesh& mesh, const unsigned char* tex, int W, int H, bool bilinear, std::vector<double>& colors);

// The geometry GetTextureSource needs for its exp-map pull-back, as the GPU library computed it
// (mof_get_array: MOF_ARR_OPPOSITE, MOF_ARR_XFORM_LINEAR, MOF_ARR_XFORM_CONSTANT).
struct EdgeTransforms {
    const int* opposite;     // [3T]
    const double* linear;    // [3T][4] row-major
    const double* constant;  // [3T][2]
};

// GetTextureSource (MeshFlow.inl:411-467). srcT[W*H] (-1 = uncovered), srcP[2*W*H]. Returns the number of
// texels whose pull-back ray missed its triangle (the reference exits on the first one).
int texture_source(const TexturedMesh& mesh, const EdgeTransforms& edges, int W, int H, int padRadius, std::vector<int>& srcT, std::vector<double>& srcP);

}  // namespace mof

#endif
