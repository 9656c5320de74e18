// PLY reader/writer for the three layouts the reference's command line touches
// (include/Misha/Ply.h:394-405 coloured vertices, :46 plain vertices, :710-714 textured faces;
// output written by OutputMesh, OpticalFlow.cpp:139-148 -> PlyWriteTriangles, Ply.inl:1148).
#ifndef MOF_PLY_IO_H
#define MOF_PLY_IO_H

#include <string>
#include <vector>

namespace mof {

struct PlyMesh {
    std::vector<float> xyz;        // 3 per vertex; the reference reads positions into float (PlyVertex<float>)
    std::vector<float> rgb;        // 3 per vertex (red/green/blue or diffuse_*), empty if the file has none
    std::vector<int> faceSize;     // vertices per face, as stored
    std::vector<int> faceIndex;    // concatenated vertex indices
    std::vector<int> uvSize;       // texcoord entries per face (empty if the file has none)
    std::vector<float> uv;         // concatenated texcoord values
    int format = 0;                // as read: 0 ascii, 1 binary_little_endian, 2 binary_big_endian
    size_t vertexCount() const { return xyz.size() / 3; }
    size_t faceCount() const { return faceSize.size(); }
};

bool ply_read(const char* file_name, PlyMesh& mesh, std::string& err);

// ASCII: float x y z, uchar red green blue, list uchar int vertex_indices; numbers printed "%g " / "%u " /
// "%d " like write_ascii_item (PlyFile.inl:2117-2160); colours are uchar by truncation (PlyFile.inl:2309-2313).
bool ply_write_colored_ascii(const char* file_name, const std::vector<float>& xyz, const std::vector<float>& rgb, const std::vector<int>& tri, std::string& err);
// The same elements as binary_little_endian records (PLY_BINARY_NATIVE on this platform): what --debug writes (OpticalFlow.cpp:458-465).
bool ply_write_colored_binary(const char* file_name, const std::vector<float>& xyz, const std::vector<float>& rgb, const std::vector<int>& tri, std::string& err);

}  // namespace mof

#endif
