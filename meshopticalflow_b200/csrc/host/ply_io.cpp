// See ply_io.h.
#include "ply_io.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

namespace mof {
namespace {

enum Type { T_I8, T_U8, T_I16, T_U16, T_I32, T_U32, T_F32, T_F64, T_BAD };

Type parse_type(const std::string& s) {
    if (s == "char" || s == "int8") return T_I8;
    if (s == "uchar" || s == "uint8") return T_U8;
    if (s == "short" || s == "int16") return T_I16;
    if (s == "ushort" || s == "uint16") return T_U16;
    if (s == "int" || s == "int32") return T_I32;
    if (s == "uint" || s == "uint32") return T_U32;
    if (s == "float" || s == "float32") return T_F32;
    if (s == "double" || s == "float64") return T_F64;
    return T_BAD;
}
int type_size(Type t) {
    switch (t) {
        case T_I8: case T_U8: return 1;
        case T_I16: case T_U16: return 2;
        case T_I32: case T_U32: case T_F32: return 4;
        case T_F64: return 8;
        default: return 0;
    }
}

struct Property {
    std::string name;
    bool isList = false;
    Type type = T_BAD, countType = T_BAD;
};
struct Element {
    std::string name;
    size_t count = 0;
    std::vector<Property> props;
};

struct Reader {
    const unsigned char* data;
    size_t size, pos;
    int format;  // 0 ascii, 1 little endian, 2 big endian
    bool ok = true;

    double binary(Type t) {
        int n = type_size(t);
        if (pos + n > size) { ok = false; return 0; }
        unsigned char b[8];
        for (int i = 0; i < n; i++) b[i] = format == 2 ? data[pos + n - 1 - i] : data[pos + i];
        pos += n;
        switch (t) {
            case T_I8: return (double)(signed char)b[0];
            case T_U8: return (double)b[0];
            case T_I16: { short v; memcpy(&v, b, 2); return v; }
            case T_U16: { unsigned short v; memcpy(&v, b, 2); return v; }
            case T_I32: { int v; memcpy(&v, b, 4); return v; }
            case T_U32: { unsigned v; memcpy(&v, b, 4); return v; }
            case T_F32: { float v; memcpy(&v, b, 4); return v; }
            case T_F64: { double v; memcpy(&v, b, 8); return v; }
            default: ok = false; return 0;
        }
    }
    double ascii() {
        while (pos < size && isspace(data[pos])) pos++;
        if (pos >= size) { ok = false; return 0; }
        char* end = nullptr;
        double v = strtod((const char*)data + pos, &end);
        if (end == (const char*)data + pos) { ok = false; return 0; }
        pos = (size_t)(end - (const char*)data);
        return v;
    }
    double next(Type t) { return format == 0 ? ascii() : binary(t); }
};

}  // namespace

bool ply_read(const char* file_name, PlyMesh& mesh, std::string& err) {
    FILE* fp = fopen(file_name, "rb");
    if (!fp) { err = std::string("Unable to read mesh ") + file_name; return false; }
    fseek(fp, 0, SEEK_END);
    long size = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    std::vector<unsigned char> bytes((size_t)(size > 0 ? size : 0) + 1, 0);  // trailing 0 keeps strtod in bounds
    if (size > 0 && fread(bytes.data(), 1, (size_t)size, fp) != (size_t)size) { fclose(fp); err = "short read"; return false; }
    fclose(fp);

    // header
    size_t pos = 0;
    std::vector<Element> elements;
    int format = -1;
    bool sawPly = false, done = false;
    while (pos < (size_t)size && !done) {
        size_t eol = pos;
        while (eol < (size_t)size && bytes[eol] != '\n') eol++;
        std::string line((const char*)bytes.data() + pos, eol - pos);
        pos = eol + 1;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ss(line);
        std::string word;
        if (!(ss >> word)) continue;
        if (word == "ply") sawPly = true;
        else if (word == "format") {
            std::string f;
            ss >> f;
            format = f == "ascii" ? 0 : f == "binary_little_endian" ? 1 : f == "binary_big_endian" ? 2 : -1;
        } else if (word == "element") {
            Element e;
            ss >> e.name >> e.count;
            elements.push_back(e);
        } else if (word == "property") {
            if (elements.empty()) { err = "PLY property before element"; return false; }
            Property p;
            std::string t;
            ss >> t;
            if (t == "list") {
                std::string ct, it;
                ss >> ct >> it >> p.name;
                p.isList = true, p.countType = parse_type(ct), p.type = parse_type(it);
            } else {
                p.type = parse_type(t);
                ss >> p.name;
            }
            if (p.type == T_BAD || (p.isList && p.countType == T_BAD)) { err = "PLY: unknown property type in: " + line; return false; }
            elements.back().props.push_back(p);
        } else if (word == "end_header") done = true;
    }
    if (!sawPly || !done || format < 0) { err = std::string("not a PLY file: ") + file_name; return false; }

    mesh = PlyMesh();
    mesh.format = format;
    Reader rd{bytes.data(), (size_t)size, pos, format};
    for (const Element& el : elements) {
        bool isVertex = el.name == "vertex", isFace = el.name == "face";
        int ix = -1, iy = -1, iz = -1, ir = -1, ig = -1, ib = -1;
        for (size_t k = 0; k < el.props.size(); k++) {
            const std::string& n = el.props[k].name;
            if (n == "x") ix = (int)k; else if (n == "y") iy = (int)k; else if (n == "z") iz = (int)k;
            // red/green/blue, then diffuse_* (later entries of ReadProperties overwrite, Ply.h:394-405)
            if (n == "red" && ir < 0) ir = (int)k;
            if (n == "green" && ig < 0) ig = (int)k;
            if (n == "blue" && ib < 0) ib = (int)k;
            if (n == "diffuse_red") ir = (int)k;
            if (n == "diffuse_green") ig = (int)k;
            if (n == "diffuse_blue") ib = (int)k;
        }
        bool hasColor = isVertex && ir >= 0 && ig >= 0 && ib >= 0;
        if (isVertex) {
            if (ix < 0 || iy < 0 || iz < 0) { err = "PLY vertex element without x y z"; return false; }
            mesh.xyz.resize(3 * el.count);
            if (hasColor) mesh.rgb.resize(3 * el.count);
        }
        std::vector<double> scalar(el.props.size());
        for (size_t i = 0; i < el.count; i++) {
            for (size_t k = 0; k < el.props.size(); k++) {
                const Property& p = el.props[k];
                if (!p.isList) { scalar[k] = rd.next(p.type); continue; }
                int n = (int)rd.next(p.countType);
                if (n < 0 || !rd.ok) { err = "PLY: bad list length"; return false; }
                bool idx = isFace && (p.name == "vertex_indices" || p.name == "vertex_index");
                bool tex = isFace && p.name == "texcoord";
                if (idx) mesh.faceSize.push_back(n);
                if (tex) mesh.uvSize.push_back(n);
                for (int q = 0; q < n; q++) {
                    double v = rd.next(p.type);
                    if (idx) mesh.faceIndex.push_back((int)v);
                    if (tex) mesh.uv.push_back((float)v);
                }
            }
            if (!rd.ok) { err = std::string("PLY: unexpected end of file in ") + file_name; return false; }
            if (isVertex) {
                mesh.xyz[3 * i] = (float)scalar[ix], mesh.xyz[3 * i + 1] = (float)scalar[iy], mesh.xyz[3 * i + 2] = (float)scalar[iz];
                if (hasColor) mesh.rgb[3 * i] = (float)scalar[ir], mesh.rgb[3 * i + 1] = (float)scalar[ig], mesh.rgb[3 * i + 2] = (float)scalar[ib];
            }
        }
    }
    if (mesh.xyz.empty()) { err = std::string("PLY without vertices: ") + file_name; return false; }
    return true;
}

bool ply_write_colored_ascii(const char* file_name, const std::vector<float>& xyz, const std::vector<float>& rgb, const std::vector<int>& tri, std::string& err) {
    FILE* fp = fopen(file_name, "w");
    if (!fp) { err = std::string("Failed to open file for writing: ") + file_name; return false; }
    size_t nv = xyz.size() / 3, nt = tri.size() / 3;
    fprintf(fp, "ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n", (int)nv);
    fprintf(fp, "property uchar red\nproperty uchar green\nproperty uchar blue\nelement face %d\nproperty list uchar int vertex_indices\nend_header\n", (int)nt);
    for (size_t i = 0; i < nv; i++) {
        fprintf(fp, "%g %g %g ", (double)xyz[3 * i], (double)xyz[3 * i + 1], (double)xyz[3 * i + 2]);
        for (int c = 0; c < 3; c++) fprintf(fp, "%u ", (unsigned int)(double)rgb[3 * i + c]);
        fputc('\n', fp);
    }
    for (size_t i = 0; i < nt; i++) fprintf(fp, "3 %d %d %d \n", tri[3 * i], tri[3 * i + 1], tri[3 * i + 2]);
    fclose(fp);
    return true;
}

bool ply_write_colored_binary(const char* file_name, const std::vector<float>& xyz, const std::vector<float>& rgb, const std::vector<int>& tri, std::string& err) {
    FILE* fp = fopen(file_name, "wb");
    if (!fp) { err = std::string("Failed to open file for writing: ") + file_name; return false; }
    size_t nv = xyz.size() / 3, nt = tri.size() / 3;
    fprintf(fp, "ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n", (int)nv);
    fprintf(fp, "property uchar red\nproperty uchar green\nproperty uchar blue\nelement face %d\nproperty list uchar int vertex_indices\nend_header\n", (int)nt);
    std::vector<unsigned char> rec(15 * nv);
    for (size_t i = 0; i < nv; i++) {
        memcpy(&rec[15 * i], &xyz[3 * i], 12);
        for (int c = 0; c < 3; c++) rec[15 * i + 12 + c] = (unsigned char)(double)rgb[3 * i + c];
    }
    fwrite(rec.data(), 1, rec.size(), fp);
    rec.resize(13 * nt);
    for (size_t i = 0; i < nt; i++) {
        rec[13 * i] = 3;
        memcpy(&rec[13 * i + 1], &tri[3 * i], 12);
    }
    fwrite(rec.data(), 1, rec.size(), fp);
    fclose(fp);
    return true;
}

}  // namespace mof
