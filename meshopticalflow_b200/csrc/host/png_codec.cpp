// See png_codec.h. PNG container per ISO/IEC 15948; compression through zlib.
#include "png_codec.h"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace mof {
namespace {

const unsigned char kSignature[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};

unsigned int be32(const unsigned char* p) {
    return ((unsigned int)p[0] << 24) | ((unsigned int)p[1] << 16) | ((unsigned int)p[2] << 8) | (unsigned int)p[3];
}
void put_be32(unsigned char* p, unsigned int v) {
    p[0] = (unsigned char)(v >> 24), p[1] = (unsigned char)(v >> 16), p[2] = (unsigned char)(v >> 8), p[3] = (unsigned char)v;
}

int paeth(int a, int b, int c) {
    int p = a + b - c;
    int pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    if (pb <= pc) return b;
    return c;
}

bool read_file(const char* file_name, std::vector<unsigned char>& bytes) {
    FILE* fp = fopen(file_name, "rb");
    if (!fp) return false;
    fseek(fp, 0, SEEK_END);
    long size = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    bytes.resize(size > 0 ? (size_t)size : 0);
    size_t got = bytes.empty() ? 0 : fread(bytes.data(), 1, bytes.size(), fp);
    fclose(fp);
    return got == bytes.size();
}

void write_chunk(FILE* fp, const char type[4], const unsigned char* data, size_t size) {
    unsigned char head[8];
    put_be32(head, (unsigned int)size);
    memcpy(head + 4, type, 4);
    fwrite(head, 1, 8, fp);
    if (size) fwrite(data, 1, size, fp);
    uLong crc = crc32(0L, Z_NULL, 0);
    crc = crc32(crc, (const Bytef*)type, 4);
    if (size) crc = crc32(crc, data, (uInt)size);
    unsigned char tail[4];
    put_be32(tail, (unsigned int)crc);
    fwrite(tail, 1, 4, fp);
}

}  // namespace

bool png_read_rgb8(const char* file_name, std::vector<unsigned char>& rgb, int& width, int& height, std::string& err) {
    std::vector<unsigned char> file;
    if (!read_file(file_name, file)) { err = std::string("Failed to open file for reading: ") + file_name; return false; }
    if (file.size() < 8 || memcmp(file.data(), kSignature, 8)) { err = "not a PNG file"; return false; }

    int bit_depth = 0, color_type = 0, interlace = 0;
    std::vector<unsigned char> idat, palette;
    bool have_header = false;
    for (size_t pos = 8; pos + 12 <= file.size();) {
        unsigned int size = be32(&file[pos]);
        const unsigned char* type = &file[pos + 4];
        const unsigned char* data = &file[pos + 8];
        if (pos + 12 + (size_t)size > file.size()) { err = "truncated PNG chunk"; return false; }
        if (!memcmp(type, "IHDR", 4)) {
            if (size < 13) { err = "bad IHDR"; return false; }
            width = (int)be32(data), height = (int)be32(data + 4);
            bit_depth = data[8], color_type = data[9], interlace = data[12];
            have_header = true;
        } else if (!memcmp(type, "PLTE", 4)) palette.assign(data, data + size);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + size);
        else if (!memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)size;
    }
    if (!have_header || width <= 0 || height <= 0) { err = "missing or empty IHDR"; return false; }
    if (interlace) { err = "interlaced PNG is not supported"; return false; }
    int channels;
    switch (color_type) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: err = "bad PNG colour type"; return false;
    }
    if (bit_depth != 1 && bit_depth != 2 && bit_depth != 4 && bit_depth != 8 && bit_depth != 16) { err = "bad PNG bit depth"; return false; }

    size_t bits_per_pixel = (size_t)channels * bit_depth;
    size_t stride = ((size_t)width * bits_per_pixel + 7) / 8;
    size_t bpp = (bits_per_pixel + 7) / 8;  // filter distance in bytes
    std::vector<unsigned char> raw((stride + 1) * (size_t)height);
    uLongf raw_size = (uLongf)raw.size();
    int z = uncompress(raw.data(), &raw_size, idat.data(), (uLong)idat.size());
    if (z != Z_OK || raw_size != raw.size()) { err = "failed to inflate PNG image data"; return false; }

    // Undo the scan-line filters in place.
    std::vector<unsigned char> zero(stride, 0);
    for (int y = 0; y < height; y++) {
        unsigned char* line = &raw[(stride + 1) * (size_t)y];
        int filter = line[0];
        unsigned char* cur = line + 1;
        const unsigned char* up = y ? cur - (stride + 1) : zero.data();
        for (size_t i = 0; i < stride; i++) {
            int a = i >= bpp ? cur[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0;
            switch (filter) {
                case 0: break;
                case 1: cur[i] = (unsigned char)(cur[i] + a); break;
                case 2: cur[i] = (unsigned char)(cur[i] + b); break;
                case 3: cur[i] = (unsigned char)(cur[i] + ((a + b) >> 1)); break;
                case 4: cur[i] = (unsigned char)(cur[i] + paeth(a, b, c)); break;
                default: err = "bad PNG filter"; return false;
            }
        }
    }

    rgb.assign((size_t)width * height * 3, 0);
    for (int y = 0; y < height; y++) {
        const unsigned char* cur = &raw[(stride + 1) * (size_t)y + 1];
        for (int x = 0; x < width; x++) {
            unsigned char sample[4] = {0, 0, 0, 0};
            for (int c = 0; c < channels; c++) {
                size_t idx = (size_t)x * channels + c;
                if (bit_depth == 8) sample[c] = cur[idx];
                else if (bit_depth == 16) sample[c] = cur[2 * idx];  // strip to the high byte
                else {
                    size_t bit = idx * bit_depth;
                    sample[c] = (unsigned char)((cur[bit >> 3] >> (8 - bit_depth - (bit & 7))) & ((1 << bit_depth) - 1));
                }
            }
            unsigned char* out = &rgb[((size_t)y * width + x) * 3];
            if (color_type == 3) {
                size_t p = (size_t)sample[0] * 3;
                if (p + 2 < palette.size()) out[0] = palette[p], out[1] = palette[p + 1], out[2] = palette[p + 2];
            } else
                for (int c = 0; c < channels && c < 3; c++) out[c] = sample[c];
        }
    }
    return true;
}

bool png_write_rgb8(const char* file_name, const unsigned char* rgb, int width, int height, std::string& err) {
    FILE* fp = fopen(file_name, "wb");
    if (!fp) { err = std::string("Failed to open file for writing: ") + file_name; return false; }
    fwrite(kSignature, 1, 8, fp);
    unsigned char ihdr[13];
    put_be32(ihdr, (unsigned int)width), put_be32(ihdr + 4, (unsigned int)height);
    ihdr[8] = 8, ihdr[9] = 2, ihdr[10] = 0, ihdr[11] = 0, ihdr[12] = 0;
    write_chunk(fp, "IHDR", ihdr, 13);

    size_t stride = (size_t)width * 3;
    std::vector<unsigned char> raw((stride + 1) * (size_t)height);
    for (int y = 0; y < height; y++) {
        raw[(stride + 1) * (size_t)y] = 0;
        memcpy(&raw[(stride + 1) * (size_t)y + 1], rgb + stride * (size_t)y, stride);
    }
    uLongf packed_size = compressBound((uLong)raw.size());
    std::vector<unsigned char> packed(packed_size);
    if (compress2(packed.data(), &packed_size, raw.data(), (uLong)raw.size(), Z_DEFAULT_COMPRESSION) != Z_OK) {
        fclose(fp);
        err = "failed to deflate PNG image data";
        return false;
    }
    write_chunk(fp, "IDAT", packed.data(), packed_size);
    write_chunk(fp, "IEND", NULL, 0);
    fclose(fp);
    return true;
}

}  // namespace mof
