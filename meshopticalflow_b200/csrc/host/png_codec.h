// PNG codec over zlib for the OpticalFlow host (no libpng in this image).
//
// Mirrors the interface of the reference's libpng wrapper
// (include/Misha/PNG.h:3-4, PNG.inl:10-128): 8-bit RGB in, 8-bit RGB out,
// rows top to bottom, tightly packed.
//
//   read : 1/2/4/8/16-bit, gray / gray+alpha / RGB / RGBA / palette, non-interlaced.
//          16-bit is stripped to 8 (PNG_TRANSFORM_STRIP_16), sub-byte samples are
//          expanded one per byte (PNG_TRANSFORM_PACKING). RGBA drops alpha
//          (PNG.inl:65-73 overwrites it with the next pixel's red). Palette
//          entries are expanded to RGB (PNG.inl:66-72). Gray images fill only the
//          first channel in the reference (the rest is uninitialised there); here
//          the missing channels are zero.
//   write: RGB8, filter 0, one IDAT.
#ifndef MOF_PNG_CODEC_H
#define MOF_PNG_CODEC_H

#include <string>
#include <vector>

namespace mof {

// Returns false and fills err on failure.
bool png_read_rgb8(const char* file_name, std::vector<unsigned char>& rgb, int& width, int& height, std::string& err);
bool png_write_rgb8(const char* file_name, const unsigned char* rgb, int width, int height, std::string& err);

}  // namespace mof

#endif
