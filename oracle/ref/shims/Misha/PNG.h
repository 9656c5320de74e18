// Stand-in for the reference's libpng wrapper (include/Misha/PNG.h:3-4): same two functions,
// backed by the repo's zlib PNG codec because libpng is not installed. Defining PNG_INCLUDED
// also turns the reference's own PNG.h (pulled in by Misha/Image.h:6) into a no-op.
// Test infrastructure only.
#ifndef PNG_INCLUDED
#define PNG_INCLUDED
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "png_codec.h"
inline void PNGWriteColor(const char* fileName, const unsigned char* pixels, int width, int height)
{
	std::string err;
	if (!mof::png_write_rgb8(fileName, pixels, width, height, err)) fprintf(stderr, "[ERROR] %s\n", err.c_str()), exit(0);
}
inline unsigned char* PNGReadColor(const char* fileName, int& width, int& height)
{
	std::vector<unsigned char> rgb;
	std::string err;
	if (!mof::png_read_rgb8(fileName, rgb, width, height, err)) fprintf(stderr, "[ERROR] %s\n", err.c_str()), exit(0);
	unsigned char* pixels = new unsigned char[rgb.size() + 4];  // +4: the reference writes alpha one byte past the end (PNG.inl:65-73)
	memcpy(pixels, rgb.data(), rgb.size());
	return pixels;
}
#endif
