// Image file decoding / encoding for the scene boundary.
//
// The reference decodes texture files with stb_image (fredholm/src/scene.cpp:7-67:
// stbi_load(..., STBI_rgb_alpha) with a vertical flip for 8-bit textures, stbi_loadf
// without a flip for the float environment map) and writes frames with
// stbi_write_png (app/controller.cpp:291-308, app/rtcamp8.cpp:286-293).  This file
// carries its own codecs for the formats those calls meet in practice: PNG (all colour
// types / bit depths / interlacing, converted to 8-bit RGBA the way stb does), baseline and
// progressive JPEG (stb's integer IDCT, chroma up-sampling and YCbCr conversion restated so
// the texels are the ones the reference sees), Radiance .hdr (RGBE) and a PNG writer.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace fredholm
{
namespace codec
{

struct Image8 {
  int width = 0, height = 0;
  int source_channels = 0;    // channels in the file (1..4)
  std::vector<uint8_t> rgba;  // width*height*4, row 0 = TOP row of the image
};

struct ImageF {
  int width = 0, height = 0;
  std::vector<float> rgba;  // width*height*4, row 0 = TOP row
};

std::vector<uint8_t> read_file_bytes(const std::string& path);  // throws std::runtime_error

// zlib stream (RFC 1950 / 1951) -> bytes; throws std::runtime_error on a corrupt stream
std::vector<uint8_t> zlib_inflate(const uint8_t* src, size_t n, size_t size_hint = 0);
// bytes -> zlib stream (LZ77 + fixed Huffman codes; valid for any inflater)
std::vector<uint8_t> zlib_deflate(const uint8_t* src, size_t n);

bool is_png(const uint8_t* p, size_t n);
bool is_jpeg(const uint8_t* p, size_t n);
bool is_hdr(const uint8_t* p, size_t n);

Image8 decode_png(const uint8_t* p, size_t n);
Image8 decode_jpeg(const uint8_t* p, size_t n);
ImageF decode_hdr(const uint8_t* p, size_t n);

// 8-bit RGBA from a PNG or JPEG file (what stbi_load(path, .., 4) returns, top row first)
Image8 load_image8(const std::string& path);
// float RGBA (what stbi_loadf(path, .., 4) returns): .hdr decoded from RGBE, 8-bit files
// through stb's gamma-2.2 "ldr to hdr" rule
ImageF load_imagef(const std::string& path);

// 8-bit RGBA (or RGB when channels == 3) -> PNG file bytes
std::vector<uint8_t> encode_png(const uint8_t* pixels, int width, int height, int channels);
void write_png(const std::string& path, const uint8_t* pixels, int width, int height, int channels);

}  // namespace codec
}  // namespace fredholm
