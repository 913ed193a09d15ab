// Image codecs of the scene boundary (see image_codec.h).
//
// PNG follows the PNG specification (RFC 2083) + DEFLATE (RFC 1951); the conversion to
// 8-bit RGBA is the one stb_image's stbi_load(.., 4) applies (reference call site
// fredholm/src/scene.cpp:15-16): low bit depths are scaled to the full 8-bit range,
// 16-bit samples keep their high byte, tRNS keys and palette alpha become the alpha
// channel, grey is replicated to RGB.
//
// JPEG follows ITU T.81 for the entropy-coded data; the lossy back end (where decoders
// legitimately differ) restates stb_image's choices so texels equal the reference's:
// a 12-bit fixed-point Loeffler/Ligtenberg/Moschytz IDCT with 2 guard bits after the
// column pass, triangle-filter chroma up-sampling for 2x1 / 1x2 / 2x2 (nearest otherwise),
// and a 20-bit fixed-point YCbCr -> RGB conversion whose green Cb term is truncated to
// its upper 16 bits.
//
// Radiance .hdr: RGBE pixels, new-style RLE scanlines; value = mantissa * 2^(e-136).
#include "image_codec.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace fredholm
{
namespace codec
{

namespace
{
[[noreturn]] void fail(const std::string& what) { throw std::runtime_error(what); }
}  // namespace

std::vector<uint8_t> read_file_bytes(const std::string& path)
{
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) fail("failed to load " + path);
  std::vector<uint8_t> data;
  uint8_t chunk[1 << 16];
  size_t got;
  while ((got = std::fread(chunk, 1, sizeof(chunk), f)) > 0) data.insert(data.end(), chunk, chunk + got);
  std::fclose(f);
  return data;
}

// =========================================================================================
// DEFLATE
// =========================================================================================
namespace
{

struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint64_t acc = 0;
  int n = 0;
  void refill()
  {
    while (n <= 56 && p < end) {
      acc |= (uint64_t)(*p++) << n;
      n += 8;
    }
  }
  uint32_t peek(int k)
  {
    if (n < k) refill();
    return (uint32_t)(acc & ((1ull << k) - 1ull));
  }
  void drop(int k)
  {
    if (n < k) fail("deflate: unexpected end of stream");
    acc >>= k;
    n -= k;
  }
  uint32_t take(int k)
  {
    if (k == 0) return 0;
    const uint32_t v = peek(k);
    drop(k);
    return v;
  }
  void align_byte()
  {
    const int r = n & 7;
    acc >>= r;
    n -= r;
  }
};

// canonical Huffman decoder: symbols sorted by (length, value), decoded length by length
struct HuffTable {
  uint16_t count[16];
  uint16_t symbol[320];
  void build(const uint8_t* lengths, int n_sym)
  {
    std::memset(count, 0, sizeof(count));
    for (int i = 0; i < n_sym; ++i) count[lengths[i]]++;
    count[0] = 0;
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + count[l];
    for (int i = 0; i < n_sym; ++i)
      if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
  }
  int decode(BitReader& br) const
  {
    uint32_t bits = br.peek(15);
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; ++len) {
      code |= (int)(bits & 1u);
      bits >>= 1;
      const int c = count[len];
      if (code - c < first) {
        br.drop(len);
        return symbol[index + (code - first)];
      }
      index += c;
      first = (first + c) << 1;
      code <<= 1;
    }
    fail("deflate: invalid Huffman code");
  }
};

const uint16_t kLenBase[29] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23, 27,
                               31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,   33,   49,   65,    97,    129,
                                193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

void inflate_block(BitReader& br, const HuffTable& lit, const HuffTable& dist, std::vector<uint8_t>& out)
{
  for (;;) {
    const int sym = lit.decode(br);
    if (sym < 256) {
      out.push_back((uint8_t)sym);
    } else if (sym == 256) {
      return;
    } else {
      const int li = sym - 257;
      if (li >= 29) fail("deflate: invalid length symbol");
      const int len = kLenBase[li] + (int)br.take(kLenExtra[li]);
      const int ds = dist.decode(br);
      if (ds >= 30) fail("deflate: invalid distance symbol");
      const size_t d = kDistBase[ds] + br.take(kDistExtra[ds]);
      if (d > out.size()) fail("deflate: distance too far back");
      const size_t start = out.size() - d;
      out.resize(out.size() + len);
      uint8_t* dst = out.data() + out.size() - len;
      const uint8_t* src = out.data() + start;
      for (int i = 0; i < len; ++i) dst[i] = src[i];  // may overlap: byte-wise forward copy
    }
  }
}

uint32_t adler32(const uint8_t* p, size_t n)
{
  uint32_t a = 1, b = 0;
  while (n > 0) {
    const size_t k = std::min<size_t>(n, 5552);
    for (size_t i = 0; i < k; ++i) {
      a += p[i];
      b += a;
    }
    a %= 65521u;
    b %= 65521u;
    p += k;
    n -= k;
  }
  return (b << 16) | a;
}

uint32_t crc32(const uint8_t* p, size_t n, uint32_t crc = 0)
{
  static uint32_t table[256];
  static bool ready = false;
  if (!ready) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
    ready = true;
  }
  crc = ~crc;
  for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
  return ~crc;
}

}  // namespace

std::vector<uint8_t> zlib_inflate(const uint8_t* src, size_t n, size_t size_hint)
{
  if (n < 2) fail("zlib: stream too short");
  const int cmf = src[0], flg = src[1];
  if ((cmf & 15) != 8 || ((cmf << 8) | flg) % 31 != 0) fail("zlib: bad header");
  if (flg & 32) fail("zlib: preset dictionary not supported");
  BitReader br{src + 2, src + n};
  std::vector<uint8_t> out;
  out.reserve(size_hint ? size_hint : n * 4);
  bool last = false;
  while (!last) {
    last = br.take(1) != 0;
    const uint32_t type = br.take(2);
    if (type == 0) {
      br.align_byte();
      const uint32_t len = br.take(16), nlen = br.take(16);
      if ((len ^ 0xffffu) != nlen) fail("deflate: stored block length mismatch");
      for (uint32_t i = 0; i < len; ++i) out.push_back((uint8_t)br.take(8));
    } else if (type == 1) {
      uint8_t l[288];
      for (int i = 0; i < 144; ++i) l[i] = 8;
      for (int i = 144; i < 256; ++i) l[i] = 9;
      for (int i = 256; i < 280; ++i) l[i] = 7;
      for (int i = 280; i < 288; ++i) l[i] = 8;
      uint8_t d[30];
      for (int i = 0; i < 30; ++i) d[i] = 5;
      HuffTable lit, dist;
      lit.build(l, 288);
      dist.build(d, 30);
      inflate_block(br, lit, dist, out);
    } else if (type == 2) {
      const int hlit = (int)br.take(5) + 257, hdist = (int)br.take(5) + 1, hclen = (int)br.take(4) + 4;
      static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      uint8_t cl[19] = {0};
      for (int i = 0; i < hclen; ++i) cl[order[i]] = (uint8_t)br.take(3);
      HuffTable clt;
      clt.build(cl, 19);
      uint8_t lengths[288 + 32];
      int i = 0;
      while (i < hlit + hdist) {
        const int s = clt.decode(br);
        if (s < 16) {
          lengths[i++] = (uint8_t)s;
        } else {
          int rep, val = 0;
          if (s == 16) {
            if (i == 0) fail("deflate: repeat without previous length");
            val = lengths[i - 1];
            rep = 3 + (int)br.take(2);
          } else if (s == 17) {
            rep = 3 + (int)br.take(3);
          } else {
            rep = 11 + (int)br.take(7);
          }
          if (i + rep > hlit + hdist) fail("deflate: too many code lengths");
          while (rep--) lengths[i++] = (uint8_t)val;
        }
      }
      HuffTable lit, dist;
      lit.build(lengths, hlit);
      dist.build(lengths + hlit, hdist);
      inflate_block(br, lit, dist, out);
    } else {
      fail("deflate: invalid block type");
    }
  }
  return out;
}

// LZ77 with a hash chain + fixed Huffman codes (one block).
std::vector<uint8_t> zlib_deflate(const uint8_t* src, size_t n)
{
  std::vector<uint8_t> out;
  out.reserve(n / 2 + 64);
  out.push_back(0x78);
  out.push_back(0x01);
  uint64_t acc = 0;
  int nbits = 0;
  auto put = [&](uint32_t v, int k) {  // LSB first
    acc |= (uint64_t)v << nbits;
    nbits += k;
    while (nbits >= 8) {
      out.push_back((uint8_t)acc);
      acc >>= 8;
      nbits -= 8;
    }
  };
  auto put_code = [&](uint32_t code, int k) {  // Huffman codes are sent MSB first
    uint32_t r = 0;
    for (int i = 0; i < k; ++i) r |= ((code >> i) & 1u) << (k - 1 - i);
    put(r, k);
  };
  auto put_lit = [&](int s) {
    if (s < 144)
      put_code(0x30 + s, 8);
    else if (s < 256)
      put_code(0x190 + (s - 144), 9);
    else if (s < 280)
      put_code(s - 256, 7);
    else
      put_code(0xC0 + (s - 280), 8);
  };
  put(1, 1);  // final block
  put(1, 2);  // fixed Huffman
  constexpr int kHashBits = 15, kWindow = 32768, kMaxChain = 32;
  std::vector<int32_t> head(1 << kHashBits, -1), prev(n ? n : 1, -1);
  auto hash3 = [&](size_t i) {
    const uint32_t v = src[i] | (src[i + 1] << 8) | (src[i + 2] << 16);
    return (v * 2654435761u) >> (32 - kHashBits);
  };
  size_t i = 0;
  while (i < n) {
    int best_len = 0, best_dist = 0;
    if (i + 3 <= n) {
      const uint32_t h = hash3(i);
      int32_t cand = head[h];
      int chain = 0;
      const int max_len = (int)std::min<size_t>(258, n - i);
      while (cand >= 0 && (int)(i - cand) <= kWindow && chain++ < kMaxChain) {
        int l = 0;
        while (l < max_len && src[cand + l] == src[i + l]) ++l;
        if (l > best_len) {
          best_len = l;
          best_dist = (int)(i - cand);
          if (l == max_len) break;
        }
        cand = prev[cand];
      }
      prev[i] = head[h];
      head[h] = (int32_t)i;
    }
    if (best_len >= 3) {
      int li = 28;
      while (kLenBase[li] > best_len) --li;
      put_lit(257 + li);
      put(best_len - kLenBase[li], kLenExtra[li]);
      int di = 29;
      while (kDistBase[di] > best_dist) --di;
      put_code(di, 5);
      put(best_dist - kDistBase[di], kDistExtra[di]);
      for (int k = 1; k < best_len; ++k) {
        const size_t j = i + k;
        if (j + 3 <= n) {
          const uint32_t h = hash3(j);
          prev[j] = head[h];
          head[h] = (int32_t)j;
        }
      }
      i += best_len;
    } else {
      put_lit(src[i]);
      ++i;
    }
  }
  put_lit(256);
  if (nbits) put(0, 8 - nbits);
  const uint32_t ad = adler32(src, n);
  out.push_back((uint8_t)(ad >> 24));
  out.push_back((uint8_t)(ad >> 16));
  out.push_back((uint8_t)(ad >> 8));
  out.push_back((uint8_t)ad);
  return out;
}

// =========================================================================================
// PNG
// =========================================================================================
namespace
{
const uint8_t kPngSig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int paeth(int a, int b, int c)
{
  const int p = a + b - c;
  const int pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  if (pa <= pb && pa <= pc) return a;
  if (pb <= pc) return b;
  return c;
}

// reverses the per-scanline filters of one (sub-)image in place; `raw` holds h rows of
// (1 + stride) bytes; returns the rows without their filter byte
void unfilter(uint8_t* raw, size_t stride, int h, int bpp, std::vector<uint8_t>& rows)
{
  rows.resize(stride * (size_t)h);
  std::vector<uint8_t> zero(stride, 0);
  for (int y = 0; y < h; ++y) {
    const uint8_t ft = raw[(stride + 1) * (size_t)y];
    const uint8_t* in = raw + (stride + 1) * (size_t)y + 1;
    uint8_t* cur = rows.data() + stride * (size_t)y;
    const uint8_t* up = y ? cur - stride : zero.data();
    for (size_t i = 0; i < stride; ++i) {
      const int a = i >= (size_t)bpp ? cur[i - bpp] : 0;
      const int b = up[i];
      const int c = i >= (size_t)bpp ? up[i - bpp] : 0;
      int v = in[i];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: fail("png: invalid filter type");
      }
      cur[i] = (uint8_t)v;
    }
  }
}
}  // namespace

bool is_png(const uint8_t* p, size_t n) { return n >= 8 && std::memcmp(p, kPngSig, 8) == 0; }

Image8 decode_png(const uint8_t* p, size_t n)
{
  if (!is_png(p, n)) fail("png: bad signature");
  size_t pos = 8;
  uint32_t W = 0, H = 0;
  int depth = 0, ctype = 0, interlace = 0;
  bool have_ihdr = false, have_trns = false;
  uint8_t palette[256][4];
  int n_pal = 0;
  uint16_t key[3] = {0, 0, 0};
  std::vector<uint8_t> idat;
  for (bool done = false; !done;) {
    if (pos + 12 > n) fail("png: truncated file");
    const uint32_t len = be32(p + pos);
    const uint8_t* tag = p + pos + 4;
    const uint8_t* body = p + pos + 8;
    if (pos + 12 + (size_t)len > n) fail("png: truncated chunk");
    auto is = [&](const char* t) { return std::memcmp(tag, t, 4) == 0; };
    if (is("IHDR")) {
      if (len != 13) fail("png: bad IHDR");
      W = be32(body);
      H = be32(body + 4);
      depth = body[8];
      ctype = body[9];
      interlace = body[12];
      if (W == 0 || H == 0 || W > (1u << 24) || H > (1u << 24)) fail("png: bad image size");
      if (body[10] != 0 || body[11] != 0 || interlace > 1) fail("png: unsupported compression / filter / interlace");
      const bool ok = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
                      (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                      ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
      if (!ok) fail("png: bad colour type / bit depth");
      have_ihdr = true;
    } else if (is("PLTE")) {
      if (len % 3 || len > 768) fail("png: bad PLTE");
      n_pal = (int)(len / 3);
      for (int i = 0; i < n_pal; ++i) {
        palette[i][0] = body[3 * i];
        palette[i][1] = body[3 * i + 1];
        palette[i][2] = body[3 * i + 2];
        palette[i][3] = 255;
      }
    } else if (is("tRNS")) {
      if (!have_ihdr) fail("png: tRNS before IHDR");
      if (ctype == 3) {
        if ((int)len > n_pal) fail("png: bad tRNS");
        for (uint32_t i = 0; i < len; ++i) palette[i][3] = body[i];
      } else if (ctype == 0 && len >= 2) {
        key[0] = (uint16_t)((body[0] << 8) | body[1]);
        have_trns = true;
      } else if (ctype == 2 && len >= 6) {
        for (int k = 0; k < 3; ++k) key[k] = (uint16_t)((body[2 * k] << 8) | body[2 * k + 1]);
        have_trns = true;
      }
    } else if (is("IDAT")) {
      idat.insert(idat.end(), body, body + len);
    } else if (is("IEND")) {
      done = true;
    } else if (!(tag[0] & 32)) {
      fail("png: unknown critical chunk");
    }
    pos += 12 + (size_t)len;
  }
  if (!have_ihdr || idat.empty()) fail("png: missing IHDR / IDAT");
  if (ctype == 3 && n_pal == 0) fail("png: palette image without PLTE");

  const int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : 4;
  const int bits_pp = channels * depth;
  const int bpp = std::max(1, bits_pp / 8);
  auto row_bytes = [&](uint32_t w) { return ((size_t)w * bits_pp + 7) / 8; };

  size_t expect = 0;
  if (!interlace) {
    expect = (row_bytes(W) + 1) * (size_t)H;
  } else {
    static const int xs[7] = {0, 4, 0, 2, 0, 1, 0}, ys[7] = {0, 0, 4, 0, 2, 0, 1}, dx[7] = {8, 8, 4, 4, 2, 2, 1},
                     dy[7] = {8, 8, 8, 4, 4, 2, 2};
    for (int k = 0; k < 7; ++k) {
      const uint32_t w = (W - xs[k] + dx[k] - 1) / dx[k], h = (H - ys[k] + dy[k] - 1) / dy[k];
      if (w && h) expect += (row_bytes(w) + 1) * (size_t)h;
    }
  }
  std::vector<uint8_t> raw = zlib_inflate(idat.data(), idat.size(), expect);
  if (raw.size() < expect) fail("png: not enough pixel data");

  Image8 img;
  img.width = (int)W;
  img.height = (int)H;
  img.rgba.assign((size_t)W * H * 4, 0);
  const int maxv = (1 << depth) - 1;
  const int scale8 = depth < 8 ? 255 / maxv : 1;  // 1-bit 0xff, 2-bit 0x55, 4-bit 0x11

  // sample c of pixel x in an unfiltered row, at the file's bit depth
  auto sample = [&](const uint8_t* row, uint32_t x, int c) -> int {
    if (depth == 8) return row[(size_t)x * channels + c];
    if (depth == 16) {
      const uint8_t* q = row + ((size_t)x * channels + c) * 2;
      return (q[0] << 8) | q[1];
    }
    const size_t bit = (size_t)x * depth;  // depth < 8 has one channel
    return (row[bit >> 3] >> (8 - depth - (bit & 7))) & maxv;
  };
  auto put_pixel = [&](const uint8_t* row, uint32_t sx, uint32_t ox, uint32_t oy) {
    uint8_t* o = img.rgba.data() + ((size_t)oy * W + ox) * 4;
    auto to8 = [&](int v) -> uint8_t { return depth == 16 ? (uint8_t)(v >> 8) : depth < 8 ? (uint8_t)(v * scale8) : (uint8_t)v; };
    if (ctype == 3) {
      const int idx = sample(row, sx, 0);
      // indices past the palette decode to opaque black like an all-zero entry
      if (idx < n_pal) {
        o[0] = palette[idx][0];
        o[1] = palette[idx][1];
        o[2] = palette[idx][2];
        o[3] = palette[idx][3];
      } else {
        o[0] = o[1] = o[2] = 0;
        o[3] = 255;
      }
    } else if (ctype == 0) {
      const int g = sample(row, sx, 0);
      o[0] = o[1] = o[2] = to8(g);
      o[3] = (have_trns && g == key[0]) ? 0 : 255;
    } else if (ctype == 4) {
      o[0] = o[1] = o[2] = to8(sample(row, sx, 0));
      o[3] = to8(sample(row, sx, 1));
    } else if (ctype == 2) {
      const int r = sample(row, sx, 0), g = sample(row, sx, 1), b = sample(row, sx, 2);
      o[0] = to8(r);
      o[1] = to8(g);
      o[2] = to8(b);
      o[3] = (have_trns && r == key[0] && g == key[1] && b == key[2]) ? 0 : 255;
    } else {
      for (int c = 0; c < 4; ++c) o[c] = to8(sample(row, sx, c));
    }
  };

  std::vector<uint8_t> rows;
  if (!interlace) {
    const size_t stride = row_bytes(W);
    unfilter(raw.data(), stride, (int)H, bpp, rows);
    for (uint32_t y = 0; y < H; ++y)
      for (uint32_t x = 0; x < W; ++x) put_pixel(rows.data() + stride * y, x, x, y);
  } else {
    static const int xs[7] = {0, 4, 0, 2, 0, 1, 0}, ys[7] = {0, 0, 4, 0, 2, 0, 1}, dx[7] = {8, 8, 4, 4, 2, 2, 1},
                     dy[7] = {8, 8, 8, 4, 4, 2, 2};
    size_t off = 0;
    for (int k = 0; k < 7; ++k) {
      const uint32_t w = (W - xs[k] + dx[k] - 1) / dx[k], h = (H - ys[k] + dy[k] - 1) / dy[k];
      if (!w || !h) continue;
      const size_t stride = row_bytes(w);
      unfilter(raw.data() + off, stride, (int)h, bpp, rows);
      for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x) put_pixel(rows.data() + stride * y, x, xs[k] + x * dx[k], ys[k] + y * dy[k]);
      off += (stride + 1) * (size_t)h;
    }
  }
  img.source_channels = (ctype == 3) ? (have_trns ? 4 : 3) : channels + ((have_trns && ctype != 3) ? 1 : 0);
  return img;
}

std::vector<uint8_t> encode_png(const uint8_t* pixels, int width, int height, int channels)
{
  if (width <= 0 || height <= 0 || (channels != 3 && channels != 4)) fail("png: bad image for encoding");
  const size_t stride = (size_t)width * channels;
  std::vector<uint8_t> raw((stride + 1) * (size_t)height);
  // filter per row: try None / Sub / Up / Paeth, keep the one with the smallest sum of |residual|
  std::vector<uint8_t> cand(stride);
  for (int y = 0; y < height; ++y) {
    const uint8_t* cur = pixels + stride * (size_t)y;
    const uint8_t* up = y ? cur - stride : nullptr;
    uint8_t* dst = raw.data() + (stride + 1) * (size_t)y;
    long best_score = -1;
    for (int ft : {0, 1, 2, 4}) {
      long score = 0;
      for (size_t i = 0; i < stride; ++i) {
        const int a = i >= (size_t)channels ? cur[i - channels] : 0;
        const int b = up ? up[i] : 0;
        const int c = (up && i >= (size_t)channels) ? up[i - channels] : 0;
        const int pred = ft == 0 ? 0 : ft == 1 ? a : ft == 2 ? b : paeth(a, b, c);
        cand[i] = (uint8_t)(cur[i] - pred);
        score += std::abs((int)(int8_t)cand[i]);
      }
      if (best_score < 0 || score < best_score) {
        best_score = score;
        dst[0] = (uint8_t)ft;
        std::memcpy(dst + 1, cand.data(), stride);
      }
    }
  }
  const std::vector<uint8_t> z = zlib_deflate(raw.data(), raw.size());
  std::vector<uint8_t> out(kPngSig, kPngSig + 8);
  auto chunk = [&](const char* tag, const uint8_t* body, size_t len) {
    const size_t at = out.size();
    out.resize(at + 12 + len);
    uint8_t* q = out.data() + at;
    q[0] = (uint8_t)(len >> 24), q[1] = (uint8_t)(len >> 16), q[2] = (uint8_t)(len >> 8), q[3] = (uint8_t)len;
    std::memcpy(q + 4, tag, 4);
    if (len) std::memcpy(q + 8, body, len);
    const uint32_t c = crc32(q + 4, len + 4);
    q[8 + len] = (uint8_t)(c >> 24), q[9 + len] = (uint8_t)(c >> 16), q[10 + len] = (uint8_t)(c >> 8), q[11 + len] = (uint8_t)c;
  };
  uint8_t ihdr[13] = {(uint8_t)(width >> 24),  (uint8_t)(width >> 16),  (uint8_t)(width >> 8),  (uint8_t)width,
                      (uint8_t)(height >> 24), (uint8_t)(height >> 16), (uint8_t)(height >> 8), (uint8_t)height,
                      8, (uint8_t)(channels == 4 ? 6 : 2), 0, 0, 0};
  chunk("IHDR", ihdr, 13);
  chunk("IDAT", z.data(), z.size());
  chunk("IEND", nullptr, 0);
  return out;
}

void write_png(const std::string& path, const uint8_t* pixels, int width, int height, int channels)
{
  const std::vector<uint8_t> bytes = encode_png(pixels, width, height, channels);
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) fail("failed to write " + path);
  const size_t n = std::fwrite(bytes.data(), 1, bytes.size(), f);
  std::fclose(f);
  if (n != bytes.size()) fail("failed to write " + path);
}

// =========================================================================================
// JPEG
// =========================================================================================
namespace
{

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct JHuff {
  // canonical JPEG Huffman table: codes of length l occupy [mincode[l], maxcode[l])
  int mincode[18], maxcode[18], first_index[18];
  uint8_t values[256];
  int n_values = 0;
  bool defined = false;
  void build(const uint8_t counts[16], const uint8_t* vals, int n)
  {
    n_values = n;
    std::memcpy(values, vals, n);
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
      first_index[l] = k;
      mincode[l] = code;
      code += counts[l - 1];
      k += counts[l - 1];
      maxcode[l] = code;
      if (code > (1 << l)) fail("jpeg: bad Huffman code lengths");
      code <<= 1;
    }
    defined = true;
  }
};

struct JComponent {
  int id = 0, h = 1, v = 1, tq = 0, hd = 0, ha = 0;
  int dc_pred = 0;
  int x = 0, y = 0;    // size in samples
  int bw = 0, bh = 0;  // size in blocks, padded to whole MCUs
  std::vector<uint8_t> plane;   // bw*8 x bh*8 samples
  std::vector<int16_t> coeff;   // progressive only: bw*bh*64
};

struct JpegDecoder {
  const uint8_t* p;
  const uint8_t* end;
  // bit reader state (MSB first, byte-stuffed)
  uint32_t bitbuf = 0;
  int bitcnt = 0;
  int marker = -1;  // marker met inside entropy-coded data
  bool nomore = false;

  int width = 0, height = 0, ncomp = 0;
  bool progressive = false;
  int hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
  JComponent comp[4];
  uint16_t dequant[4][64];
  JHuff hdc[4], hac[4];
  int restart_interval = 0, todo = 0;
  bool jfif = false;
  int app14_transform = -1;
  int rgb_ids = 0;
  // scan parameters
  int scan_n = 0, order[4];
  int spec_start = 0, spec_end = 63, succ_high = 0, succ_low = 0, eob_run = 0;

  int get8()
  {
    if (p >= end) return 0;
    return *p++;
  }
  int get16()
  {
    const int a = get8();
    return (a << 8) | get8();
  }

  void fill_bits()
  {
    while (bitcnt <= 24) {
      int b = nomore ? 0 : get8();
      if (b == 0xff && !nomore) {
        int c = get8();
        while (c == 0xff) c = get8();
        if (c != 0) {
          marker = c;
          nomore = true;
          b = 0;
        }
      }
      bitbuf |= (uint32_t)b << (24 - bitcnt);
      bitcnt += 8;
    }
  }
  int get_bits(int k)
  {
    if (k == 0) return 0;
    if (bitcnt < k) fill_bits();
    const int v = (int)(bitbuf >> (32 - k));
    bitbuf <<= k;
    bitcnt -= k;
    return v;
  }
  int get_bit() { return get_bits(1); }
  int decode_huff(const JHuff& h)
  {
    if (!h.defined) fail("jpeg: missing Huffman table");
    if (bitcnt < 16) fill_bits();
    for (int l = 1; l <= 16; ++l) {
      const int code = (int)(bitbuf >> (32 - l));
      if (code < h.maxcode[l] && code >= h.mincode[l]) {
        bitbuf <<= l;
        bitcnt -= l;
        const int idx = h.first_index[l] + code - h.mincode[l];
        if (idx >= h.n_values) fail("jpeg: bad Huffman code");
        return h.values[idx];
      }
    }
    fail("jpeg: bad Huffman code");
  }
  // n-bit magnitude with JPEG's sign convention
  int receive_extend(int n)
  {
    if (n == 0) return 0;
    const int v = get_bits(n);
    return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
  }
  void reset_entropy()
  {
    bitbuf = 0;
    bitcnt = 0;
    nomore = false;
    marker = -1;
    for (int i = 0; i < 4; ++i) comp[i].dc_pred = 0;
    todo = restart_interval ? restart_interval : 0x7fffffff;
    eob_run = 0;
  }

  // ---- block decoders -------------------------------------------------------------------
  void block_baseline(int16_t* d, JComponent& c)
  {
    std::memset(d, 0, 64 * sizeof(int16_t));
    const uint16_t* q = dequant[c.tq];
    const int t = decode_huff(hdc[c.hd]);
    if (t > 15) fail("jpeg: bad DC size");
    c.dc_pred += receive_extend(t);
    d[0] = (int16_t)(c.dc_pred * q[0]);
    for (int k = 1; k < 64;) {
      const int rs = decode_huff(hac[c.ha]);
      const int r = rs >> 4, s = rs & 15;
      if (s == 0) {
        if (rs != 0xf0) break;
        k += 16;
      } else {
        k += r;
        if (k > 63) fail("jpeg: bad AC run");
        const int z = kZigzag[k++];
        d[z] = (int16_t)(receive_extend(s) * q[z]);
      }
    }
  }
  void block_prog_dc(int16_t* d, JComponent& c)
  {
    if (spec_end != 0) fail("jpeg: DC and AC coefficients merged in a progressive scan");
    if (succ_high == 0) {
      std::memset(d, 0, 64 * sizeof(int16_t));
      const int t = decode_huff(hdc[c.hd]);
      if (t > 15) fail("jpeg: bad DC size");
      c.dc_pred += receive_extend(t);
      d[0] = (int16_t)(c.dc_pred * (1 << succ_low));
    } else if (get_bit()) {
      d[0] = (int16_t)(d[0] + (1 << succ_low));
    }
  }
  void block_prog_ac(int16_t* d, JComponent& c)
  {
    if (spec_start == 0) fail("jpeg: DC and AC coefficients merged in a progressive scan");
    if (succ_high == 0) {
      const int shift = succ_low;
      if (eob_run) {
        --eob_run;
        return;
      }
      for (int k = spec_start; k <= spec_end;) {
        const int rs = decode_huff(hac[c.ha]);
        const int r = rs >> 4, s = rs & 15;
        if (s == 0) {
          if (r < 15) {
            eob_run = (1 << r);
            if (r) eob_run += get_bits(r);
            --eob_run;
            break;
          }
          k += 16;
        } else {
          k += r;
          if (k > 63) fail("jpeg: bad AC run");
          const int z = kZigzag[k++];
          d[z] = (int16_t)(receive_extend(s) * (1 << shift));
        }
      }
    } else {
      const int16_t bit = (int16_t)(1 << succ_low);
      auto refine = [&](int16_t& v) {
        if (get_bit() && (v & bit) == 0) v = (int16_t)(v > 0 ? v + bit : v - bit);
      };
      if (eob_run) {
        --eob_run;
        for (int k = spec_start; k <= spec_end; ++k) {
          int16_t& v = d[kZigzag[k]];
          if (v != 0) refine(v);
        }
        return;
      }
      int k = spec_start;
      do {
        const int rs = decode_huff(hac[c.ha]);
        int r = rs >> 4, s = rs & 15;
        if (s == 0) {
          if (r < 15) {
            eob_run = (1 << r) - 1;
            if (r) eob_run += get_bits(r);
            r = 64;  // run to the end of the band, refining on the way
          }
          // r == 15: skip 16 zero coefficients (refining non-zero ones passed over)
        } else {
          if (s != 1) fail("jpeg: bad refinement code");
          s = get_bit() ? bit : -bit;
        }
        while (k <= spec_end) {
          int16_t& v = d[kZigzag[k++]];
          if (v != 0) {
            refine(v);
          } else {
            if (r == 0) {
              v = (int16_t)s;
              break;
            }
            --r;
          }
        }
      } while (k <= spec_end);
    }
  }

  // ---- inverse DCT: LLM with 12-bit constants, 2 guard bits between the passes --------------
  static int fx(double c) { return (int)(c * 4096 + 0.5); }
  static void idct_1d(const int s[8], int out_even[4], int out_odd[4])
  {
    // even part
    const int z1 = (s[2] + s[6]) * fx(0.5411961f);
    const int e2 = z1 + s[6] * fx(-1.847759065f);
    const int e3 = z1 + s[2] * fx(0.765366865f);
    const int e0 = (s[0] + s[4]) * 4096;
    const int e1 = (s[0] - s[4]) * 4096;
    out_even[0] = e0 + e3;
    out_even[3] = e0 - e3;
    out_even[1] = e1 + e2;
    out_even[2] = e1 - e2;
    // odd part
    int o0 = s[7], o1 = s[5], o2 = s[3], o3 = s[1];
    const int z3 = o0 + o2, z4 = o1 + o3, z1o = o0 + o3, z2o = o1 + o2;
    const int z5 = (z3 + z4) * fx(1.175875602f);
    o0 *= fx(0.298631336f);
    o1 *= fx(2.053119869f);
    o2 *= fx(3.072711026f);
    o3 *= fx(1.501321110f);
    const int a = z5 + z1o * fx(-0.899976223f);
    const int b = z5 + z2o * fx(-2.562915447f);
    const int c = z3 * fx(-1.961570560f);
    const int d = z4 * fx(-0.390180644f);
    out_odd[3] = o3 + a + d;
    out_odd[2] = o2 + b + c;
    out_odd[1] = o1 + b + d;
    out_odd[0] = o0 + a + c;
  }
  static uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
  static void idct_block(uint8_t* out, int stride, const int16_t d[64])
  {
    int tmp[64];
    for (int col = 0; col < 8; ++col) {
      int s[8], ev[4], od[4];
      for (int r = 0; r < 8; ++r) s[r] = d[8 * r + col];
      idct_1d(s, ev, od);
      for (int k = 0; k < 4; ++k) {
        const int e = ev[k] + 512;
        tmp[8 * k + col] = (e + od[3 - k]) >> 10;
        tmp[8 * (7 - k) + col] = (e - od[3 - k]) >> 10;
      }
    }
    for (int row = 0; row < 8; ++row) {
      int ev[4], od[4];
      idct_1d(tmp + 8 * row, ev, od);
      uint8_t* o = out + (size_t)stride * row;
      for (int k = 0; k < 4; ++k) {
        const int e = ev[k] + 65536 + (128 << 17);
        o[k] = clamp8((e + od[3 - k]) >> 17);
        o[7 - k] = clamp8((e - od[3 - k]) >> 17);
      }
    }
  }

  // ---- headers ---------------------------------------------------------------------------
  void read_dqt(int len)
  {
    while (len > 0) {
      const int q = get8();
      const int prec = q >> 4, t = q & 15;
      if (prec > 1 || t > 3) fail("jpeg: bad DQT");
      for (int i = 0; i < 64; ++i) dequant[t][kZigzag[i]] = (uint16_t)(prec ? get16() : get8());
      len -= prec ? 129 : 65;
    }
  }
  void read_dht(int len)
  {
    while (len > 0) {
      const int q = get8();
      const int tc = q >> 4, th = q & 15;
      if (tc > 1 || th > 3) fail("jpeg: bad DHT");
      uint8_t counts[16];
      int n = 0;
      for (int i = 0; i < 16; ++i) {
        counts[i] = (uint8_t)get8();
        n += counts[i];
      }
      if (n > 256) fail("jpeg: bad DHT");
      uint8_t vals[256];
      for (int i = 0; i < n; ++i) vals[i] = (uint8_t)get8();
      (tc == 0 ? hdc[th] : hac[th]).build(counts, vals, n);
      len -= 17 + n;
    }
  }
  void read_marker_segment(int m)
  {
    if (m == 0xDD) {
      if (get16() != 4) fail("jpeg: bad DRI");
      restart_interval = get16();
      return;
    }
    int len = get16();
    if (len < 2) fail("jpeg: bad segment length");
    len -= 2;
    if (m == 0xDB) {
      read_dqt(len);
    } else if (m == 0xC4) {
      read_dht(len);
    } else if (m == 0xE0 && len >= 5) {
      static const char tag[5] = {'J', 'F', 'I', 'F', 0};
      bool ok = true;
      for (int i = 0; i < 5; ++i)
        if (get8() != tag[i]) ok = false;
      if (ok) jfif = true;
      p = std::min(end, p + (len - 5));
    } else if (m == 0xEE && len >= 12) {
      static const char tag[6] = {'A', 'd', 'o', 'b', 'e', 0};
      bool ok = true;
      for (int i = 0; i < 6; ++i)
        if (get8() != tag[i]) ok = false;
      int used = 6;
      if (ok) {
        get8();
        get16();
        get16();
        app14_transform = get8();
        used = 12;
      }
      p = std::min(end, p + (len - used));
    } else if ((m >= 0xE0 && m <= 0xEF) || m == 0xFE) {
      p = std::min(end, p + len);
    } else {
      fail("jpeg: unsupported marker");
    }
  }
  int next_marker()
  {
    if (marker >= 0) {
      const int m = marker;
      marker = -1;
      return m;
    }
    int c = get8();
    if (c != 0xff) return -1;
    while (c == 0xff) c = get8();
    return c;
  }
  void read_frame_header()
  {
    const int len = get16();
    if (get8() != 8) fail("jpeg: only 8-bit samples are supported");
    height = get16();
    width = get16();
    ncomp = get8();
    if (width == 0 || height == 0) fail("jpeg: bad image size");
    if (ncomp != 1 && ncomp != 3 && ncomp != 4) fail("jpeg: bad component count");
    if (len != 8 + 3 * ncomp) fail("jpeg: bad SOF length");
    rgb_ids = 0;
    static const char rgb[3] = {'R', 'G', 'B'};
    for (int i = 0; i < ncomp; ++i) {
      JComponent& c = comp[i];
      c.id = get8();
      if (ncomp == 3 && c.id == rgb[i]) ++rgb_ids;
      const int q = get8();
      c.h = q >> 4;
      c.v = q & 15;
      c.tq = get8();
      if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) fail("jpeg: bad component parameters");
      hmax = std::max(hmax, c.h);
      vmax = std::max(vmax, c.v);
    }
    for (int i = 0; i < ncomp; ++i)
      if (hmax % comp[i].h || vmax % comp[i].v) fail("jpeg: fractional sampling ratios are not supported");
    mcux = (width + 8 * hmax - 1) / (8 * hmax);
    mcuy = (height + 8 * vmax - 1) / (8 * vmax);
    for (int i = 0; i < ncomp; ++i) {
      JComponent& c = comp[i];
      c.x = (width * c.h + hmax - 1) / hmax;
      c.y = (height * c.v + vmax - 1) / vmax;
      c.bw = mcux * c.h;
      c.bh = mcuy * c.v;
      c.plane.assign((size_t)c.bw * 8 * c.bh * 8, 0);
      if (progressive) c.coeff.assign((size_t)c.bw * c.bh * 64, 0);
    }
  }
  void read_scan_header()
  {
    const int len = get16();
    scan_n = get8();
    if (scan_n < 1 || scan_n > 4 || scan_n > ncomp) fail("jpeg: bad SOS component count");
    if (len != 6 + 2 * scan_n) fail("jpeg: bad SOS length");
    for (int i = 0; i < scan_n; ++i) {
      const int id = get8(), q = get8();
      int which = 0;
      for (; which < ncomp; ++which)
        if (comp[which].id == id) break;
      if (which == ncomp) fail("jpeg: SOS names an unknown component");
      comp[which].hd = q >> 4;
      comp[which].ha = q & 15;
      if (comp[which].hd > 3 || comp[which].ha > 3) fail("jpeg: bad Huffman table index");
      order[i] = which;
    }
    spec_start = get8();
    spec_end = get8();
    const int a = get8();
    succ_high = a >> 4;
    succ_low = a & 15;
    if (progressive) {
      if (spec_start > 63 || spec_end > 63 || spec_start > spec_end || succ_high > 13 || succ_low > 13) fail("jpeg: bad SOS");
    } else {
      if (spec_start != 0 || succ_high != 0 || succ_low != 0) fail("jpeg: bad SOS");
      spec_end = 63;
    }
  }

  // returns false when the restart marker expected at the end of an interval is missing
  bool end_of_mcu()
  {
    if (--todo > 0) return true;
    if (bitcnt < 24) fill_bits();
    if (marker < 0xD0 || marker > 0xD7) return false;
    reset_entropy();
    return true;
  }

  void decode_scan()
  {
    reset_entropy();
    int16_t block[64];
    if (scan_n == 1) {
      JComponent& c = comp[order[0]];
      const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
      for (int j = 0; j < h; ++j)
        for (int i = 0; i < w; ++i) {
          if (!progressive) {
            block_baseline(block, c);
            idct_block(c.plane.data() + ((size_t)j * 8 * c.bw * 8 + (size_t)i * 8), c.bw * 8, block);
          } else {
            int16_t* d = c.coeff.data() + 64 * ((size_t)i + (size_t)j * c.bw);
            if (spec_start == 0)
              block_prog_dc(d, c);
            else
              block_prog_ac(d, c);
          }
          if (!end_of_mcu()) return;
        }
    } else {
      for (int j = 0; j < mcuy; ++j)
        for (int i = 0; i < mcux; ++i) {
          for (int k = 0; k < scan_n; ++k) {
            JComponent& c = comp[order[k]];
            for (int y = 0; y < c.v; ++y)
              for (int x = 0; x < c.h; ++x) {
                const int bx = i * c.h + x, by = j * c.v + y;
                if (!progressive) {
                  block_baseline(block, c);
                  idct_block(c.plane.data() + ((size_t)by * 8 * c.bw * 8 + (size_t)bx * 8), c.bw * 8, block);
                } else {
                  block_prog_dc(c.coeff.data() + 64 * ((size_t)bx + (size_t)by * c.bw), c);
                }
              }
          }
          if (!end_of_mcu()) return;
        }
    }
  }

  void finish_progressive()
  {
    for (int n = 0; n < ncomp; ++n) {
      JComponent& c = comp[n];
      const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
      const uint16_t* q = dequant[c.tq];
      for (int j = 0; j < h; ++j)
        for (int i = 0; i < w; ++i) {
          int16_t* d = c.coeff.data() + 64 * ((size_t)i + (size_t)j * c.bw);
          for (int k = 0; k < 64; ++k) d[k] = (int16_t)(d[k] * q[k]);
          idct_block(c.plane.data() + ((size_t)j * 8 * c.bw * 8 + (size_t)i * 8), c.bw * 8, d);
        }
    }
  }

  void decode()
  {
    std::memset(dequant, 0, sizeof(dequant));
    if (get8() != 0xff || get8() != 0xD8) fail("jpeg: no SOI");
    int m = next_marker();
    while (!(m == 0xC0 || m == 0xC1 || m == 0xC2)) {
      if (m < 0 || p >= end) fail("jpeg: no SOF");
      if (m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) fail("jpeg: unsupported coding process");
      read_marker_segment(m);
      m = next_marker();
      while (m < 0) {
        if (p >= end) fail("jpeg: no SOF");
        m = next_marker();
      }
    }
    progressive = m == 0xC2;
    read_frame_header();
    m = next_marker();
    while (m != 0xD9) {
      if (m == 0xDA) {
        read_scan_header();
        decode_scan();
        if (marker < 0) {
          // skip any padding up to the next marker
          while (p < end) {
            if (get8() == 0xff) {
              const int c = p < end ? *p : 0;
              if (c != 0 && c != 0xff) {
                marker = get8();
                break;
              }
            }
          }
        }
      } else if (m == 0xDC) {
        const int len = get16();
        const int nl = get16();
        if (len != 4 || nl != height) fail("jpeg: bad DNL");
      } else if (m >= 0) {
        read_marker_segment(m);
      }
      if (p >= end && marker < 0) break;
      m = next_marker();
    }
    if (progressive) finish_progressive();
  }
};

// chroma up-sampling of one output row (see the file comment)
const uint8_t* upsample_row(std::vector<uint8_t>& line, const uint8_t* near, const uint8_t* far, int w, int hs, int vs)
{
  auto d4 = [](int v) { return (uint8_t)(v >> 2); };
  auto d16 = [](int v) { return (uint8_t)(v >> 4); };
  uint8_t* out = line.data();
  if (hs == 1 && vs == 1) return near;
  if (hs == 1 && vs == 2) {
    for (int i = 0; i < w; ++i) out[i] = d4(3 * near[i] + far[i] + 2);
    return out;
  }
  if (hs == 2 && vs == 1) {
    if (w == 1) {
      out[0] = out[1] = near[0];
      return out;
    }
    out[0] = near[0];
    out[1] = d4(near[0] * 3 + near[1] + 2);
    int i = 1;
    for (; i < w - 1; ++i) {
      const int n = 3 * near[i] + 2;
      out[2 * i] = d4(n + near[i - 1]);
      out[2 * i + 1] = d4(n + near[i + 1]);
    }
    out[2 * i] = d4(near[w - 2] * 3 + near[w - 1] + 2);
    out[2 * i + 1] = near[w - 1];
    return out;
  }
  if (hs == 2 && vs == 2) {
    if (w == 1) {
      out[0] = out[1] = d4(3 * near[0] + far[0] + 2);
      return out;
    }
    int cur = 3 * near[0] + far[0];
    out[0] = d4(cur + 2);
    for (int i = 1; i < w; ++i) {
      const int prev = cur;
      cur = 3 * near[i] + far[i];
      out[2 * i - 1] = d16(3 * prev + cur + 8);
      out[2 * i] = d16(3 * cur + prev + 8);
    }
    out[2 * w - 1] = d4(cur + 2);
    return out;
  }
  for (int i = 0; i < w; ++i)
    for (int j = 0; j < hs; ++j) out[i * hs + j] = near[i];
  return out;
}

uint8_t mul8(uint8_t x, uint8_t y)
{
  const unsigned t = (unsigned)x * y + 128;
  return (uint8_t)((t + (t >> 8)) >> 8);
}

}  // namespace

bool is_jpeg(const uint8_t* p, size_t n) { return n >= 3 && p[0] == 0xff && p[1] == 0xD8 && p[2] == 0xff; }

Image8 decode_jpeg(const uint8_t* p, size_t n)
{
  JpegDecoder z;
  z.p = p;
  z.end = p + n;
  z.decode();
  Image8 img;
  img.width = z.width;
  img.height = z.height;
  img.source_channels = z.ncomp >= 3 ? 3 : 1;
  img.rgba.assign((size_t)z.width * z.height * 4, 255);
  const bool is_rgb = z.ncomp == 3 && (z.rgb_ids == 3 || (z.app14_transform == 0 && !z.jfif));

  struct Resampler {
    int hs, vs, ystep, w_lores, ypos;
    const uint8_t *line0, *line1;
    std::vector<uint8_t> buf;
  } rs[4];
  for (int k = 0; k < z.ncomp; ++k) {
    Resampler& r = rs[k];
    r.hs = z.hmax / z.comp[k].h;
    r.vs = z.vmax / z.comp[k].v;
    r.ystep = r.vs >> 1;
    r.w_lores = (z.width + r.hs - 1) / r.hs;
    r.ypos = 0;
    r.line0 = r.line1 = z.comp[k].plane.data();
    r.buf.assign((size_t)z.width + 8, 0);
  }
  const uint8_t* row[4] = {nullptr, nullptr, nullptr, nullptr};
  auto fixed = [](float c) { return ((int)(c * 4096.0f + 0.5f)) << 8; };
  const int cr_r = fixed(1.40200f), cr_g = -fixed(0.71414f), cb_g = -fixed(0.34414f), cb_b = fixed(1.77200f);
  for (int j = 0; j < z.height; ++j) {
    for (int k = 0; k < z.ncomp; ++k) {
      Resampler& r = rs[k];
      const bool bottom = r.ystep >= (r.vs >> 1);
      row[k] = upsample_row(r.buf, bottom ? r.line1 : r.line0, bottom ? r.line0 : r.line1, r.w_lores, r.hs, r.vs);
      if (++r.ystep >= r.vs) {
        r.ystep = 0;
        r.line0 = r.line1;
        if (++r.ypos < z.comp[k].y) r.line1 += (size_t)z.comp[k].bw * 8;
      }
    }
    uint8_t* out = img.rgba.data() + (size_t)j * z.width * 4;
    auto ycc = [&](int i, uint8_t* o) {
      const int yf = (row[0][i] << 20) + (1 << 19);
      const int cr = row[2][i] - 128, cb = row[1][i] - 128;
      int r = yf + cr * cr_r;
      int g = yf + cr * cr_g + (int)((unsigned)(cb * cb_g) & 0xffff0000u);
      int b = yf + cb * cb_b;
      r >>= 20;
      g >>= 20;
      b >>= 20;
      o[0] = JpegDecoder::clamp8(r);
      o[1] = JpegDecoder::clamp8(g);
      o[2] = JpegDecoder::clamp8(b);
    };
    for (int i = 0; i < z.width; ++i) {
      uint8_t* o = out + 4 * i;
      if (z.ncomp == 1) {
        o[0] = o[1] = o[2] = row[0][i];
      } else if (z.ncomp == 3) {
        if (is_rgb) {
          o[0] = row[0][i];
          o[1] = row[1][i];
          o[2] = row[2][i];
        } else {
          ycc(i, o);
        }
      } else {  // four components: CMYK / YCCK (Adobe) or YCbCr + ignored fourth channel
        const uint8_t m = row[3][i];
        if (z.app14_transform == 0) {
          o[0] = mul8(row[0][i], m);
          o[1] = mul8(row[1][i], m);
          o[2] = mul8(row[2][i], m);
        } else if (z.app14_transform == 2) {
          ycc(i, o);
          o[0] = mul8(255 - o[0], m);
          o[1] = mul8(255 - o[1], m);
          o[2] = mul8(255 - o[2], m);
        } else {
          ycc(i, o);
        }
      }
      o[3] = 255;
    }
  }
  return img;
}

// =========================================================================================
// Radiance HDR
// =========================================================================================
bool is_hdr(const uint8_t* p, size_t n)
{
  return (n >= 11 && std::memcmp(p, "#?RADIANCE\n", 11) == 0) || (n >= 7 && std::memcmp(p, "#?RGBE\n", 7) == 0);
}

ImageF decode_hdr(const uint8_t* p, size_t n)
{
  if (!is_hdr(p, n)) fail("hdr: not a Radiance file");
  size_t pos = 0;
  auto line = [&]() {
    std::string s;
    while (pos < n && p[pos] != '\n') s.push_back((char)p[pos++]);
    if (pos < n) ++pos;
    return s;
  };
  line();
  bool format_ok = false;
  for (;;) {
    const std::string s = line();
    if (s.empty()) break;
    if (s == "FORMAT=32-bit_rle_rgbe") format_ok = true;
    if (pos >= n) break;
  }
  if (!format_ok) fail("hdr: unsupported format");
  const std::string res = line();
  int H = 0, W = 0;
  if (std::sscanf(res.c_str(), "-Y %d +X %d", &H, &W) != 2 || W <= 0 || H <= 0) fail("hdr: unsupported data layout");
  ImageF img;
  img.width = W;
  img.height = H;
  img.rgba.assign((size_t)W * H * 4, 1.0f);
  auto convert = [](const uint8_t* rgbe, float* o) {
    if (rgbe[3] != 0) {
      const float f = (float)std::ldexp(1.0f, (int)rgbe[3] - (128 + 8));
      o[0] = rgbe[0] * f;
      o[1] = rgbe[1] * f;
      o[2] = rgbe[2] * f;
    } else {
      o[0] = o[1] = o[2] = 0.0f;
    }
    o[3] = 1.0f;
  };
  auto need = [&](size_t k) {
    if (pos + k > n) fail("hdr: truncated file");
  };
  std::vector<uint8_t> scan((size_t)W * 4);
  bool flat = W < 8 || W >= 32768;
  for (int y = 0; y < H && !flat; ++y) {
    need(4);
    const int c1 = p[pos], c2 = p[pos + 1], len = p[pos + 2];
    if (c1 != 2 || c2 != 2 || (len & 0x80)) {
      if (y != 0) fail("hdr: mixed scanline encodings");
      flat = true;  // not run-length encoded: the whole file is flat RGBE
      break;
    }
    if (((len << 8) | p[pos + 3]) != W) fail("hdr: invalid decoded scanline length");
    pos += 4;
    for (int k = 0; k < 4; ++k) {
      int i = 0;
      while (i < W) {
        need(1);
        int count = p[pos++];
        if (count > 128) {
          count -= 128;
          need(1);
          const uint8_t v = p[pos++];
          if (count == 0 || i + count > W) fail("hdr: corrupt run");
          for (int q = 0; q < count; ++q) scan[(size_t)(i++) * 4 + k] = v;
        } else {
          if (count == 0 || i + count > W) fail("hdr: corrupt run");
          need((size_t)count);
          for (int q = 0; q < count; ++q) scan[(size_t)(i++) * 4 + k] = p[pos++];
        }
      }
    }
    for (int x = 0; x < W; ++x) convert(scan.data() + 4 * (size_t)x, img.rgba.data() + ((size_t)y * W + x) * 4);
  }
  if (flat) {
    need((size_t)W * H * 4);
    for (size_t i = 0; i < (size_t)W * H; ++i) convert(p + pos + 4 * i, img.rgba.data() + 4 * i);
  }
  return img;
}

// =========================================================================================
Image8 load_image8(const std::string& path)
{
  const std::vector<uint8_t> bytes = read_file_bytes(path);
  try {
    if (is_png(bytes.data(), bytes.size())) return decode_png(bytes.data(), bytes.size());
    if (is_jpeg(bytes.data(), bytes.size())) return decode_jpeg(bytes.data(), bytes.size());
  } catch (const std::runtime_error& e) {
    fail("failed to load " + path + ": " + e.what());
  }
  fail("failed to load " + path + ": unsupported image format (PNG and JPEG are supported)");
}

ImageF load_imagef(const std::string& path)
{
  const std::vector<uint8_t> bytes = read_file_bytes(path);
  try {
    if (is_hdr(bytes.data(), bytes.size())) return decode_hdr(bytes.data(), bytes.size());
  } catch (const std::runtime_error& e) {
    fail("failed to load " + path + ": " + e.what());
  }
  // 8-bit files become "HDR" through a gamma-2.2 curve on the colour channels, alpha stays linear
  const Image8 ldr = load_image8(path);
  ImageF img;
  img.width = ldr.width;
  img.height = ldr.height;
  img.rgba.resize(ldr.rgba.size());
  for (size_t i = 0; i < ldr.rgba.size(); i += 4) {
    for (int c = 0; c < 3; ++c) img.rgba[i + c] = (float)(std::pow(ldr.rgba[i + c] / 255.0f, 2.2f) * 1.0f);
    img.rgba[i + 3] = ldr.rgba[i + 3] / 255.0f;
  }
  return img;
}

}  // namespace codec
}  // namespace fredholm
