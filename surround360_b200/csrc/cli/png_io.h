// Minimal PNG codec on top of zlib for the command-line tool (host side, off the hot path).
// Reads non-interlaced 8/16-bit grey, grey+alpha, RGB, RGBA and palette PNGs into 8-bit BGR the way
// cv::imread(path, CV_LOAD_IMAGE_COLOR) does (16-bit samples keep their high byte, alpha is dropped,
// grey is replicated); writes 8-bit BGR / BGRA images. For Raw2Rgb: grey frames at their own depth in, 16-bit BGR out.
#pragma once
#include <zlib.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace pngio {

struct Image {
  int w = 0, h = 0, channels = 0;  // channels: 3 = BGR, 4 = BGRA
  std::vector<uint8_t> data;
};

inline uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

inline std::vector<uint8_t> readFile(const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("failed to load image: " + path);
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> buf((size_t)n);
  if (n > 0 && std::fread(buf.data(), 1, (size_t)n, f) != (size_t)n) { std::fclose(f); throw std::runtime_error("short read: " + path); }
  std::fclose(f);
  return buf;
}

inline int paeth(int a, int b, int c) {
  int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// -> BGR8 (cv::imread CV_LOAD_IMAGE_COLOR), or with keepAlpha -> BGRA8 (cv::imread flag -1 on an 8-bit RGBA file; alpha 255 when the file has none)
inline Image readColor(const std::string& path, bool keepAlpha = false) {
  std::vector<uint8_t> buf = readFile(path);
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (buf.size() < 8 || std::memcmp(buf.data(), sig, 8)) throw std::runtime_error("not a PNG file (only PNG input is supported by this tool): " + path);
  size_t pos = 8;
  int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> idat, palette;
  while (pos + 12 <= buf.size()) {
    const uint32_t len = be32(&buf[pos]);
    const std::string type((const char*)&buf[pos + 4], 4);
    const uint8_t* d = &buf[pos + 8];
    if (pos + 12 + len > buf.size()) throw std::runtime_error("truncated PNG: " + path);
    if (type == "IHDR") { w = (int)be32(d); h = (int)be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12]; }
    else if (type == "PLTE") palette.assign(d, d + len);
    else if (type == "IDAT") idat.insert(idat.end(), d, d + len);
    else if (type == "IEND") break;
    pos += 12 + len;
  }
  if (w <= 0 || h <= 0) throw std::runtime_error("bad PNG header: " + path);
  if (interlace) throw std::runtime_error("interlaced PNG is not supported: " + path);
  if (!(depth == 8 || depth == 16) && ctype != 3) throw std::runtime_error("unsupported PNG bit depth: " + path);
  const int spp = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
  if (!spp) throw std::runtime_error("unsupported PNG colour type: " + path);
  if (ctype == 3 && depth != 8) throw std::runtime_error("only 8-bit palette PNG is supported: " + path);
  const int bps = depth / 8, bpp = spp * bps;
  const size_t stride = (size_t)w * bpp;
  std::vector<uint8_t> raw((stride + 1) * (size_t)h);
  uLongf outLen = (uLongf)raw.size();
  if (uncompress(raw.data(), &outLen, idat.data(), (uLong)idat.size()) != Z_OK || outLen != raw.size())
    throw std::runtime_error("PNG inflate failed: " + path);
  std::vector<uint8_t> prev(stride, 0), cur(stride);
  Image img;
  const int oc = keepAlpha ? 4 : 3;
  img.w = w; img.h = h; img.channels = oc;
  img.data.resize((size_t)w * h * oc);
  for (int y = 0; y < h; ++y) {
    const uint8_t* in = &raw[(stride + 1) * (size_t)y];
    const int ft = in[0];
    ++in;
    for (size_t i = 0; i < stride; ++i) {
      const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
      int v = in[i];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: throw std::runtime_error("bad PNG filter: " + path);
      }
      cur[i] = (uint8_t)v;
    }
    uint8_t* out = &img.data[(size_t)y * w * oc];
    for (int x = 0; x < w; ++x) {
      const uint8_t* p = &cur[(size_t)x * bpp];
      uint8_t r, g, b;
      if (ctype == 0 || ctype == 4) r = g = b = p[0];                // high byte first for 16-bit
      else if (ctype == 3) { const size_t k = (size_t)p[0] * 3; if (k + 2 >= palette.size()) throw std::runtime_error("bad palette index: " + path); r = palette[k]; g = palette[k + 1]; b = palette[k + 2]; }
      else { r = p[0]; g = p[bps]; b = p[2 * bps]; }
      out[x * oc] = b; out[x * oc + 1] = g; out[x * oc + 2] = r;
      if (keepAlpha) out[x * oc + 3] = ctype == 6 ? p[3 * bps] : ctype == 4 ? p[bps] : 255;
    }
    std::swap(prev, cur);
  }
  return img;
}

inline void putChunk(std::vector<uint8_t>& out, const char* type, const uint8_t* data, size_t len) {
  uint8_t hdr[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len, (uint8_t)type[0], (uint8_t)type[1], (uint8_t)type[2], (uint8_t)type[3]};
  out.insert(out.end(), hdr, hdr + 8);
  if (len) out.insert(out.end(), data, data + len);
  uLong crc = crc32(0L, hdr + 4, 4);
  if (len) crc = crc32(crc, data, (uInt)len);
  uint8_t c[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
  out.insert(out.end(), c, c + 4);
}

// data: BGR8 (channels 3) or BGRA8 (channels 4), tightly packed
inline void write(const std::string& path, const uint8_t* data, int w, int h, int channels, int level = 1) {
  if (channels != 3 && channels != 4) throw std::runtime_error("png write: 3 or 4 channels only");
  const size_t stride = (size_t)w * channels;
  std::vector<uint8_t> raw((stride + 1) * (size_t)h);
  for (int y = 0; y < h; ++y) {
    uint8_t* o = &raw[(stride + 1) * (size_t)y];
    *o++ = 1;  // Sub filter
    const uint8_t* in = data + stride * y;
    for (int x = 0; x < w; ++x)
      for (int c = 0; c < channels; ++c) {
        const int sc = c < 3 ? 2 - c : 3;  // BGR(A) -> RGB(A)
        const int cur = in[x * channels + sc], left = x ? in[(x - 1) * channels + sc] : 0;
        o[x * channels + c] = (uint8_t)(cur - left);
      }
  }
  uLongf bound = compressBound((uLong)raw.size());
  std::vector<uint8_t> z(bound);
  if (compress2(z.data(), &bound, raw.data(), (uLong)raw.size(), level) != Z_OK) throw std::runtime_error("png deflate failed");
  std::vector<uint8_t> out;
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  out.insert(out.end(), sig, sig + 8);
  uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w, (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h,
                      8, (uint8_t)(channels == 3 ? 2 : 6), 0, 0, 0};
  putChunk(out, "IHDR", ihdr, 13);
  putChunk(out, "IDAT", z.data(), bound);
  putChunk(out, "IEND", nullptr, 0);
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("failed to write image: " + path);
  const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
  std::fclose(f);
  if (!ok) throw std::runtime_error("failed to write image: " + path);
}

// Grey PNG (colour type 0, 8 or 16 bit) at its own depth, as cv::imread(path, CV_LOAD_IMAGE_GRAYSCALE | CV_LOAD_IMAGE_ANYDEPTH) returns a
// raw Bayer frame (Raw2Rgb.cpp:402-404). depthBits: 8 or 16; samples: host-endian, row major.
struct Grey {
  int w = 0, h = 0, depthBits = 0;
  std::vector<uint8_t> data;
};
inline Grey readGrey(const std::string& path) {
  std::vector<uint8_t> buf = readFile(path);
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (buf.size() < 8 || std::memcmp(buf.data(), sig, 8)) throw std::runtime_error("not a PNG file (only .png and .raw input are supported by this tool): " + path);
  size_t pos = 8;
  int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> idat;
  while (pos + 12 <= buf.size()) {
    const uint32_t len = be32(&buf[pos]);
    const std::string type((const char*)&buf[pos + 4], 4);
    const uint8_t* d = &buf[pos + 8];
    if (pos + 12 + len > buf.size()) throw std::runtime_error("truncated PNG: " + path);
    if (type == "IHDR") { w = (int)be32(d); h = (int)be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12]; }
    else if (type == "IDAT") idat.insert(idat.end(), d, d + len);
    else if (type == "IEND") break;
    pos += 12 + len;
  }
  if (w <= 0 || h <= 0) throw std::runtime_error("bad PNG header: " + path);
  if (interlace || ctype != 0 || !(depth == 8 || depth == 16)) throw std::runtime_error("expected a non-interlaced 8 or 16 bit grey PNG (a raw Bayer frame): " + path);
  const int bpp = depth / 8;
  const size_t stride = (size_t)w * bpp;
  std::vector<uint8_t> raw((stride + 1) * (size_t)h);
  uLongf outLen = (uLongf)raw.size();
  if (uncompress(raw.data(), &outLen, idat.data(), (uLong)idat.size()) != Z_OK || outLen != raw.size()) throw std::runtime_error("PNG inflate failed: " + path);
  Grey g;
  g.w = w; g.h = h; g.depthBits = depth;
  g.data.resize(stride * (size_t)h);
  std::vector<uint8_t> prev(stride, 0), cur(stride);
  for (int y = 0; y < h; ++y) {
    const uint8_t* in = &raw[(stride + 1) * (size_t)y];
    const int ft = *in++;
    for (size_t i = 0; i < stride; ++i) {
      const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
      int v = in[i];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: throw std::runtime_error("bad PNG filter: " + path);
      }
      cur[i] = (uint8_t)v;
    }
    uint8_t* out = &g.data[stride * (size_t)y];
    if (bpp == 1) std::memcpy(out, cur.data(), stride);
    else for (int x = 0; x < w; ++x) { const uint16_t v = (uint16_t)(cur[2 * x] << 8 | cur[2 * x + 1]); std::memcpy(out + 2 * x, &v, 2); }   // big endian -> host
    std::swap(prev, cur);
  }
  return g;
}

// data: BGR, 16 bits per sample (host endian), tightly packed -> 16-bit RGB PNG (what cv::imwrite makes of a CV_16UC3 Mat)
inline void write16(const std::string& path, const uint16_t* data, int w, int h, int level = 1) {
  const size_t stride = (size_t)w * 6;
  std::vector<uint8_t> raw((stride + 1) * (size_t)h);
  for (int y = 0; y < h; ++y) {
    uint8_t* o = &raw[(stride + 1) * (size_t)y];
    *o++ = 0;  // no filter
    const uint16_t* in = data + (size_t)w * 3 * y;
    for (int x = 0; x < w; ++x)
      for (int c = 0; c < 3; ++c) { const uint16_t v = in[x * 3 + 2 - c]; o[x * 6 + 2 * c] = (uint8_t)(v >> 8); o[x * 6 + 2 * c + 1] = (uint8_t)v; }
  }
  uLongf bound = compressBound((uLong)raw.size());
  std::vector<uint8_t> z(bound);
  if (compress2(z.data(), &bound, raw.data(), (uLong)raw.size(), level) != Z_OK) throw std::runtime_error("png deflate failed");
  std::vector<uint8_t> out;
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  out.insert(out.end(), sig, sig + 8);
  uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w, (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h, 16, 2, 0, 0, 0};
  putChunk(out, "IHDR", ihdr, 13);
  putChunk(out, "IDAT", z.data(), bound);
  putChunk(out, "IEND", nullptr, 0);
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("failed to write image: " + path);
  const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
  std::fclose(f);
  if (!ok) throw std::runtime_error("failed to write image: " + path);
}

}  // namespace pngio
