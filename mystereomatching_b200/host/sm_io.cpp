// sm_io.cpp -- see sm_io.h.  PNG through zlib (inflate / deflate), no other dependency.
#include "sm_io.h"

#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace smio {
namespace {

bool fail(std::string* err, const std::string& msg) {
  if (err) *err = msg;
  return false;
}

bool slurp(const std::string& path, std::vector<uint8_t>& buf) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  buf.resize(n > 0 ? (size_t)n : 0);
  const bool ok = n >= 0 && std::fread(buf.data(), 1, buf.size(), f) == buf.size();
  std::fclose(f);
  return ok;
}

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// Decoded file: 8-bit samples, `c` channels in FILE order (RGB for colour), alpha already dropped.
struct Raw {
  int h = 0, w = 0, c = 0;      // c = 1 (gray) or 3 (RGB)
  std::vector<uint8_t> px;
  std::vector<uint16_t> px16;   // 16-bit colour PNG only: the full samples (libpng converts to gray BEFORE it strips to 8 bits)
  bool is_png = false;
};

int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

bool decode_png(const std::vector<uint8_t>& f, Raw& out, std::string* err) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (f.size() < 8 + 25 || std::memcmp(f.data(), sig, 8) != 0) return fail(err, "not a PNG file");
  size_t pos = 8;
  int W = 0, H = 0, depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> idat, plte;
  bool have_ihdr = false;
  while (pos + 12 <= f.size()) {
    const uint32_t len = be32(&f[pos]);
    const char* type = (const char*)&f[pos + 4];
    if (pos + 12 + len > f.size()) return fail(err, "truncated PNG chunk");
    const uint8_t* d = &f[pos + 8];
    if (!std::memcmp(type, "IHDR", 4)) {
      if (len < 13) return fail(err, "bad IHDR");
      W = (int)be32(d); H = (int)be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12];
      have_ihdr = true;
    } else if (!std::memcmp(type, "PLTE", 4)) plte.assign(d, d + len);
    else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), d, d + len);
    else if (!std::memcmp(type, "IEND", 4)) break;
    pos += 12 + len;
  }
  if (!have_ihdr || W <= 0 || H <= 0) return fail(err, "PNG without IHDR");
  if (interlace) return fail(err, "interlaced PNG is not supported");
  const int nch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
  if (!nch || !(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4))))
    return fail(err, "unsupported PNG colour type / bit depth");
  if (ctype == 3 && plte.size() < 3) return fail(err, "palette PNG without PLTE");
  const size_t rowB = ((size_t)W * nch * depth + 7) / 8;
  const int bpp = std::max(1, nch * depth / 8);
  std::vector<uint8_t> raw((rowB + 1) * (size_t)H);
  uLongf rawLen = (uLongf)raw.size();
  if (uncompress(raw.data(), &rawLen, idat.data(), (uLong)idat.size()) != Z_OK || rawLen != raw.size())
    return fail(err, "PNG inflate failed");
  // unfilter in place
  std::vector<uint8_t> prev(rowB, 0);
  for (int y = 0; y < H; y++) {
    uint8_t* line = &raw[(rowB + 1) * (size_t)y];
    const int ft = line[0];
    uint8_t* cur = line + 1;
    for (size_t i = 0; i < rowB; i++) {
      const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
      int v = cur[i];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: return fail(err, "bad PNG filter type");
      }
      cur[i] = (uint8_t)v;
    }
    std::memcpy(prev.data(), cur, rowB);
  }
  out.h = H; out.w = W; out.is_png = true;
  out.c = (ctype == 0 || ctype == 4) ? 1 : 3;
  out.px.resize((size_t)H * W * out.c);
  for (int y = 0; y < H; y++) {
    const uint8_t* cur = &raw[(rowB + 1) * (size_t)y + 1];
    uint8_t* o = &out.px[(size_t)y * W * out.c];
    for (int x = 0; x < W; x++) {
      auto sample = [&](int k) -> int {   // k-th sample of pixel x, reduced to 8 bits as libpng does for OpenCV
        if (depth == 8) return cur[(size_t)x * nch + k];
        if (depth == 16) return cur[((size_t)x * nch + k) * 2];           // png_set_strip_16: the high byte
        const int per = 8 / depth, idx = x, byte = cur[idx / per], sh = 8 - depth * (idx % per + 1);
        return (byte >> sh) & ((1 << depth) - 1);
      };
      if (ctype == 0) o[x] = depth < 8 ? (uint8_t)(sample(0) * 255 / ((1 << depth) - 1)) : (uint8_t)sample(0);
      else if (ctype == 4) o[x] = (uint8_t)sample(0);
      else if (ctype == 3) {
        const int i = sample(0);
        for (int k = 0; k < 3; k++) o[3 * x + k] = (size_t)(3 * i + k) < plte.size() ? plte[3 * i + k] : 0;
      } else {
        for (int k = 0; k < 3; k++) o[3 * x + k] = (uint8_t)sample(k);
        if (depth == 16) {
          if (out.px16.empty()) out.px16.resize((size_t)H * W * 3);
          for (int k = 0; k < 3; k++) {
            const uint8_t* s16 = &cur[((size_t)x * nch + k) * 2];
            out.px16[((size_t)y * W + x) * 3 + k] = (uint16_t)((s16[0] << 8) | s16[1]);
          }
        }
      }
    }
  }
  return true;
}

bool decode_pnm(const std::vector<uint8_t>& f, Raw& out, std::string* err) {
  size_t pos = 0;
  auto token = [&](std::string& t) -> bool {
    t.clear();
    for (;;) {
      while (pos < f.size() && std::isspace(f[pos])) pos++;
      if (pos < f.size() && f[pos] == '#') { while (pos < f.size() && f[pos] != '\n') pos++; continue; }
      break;
    }
    while (pos < f.size() && !std::isspace(f[pos])) t.push_back((char)f[pos++]);
    return !t.empty();
  };
  std::string magic, sw, sh, smax;
  if (!token(magic) || (magic != "P5" && magic != "P6")) return fail(err, "not a binary PGM / PPM file");
  if (!token(sw) || !token(sh) || !token(smax)) return fail(err, "bad PNM header");
  const int W = std::atoi(sw.c_str()), H = std::atoi(sh.c_str()), maxv = std::atoi(smax.c_str());
  pos++;   // the single whitespace after maxval
  const int c = magic == "P5" ? 1 : 3, bps = maxv > 255 ? 2 : 1;
  if (W <= 0 || H <= 0 || pos + (size_t)W * H * c * bps > f.size()) return fail(err, "truncated PNM data");
  out.h = H; out.w = W; out.c = c; out.is_png = false;
  out.px.resize((size_t)W * H * c);
  for (size_t i = 0; i < out.px.size(); i++) out.px[i] = f[pos + i * bps];   // 16-bit: the high (first) byte
  return true;
}

bool decode(const std::string& path, Raw& out, std::string* err) {
  std::vector<uint8_t> f;
  if (!slurp(path, f)) return fail(err, "cannot read " + path);
  if (f.size() >= 2 && f[0] == 'P' && (f[1] == '5' || f[1] == '6')) return decode_pnm(f, out, err);
  return decode_png(f, out, err);
}

}  // namespace

bool imread_color(const std::string& path, Image& out, std::string* err) {
  out = Image();
  Raw r;
  if (!decode(path, r, err)) return false;
  out.h = r.h; out.w = r.w; out.c = 3;
  out.data.resize((size_t)r.h * r.w * 3);
  for (size_t i = 0; i < (size_t)r.h * r.w; i++) {
    if (r.c == 1) out.data[3 * i] = out.data[3 * i + 1] = out.data[3 * i + 2] = r.px[i];
    else { out.data[3 * i] = r.px[3 * i + 2]; out.data[3 * i + 1] = r.px[3 * i + 1]; out.data[3 * i + 2] = r.px[3 * i]; }   // RGB -> BGR
  }
  return true;
}

bool imread_gray(const std::string& path, Image& out, std::string* err) {
  out = Image();
  Raw r;
  if (!decode(path, r, err)) return false;
  out.h = r.h; out.w = r.w; out.c = 1;
  out.data.resize((size_t)r.h * r.w);
  for (size_t i = 0; i < out.data.size(); i++) {
    if (r.c == 1) { out.data[i] = r.px[i]; continue; }
    if (!r.px16.empty()) {   // libpng's 16-bit rgb_to_gray rounds ((.. + 16384) >> 15), then png_set_strip_16 keeps the high byte
      const uint32_t R16 = r.px16[3 * i], G16 = r.px16[3 * i + 1], B16 = r.px16[3 * i + 2];
      out.data[i] = (uint8_t)((((9797u * R16 + 19234u * G16 + 3737u * B16 + 16384u) >> 15) & 0xffffu) >> 8);
      continue;
    }
    const int R = r.px[3 * i], G = r.px[3 * i + 1], B = r.px[3 * i + 2];
    out.data[i] = r.is_png ? (uint8_t)((9797 * R + 19234 * G + 3737 * B) >> 15)             // libpng rgb_to_gray(0.299, 0.587)
                           : (uint8_t)((4899 * R + 9617 * G + 1868 * B + 8192) >> 14);      // imgcodecs icvCvt_BGR2Gray
  }
  return true;
}

bool read_pfm(const std::string& path, ImageF& out, std::string* err) {
  out = ImageF();
  std::vector<uint8_t> f;
  if (!slurp(path, f)) return fail(err, "cannot read " + path);
  int W = 0, H = 0, used = 0;
  float scale = 0.f;
  char magic[3] = {0, 0, 0};
  if (std::sscanf((const char*)f.data(), "%2s %d %d %f%n", magic, &W, &H, &scale, &used) != 4 || std::strcmp(magic, "Pf") != 0)
    return fail(err, "not a single-channel PFM file");
  size_t pos = (size_t)used + 1;
  if (W <= 0 || H <= 0 || pos + (size_t)W * H * 4 > f.size()) return fail(err, "truncated PFM data");
  const bool little = scale < 0.f;
  out.h = H; out.w = W;
  out.data.resize((size_t)W * H);
  for (int y = 0; y < H; y++)      // rows are stored bottom to top
    for (int x = 0; x < W; x++) {
      const uint8_t* p = &f[pos + ((size_t)(H - 1 - y) * W + x) * 4];
      uint32_t u = little ? ((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24)) : be32(p);
      float v;
      std::memcpy(&v, &u, 4);
      out.data[(size_t)y * W + x] = v;
    }
  return true;
}

bool write_pfm(const std::string& path, const float* data, int h, int w, std::string* err) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return fail(err, "cannot write " + path);
  std::fprintf(f, "Pf\n%d %d\n-1.0\n", w, h);
  for (int y = h - 1; y >= 0; y--) std::fwrite(data + (size_t)y * w, 4, (size_t)w, f);
  std::fclose(f);
  return true;
}

bool write_pnm(const std::string& path, const uint8_t* data, int h, int w, int c, std::string* err) {
  if (c != 1 && c != 3) return fail(err, "write_pnm: 1 or 3 channels");
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return fail(err, "cannot write " + path);
  std::fprintf(f, "%s\n%d %d\n255\n", c == 1 ? "P5" : "P6", w, h);
  if (c == 1) std::fwrite(data, 1, (size_t)h * w, f);
  else
    for (size_t i = 0; i < (size_t)h * w; i++) { const uint8_t rgb[3] = {data[3 * i + 2], data[3 * i + 1], data[3 * i]}; std::fwrite(rgb, 1, 3, f); }
  std::fclose(f);
  return true;
}

bool write_png(const std::string& path, const uint8_t* data, int h, int w, int c, std::string* err) {
  if (c != 1 && c != 3) return fail(err, "write_png: 1 or 3 channels");
  std::vector<uint8_t> raw(((size_t)w * c + 1) * h);
  for (int y = 0; y < h; y++) {
    uint8_t* line = &raw[((size_t)w * c + 1) * y];
    line[0] = 0;   // filter type None
    const uint8_t* s = data + (size_t)y * w * c;
    if (c == 1) std::memcpy(line + 1, s, (size_t)w);
    else
      for (int x = 0; x < w; x++) { line[1 + 3 * x] = s[3 * x + 2]; line[2 + 3 * x] = s[3 * x + 1]; line[3 + 3 * x] = s[3 * x]; }   // BGR -> RGB
  }
  uLongf zlen = compressBound((uLong)raw.size());
  std::vector<uint8_t> z(zlen);
  if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return fail(err, "deflate failed");
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return fail(err, "cannot write " + path);
  auto chunk = [&](const char* type, const uint8_t* d, uint32_t len) {
    uint8_t hd[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len, (uint8_t)type[0], (uint8_t)type[1],
                     (uint8_t)type[2], (uint8_t)type[3]};
    std::fwrite(hd, 1, 8, f);
    if (len) std::fwrite(d, 1, len, f);
    uLong crc = crc32(0L, hd + 4, 4);
    if (len) crc = crc32(crc, d, len);
    const uint8_t cr[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
    std::fwrite(cr, 1, 4, f);
  };
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  std::fwrite(sig, 1, 8, f);
  const uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w, (uint8_t)(h >> 24), (uint8_t)(h >> 16),
                            (uint8_t)(h >> 8), (uint8_t)h, 8, (uint8_t)(c == 1 ? 0 : 2), 0, 0, 0};
  chunk("IHDR", ihdr, 13);
  chunk("IDAT", z.data(), (uint32_t)zlen);
  chunk("IEND", nullptr, 0);
  std::fclose(f);
  return true;
}

// main_.cpp:31-39
const MiddleburyEntry kMiddlebury[33] = {
    {"tsukuba", "scene1.row3.col3", "scene1.row3.col4", "truedisp.row3.col3", 16, 15},
    {"venus", "im2", "im6", "disp2", 8, 19},
    {"teddy", "im2", "im6", "disp2", 4, 59},
    {"cones", "im2", "im6", "disp2", 4, 59},
    {"Art", "view1", "view5", "disp1", 3, 85}, {"Books", "view1", "view5", "disp1", 3, 85}, {"Dolls", "view1", "view5", "disp1", 3, 85},
    {"Laundry", "view1", "view5", "disp1", 3, 85}, {"Moebius", "view1", "view5", "disp1", 3, 85}, {"Reindeer", "view1", "view5", "disp1", 3, 85},
    {"Aloe", "view1", "view5", "disp1", 3, 85}, {"Baby1", "view1", "view5", "disp1", 3, 85}, {"Baby2", "view1", "view5", "disp1", 3, 85},
    {"Baby3", "view1", "view5", "disp1", 3, 85}, {"Bowling1", "view1", "view5", "disp1", 3, 85}, {"Bowling2", "view1", "view5", "disp1", 3, 85},
    {"Cloth1", "view1", "view5", "disp1", 3, 85}, {"Cloth2", "view1", "view5", "disp1", 3, 85}, {"Cloth3", "view1", "view5", "disp1", 3, 85},
    {"Cloth4", "view1", "view5", "disp1", 3, 85}, {"Flowerpots", "view1", "view5", "disp1", 3, 85}, {"Lampshade1", "view1", "view5", "disp1", 3, 85},
    {"Lampshade2", "view1", "view5", "disp1", 3, 85}, {"Midd1", "view1", "view5", "disp1", 3, 85}, {"Midd2", "view1", "view5", "disp1", 3, 85},
    {"Monopoly", "view1", "view5", "disp1", 3, 85}, {"Plastic", "view1", "view5", "disp1", 3, 85}, {"Rocks1", "view1", "view5", "disp1", 3, 85},
    {"Rocks2", "view1", "view5", "disp1", 3, 85}, {"Wood1", "view1", "view5", "disp1", 3, 85}, {"Wood2", "view1", "view5", "disp1", 3, 85},
    {"Katzaa", "left_matlab_valid_resize", "right_matlab_valid_resize", "all", 5, 80},
    {"Michmoret", "left_matlab_valid_resize", "right_matlab_valid_resize", "all", 5, 80},
};

const MiddleburyEntry* middlebury_find(const std::string& object) {
  for (const auto& e : kMiddlebury)
    if (object == e.object) return &e;
  return nullptr;
}

bool load_middlebury(const std::string& root, const std::string& object, StereoPair& out, std::string* err, const std::string& ext) {
  const MiddleburyEntry* e = middlebury_find(object);
  if (!e) return fail(err, "unknown Middlebury object " + object);
  const std::string dir = root + object + "/";
  const std::string left = dir + e->left + ext, right = dir + e->right + ext;
  out = StereoPair();
  // main_.cpp:91-103: colour and gray versions of both views are read from the same files; missing images are fatal
  if (!imread_color(left, out.I1_c, err) || !imread_color(right, out.I2_c, err) || !imread_gray(left, out.I1_g, err) ||
      !imread_gray(right, out.I2_g, err))
    return false;
  if (out.I1_c.h != out.I2_c.h || out.I1_c.w != out.I2_c.w) return fail(err, "left / right sizes differ");
  // main_.cpp:105-118: masks and ground truth are optional ("can't read mask img" does not stop the run)
  std::string ignore;
  imread_gray(dir + "all" + ext, out.all_mask, &ignore);
  imread_gray(dir + "nonocc" + ext, out.nonocc_mask, &ignore);
  imread_gray(dir + "disc" + ext, out.disc_mask, &ignore);
  Image dt;
  if (imread_gray(dir + e->disp + ext, dt, &ignore)) {
    out.DT.h = dt.h; out.DT.w = dt.w;
    out.DT.data.resize(dt.data.size());
    // DT.convertTo(DT, CV_32F, 1.0 / coeff), main_.cpp:126-129: the quotient is a double, OpenCV's 8u -> 32f kernel
    // narrows it to float and multiplies in float (core/src/convert_scale.simd.hpp, cvt_32f)
    const float alpha = (float)(1.0 / e->disp_reduceCoeff);
    for (size_t i = 0; i < dt.data.size(); i++) out.DT.data[i] = (float)dt.data[i] * alpha;
  }
  out.maxdisp = e->maxdisp;
  return true;
}

void disp_to_bgr(const int16_t* disp, int h, int w, int DISP_OCC, int DISP_MIS, int DISP_PKR, std::vector<uint8_t>& bgr, const float* dt,
                 const uint8_t* all_mask, int err_ip_dispV, int cor_ip_dispV) {
  const size_t n = (size_t)h * w;
  bgr.assign(n * 3, 0);
  short dmax = 0, dmin = std::numeric_limits<short>::max();
  for (size_t i = 0; i < n; i++) {
    const short v = disp[i];
    if (v < dmin && v >= 0) dmin = v;
    if (v > dmax) dmax = v;
  }
  const float distance = (float)(dmax - dmin);
  const float ratio = (float)(255.0 / distance);
  for (size_t i = 0; i < n; i++) {
    uint8_t* o = &bgr[3 * i];
    const short v = disp[i];
    if (v >= 0) {
      const float d = ratio * (float)(v - dmin);
      o[0] = o[1] = o[2] = static_cast<uint8_t>(d);
    } else if (v == DISP_OCC) { o[0] = 255; }
    else if (v == DISP_MIS) { o[2] = 255; }
    else if (v == DISP_PKR) { o[1] = 255; o[2] = 255; }
    else if (v == err_ip_dispV) { o[0] = 255; o[2] = 255; }
    else if (v == cor_ip_dispV) { o[0] = 255; o[1] = 255; }
  }
  if (dt && all_mask)
    for (size_t i = 0; i < n; i++)
      if (all_mask[i] > 0 && std::fabs(dt[i] - (float)disp[i]) > 1.f) { bgr[3 * i] = 0; bgr[3 * i + 1] = 0; bgr[3 * i + 2] = 255; }
}

}  // namespace smio

// ---- plain-C view of the same functions (ctypes binding of tests/test_host_io.py; a C caller of the frame stream can
// use them to feed sm_stream_submit from files).  Buffers are caller-allocated; *_dims first, then the read.
extern "C" {
int smio_image_dims(const char* path, int* h, int* w, int* c_file) {
  smio::Image im;
  if (!smio::imread_color(path, im)) return -1;
  *h = im.h; *w = im.w; *c_file = 3;
  return 0;
}
// flag 1: BGR (3 bytes / pixel), flag 0: gray (1 byte / pixel) -- cv::imread's flag
int smio_imread(const char* path, int flag, uint8_t* out, int h, int w) {
  smio::Image im;
  if (!(flag ? smio::imread_color(path, im) : smio::imread_gray(path, im))) return -1;
  if (im.h != h || im.w != w) return -2;
  std::memcpy(out, im.data.data(), im.data.size());
  return 0;
}
int smio_imwrite_png(const char* path, const uint8_t* data, int h, int w, int c) { return smio::write_png(path, data, h, w, c) ? 0 : -1; }
int smio_imwrite_pnm(const char* path, const uint8_t* data, int h, int w, int c) { return smio::write_pnm(path, data, h, w, c) ? 0 : -1; }
int smio_pfm_dims(const char* path, int* h, int* w) {
  smio::ImageF f;
  if (!smio::read_pfm(path, f)) return -1;
  *h = f.h; *w = f.w;
  return 0;
}
int smio_read_pfm(const char* path, float* out, int h, int w) {
  smio::ImageF f;
  if (!smio::read_pfm(path, f)) return -1;
  if (f.h != h || f.w != w) return -2;
  std::memcpy(out, f.data.data(), f.data.size() * 4);
  return 0;
}
int smio_write_pfm(const char* path, const float* data, int h, int w) { return smio::write_pfm(path, data, h, w) ? 0 : -1; }
void smio_disp_to_bgr(const int16_t* disp, int h, int w, int occ, int mis, int pkr, uint8_t* bgr, const float* dt, const uint8_t* all_mask) {
  std::vector<uint8_t> v;
  smio::disp_to_bgr(disp, h, w, occ, mis, pkr, v, dt, all_mask);
  std::memcpy(bgr, v.data(), v.size());
}
// load_middlebury: sizes first (h, w, maxdisp, which optional files were found as bits 0..3 = all, nonocc, disc, DT)
int smio_middlebury_probe(const char* root, const char* object, const char* ext, int* h, int* w, int* maxdisp, int* found) {
  smio::StereoPair p;
  if (!smio::load_middlebury(root, object, p, nullptr, ext)) return -1;
  *h = p.I1_c.h; *w = p.I1_c.w; *maxdisp = p.maxdisp;
  *found = (!p.all_mask.empty()) | (!p.nonocc_mask.empty() << 1) | (!p.disc_mask.empty() << 2) | (!p.DT.data.empty() << 3);
  return 0;
}
int smio_middlebury_load(const char* root, const char* object, const char* ext, uint8_t* I1c, uint8_t* I2c, uint8_t* I1g, uint8_t* I2g,
                         uint8_t* all_mask, uint8_t* nonocc, uint8_t* disc, float* DT) {
  smio::StereoPair p;
  if (!smio::load_middlebury(root, object, p, nullptr, ext)) return -1;
  auto put = [](uint8_t* dst, const smio::Image& im) { if (dst && !im.empty()) std::memcpy(dst, im.data.data(), im.data.size()); };
  put(I1c, p.I1_c); put(I2c, p.I2_c); put(I1g, p.I1_g); put(I2g, p.I2_g);
  put(all_mask, p.all_mask); put(nonocc, p.nonocc_mask); put(disc, p.disc_mask);
  if (DT && !p.DT.data.empty()) std::memcpy(DT, p.DT.data.data(), p.DT.data.size() * 4);
  return 0;
}
}
