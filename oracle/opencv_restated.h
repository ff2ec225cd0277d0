// TEST INFRASTRUCTURE.  The three OpenCV imgproc functions discontinuityAdjust calls (stereoMatching.cpp:6061-6064),
// restated for 8-bit single-channel images.  OpenCV's sources are a dependency that is absent from /root/reference (the
// reference links opencv 4.x, stereoMatching.h:15-18); the algorithms below are the published ones and are PINNED against
// cv2 4.13 outputs generated here (tests/golden/make_da_golden.py -> tests/golden/da_ref.npz; tests/test_oracle_golden.py):
//   equalizeHist            histogram, first non-empty bin i0, scale = 255.f / (total - hist[i0]) in float,
//                           lut[i] = saturate_cast<uchar>(sum * scale) (round half to even); a constant image is kept
//   GaussianBlur(3x3, s=4)  the fixed-point path of 8-bit images: kernel {84, 88, 84} / 256 (Q8; getGaussianKernel gives
//                           84.44, 87.12, 84.44 and the fixed-point generator makes the sum exactly 256), horizontal pass
//                           kept in Q8, vertical pass (.. + 2^15) >> 16, BORDER_REFLECT_101
//   Canny(low, high, 3, L1) Sobel 3x3 with BORDER_REPLICATE, magnitude |dx| + |dy| (zero outside the image), non-maximum
//                           suppression with the TG22 fixed-point sector test, hysteresis over the 8-neighbourhood
// Used by stereo_oracle.cpp (orc_disc_adjust) and, through cv_standin.h, by the reference's own discontinuityAdjust body
// compiled into oracle/_ref/libsmref.so.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

namespace orc_cv {

inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}

inline void equalize_hist(const uint8_t* src, uint8_t* dst, int H, int W) {
  long hist[256] = {0};
  const long total = (long)H * W;
  for (long i = 0; i < total; i++) hist[src[i]]++;
  int i0 = 0;
  while (!hist[i0]) ++i0;
  if (hist[i0] == total) {
    for (long i = 0; i < total; i++) dst[i] = (uint8_t)i0;
    return;
  }
  const float scale = 255.f / (float)(total - hist[i0]);
  uint8_t lut[256] = {0};
  long sum = 0;
  for (int i = i0 + 1; i < 256; i++) {
    sum += hist[i];
    long r = lrintf((float)sum * scale);   // cvRound under the default rounding mode
    lut[i] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
  }
  for (long i = 0; i < total; i++) dst[i] = lut[src[i]];
}

inline void gauss3_sigma4(const uint8_t* src, uint8_t* dst, int H, int W) {
  const int A = 84, B = 88;
  std::vector<int> hq((size_t)H * W);
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      const uint8_t* r = src + (size_t)v * W;
      hq[(size_t)v * W + u] = A * ((int)r[reflect101(u - 1, W)] + (int)r[reflect101(u + 1, W)]) + B * (int)r[u];
    }
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      const long s = (long)A * (hq[(size_t)reflect101(v - 1, H) * W + u] + hq[(size_t)reflect101(v + 1, H) * W + u]) +
                     (long)B * hq[(size_t)v * W + u];
      dst[(size_t)v * W + u] = (uint8_t)((s + (1 << 15)) >> 16);
    }
}

// map: 1 = no edge, 0 = candidate below `high`, 2 = edge; dst = 255 on edges.  In place (src == dst) is allowed.
inline void canny3_l1(const uint8_t* src, uint8_t* dst, int H, int W, int low, int high) {
  std::vector<int> dx((size_t)H * W), dy((size_t)H * W), mag((size_t)(H + 2) * (W + 2), 0);
  auto px = [&](int v, int u) { v = v < 0 ? 0 : v >= H ? H - 1 : v; u = u < 0 ? 0 : u >= W ? W - 1 : u; return (int)src[(size_t)v * W + u]; };
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      const int gx = (px(v - 1, u + 1) + 2 * px(v, u + 1) + px(v + 1, u + 1)) - (px(v - 1, u - 1) + 2 * px(v, u - 1) + px(v + 1, u - 1));
      const int gy = (px(v + 1, u - 1) + 2 * px(v + 1, u) + px(v + 1, u + 1)) - (px(v - 1, u - 1) + 2 * px(v - 1, u) + px(v - 1, u + 1));
      dx[(size_t)v * W + u] = gx; dy[(size_t)v * W + u] = gy;
      mag[(size_t)(v + 1) * (W + 2) + u + 1] = std::abs(gx) + std::abs(gy);
    }
  const int TG22 = (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5);
  std::vector<uint8_t> map((size_t)H * W, 1);
  std::vector<int> stack;
  auto M = [&](int v, int u) { return mag[(size_t)(v + 1) * (W + 2) + u + 1]; };
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      const int m = M(v, u);
      if (m <= low) continue;
      const int xs = dx[(size_t)v * W + u], ys = dy[(size_t)v * W + u];
      const long x = std::abs(xs), y = (long)std::abs(ys) << 15, tg22x = x * TG22;
      bool ok;
      if (y < tg22x) ok = m > M(v, u - 1) && m >= M(v, u + 1);
      else {
        const long tg67x = tg22x + (x << 16);
        if (y > tg67x) ok = m > M(v - 1, u) && m >= M(v + 1, u);
        else {
          const int s = (xs ^ ys) < 0 ? -1 : 1;
          ok = m > M(v - 1, u - s) && m > M(v + 1, u + s);
        }
      }
      if (!ok) continue;
      if (m > high) { map[(size_t)v * W + u] = 2; stack.push_back(v * W + u); }
      else map[(size_t)v * W + u] = 0;
    }
  while (!stack.empty()) {
    const int p = stack.back(); stack.pop_back();
    const int v = p / W, u = p % W;
    for (int dv = -1; dv <= 1; dv++)
      for (int du = -1; du <= 1; du++) {
        const int a = v + dv, b = u + du;
        if (a < 0 || a >= H || b < 0 || b >= W) continue;
        if (map[(size_t)a * W + b] == 0) { map[(size_t)a * W + b] = 2; stack.push_back(a * W + b); }
      }
  }
  for (size_t i = 0; i < (size_t)H * W; i++) dst[i] = map[i] == 2 ? 255 : 0;
}

}  // namespace orc_cv
