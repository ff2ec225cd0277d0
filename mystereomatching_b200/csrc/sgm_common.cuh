// Scan-line geometry shared by the float (sgm.cu) and the 16-bit integer (sgm_u16.cu) path kernels.
#pragma once
#include "common.cuh"

// sgm()'s direction table (stereoMatching.cpp:6207-6208): offset TO THE PREDECESSOR of path i
static const int SGM_RV[8] = {+1, -1, 0, 0, +1, +1, -1, -1};
static const int SGM_RU[8] = {0, 0, +1, -1, -1, +1, +1, -1};

// Scan-line geometry.  (mv,mu) = direction of travel = -(rv,ru).
struct sgm_geom {
  int H, W, mv, mu, nLines;
};

__device__ __forceinline__ void line_start(const sgm_geom& g, int k, int& v, int& u, int& len) {
  if (g.mv == 0) {  // horizontal: line = row
    v = k; u = g.mu > 0 ? 0 : g.W - 1; len = g.W;
  } else if (g.mu == 0) {  // vertical: line = column
    u = k; v = g.mv > 0 ? 0 : g.H - 1; len = g.H;
  } else {  // diagonal: W lines start on the first row, H-1 more on the entry column
    if (k < g.W) {
      u = k; v = g.mv > 0 ? 0 : g.H - 1;
    } else {
      int j = k - g.W + 1;
      v = g.mv > 0 ? j : g.H - 1 - j;
      u = g.mu > 0 ? 0 : g.W - 1;
    }
    int lv = g.mv > 0 ? g.H - v : v + 1, lu = g.mu > 0 ? g.W - u : u + 1;
    len = min(lv, lu);
  }
}

