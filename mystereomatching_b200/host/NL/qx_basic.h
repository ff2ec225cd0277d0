// qx_basic.h -- the handful of qx_basic helpers a caller of the NL/ classes needs (reference NL/qx_basic.h:163-398:
// contiguous N-D arrays addressed through pointer tables; a[0] / a[0][0] is the flat buffer).  Own implementation.
#pragma once
#include <cstdlib>
#include <cstring>

inline unsigned char** qx_allocu(int r, int c) {
  unsigned char** a = (unsigned char**)malloc(sizeof(unsigned char*) * r);
  a[0] = (unsigned char*)malloc((size_t)r * c);
  for (int i = 1; i < r; i++) a[i] = a[0] + (size_t)i * c;
  return a;
}
inline void qx_freeu(unsigned char** a) { if (a) { free(a[0]); free(a); } }
inline unsigned char*** qx_allocu_3(int n, int r, int c) {
  unsigned char*** a = (unsigned char***)malloc(sizeof(unsigned char**) * n);
  a[0] = (unsigned char**)malloc(sizeof(unsigned char*) * (size_t)n * r);
  a[0][0] = (unsigned char*)malloc((size_t)n * r * c);
  for (int i = 0; i < n; i++) {
    a[i] = a[0] + (size_t)i * r;
    for (int j = 0; j < r; j++) a[i][j] = a[0][0] + ((size_t)i * r + j) * c;
  }
  return a;
}
inline void qx_freeu_3(unsigned char*** a) { if (a) { free(a[0][0]); free(a[0]); free(a); } }
