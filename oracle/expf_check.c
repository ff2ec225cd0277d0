/* TEST INFRASTRUCTURE.  Exhaustive host check of the expf restatement the CUDA library uses (smd_expf_host,
 * mystereomatching_b200/csrc/common.cuh) against this machine's libm: every float in [lo, hi] given on the command line
 * (default [-320, 100]: 2.25e9 inputs, ~17 s).  The restatement follows glibc's sysdeps/ieee754/flt-32/e_expf.c
 * (>= 2.27; the x86-64 FMA ifunc variant, whose multiply-adds are contracted).  Build:
 *   gcc -O2 -ffp-contract=off -mfma oracle/expf_check.c -o /tmp/expf_check -lm
 * Exit code 0 and "mismatches 0" = the device function reproduces the host's expf bit for bit on that range. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static const uint64_t T[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, 0x3fef72b83c7d517bull,
    0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, 0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull,
    0x3feedea64c123422ull, 0x3feece086061892dull, 0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull,
    0x3feea47eb03a5585ull, 0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, 0x3feee89f995ad3adull,
    0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, 0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full,
    0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};
static inline double asd(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t asu(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
float sm_expf_restated(float x) {
  if (x < -0x1.9fe368p6f) return 0.0f;
  if (!(x <= 0x1.62e42ep6f)) return x > 0.f ? INFINITY : x + x;
  const double N = 32, InvLn2N = 0x1.71547652b82fep+0 * N, SHIFT = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-5 / N / N / N, C1 = 0x1.ebfce50fac4f3p-3 / N / N, C2 = 0x1.62e42ff0c52d6p-1 / N;
  double xd = x, z = InvLn2N * xd, kd = z + SHIFT;
  uint64_t ki = asu(kd);
  kd -= SHIFT;
  double r = fma(InvLn2N, xd, -kd);
  uint64_t t = T[ki % 32] + (ki << 47);
  double s = asd(t);
  z = fma(C0, r, C1);
  double r2 = r * r, y = fma(C2, r, 1.0);
  y = fma(z, r2, y);
  y = y * s;
  return (float)y;
}
/* batch form for tests/test_expf_emulation.py: out_restated[i], out_libm[i] for x[i] */
void sm_expf_both(const float* x, long n, float* out_restated, float* out_libm) {
  for (long i = 0; i < n; i++) { out_restated[i] = sm_expf_restated(x[i]); out_libm[i] = expf(x[i]); }
}
#ifndef EXPF_CHECK_NO_MAIN
int main(int argc, char** argv) {
  float lo = argc > 1 ? (float)atof(argv[1]) : -320.f, hi = argc > 2 ? (float)atof(argv[2]) : 100.f;
  long bad = 0, n = 0;
  for (int neg = 0; neg < 2; neg++) {
    float lim = neg ? -lo : hi;
    if (lim < 0) continue;
    uint32_t ulim; memcpy(&ulim, &lim, 4);
    for (uint32_t u = 0; u <= ulim; u++) {
      uint32_t w = u | (neg ? 0x80000000u : 0u);
      float x; memcpy(&x, &w, 4);
      float a = expf(x), b = sm_expf_restated(x);
      uint32_t ua, ub; memcpy(&ua, &a, 4); memcpy(&ub, &b, 4);
      if (ua != ub) { if (bad < 10) printf("x=%a libm=%a restated=%a\n", x, a, b); bad++; }
      n++;
    }
  }
  printf("checked %ld floats in [%g, %g], mismatches %ld\n", n, lo, hi, bad);
  return bad != 0;
}
#endif
