// Microbenchmark: what HBM bandwidth does the CBCA access pattern allow?  Each warp streams one (line, 32-d chunk)
// of an [H][W][D] float volume along x (DIR 0: stride D floats) or along v (DIR 1: stride W*D floats), 128 bytes
// per position, copying in -> out, with the cost stream staged by cp.async NB blocks of 8 positions ahead.
// Occupancy is set by padding dynamic shared memory.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define U 8
__device__ __forceinline__ void cp4(uint32_t d, const void* s) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(s) : "memory"); }
__device__ __forceinline__ void cp16(uint32_t d, const void* s) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(s) : "memory"); }
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void waitg() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t lds(uint32_t a) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }

// MODE 0: cp.async 4B staged; MODE 1: plain register loads one block ahead (LDG.32), MODE 2: no load staging, LDG+STG unrolled by U*NB
template <int DIR, int NB, int MODE>
__global__ void __launch_bounds__(64) k(const float* __restrict__ in, float* __restrict__ out, int H, int W, int D, int nChunk, int nLines, int stageOff) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long task = (long long)blockIdx.x * 2 + warp;
  if (task >= (long long)nLines * nChunk) return;
  const int line = (int)(task / nChunk), chunk = (int)(task % nChunk);
  const int N = DIR == 0 ? W : H;
  const size_t step = DIR == 0 ? (size_t)D : (size_t)W * D;
  const size_t e0 = (DIR == 0 ? (size_t)line * W * D : (size_t)line * D) + chunk * 32 + lane;
  const float* pin = in + e0;
  float* pout = out + e0;
  if (MODE == 0) {
    constexpr int NST = NB + 1;
    const uint32_t st = (uint32_t)__cvta_generic_to_shared(sm) + stageOff + warp * NST * U * 128 + lane * 4;
    for (int s = 0; s < NB; s++) {
      for (int i = 0; i < U; i++) if (s * U + i < N) cp4(st + (s * U + i) * 128, pin + (size_t)(s * U + i) * step);
      commit();
    }
    int rd = 0, wr = NB;
    for (int xb = 0; xb < N; xb += U) {
      waitg<NB - 1>();
      float v[U];
#pragma unroll
      for (int i = 0; i < U; i++) v[i] = __uint_as_float(lds(st + (rd * U + i) * 128));
#pragma unroll
      for (int i = 0; i < U; i++) if (xb + i + NB * U < N) cp4(st + (wr * U + i) * 128, pin + (size_t)(xb + i + NB * U) * step);
      commit();
#pragma unroll
      for (int i = 0; i < U; i++) if (xb + i < N) pout[(size_t)(xb + i) * step] = v[i] + 1.0f;
      if (++rd == NST) rd = 0;
      if (++wr == NST) wr = 0;
    }
    waitg<0>();
  } else {
    for (int xb = 0; xb < N; xb += U * NB) {
      float v[U * NB];
#pragma unroll
      for (int i = 0; i < U * NB; i++) v[i] = xb + i < N ? __ldg(pin + (size_t)(xb + i) * step) : 0.f;
#pragma unroll
      for (int i = 0; i < U * NB; i++) if (xb + i < N) pout[(size_t)(xb + i) * step] = v[i] + 1.0f;
    }
  }
}

template <int DIR, int NB, int MODE>
float run(const float* in, float* out, int H, int W, int D, int smemPad) {
  int nChunk = D / 32, nLines = DIR == 0 ? H : W;
  long long tasks = (long long)nLines * nChunk;
  int grid = (int)((tasks + 1) / 2);
  size_t smem = smemPad + 2 * (NB + 1) * U * 128;
  cudaFuncSetAttribute(k<DIR, NB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int r = 0; r < 2; r++) k<DIR, NB, MODE><<<grid, 64, smem>>>(in, out, H, W, D, nChunk, nLines, smemPad);
  cudaEventRecord(a);
  for (int r = 0; r < 3; r++) k<DIR, NB, MODE><<<grid, 64, smem>>>(in, out, H, W, D, nChunk, nLines, smemPad);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("ERR %s\n", cudaGetErrorString(e));
  return ms / 3;
}

int main() {
  const int H = 1080, W = 1920, D = 256;
  size_t n = (size_t)H * W * D;
  float *in, *out; cudaMalloc(&in, n * 4); cudaMalloc(&out, n * 4); cudaMemset(in, 0, n * 4);
  const double gb = 2.0 * n * 4 / 1e9;
  int pads[] = {0, 10 * 1024, 20 * 1024, 44 * 1024};   // -> warps/SM limited by smem: 64, ~36, ~20, ~10
  for (int p : pads) {
    float t;
    t = run<0, 6, 0>(in, out, H, W, D, p); printf("H cpasync NB=6 pad=%5d : %.3f ms %.0f GB/s\n", p, t, gb / t * 1e3);
    t = run<1, 6, 0>(in, out, H, W, D, p); printf("V cpasync NB=6 pad=%5d : %.3f ms %.0f GB/s\n", p, t, gb / t * 1e3);
    t = run<0, 2, 0>(in, out, H, W, D, p); printf("H cpasync NB=2 pad=%5d : %.3f ms %.0f GB/s\n", p, t, gb / t * 1e3);
    t = run<1, 2, 0>(in, out, H, W, D, p); printf("V cpasync NB=2 pad=%5d : %.3f ms %.0f GB/s\n", p, t, gb / t * 1e3);
    t = run<0, 2, 1>(in, out, H, W, D, p); printf("H ldg16    pad=%5d : %.3f ms %.0f GB/s\n", p, t, gb / t * 1e3);
    t = run<1, 2, 1>(in, out, H, W, D, p); printf("V ldg16    pad=%5d : %.3f ms %.0f GB/s\n", p, t, gb / t * 1e3);
    t = run<0, 4, 1>(in, out, H, W, D, p); printf("H ldg32    pad=%5d : %.3f ms %.0f GB/s\n", p, t, gb / t * 1e3);
    t = run<1, 4, 1>(in, out, H, W, D, p); printf("V ldg32    pad=%5d : %.3f ms %.0f GB/s\n", p, t, gb / t * 1e3);
  }
  return 0;
}
