// K8 + K9: non-local (minimum-spanning-tree) cost aggregation.
//
//   qx_mst_kruskals_image::mst     NL/qx_mst_kruskals_image.cpp:167-277   guidance median, edges, Kruskal, BFS rooting
//   qx_tree_filter::update_table   NL/qx_tree_filter.cpp:21-25            w[i] = exp(-i / (255 * max(0.01, sigma)))
//   qx_tree_filter::filter         NL/qx_tree_filter.cpp:61-117           leaf-to-root then root-to-leaf recurrences, f64
//   NLCCA::aggreCV                 NL/NLCCA.cpp:27-95                     f32 -> f64, sigma = 0.1, filter, f64 -> f32
//   StereoMatching::NL             stereoMatching.cpp:4892-4917           aggreCV(vm[0]), aggreCV(ones), divide
//
// Tree identity.  The reference runs Kruskal over a STABLE counting sort of the 4-connected grid edges, i.e. it
// builds the unique minimum spanning tree under the strict total order key(e) = (weight(e), index(e)) with the
// reference's edge enumeration (all horizontal edges row-major, then all vertical edges column-major,
// NL/qx_mst_kruskals_image.cpp:46-69).  Any algorithm that uses the same total order yields the same tree; here it
// is Boruvka: every component picks its minimum outgoing edge by 64-bit atomicMin on that key, components hook
// (mutual picks keep the smaller label as the root), labels are flattened by pointer jumping; <= log2(N) rounds.
// The tree is then rooted at pixel 0 (NL/qx_mst_kruskals_image.cpp:233) WITHOUT a level-by-level walk: parent, depth
// and the level order follow from the tree's Euler tour by list ranking (pointer jumping), an inclusive scan and a
// stable sort by depth (nl_root_euler; O(log N) data-parallel steps; scan and radix sort are kernels of this file).
// A node's children are kept in increasing key order, which is the order Kruskal appended them to the adjacency
// list and therefore the order in which the reference's leaf-to-root pass adds them (bit-identical f64 sums).
//
// Filter.  Leaf-to-root over the levels deepest-1 .. 0 (A[p] = cost[p] + sum_children w(c) * A[c], children already
// final), then root-to-leaf over levels 1 .. deepest (out[v] = w * (out[parent] - w * A[v]) + A[v], written over A[v]).
// Planes never interact, so the work volume is split BY PLANE: one CTA per plane walks all levels on one SM with
// __syncthreads between levels (k_tf_cta_fast / k_tf_cta; any barrier across SMs costs 2.6 us per level), the
// level-ordered records and the plane's values streamed through shared memory by the TMA unit.  The all-ones volume of
// StereoMatching::NL has identical planes, so its filter result is ONE extra plane (index D) carried through the
// same two sweeps; the division happens in the final conversion.  Time is bound by tree depth x per-level latency,
// not by bandwidth.
#include <math.h>

#include <stdlib.h>


#include "common.cuh"

// ------------------------------------------------------------------ small helpers
__device__ __forceinline__ void edge_ends(int e, int H, int W, int& a, int& b) {
  const int nh = H * (W - 1);
  if (e < nh) {
    const int y = e / (W - 1), x = e - y * (W - 1);
    a = y * W + x; b = a + 1;
  } else {
    const int r = e - nh, x = r / (H - 1), y = r - x * (H - 1);
    a = y * W + x; b = a + W;
  }
}
// index of the edge between pixel p and its right (dir 0) or lower (dir 1) neighbour
__device__ __forceinline__ int edge_index(int y, int x, int dir, int H, int W) {
  return dir == 0 ? y * (W - 1) + x : H * (W - 1) + x * (H - 1) + y;
}

__global__ void k_edge_weights(const uint8_t* __restrict__ img, int H, int W, int cn, uint8_t* __restrict__ ew) {
  const int E = H * (W - 1) + W * (H - 1);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    int a, b;
    edge_ends(e, H, W, a, b);
    int m = 0;
    for (int c = 0; c < cn; c++) m = max(m, abs((int)img[(size_t)a * cn + c] - (int)img[(size_t)b * cn + c]));
    ew[e] = (uint8_t)m;
  }
}

// ------------------------------------------------------------------ Boruvka
#define KEY_NONE 0xFFFFFFFFFFFFFFFFull

__global__ void k_bor_init(int N, int E, int* comp, int* link, unsigned long long* best, uint8_t* inMST) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) { comp[i] = i; link[i] = i; best[i] = KEY_NONE; }
  if (i < E) inMST[i] = 0;
}

__global__ void k_bor_find(const uint8_t* __restrict__ ew, const int* __restrict__ comp, int H, int W,
                           unsigned long long* best) {
  const int E = H * (W - 1) + W * (H - 1);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    int a, b;
    edge_ends(e, H, W, a, b);
    const int ca = comp[a], cb = comp[b];
    if (ca == cb) continue;
    const unsigned long long key = ((unsigned long long)ew[e] << 32) | (unsigned)e;
    if (key < best[ca]) atomicMin(&best[ca], key);   // plain pre-check keeps late rounds off the hot addresses
    if (key < best[cb]) atomicMin(&best[cb], key);
  }
}

__global__ void k_bor_hook(int N, int H, int W, const int* __restrict__ comp, const unsigned long long* __restrict__ best,
                           int* __restrict__ link, uint8_t* __restrict__ inMST, int* __restrict__ nHooks) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N || comp[c] != c) return;
  const unsigned long long k = best[c];
  if (k == KEY_NONE) return;
  const int e = (int)(k & 0xFFFFFFFFu);
  int a, b;
  edge_ends(e, H, W, a, b);
  const int ca = comp[a], other = ca == c ? comp[b] : ca;
  inMST[e] = 1;
  const bool mutual = best[other] == k;
  if (!mutual || c > other) link[c] = other;   // of a mutual pair the smaller label stays the root
  atomicAdd(nHooks, 1);
}

__global__ void k_bor_jump(int N, int* __restrict__ link, int* __restrict__ changed) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  const int p = link[c];
  const int g = link[p];
  if (g != p) { link[c] = g; *changed = 1; }
}

// Every hooked root follows its links to the root of the merged tree in ONE launch (the links of a round form a
// forest: a component hooks along its minimum edge, keys are unique, the mutual pair keeps the smaller label as root).
// Concurrent writers only ever replace a link by an ancestor further up, so a reader that sees either value still
// walks to the same root.  This replaces the pointer-jumping loop, which cost a host round trip per iteration.
__global__ void k_bor_chase(int N, int* link) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  volatile int* vl = link;
  int r = vl[c];
  if (r == c) return;
  for (int g = vl[r]; g != r; g = vl[r]) r = g;
  link[c] = r;
}

__global__ void k_bor_relabel(int N, int* __restrict__ comp, int* __restrict__ link, unsigned long long* __restrict__ best) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  comp[v] = link[comp[v]];
  best[v] = KEY_NONE;
}

__global__ void k_bor_fixlink(int N, int* __restrict__ link) {
  // after relabelling, comp holds roots only: start the next round from the identity again
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < N) link[v] = v;
}

// ------------------------------------------------------------------ adjacency (<= 4 tree neighbours, key order)
struct nl_adj {   // per node
  int nbr[4];
};

__global__ void k_tree_adj(int H, int W, const uint8_t* __restrict__ ew, const uint8_t* __restrict__ inMST,
                           int* __restrict__ nbr /*[N][4]*/, uint8_t* __restrict__ nbw /*[N][4]*/,
                           uint8_t* __restrict__ deg) {
  const int N = H * W;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  const int y = v / W, x = v - y * W;
  unsigned long long key[4];
  int to[4], n = 0;
  auto consider = [&](int e, int other) {
    if (inMST[e]) { key[n] = ((unsigned long long)ew[e] << 32) | (unsigned)e; to[n] = other; n++; }
  };
  if (x > 0) consider(edge_index(y, x - 1, 0, H, W), v - 1);
  if (x < W - 1) consider(edge_index(y, x, 0, H, W), v + 1);
  if (y > 0) consider(edge_index(y - 1, x, 1, H, W), v - W);
  if (y < H - 1) consider(edge_index(y, x, 1, H, W), v + W);
  // insertion sort by key (<= 4 entries)
  for (int i = 1; i < n; i++) {
    const unsigned long long k = key[i];
    const int t = to[i];
    int j = i - 1;
    while (j >= 0 && key[j] > k) { key[j + 1] = key[j]; to[j + 1] = to[j]; j--; }
    key[j + 1] = k; to[j + 1] = t;
  }
  for (int i = 0; i < 4; i++) {
    nbr[(size_t)v * 4 + i] = i < n ? to[i] : -1;
    nbw[(size_t)v * 4 + i] = i < n ? (uint8_t)(key[i] >> 32) : 0;
  }
  deg[v] = (uint8_t)n;
}

// ------------------------------------------------------------------ level-synchronous kernels
// A tree level holds a few dozen to a few thousand nodes and image MSTs are deep (4366 levels at 640x480, 17051 at
// 1920x1080), so rooting and filtering are latency bound: time = levels x (cost of handing a level's results to the
// next).  Measured on B200: any barrier ACROSS SMs costs 2.6-2.8 us per level (hardware cluster barrier of 8 CTAs and
// a grid-wide atomic barrier alike: store acknowledge + release/acquire + dependent L2 load).  Both kernels therefore
// keep every dependency inside ONE SM and synchronise with __syncthreads only:
//   * k_tf_sweeps: the volume is split by PLANE (planes never interact), one CTA per plane walking all levels, the
//     values of the adjacent level in shared memory, records and own values fetched four levels ahead, level bounds in
//     shared memory: ~0.9 us per level (what remains is the dependent instruction chain of the few working threads).
#define NL_CTA 1024
#define NL_TF_CTA 512
#define NL_TF_CAP 4096   // nodes of a level kept in shared memory (x planes per CTA; 2 buffers of doubles: 64 KB)
#define NL_TF_LS 32768   // level bounds kept in shared memory (128 KB); deeper trees read the rest from global memory

struct nl_sync {
  int cnt[3];           // BFS: nodes appended per level, rotating (level % 3)
  int nlevels;          // BFS result: number of levels
};


// ------------------------------------------------------------------ rooting without a level-by-level walk
// A level-by-level BFS costs one dependent global access per LEVEL (0.75 us x 4366 levels at 640x480, x 17051 at 1080p).  The
// same parent / depth / level order follow from the tree's EULER TOUR in O(log N) data-parallel steps:
//   * directed edge (u -> k-th neighbour v) has id 4u + k; its successor in the tour is (v -> the neighbour after u in
//     v's adjacency list, cyclically); the tour starts with edge (0 -> first neighbour of 0), the edge whose successor
//     would be that one ends it;
//   * list ranking by pointer jumping gives every edge its position in the tour;
//   * an edge is DOWNWARD (u is v's parent) iff it comes before its reverse; parent[v] = u, wpar[v] = its weight;
//   * +1 for downward, -1 for upward edges, inclusive scan over the tour: the value at a downward edge is depth(v);
//   * a stable radix sort of the nodes by depth (node order within a level) is the level order.
// Identical parent / weight / rank to the BFS (a rooted tree has one parent function); `order` is grouped by level,
// which is all the filter needs (its results do not depend on the order inside a level).
#define ET_END (-1)
// (successor, distance) of a directed edge live in ONE int2: a jump gathers one 8-byte entry instead of two 4-byte ones from
// two arrays (each a 32-byte sector of its own)
__global__ void k_et_init(int N, const int4* __restrict__ nbr, const uint8_t* __restrict__ deg, int2* __restrict__ sd) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 4 * N) return;
  const int u = e >> 2, k = e & 3;
  int s = ET_END, d = 0;
  if (k < deg[u]) {
    const int4 a = nbr[u];
    const int v = k == 0 ? a.x : (k == 1 ? a.y : (k == 2 ? a.z : a.w));
    const int4 b = nbr[v];
    const int dv = deg[v];
    const int j = b.x == u ? 0 : (b.y == u ? 1 : (b.z == u ? 2 : 3));
    const int nj = j + 1 == dv ? 0 : j + 1;
    s = 4 * v + nj;
    if (s == 0) s = ET_END;   // edge (0, 0) starts the tour: whoever precedes it is the last edge
    d = s == ET_END ? 0 : 1;
  }
  sd[e] = make_int2(s, d);
}
__global__ void k_et_jump(int n, const int2* __restrict__ sd, int2* __restrict__ sd2) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  int2 a = sd[e];
  if (a.x != ET_END) { const int2 b = sd[a.x]; a.y += b.y; a.x = b.x; }
  sd2[e] = a;
}
// dist = number of edges after e in the tour; T = 2(N-1) edges; position = T - 1 - dist
__global__ void k_et_classify(int N, int T, const int4* __restrict__ nbr, const uchar4* __restrict__ nbw,
                              const uint8_t* __restrict__ deg, const int2* __restrict__ sd, int* __restrict__ parent,
                              uint8_t* __restrict__ wpar, int* __restrict__ pm, int* __restrict__ node_at) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 4 * N) return;
  const int u = e >> 2, k = e & 3;
  if (k >= deg[u]) return;
  const int4 a = nbr[u];
  const uchar4 w = nbw[u];
  const int v = k == 0 ? a.x : (k == 1 ? a.y : (k == 2 ? a.z : a.w));
  const int wk = k == 0 ? w.x : (k == 1 ? w.y : (k == 2 ? w.z : w.w));
  const int4 b = nbr[v];
  const int j = b.x == u ? 0 : (b.y == u ? 1 : (b.z == u ? 2 : 3));
  const int pe = T - 1 - sd[e].y, pr = T - 1 - sd[4 * v + j].y;
  const bool down = pe < pr;
  pm[pe] = down ? 1 : -1;
  node_at[pe] = down ? v : -1;
  if (down) { parent[v] = u; wpar[v] = (uint8_t)wk; }
}
// after the inclusive scan: depth of the node entered by the downward edge at tour position p
__global__ void k_et_depth(int T, const int* __restrict__ scan, const int* __restrict__ node_at, int* __restrict__ rank) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > T) return;
  if (i == 0) { rank[0] = 0; return; }
  const int p = i - 1, v = node_at[p];
  if (v >= 0) rank[v] = scan[p];
}

// ------------------------------------------------------------------ scan and stable radix sort (the two library-shaped steps
// of the rooting, written out: a three-kernel tile scan and an LSD radix sort, 8 bits per pass)
#define NLP_TB 256
#define NLP_IT 8
#define NLP_TILE (NLP_TB * NLP_IT)

// inclusive (EXCL = false) or exclusive scan of every 2048-element tile on its own; sums[tile] = the tile's total
template <bool EXCL>
__global__ void __launch_bounds__(NLP_TB) k_scan_tiles(const int* __restrict__ in, int* __restrict__ out, int n, int* __restrict__ sums) {
  __shared__ int wsum[NLP_TB / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int base = blockIdx.x * NLP_TILE + tid * NLP_IT;
  int v[NLP_IT], run = 0;
#pragma unroll
  for (int j = 0; j < NLP_IT; j++) { v[j] = base + j < n ? in[base + j] : 0; run += v[j]; }
  int inc = run;   // inclusive scan of the thread totals inside the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; w++) woff += wsum[w];
  int acc = woff + inc - run;   // exclusive prefix of this thread inside the tile
#pragma unroll
  for (int j = 0; j < NLP_IT; j++) {
    const int e = acc;
    acc += v[j];
    if (base + j < n) out[base + j] = EXCL ? e : acc;
  }
  if (tid == NLP_TB - 1) sums[blockIdx.x] = acc;
}
// exclusive scan of the tile totals in place (one block; any number of tiles, 1024 at a time with a carry)
__global__ void __launch_bounds__(1024) k_scan_sums(int* __restrict__ sums, int nt) {
  __shared__ int wsum[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int c0 = 0; c0 < nt; c0 += 1024) {
    const int i = c0 + tid;
    const int x = i < nt ? sums[i] : 0;
    int inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; w++) woff += wsum[w];
    const int c = carry;
    if (i < nt) sums[i] = c + woff + inc - x;
    __syncthreads();
    if (tid == 1023) carry = c + woff + inc;
    __syncthreads();
  }
}
__global__ void k_scan_add(int* __restrict__ out, int n, const int* __restrict__ sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += sums[i / NLP_TILE];
}
// out = scan(in) over n ints; sums: ceil(n / 2048) ints of scratch
template <bool EXCL>
static int nl_scan(sm_ctx* ctx, const int* in, int* out, int n, int* sums) {
  const int nt = sm_div_up(n, NLP_TILE);
  SM_LAUNCH(ctx, k_scan_tiles<EXCL>, nt, NLP_TB, 0, in, out, n, sums);
  if (nt > 1) {
    SM_LAUNCH(ctx, k_scan_sums, 1, 1024, 0, sums, nt);
    SM_LAUNCH(ctx, k_scan_add, sm_div_up(n, 256), 256, 0, out, n, sums);
  }
  return SM_OK;
}

// one 8-bit pass of a stable LSD radix sort: per-tile digit histogram, hist[digit][tile]
__global__ void __launch_bounds__(NLP_TB) k_rs_hist(const unsigned* __restrict__ keys, int n, int shift, int* __restrict__ hist, int nt) {
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * NLP_TILE;
  for (int j = 0; j < NLP_IT; j++) {
    const int i = base + j * NLP_TB + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nt + blockIdx.x] = h[threadIdx.x];
}
// scatter with the scanned histogram: an item goes to offs[digit][tile] + (items of the same digit before it in the tile).
// Rounds of 256 consecutive items; inside a round the rank comes from a warp match + the counts of the warps before.
__global__ void __launch_bounds__(NLP_TB) k_rs_scatter(const unsigned* __restrict__ keys, const int* __restrict__ vals, unsigned* __restrict__ keys2,
                                                      int* __restrict__ vals2, int n, int shift, const int* __restrict__ offs, int nt) {
  __shared__ int basePos[256];
  __shared__ int cnt[NLP_TB / 32][256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  basePos[tid] = offs[(size_t)tid * nt + blockIdx.x];
  const int base = blockIdx.x * NLP_TILE;
  for (int j = 0; j < NLP_IT; j++) {
    const int i = base + j * NLP_TB + tid;
    const bool valid = i < n;
    const unsigned key = valid ? keys[i] : 0u;
    const int val = valid ? vals[i] : 0;
    const unsigned dg = valid ? ((key >> shift) & 255u) : 256u;   // 256: the items past the end group among themselves
#pragma unroll
    for (int w = 0; w < NLP_TB / 32; w++) cnt[w][tid] = 0;
    __syncthreads();
    const unsigned peers = __match_any_sync(0xffffffffu, dg);
    const int lrank = __popc(peers & ((1u << lane) - 1u));
    if (valid && lrank == 0) cnt[warp][dg] = __popc(peers);
    __syncthreads();
    int pos = 0;
    if (valid) {
      int before = 0;
      for (int w = 0; w < warp; w++) before += cnt[w][dg];
      pos = basePos[dg] + before + lrank;
    }
    __syncthreads();
    int add = 0;
#pragma unroll
    for (int w = 0; w < NLP_TB / 32; w++) add += cnt[w][tid];
    basePos[tid] += add;
    __syncthreads();
    if (valid) { keys2[pos] = key; vals2[pos] = val; }
  }
}
__global__ void k_rs_init(int N, const int* __restrict__ rank, unsigned* __restrict__ keys, int* __restrict__ vals) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < N) { keys[v] = (unsigned)rank[v]; vals[v] = v; }
}

// rooted tree from the adjacency by the Euler tour; returns SM_OK and fills parent / wpar / rank / order /
// level_start / sync->nlevels (order: grouped by level)
struct nl_tree;
static int nl_root_euler(sm_ctx* ctx, int N, const int* nbr, const uint8_t* nbw, const uint8_t* deg, int* parent,
                         uint8_t* wpar, int* rank, int* order, int* level_start, nl_sync* sync);

// children (in key order) from parent / weight arrays alone -- lets sm_tree_filter accept any rooted tree
__global__ void k_children_from_parent(int H, int W, const int* __restrict__ parent, const uint8_t* __restrict__ wpar,
                                       int* __restrict__ child /*[N][4]*/, uint8_t* __restrict__ nchild) {
  const int N = H * W;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  const int y = v / W, x = v - y * W;
  unsigned long long key[4];
  int to[4], n = 0;
  auto consider = [&](int c, int e) {
    if (c != v && parent[c] == v && c != parent[v]) { key[n] = ((unsigned long long)wpar[c] << 32) | (unsigned)e; to[n] = c; n++; }
  };
  if (x > 0) consider(v - 1, edge_index(y, x - 1, 0, H, W));
  if (x < W - 1) consider(v + 1, edge_index(y, x, 0, H, W));
  if (y > 0) consider(v - W, edge_index(y - 1, x, 1, H, W));
  if (y < H - 1) consider(v + W, edge_index(y, x, 1, H, W));
  for (int i = 1; i < n; i++) {
    const unsigned long long k = key[i];
    const int t = to[i];
    int j = i - 1;
    while (j >= 0 && key[j] > k) { key[j + 1] = key[j]; to[j + 1] = to[j]; j--; }
    key[j + 1] = k; to[j + 1] = t;
  }
  for (int i = 0; i < 4; i++) child[(size_t)v * 4 + i] = i < n ? to[i] : -1;
  nchild[v] = (uint8_t)n;
}

__global__ void k_level_bounds(int N, const int* __restrict__ rank, const int* __restrict__ order,
                               int* __restrict__ level_start, nl_sync* s) {
  // order is grouped by rank: level l starts where the rank changes to l
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const int r = rank[order[i]];
    if (i == 0 || rank[order[i - 1]] != r) level_start[r] = i;
    if (i == N - 1) { level_start[r + 1] = N; s->nlevels = r + 1; }
  }
}

// ------------------------------------------------------------------ tree filter
// A: Dp planes of N doubles (Dp = D, or D+1 with the all-ones plane at index D), element (v, d) at A[v * sn + d * sd]:
// plane-major (sn = 1, sd = N) for the work buffer the filter owns, node-major (sn = D, sd = 1) in place for
// qx_tree_filter::filter's own layout.
// Both directions go through a 32-position x 64-plane tile in shared memory: a pixel's D floats are contiguous in the
// volume and a plane's positions are contiguous in A, so each side is read / written in 128-256-byte runs (one thread per
// element of A gathered 4 bytes from 32 different pixels per warp instruction: 0.124 + 0.101 ms per frame at 640x480 D=64).
#define TFX_Q 32
#define TFX_D 64
__global__ void __launch_bounds__(256)
    k_tf_load(const float* __restrict__ vol, double* __restrict__ A, size_t N, int D, int Dp, const int* __restrict__ order) {
  // work buffer: plane-major AND level-ordered, element (position q of `order`, plane d) at A[d * N + q]
  __shared__ float tile[TFX_D][TFX_Q + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (size_t q0 = (size_t)blockIdx.x * TFX_Q; q0 < N; q0 += (size_t)gridDim.x * TFX_Q) {
    if (Dp > D && warp == 0 && q0 + lane < N) A[(size_t)D * N + q0 + lane] = 1.0;
    for (int d0 = 0; d0 < D; d0 += TFX_D) {
      const int nd = min(TFX_D, D - d0);
      for (int k = warp; k < TFX_Q; k += 8) {          // a pixel's run of planes, coalesced
        const size_t q = q0 + k;
        if (q < N) {
          const float* src = vol + (size_t)order[q] * D + d0;
          for (int d = lane; d < nd; d += 32) tile[d][k] = src[d];
        }
      }
      __syncthreads();
      for (int d = warp; d < nd; d += 8)               // a plane's run of positions, coalesced
        if (q0 + lane < N) A[(size_t)(d0 + d) * N + q0 + lane] = (double)tile[d][lane];
      __syncthreads();
    }
  }
}

// Level-ordered records of the rooted tree, so that a level's static data is read contiguously (and one level
// ahead) instead of through order -> node -> child -> value chains:
//   position i of `order` (node v = order[i]):  rnc[i] = number of children, rcp[i][k] = position of child k INSIDE
//   its level (the next deeper one), rcw[i] = the four child weights packed, rpp[i] = position of the parent inside
//   its level, rw[i] = weight of the edge to the parent.
__global__ void k_tf_positions(int N, const int* __restrict__ order, int* __restrict__ pos) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) pos[order[i]] = i;
}
__global__ void k_tf_records(int N, const int* __restrict__ order, const int* __restrict__ pos, const int* __restrict__ rank,
                             const int* __restrict__ level_start, const int* __restrict__ parent,
                             const uint8_t* __restrict__ wpar, const int* __restrict__ child, const uint8_t* __restrict__ nchild,
                             int4* __restrict__ rcp, uint32_t* __restrict__ rcw, uint8_t* __restrict__ rnc, int* __restrict__ rpp,
                             uint8_t* __restrict__ rw, uint4* __restrict__ rup, int2* __restrict__ rdn) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const int v = order[i], l = rank[v], nc = nchild[v];
    int cp[4] = {0, 0, 0, 0};
    uint32_t cw = 0;
    for (int k = 0; k < nc; k++) {
      const int c = child[(size_t)v * 4 + k];
      cp[k] = pos[c] - level_start[l + 1];
      cw |= (uint32_t)wpar[c] << (8 * k);
    }
    rcp[i] = make_int4(cp[0], cp[1], cp[2], cp[3]);
    rcw[i] = cw;
    rnc[i] = (uint8_t)nc;
    rpp[i] = l > 0 ? pos[parent[v]] - level_start[l - 1] : 0;
    rw[i] = wpar[v];
    rup[2 * (size_t)i] = make_uint4((unsigned)cp[0], (unsigned)cp[1], (unsigned)cp[2], (unsigned)cp[3]);
    rup[2 * (size_t)i + 1] = make_uint4(cw, (unsigned)nc, 0u, 0u);
    rdn[i] = make_int2(rpp[i], (int)wpar[v]);
  }
}

// The two sweeps.  Planes never interact, so the volume is split BY PLANE: a CTA owns a few planes of ALL nodes and
// walks the levels alone, one __syncthreads per level; no barrier across SMs is needed at all (the cluster-wide
// version cost 2.6 us per level).  A level only consumes the level next to it, which this CTA has just produced:
// those values are kept in shared memory (two level buffers of NL_TF_CAP nodes; wider levels fall back to global
// memory for the excess), and each thread fetches its first item of the NEXT level -- record and own value, which
// do not depend on the level in progress -- before it works on the current one.  What is left per level is a
// barrier, a shared-memory read and the arithmetic.
// Per (node, plane) the operations and their order are exactly the reference's, so the result is bit-identical.
struct nl_up_item { int v; int nc; int4 cp; uint32_t cw; double own; };
struct nl_dn_item { int v; int pp; uint32_t w; double own; };
#define NL_TF_PF 4   // levels fetched ahead (first item of every thread)

// LEVELORD: A is the filter's own work buffer, plane-major and level-ordered (A[d * N + position]): every address of a
// level is known from the level bounds alone.  Otherwise A is the caller's node-major array (A[v * sn + d * sd]) and
// the node index comes from `order`.
template <bool LEVELORD, bool NP1>
__global__ void __launch_bounds__(NL_TF_CTA)
    k_tf_sweeps(double* A, size_t sn, size_t sd, int Dp, int planesPerCta, int cap, const int* __restrict__ order,
                const int* __restrict__ level_start, const int4* __restrict__ rcp, const uint32_t* __restrict__ rcw,
                const uint8_t* __restrict__ rnc, const int* __restrict__ rpp, const uint8_t* __restrict__ rw,
                const double* __restrict__ table, const nl_sync* __restrict__ s) {
  extern __shared__ double lv[];   // [2][cap * planesPerCta] values of two adjacent levels, then NL_TF_LS level bounds
  __shared__ double tab[256];
  const int nlevels = s->nlevels;
  // the level bounds are read by every warp at every level: from shared memory, not through L1/L2
  int* ls = reinterpret_cast<int*>(lv + 2 * (size_t)cap * planesPerCta);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = table[i];
  for (int i = threadIdx.x; i <= nlevels && i < NL_TF_LS; i += blockDim.x) ls[i] = level_start[i];
  __syncthreads();
  const uint32_t lsAddr = (uint32_t)__cvta_generic_to_shared(ls);
  auto lstart = [&](int l) -> int {
    if (l >= NL_TF_LS) return level_start[l];
    int v;   // explicit ld.shared: through the cast pointer the compiler emits a generic load
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(lsAddr + (uint32_t)l * 4) : "memory");
    return v;
  };
  const int dlo = blockIdx.x * planesPerCta, np = NP1 ? 1 : min(planesPerCta, Dp - dlo);   // NP1: one plane per CTA
  if (np <= 0) return;
  const int tid = threadIdx.x, nth = blockDim.x;
  const size_t lvB = (size_t)cap * np;
  auto at = [&](int d, int q, int v) -> double* {   // element of plane d: position q of `order` = node v
    return LEVELORD ? A + (size_t)d * sd + q : A + (size_t)d * sd + (size_t)v * sn;
  };

  // ---- leaf to root: backup[p] += sum over children (adjacency order) of w(c) * backup[c]
  auto loadUp = [&](int l, int t, nl_up_item& it) {
    it.v = -1; it.nc = 0;
    if (l < 0) return;
    const int lo = lstart(l), cnt = lstart(l + 1) - lo;
    if (t < cnt * np) {
      const int i = NP1 ? t : t / np, d = dlo + (t - i * np);
      it.v = LEVELORD ? 0 : order[lo + i];
      it.nc = rnc[lo + i];
      it.cp = rcp[lo + i];
      it.cw = rcw[lo + i];
      it.own = *at(d, lo + i, it.v);   // this node's own cost: written before the kernel
    }
  };
  {
    // The items fetched ahead live in four NAMED register sets used in rotation (the level loop is unrolled by four):
    // shifting them through an array would copy registers whose loads are still in flight, and such a copy waits.
    auto stepUp = [&](int l, nl_up_item& slot) {
      if (l < 0) return;
      const int lo = lstart(l), cnt = lstart(l + 1) - lo;
      double* cur = lv + (size_t)(l & 1) * lvB;
      const double* kid = lv + (size_t)((l + 1) & 1) * lvB;
      nl_up_item it = slot;
      loadUp(l - NL_TF_PF, tid, slot);   // in flight while the next levels are computed
      for (int t = tid; t < cnt * np; t += nth) {
        if (t != tid) loadUp(l, t, it);
        const int i = NP1 ? t : t / np, dd = t - i * np, d = dlo + dd;
        double acc = it.own;
        const int cp[4] = {it.cp.x, it.cp.y, it.cp.z, it.cp.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {   // unrolled: cp[] stays in registers (a node has at most 4 children)
          if (k < it.nc) {
            const int p = cp[k];
            double val;
            if (p < cap) val = kid[(size_t)p * np + dd];
            else { const int q = lstart(l + 1) + p; val = *at(d, q, LEVELORD ? 0 : order[q]); }
            acc += val * tab[(it.cw >> (8 * k)) & 0xff];
          }
        }
        if (it.nc) *at(d, lo + i, it.v) = acc;
        if (i < cap) cur[(size_t)i * np + dd] = acc;
      }
      __syncthreads();
    };
    nl_up_item p0, p1, p2, p3;
    loadUp(nlevels - 1, tid, p0); loadUp(nlevels - 2, tid, p1); loadUp(nlevels - 3, tid, p2); loadUp(nlevels - 4, tid, p3);
    for (int l = nlevels - 1; l >= 0; l -= 4) { stepUp(l, p0); stepUp(l - 1, p1); stepUp(l - 2, p2); stepUp(l - 3, p3); }
  }
  // ---- root to leaf: cost[i] = w * (cost[parent] - w * backup[i]) + backup[i]   (the root keeps backup)
  auto loadDn = [&](int l, int t, nl_dn_item& it) {
    it.v = -1;
    if (l >= nlevels) return;
    const int lo = lstart(l), cnt = lstart(l + 1) - lo;
    if (t < cnt * np) {
      const int i = NP1 ? t : t / np, d = dlo + (t - i * np);
      it.v = LEVELORD ? 0 : order[lo + i];
      it.pp = rpp[lo + i];
      it.w = rw[lo + i];
      it.own = *at(d, lo + i, it.v);   // backup[i]: written by the sweep above, at least NL_TF_PF barriers ago
    }
  };
  // level 0 (the root) is final already and sits in lv[0] from the sweep above.  The first NL_TF_PF levels are
  // fetched only now: their backup values were written by the last iterations of the sweep above.
  {
    auto stepDn = [&](int l, nl_dn_item& slot) {
      if (l >= nlevels) return;
      const int lo = lstart(l), cnt = lstart(l + 1) - lo;
      double* cur = lv + (size_t)(l & 1) * lvB;
      const double* par = lv + (size_t)((l - 1) & 1) * lvB;
      nl_dn_item it = slot;
      loadDn(l + NL_TF_PF, tid, slot);
      for (int t = tid; t < cnt * np; t += nth) {
        if (t != tid) loadDn(l, t, it);
        const int i = NP1 ? t : t / np, dd = t - i * np, d = dlo + dd;
        const double w = tab[it.w];
        const double b = it.own;
        double pv;
        if (it.pp < cap) pv = par[(size_t)it.pp * np + dd];
        else { const int q = lstart(l - 1) + it.pp; pv = *at(d, q, LEVELORD ? 0 : order[q]); }
        const double r = w * (pv - w * b) + b;
        *at(d, lo + i, it.v) = r;
        if (i < cap) cur[(size_t)i * np + dd] = r;
      }
      __syncthreads();
    };
    nl_dn_item p0, p1, p2, p3;
    loadDn(1, tid, p0); loadDn(2, tid, p1); loadDn(3, tid, p2); loadDn(4, tid, p3);
    for (int l = 1; l < nlevels; l += 4) { stepDn(l, p0); stepDn(l + 1, p1); stepDn(l + 2, p2); stepDn(l + 3, p3); }
  }
}

// ---- one CTA per plane, streamed -----------------------------------------------------------------------------------
// k_tf_sweeps spends ~1700 cycles per level although a level is 70 nodes (at most a few hundred): its per-thread
// register prefetch, 64-bit addressing and level-bound lookups cost every warp ~70 instructions per level, and the
// records come through dependent global loads.  Here the static records and the plane's own values -- level-ordered,
// i.e. read LINEARLY by the two sweeps (descending, then ascending) -- are streamed through a 3-stage shared-memory
// ring by cp.async in chunks of 1024 positions, independent of the level bounds, so no global load sits in the
// per-level chain; a level is processed piece by piece along the chunk boundaries (one node per thread), the values of
// the adjacent level stay in two shared-memory level buffers, all level bounds are in shared memory, and a level costs
// one __syncthreads.  Per (node, plane) the operations and their order are those of k_tf_sweeps, i.e. the
// reference's: bit-identical.
#define TFW_T 256          // threads per plane
#define TFW_C 1024         // positions per chunk
#define TFW_S 3            // stages
#define TFW_CAP 1024       // nodes of a level kept in shared memory (wider levels read the excess back from global memory)
__device__ __forceinline__ void tfw_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void tfw_cp8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void tfw_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tfw_wait_oldest() { asm volatile("cp.async.wait_group %0;" ::"n"(TFW_S - 1) : "memory"); }
__device__ __forceinline__ void tfw_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ double tfw_ldcg(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(TFW_T)
    k_tf_cta(double* __restrict__ A, int N, int Dp, int lsCap, const int* __restrict__ level_start,
             const uint4* __restrict__ rup, const int2* __restrict__ rdn, const double* __restrict__ table,
             const nl_sync* __restrict__ s) {
  extern __shared__ __align__(16) uint8_t tfw_smem[];
  __shared__ double tab[256];
  const int tid = threadIdx.x, nth = blockDim.x;
  const int d = blockIdx.x;
  const int nlevels = s->nlevels;
  double* lvb = reinterpret_cast<double*>(tfw_smem);                                              // [2][CAP]
  uint4* ringRec = reinterpret_cast<uint4*>(tfw_smem + 2 * (size_t)TFW_CAP * 8);                   // [S*C][2] (up) / int2 [S*C] (down)
  double* ringOwn = reinterpret_cast<double*>(tfw_smem + 2 * (size_t)TFW_CAP * 8 + (size_t)TFW_S * TFW_C * 32);   // [S*C]
  int* ls = reinterpret_cast<int*>(tfw_smem + 2 * (size_t)TFW_CAP * 8 + (size_t)TFW_S * TFW_C * 40);            // [lsCap]
  for (int i = tid; i < 256; i += nth) tab[i] = table[i];
  for (int i = tid; i <= nlevels && i < lsCap; i += nth) ls[i] = level_start[i];
  __syncthreads();
  auto lstart = [&](int l) -> int { return l < lsCap ? ls[l] : level_start[l]; };
  const uint32_t recA = (uint32_t)__cvta_generic_to_shared(ringRec), ownA = (uint32_t)__cvta_generic_to_shared(ringOwn);
  double* Ad = A + (size_t)d * N;
  const int kTop = (N - 1) / TFW_C;
  auto ring_index = [&](int q) -> int { return ((q / TFW_C) % TFW_S) * TFW_C + (q % TFW_C); };

  // ------------------------------------------------------------------ leaf to root (positions descending)
  auto issue_up = [&](int k) {
    if (k >= 0 && k <= kTop) {
      const int st = k % TFW_S, base = k * TFW_C;
      for (int i = tid; i < TFW_C * 2; i += nth)
        if (base + (i >> 1) < N) tfw_cp16(recA + (uint32_t)(st * TFW_C * 2 + i) * 16u, rup + (size_t)base * 2 + i);
      for (int i = tid; i < TFW_C; i += nth)
        if (base + i < N) tfw_cp8(ownA + (uint32_t)(st * TFW_C + i) * 8u, Ad + base + i);
    }
    tfw_commit();   // an empty group keeps the count uniform
  };
  for (int j = 0; j < TFW_S; j++) issue_up(kTop - j);
  int curChunk = kTop + 1;   // chunks >= curChunk are done with; chunk curChunk - 1 is the oldest group in flight
  {
    int hi = N;              // = level_start[nlevels]
    for (int l = nlevels - 1; l >= 0; l--) {
      const int lo = lstart(l);
      double* cur = lvb + (size_t)(l & 1) * TFW_CAP;
      const double* kid = lvb + (size_t)((l + 1) & 1) * TFW_CAP;
      const int kidLo = hi;  // first position of level l + 1
      for (int c = (hi - 1) / TFW_C; c >= lo / TFW_C && hi > lo; c--) {
        while (curChunk > c) {           // first touch of chunk curChunk - 1: it is the oldest group in flight
          if (curChunk <= kTop) { __syncthreads(); issue_up(curChunk - TFW_S); }   // chunk curChunk is done with: refill its stage
          tfw_wait_oldest();
          __syncthreads();
          curChunk--;
        }
        const int p0 = max(lo, c * TFW_C), p1 = min(hi, (c + 1) * TFW_C);
        for (int q = p0 + tid; q < p1; q += nth) {
          const int ri = ring_index(q);
          const uint4 r0 = ringRec[2 * ri], r1 = ringRec[2 * ri + 1];
          double acc = ringOwn[ri];
          const int nc = (int)r1.y;
          const int cp[4] = {(int)r0.x, (int)r0.y, (int)r0.z, (int)r0.w};
#pragma unroll
          for (int k = 0; k < 4; k++) {
            if (k < nc) {
              const int p = cp[k];
              const double val = p < TFW_CAP ? kid[p] : tfw_ldcg(Ad + kidLo + p);
              acc += val * tab[(r1.x >> (8 * k)) & 0xff];
            }
          }
          if (nc) Ad[q] = acc;
          if (q - lo < TFW_CAP) cur[q - lo] = acc;
        }
      }
      __syncthreads();   // the level's values are visible to the CTA before the next level reads them
      hi = lo;
    }
  }
  tfw_wait_all();
  __threadfence();   // the backup values written above are read back through cp.async below
  __syncthreads();

  // ------------------------------------------------------------------ root to leaf (positions ascending)
  auto issue_dn = [&](int k) {
    if (k >= 0 && k <= kTop) {
      const int st = k % TFW_S, base = k * TFW_C;
      for (int i = tid; i < TFW_C; i += nth)
        if (base + i < N) {
          tfw_cp8(recA + (uint32_t)(st * TFW_C + i) * 8u, rdn + base + i);
          tfw_cp8(ownA + (uint32_t)(st * TFW_C + i) * 8u, Ad + base + i);
        }
    }
    tfw_commit();
  };
  const int2* ringDn = reinterpret_cast<const int2*>(ringRec);
  for (int j = 0; j < TFW_S; j++) issue_dn(j);
  curChunk = -1;             // chunks <= curChunk are done with; chunk curChunk + 1 is the oldest group in flight
  {
    int lo = 0, prevLo = 0;
    for (int l = 0; l < nlevels; l++) {
      const int hi = lstart(l + 1);
      double* cur = lvb + (size_t)(l & 1) * TFW_CAP;
      const double* par = lvb + (size_t)((l - 1) & 1) * TFW_CAP;
      for (int c = lo / TFW_C; c <= (hi - 1) / TFW_C && hi > lo; c++) {
        while (curChunk < c) {
          if (curChunk >= 0) { __syncthreads(); issue_dn(curChunk + TFW_S); }
          tfw_wait_oldest();
          __syncthreads();
          curChunk++;
        }
        if (l == 0) continue;            // the root keeps its backup value (already in A and in the level buffer)
        const int p0 = max(lo, c * TFW_C), p1 = min(hi, (c + 1) * TFW_C);
        for (int q = p0 + tid; q < p1; q += nth) {
          const int ri = ring_index(q);
          const int2 rec = ringDn[ri];
          const double b = ringOwn[ri];
          const double w = tab[rec.y];
          const double pv = rec.x < TFW_CAP ? par[rec.x] : tfw_ldcg(Ad + prevLo + rec.x);
          const double r = w * (pv - w * b) + b;
          Ad[q] = r;
          if (q - lo < TFW_CAP) cur[q - lo] = r;
        }
      }
      __syncthreads();
      prevLo = lo;
      lo = hi;
    }
  }
  tfw_wait_all();
}

// ---- one CTA per plane, streamed, lean level loop ------------------------------------------------------------------
// ncu on k_tf_cta (profiles/r01_nl_fast_ncu.md): every warp executes ~90 instructions of loop bookkeeping per level
// (chunk range of the level, ring index by division, bounds with an l < lsCap test, uniform-register traffic) before the
// ~60 of the node itself, at 0.17 IPC -- a single dependent chain, so time per level = instructions on that chain x ~6
// cycles.  Nearly every level is narrow (<= one node per thread), lies inside the chunk the sweep is already in, and
// has its neighbour level wholly in the shared-memory level buffer.  Whether that holds is a property of the level
// bounds alone, so it is decided once, when the bounds are copied to shared memory (two flag bits in the entry), and such
// a level runs a straight-line body: position = lo + tid, ring slot = position + a per-chunk offset, children read
// without the wide-level fallback (no branch inside), one barrier.  Every other level (chunk change, wide level) takes
// k_tf_cta's general code.  Per (node, plane) the operations and their order are unchanged: bit-identical.
__device__ __forceinline__ uint4 tff_lds16(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint2 tff_lds8(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t tff_lds4(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ double tff_ldsd(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void tff_stsd(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
#define TFF_UP 0x80000000u
#define TFF_DN 0x40000000u
#define TFF_LO 0x3fffffff
__device__ __forceinline__ void tff_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tff_mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tff_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tff_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__global__ void __launch_bounds__(TFW_T)
    k_tf_cta_fast(double* __restrict__ A, int N, int Dp, int lsCap, const int* __restrict__ level_start,
                  const uint4* __restrict__ rup, const int2* __restrict__ rdn, const double* __restrict__ table,
                  const nl_sync* __restrict__ s) {
  extern __shared__ __align__(16) uint8_t tfw_smem[];
  __shared__ double tab[256];
  __shared__ __align__(8) uint64_t bars[2 * TFW_S];   // one mbarrier per ring stage and sweep
  const int tid = threadIdx.x, nth = blockDim.x;
  const int d = blockIdx.x;
  const int nlevels = s->nlevels;   // the launcher guarantees nlevels < lsCap: every bound sits in shared memory
  double* lvb = reinterpret_cast<double*>(tfw_smem);                                              // [2][CAP]
  uint4* ringRec = reinterpret_cast<uint4*>(tfw_smem + 2 * (size_t)TFW_CAP * 8);                   // [S*C][2] (up) / int2 [S*C] (down)
  double* ringOwn = reinterpret_cast<double*>(tfw_smem + 2 * (size_t)TFW_CAP * 8 + (size_t)TFW_S * TFW_C * 32);   // [S*C]
  // [2 + nlevels + 1 + 2]: two pad entries on either side, so that the bounds fetched two levels ahead need no clamp
  uint32_t* ls = reinterpret_cast<uint32_t*>(tfw_smem + 2 * (size_t)TFW_CAP * 8 + (size_t)TFW_S * TFW_C * 40) + 2;
  for (int i = tid; i < 256; i += nth) tab[i] = table[i];
  if (tid < 2) { ls[-1 - tid] = 0u; ls[nlevels + 1 + tid] = (uint32_t)N; }
  for (int l = tid; l <= nlevels; l += nth) {
    const int lo = level_start[l];
    uint32_t e = (uint32_t)lo;
    if (l < nlevels) {
      const int hi = level_start[l + 1];
      const bool one = hi > lo && hi - lo <= nth && lo / TFW_C == (hi - 1) / TFW_C;
      // up sweep: level l + 1 was processed just before (its last piece lies in chunk hi / C) and feeds this one
      if (one && l + 1 < nlevels && hi / TFW_C == lo / TFW_C && level_start[l + 2] - hi <= TFW_CAP) e |= TFF_UP;
      // down sweep: level l - 1 was processed just before (its last piece lies in chunk (lo - 1) / C)
      if (one && l >= 1 && (lo - 1) / TFW_C == lo / TFW_C && lo - level_start[l - 1] <= TFW_CAP) e |= TFF_DN;
    }
    ls[l] = e;
  }
  const uint32_t barA = (uint32_t)__cvta_generic_to_shared(bars);
  if (tid == 0) {
    for (int i = 0; i < 2 * TFW_S; i++) tff_mbar_init(barA + i * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const uint32_t recA = (uint32_t)__cvta_generic_to_shared(ringRec), ownA = (uint32_t)__cvta_generic_to_shared(ringOwn);
  // the straight-line level body addresses shared memory by 32-bit window offsets (through generic pointers the compiler
  // rebuilt the window base from SR_CgaCtaId at every level, on the dependent chain)
  // (and passes them through an opaque move: otherwise they are rematerialised inside the loop just the same)
  auto opaque = [](uint32_t x) -> uint32_t { uint32_t y; asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x)); return y; };
  const uint32_t lvbA = opaque((uint32_t)__cvta_generic_to_shared(lvb)), tabA = opaque((uint32_t)__cvta_generic_to_shared(tab));
  const uint32_t lsA = opaque((uint32_t)__cvta_generic_to_shared(ls)), myA = opaque((uint32_t)tid * 8u);
  const uint32_t recF = opaque(recA), ownF = opaque(ownA);
  const int tidF = (int)opaque((uint32_t)tid);
  double* Ad = A + (size_t)d * N;
  asm volatile("" : "+l"(Ad));   // one 64-bit base: a store address is then a single IMAD.WIDE
  const int kTop = (N - 1) / TFW_C;
  auto ring_index = [&](int q) -> int { return ((q / TFW_C) % TFW_S) * TFW_C + (q % TFW_C); };

  // ------------------------------------------------------------------ leaf to root (positions descending)
  // a chunk = 1024 consecutive positions = one contiguous run of records and one of values: two bulk copies by the TMA
  // unit onto the stage's mbarrier, issued by one thread (the cp.async version spent ~0.9 us of all eight warps per chunk
  // on 3072 per-thread copies).  N is even (launcher), so every run is a multiple of 16 bytes and 16-byte aligned.
  auto issue_up = [&](int k) {
    if (k >= 0 && k <= kTop) {
      const int st = k % TFW_S, base = k * TFW_C, cnt = min(TFW_C, N - base);
      const uint32_t bar = barA + st * 8;
      tff_mbar_expect(bar, (uint32_t)cnt * 40u);
      tff_bulk(recA + (uint32_t)(st * TFW_C) * 32u, rup + (size_t)base * 2, (uint32_t)cnt * 32u, bar);
      tff_bulk(ownA + (uint32_t)(st * TFW_C) * 8u, Ad + base, (uint32_t)cnt * 8u, bar);
    }
  };
  if (tid == 0)
    for (int j = 0; j < TFW_S; j++) issue_up(kTop - j);
  int curChunk = kTop + 1;   // chunks >= curChunk are done with; chunk curChunk - 1 is the oldest group in flight
  int ringOff = 0;           // ring slot of position q inside the current chunk = q + ringOff
  {
    int hi = N;              // = level_start[nlevels]
    uint32_t e = ls[nlevels - 1], eN = ls[nlevels - 2];
    uint32_t lsP = lsA + (uint32_t)(nlevels - 3) * 4u;   // the entry fetched two levels ahead
    uint32_t curA = lvbA + (uint32_t)((nlevels - 1) & 1) * (TFW_CAP * 8u), kidA = lvbA + (uint32_t)(nlevels & 1) * (TFW_CAP * 8u);
    for (int l = nlevels - 1; l >= 0; l--) {
      const uint32_t eNN = tff_lds4(lsP);
      const int lo = (int)(e & TFF_LO);
      if (e & TFF_UP) {
        const int q = lo + tidF;
        if (q < hi) {
          const uint32_t ri = (uint32_t)(q + ringOff);
          const uint4 r0 = tff_lds16(recF + ri * 32u);
          const uint2 r1 = tff_lds8(recF + ri * 32u + 16u);
          const double own = tff_ldsd(ownF + ri * 8u);
          const double v0 = tff_ldsd(kidA + r0.x * 8u), w0 = tff_ldsd(tabA + (r1.x & 0xffu) * 8u);
          const double v1 = tff_ldsd(kidA + r0.y * 8u), w1 = tff_ldsd(tabA + ((r1.x >> 8) & 0xffu) * 8u);
          const double v2 = tff_ldsd(kidA + r0.z * 8u), w2 = tff_ldsd(tabA + ((r1.x >> 16) & 0xffu) * 8u);
          const double v3 = tff_ldsd(kidA + r0.w * 8u), w3 = tff_ldsd(tabA + (r1.x >> 24) * 8u);
          const int nc = (int)r1.y;
          // x + (-0.0) == x for every x (either zero included): a missing child adds -0.0, so the four additions form
          // one unconditional chain and the selects sit beside it, not in it
          const double t0 = nc > 0 ? v0 * w0 : -0.0;
          const double t1 = nc > 1 ? v1 * w1 : -0.0;
          const double t2 = nc > 2 ? v2 * w2 : -0.0;
          const double t3 = nc > 3 ? v3 * w3 : -0.0;
          double acc = own + t0;
          acc += t1;
          acc += t2;
          acc += t3;
          tff_stsd(curA + myA, acc);
          if (nc) asm volatile("st.global.f64 [%0], %1;" ::"l"(Ad + q), "d"(acc) : "memory");
        }
      } else {
        double* cur = lvb + (size_t)(l & 1) * TFW_CAP;
        const double* kid = lvb + (size_t)((l + 1) & 1) * TFW_CAP;
        const int kidLo = hi;  // first position of level l + 1
        for (int c = (hi - 1) / TFW_C; c >= lo / TFW_C && hi > lo; c--) {
          while (curChunk > c) {           // first touch of chunk curChunk - 1
            if (curChunk <= kTop) {          // chunk curChunk is done with: refill its stage (read only, no proxy fence)
              __syncthreads();
              if (tid == 0) issue_up(curChunk - TFW_S);
            }
            curChunk--;
            tff_mbar_wait(barA + (curChunk % TFW_S) * 8, (uint32_t)((kTop - curChunk) / TFW_S) & 1u);
          }
          const int p0 = max(lo, c * TFW_C), p1 = min(hi, (c + 1) * TFW_C);
          for (int q = p0 + tid; q < p1; q += nth) {
            const int ri = ring_index(q);
            const uint4 r0 = ringRec[2 * ri], r1 = ringRec[2 * ri + 1];
            double acc = ringOwn[ri];
            const int nc = (int)r1.y;
            const int cp[4] = {(int)r0.x, (int)r0.y, (int)r0.z, (int)r0.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
              if (k < nc) {
                const int p = cp[k];
                const double val = p < TFW_CAP ? kid[p] : tfw_ldcg(Ad + kidLo + p);
                acc += val * tab[(r1.x >> (8 * k)) & 0xff];
              }
            }
            if (nc) Ad[q] = acc;
            if (q - lo < TFW_CAP) cur[q - lo] = acc;
          }
        }
        ringOff = (curChunk % TFW_S) * TFW_C - curChunk * TFW_C;
      }
      // rotate the bounds here, not at the top: the entry loaded at the top has arrived by now
      e = eN;
      asm volatile("mov.u32 %0, %1;" : "=r"(eN) : "r"(eNN));
      __syncthreads();   // the level's values are visible to the CTA before the next level reads them
      hi = lo;
      lsP -= 4u;
      { const uint32_t t = curA; curA = kidA; kidA = t; }
    }
  }
  // the backup values written above (generic proxy) are read back by the TMA unit (async proxy) below
  __threadfence();
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncthreads();

  // ------------------------------------------------------------------ root to leaf (positions ascending)
  auto issue_dn = [&](int k) {
    if (k >= 0 && k <= kTop) {
      const int st = k % TFW_S, base = k * TFW_C, cnt = min(TFW_C, N - base);
      const uint32_t bar = barA + (TFW_S + st) * 8;
      tff_mbar_expect(bar, (uint32_t)cnt * 16u);
      tff_bulk(recA + (uint32_t)(st * TFW_C) * 8u, rdn + base, (uint32_t)cnt * 8u, bar);
      tff_bulk(ownA + (uint32_t)(st * TFW_C) * 8u, Ad + base, (uint32_t)cnt * 8u, bar);
    }
  };
  const int2* ringDn = reinterpret_cast<const int2*>(ringRec);
  if (tid == 0)
    for (int j = 0; j < TFW_S; j++) issue_dn(j);
  curChunk = -1;             // chunks <= curChunk are done with; chunk curChunk + 1 is the oldest group in flight
  {
    int lo = 0, prevLo = 0;
    uint32_t e = ls[0], eH = ls[1], eHN = ls[2];
    uint32_t lsP = lsA + 3u * 4u;   // the entry fetched two levels ahead
    uint32_t curA = lvbA, parA = lvbA + TFW_CAP * 8u;   // level l -> buffer l & 1; level l - 1 -> the other one
    for (int l = 0; l < nlevels; l++) {
      const uint32_t eHNN = tff_lds4(lsP);
      const int hi = (int)(eH & TFF_LO);
      if (e & TFF_DN) {
        const int q = lo + tidF;
        if (q < hi) {
          const uint32_t ri = (uint32_t)(q + ringOff);
          const uint2 rec = tff_lds8(recF + ri * 8u);
          const double b = tff_ldsd(ownF + ri * 8u);
          const double w = tff_ldsd(tabA + rec.y * 8u);
          const double pv = tff_ldsd(parA + rec.x * 8u);
          const double r = w * (pv - w * b) + b;
          tff_stsd(curA + myA, r);
          asm volatile("st.global.f64 [%0], %1;" ::"l"(Ad + q), "d"(r) : "memory");
        }
      } else {
        double* cur = lvb + (size_t)(l & 1) * TFW_CAP;
        const double* par = lvb + (size_t)((l - 1) & 1) * TFW_CAP;
        for (int c = lo / TFW_C; c <= (hi - 1) / TFW_C && hi > lo; c++) {
          while (curChunk < c) {
            if (curChunk >= 0) {
              __syncthreads();
              if (tid == 0) issue_dn(curChunk + TFW_S);
            }
            curChunk++;
            tff_mbar_wait(barA + (TFW_S + curChunk % TFW_S) * 8, (uint32_t)(curChunk / TFW_S) & 1u);
          }
          if (l == 0) continue;            // the root keeps its backup value (already in A and in the level buffer)
          const int p0 = max(lo, c * TFW_C), p1 = min(hi, (c + 1) * TFW_C);
          for (int q = p0 + tid; q < p1; q += nth) {
            const int ri = ring_index(q);
            const int2 rec = ringDn[ri];
            const double b = ringOwn[ri];
            const double w = tab[rec.y];
            const double pv = rec.x < TFW_CAP ? par[rec.x] : tfw_ldcg(Ad + prevLo + rec.x);
            const double r = w * (pv - w * b) + b;
            Ad[q] = r;
            if (q - lo < TFW_CAP) cur[q - lo] = r;
          }
        }
        ringOff = (curChunk % TFW_S) * TFW_C - curChunk * TFW_C;
      }
      e = eH; eH = eHN;
      asm volatile("mov.u32 %0, %1;" : "=r"(eHN) : "r"(eHNN));
      __syncthreads();
      prevLo = lo;
      lo = hi;
      lsP += 4u;
      { const uint32_t t = curA; curA = parA; parA = t; }
    }
  }
}

__global__ void __launch_bounds__(256)
    k_tf_store(const double* __restrict__ A, float* __restrict__ vol, size_t N, int D, int Dp, const int* __restrict__ order) {
  __shared__ float tile[TFX_D][TFX_Q + 1];
  __shared__ float wet[TFX_Q];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (size_t q0 = (size_t)blockIdx.x * TFX_Q; q0 < N; q0 += (size_t)gridDim.x * TFX_Q) {
    if (Dp > D && warp == 0) wet[lane] = q0 + lane < N ? (float)A[(size_t)D * N + q0 + lane] : 1.0f;
    for (int d0 = 0; d0 < D; d0 += TFX_D) {
      const int nd = min(TFX_D, D - d0);
      for (int d = warp; d < nd; d += 8)
        if (q0 + lane < N) tile[d][lane] = (float)A[(size_t)(d0 + d) * N + q0 + lane];   // NLCCA::aggreCV: (float)nlcP[d]
      __syncthreads();
      for (int k = warp; k < TFX_Q; k += 8) {
        const size_t q = q0 + k;
        if (q < N) {
          float* dst = vol + (size_t)order[q] * D + d0;
          const float w = Dp > D ? wet[k] : 1.0f;
          for (int d = lane; d < nd; d += 32) {
            const float x = tile[d][k];
            dst[d] = Dp > D ? x / w : x;                                                  // StereoMatching::NL: vm[0] /= wetNL
          }
        }
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------ host side
struct nl_tree {   // device buffers of one rooted tree (owned by the ctx scratch slots)
  int *parent, *rank, *order, *level_start, *child;
  uint8_t *wpar, *nchild;
  nl_sync* sync;
  // level-ordered records (k_tf_records)
  int4* rcp; uint32_t* rcw; int* rpp; int* pos; uint8_t *rnc, *rw;
  uint4* rup;   // [N][2]: {cp0, cp1, cp2, cp3}, {cw, nc, 0, 0}   (k_tf_warp: one 32-byte record per position)
  int2* rdn;    // [N]: {pp, w}
};


__global__ void k_root_single(int* parent, uint8_t* wpar, int* rank, int* order, int* level_start, nl_sync* s) {
  parent[0] = 0; wpar[0] = 0; rank[0] = 0; order[0] = 0; level_start[0] = 0; level_start[1] = 1; s->nlevels = 1;
}
__global__ void k_root_fix(int* parent, uint8_t* wpar) { parent[0] = 0; wpar[0] = 0; }

static int nl_root_euler(sm_ctx* ctx, int N, const int* nbr, const uint8_t* nbw, const uint8_t* deg, int* parent,
                         uint8_t* wpar, int* rank, int* order, int* level_start, nl_sync* sync) {
  if (N == 1) { SM_LAUNCH(ctx, k_root_single, 1, 1, 0, parent, wpar, rank, order, level_start, sync); return SM_OK; }
  const int T = 2 * (N - 1), n4 = 4 * N, TB = 256;
  const int nt = sm_div_up(N, NLP_TILE);            // tiles of a sort pass
  const int nh = 256 * nt;                          // histogram entries of a pass
  // (succ, dist) x 2 (4N int2 each) ; pm | node_at | scan (T) ; keys | keys2 | vals | vals2 (N) ;
  // hist | offs (256 nt) ; sums (tiles of the longest scan)
  const int nsums = max(sm_div_up(T, NLP_TILE), sm_div_up(nh, NLP_TILE)) + 8;
  const size_t ints = (size_t)4 * n4 + (size_t)3 * T + (size_t)4 * N + (size_t)2 * nh + nsums + 64;
  void* p;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_NLEULER, ints * 4, &p));
  int2* sdA = (int2*)p; int2* sdB = sdA + n4;   // (succ, dist) ping-pong: 4 n4 ints (the scratch base is 256-byte aligned)
  int* pm = (int*)(sdB + n4); int* node_at = pm + T; int* scan = node_at + T;
  unsigned* keys = (unsigned*)(scan + T); unsigned* keys2 = keys + N;
  int* vals = (int*)(keys2 + N); int* vals2 = vals + N;
  int* hist = vals2 + N; int* offs = hist + nh; int* sums = offs + nh;
  // ~50 dependent launches of a few microseconds each, every argument fixed by (N, the buffers): replayed as a CUDA graph
  // from the third frame on (smi_graphed)
  unsigned long long key = smi_key_mix(14695981039346656037ull, (unsigned long long)N);
  for (const void* q : {(const void*)p, (const void*)nbr, (const void*)nbw, (const void*)deg, (const void*)parent, (const void*)wpar,
                        (const void*)rank, (const void*)order, (const void*)level_start, (const void*)sync})
    key = smi_key_mix(key, (unsigned long long)(uintptr_t)q);
  unsigned* const keys0 = keys; unsigned* const keys20 = keys2; int* const vals0 = vals; int* const vals20 = vals2;
  return smi_graphed(ctx, ctx->graphs[SM_GRAPH_EULER], key, [&]() -> int {
  // (the body may run twice -- capture, then an eager retry: it starts from the same buffers each time)
  int2 *sd = sdA, *sd2 = sdB;
  int *vals = vals0, *vals2 = vals20;
  unsigned *keys = keys0, *keys2 = keys20;
  SM_LAUNCH(ctx, k_et_init, sm_div_up(n4, TB), TB, 0, N, (const int4*)nbr, deg, sd);
  for (long long span = 1; span < T; span *= 2) {   // list ranking: ceil(log2 T) jumps
    SM_LAUNCH(ctx, k_et_jump, sm_div_up(n4, TB), TB, 0, n4, (const int2*)sd, sd2);
    int2* t = sd; sd = sd2; sd2 = t;
  }
  SM_LAUNCH(ctx, k_et_classify, sm_div_up(n4, TB), TB, 0, N, T, (const int4*)nbr, (const uchar4*)nbw, deg, (const int2*)sd, parent, wpar,
            pm, node_at);
  SM_LAUNCH(ctx, k_root_fix, 1, 1, 0, parent, wpar);
  SM_TRY(nl_scan<false>(ctx, pm, scan, T, sums));   // +1 / -1 along the tour: the value at a downward edge is the depth
  SM_LAUNCH(ctx, k_et_depth, sm_div_up(T + 1, TB), TB, 0, T, scan, node_at, rank);
  // level order = the nodes sorted by depth (stable, from node order): depth < N, 8 bits per pass
  SM_LAUNCH(ctx, k_rs_init, sm_div_up(N, TB), TB, 0, N, rank, keys, vals);
  int bits = 1;
  while (bits < 31 && (1ll << bits) < (long long)N) bits++;
  const int passes = sm_div_up(bits, 8);
  for (int ps = 0; ps < passes; ps++) {
    int* vout = ps == passes - 1 ? order : vals2;   // the last pass writes the level order itself
    SM_LAUNCH(ctx, k_rs_hist, nt, NLP_TB, 0, keys, N, 8 * ps, hist, nt);
    SM_TRY(nl_scan<true>(ctx, hist, offs, nh, sums));
    SM_LAUNCH(ctx, k_rs_scatter, nt, NLP_TB, 0, keys, vals, keys2, vout, N, 8 * ps, offs, nt);
    unsigned* tk = keys; keys = keys2; keys2 = tk;
    int* tv = vals; vals = vals2; vals2 = tv;
  }
  SM_LAUNCH(ctx, k_level_bounds, min(sm_div_up(N, TB), ctx->num_sms * 8), TB, 0, N, rank, order, level_start, sync);
  return SM_OK;
  });
}

// MST + rooting.  img: [H][W][cn] u8 (already median-filtered).  Fills t (buffers must be allocated).
static int nl_build_tree(sm_ctx* ctx, const uint8_t* d_img, int H, int W, int cn, nl_tree& t) {
  const int N = H * W, E = H * (W - 1) + W * (H - 1);
  void *p_ew, *p_in, *p_comp, *p_link, *p_best, *p_nbr, *p_nbw, *p_deg, *p_cnt;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_MISC0, (size_t)(E > 0 ? E : 1), &p_ew));
  SM_TRY(sm_scratch_get(ctx, SM_SCR_MISC1, (size_t)(E > 0 ? E : 1), &p_in));
  SM_TRY(sm_scratch_get(ctx, SM_SCR_MISC2, (size_t)N * 8, &p_comp));   // comp | link
  SM_TRY(sm_scratch_get(ctx, SM_SCR_MISC3, (size_t)N * 8, &p_best));
  SM_TRY(sm_scratch_get(ctx, SM_SCR_MISC4, (size_t)N * 16 + (size_t)N * 4 + N, &p_nbr));
  SM_TRY(sm_scratch_get(ctx, SM_SCR_MISC5, 256, &p_cnt));
  uint8_t *ew = (uint8_t*)p_ew, *inMST = (uint8_t*)p_in;
  int *comp = (int*)p_comp, *link = comp + N;
  unsigned long long* best = (unsigned long long*)p_best;
  int* nbr = (int*)p_nbr;
  p_nbw = (uint8_t*)p_nbr + (size_t)N * 16;
  p_deg = (uint8_t*)p_nbw + (size_t)N * 4;
  uint8_t *nbw = (uint8_t*)p_nbw, *deg = (uint8_t*)p_deg;
  int* cnt = (int*)p_cnt;   // [0] hooks, [1] changed

  const int TB = 256, gN = sm_div_up(N, TB), gE = max(1, min(sm_div_up(E, TB), ctx->num_sms * 16));
  if (E > 0) SM_LAUNCH(ctx, k_edge_weights, gE, TB, 0, d_img, H, W, cn, ew);
  SM_LAUNCH(ctx, k_bor_init, sm_div_up(max(N, E), TB), TB, 0, N, E, comp, link, best, inMST);
  // A round at least halves the number of components, image trees usually need 8-12.  Asking the device after EVERY round
  // whether anything hooked costs a host round trip each (the GPU idles ~20 us); a round on a finished forest is a no-op
  // (no edge leaves a component: nothing hooks, labels and links stay), so the first rounds run unconditionally and the
  // question is asked after every second round from the sixth on.
  auto half_round_a = [&]() -> int {   // find + hook
    SM_CUDA(cudaMemsetAsync(cnt, 0, 8, ctx->stream));
    SM_LAUNCH(ctx, k_bor_find, gE, TB, 0, ew, comp, H, W, best);
    SM_LAUNCH(ctx, k_bor_hook, gN, TB, 0, N, H, W, comp, best, link, inMST, cnt);
    return SM_OK;
  };
  auto half_round_b = [&]() -> int {   // every label points at its root
    SM_LAUNCH(ctx, k_bor_chase, gN, TB, 0, N, link);
    SM_LAUNCH(ctx, k_bor_relabel, gN, TB, 0, N, comp, link, best);
    SM_LAUNCH(ctx, k_bor_fixlink, gN, TB, 0, N, link);
    return SM_OK;
  };
  // the unconditional rounds 0..4 and the first half of round 5 are one fixed chain of 28 launches: a CUDA graph
  // (smi_graphed); the rounds with a question to the device follow as they were
  int round = 0;
  if (E > 0) {
    unsigned long long key = smi_key_mix(smi_key_mix(14695981039346656037ull, (unsigned long long)N), ((unsigned long long)H << 32) | (unsigned)W);
    for (const void* q : {(const void*)ew, (const void*)inMST, (const void*)comp, (const void*)best, (const void*)cnt})
      key = smi_key_mix(key, (unsigned long long)(uintptr_t)q);
    SM_TRY(smi_graphed(ctx, ctx->graphs[SM_GRAPH_BORUVKA], key, [&]() -> int {
      for (int r = 0; r < 5; r++) { SM_TRY(half_round_a()); SM_TRY(half_round_b()); }
      return half_round_a();
    }));
    round = 5;
  }
  for (; E > 0 && round < 40; round++) {
    if (round > 5) SM_TRY(half_round_a());
    if (round >= 5 && (round & 1)) {
      int h_cnt[2] = {0, 0};
      SM_CUDA(cudaMemcpyAsync(h_cnt, cnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
      SM_CUDA(cudaStreamSynchronize(ctx->stream));
      if (h_cnt[0] == 0) break;   // no component has an outgoing edge: the forest is the spanning tree
    }
    SM_TRY(half_round_b());
  }
  SM_LAUNCH(ctx, k_tree_adj, gN, TB, 0, H, W, ew, inMST, nbr, nbw, deg);
  SM_CUDA(cudaMemsetAsync(t.sync, 0, sizeof(nl_sync), ctx->stream));
  // rooting: Euler tour + list ranking (nl_root_euler).  (The level-by-level BFS it replaced -- one CTA, frontier in shared
  // memory, 5.15 ms for MST + rooting at 640x480 against 0.78 ms -- is gone; DESIGN.md keeps the numbers.)
  return nl_root_euler(ctx, N, nbr, nbw, deg, t.parent, t.wpar, t.rank, t.order, t.level_start, t.sync);
}

// scratch layout of the tree arrays the filter derives (children, levels, barrier state)
static int nl_tree_scratch(sm_ctx* ctx, int N, nl_tree& t, bool own_arrays) {
  // SM_SCR_ARM0: child[N][4] | level_start[N+2] ; SM_SCR_ARM1: nchild[N] | sync | (parent | rank | order | wpar)
  void *p0, *p1;
  const size_t s1 = (size_t)N + 64 + (own_arrays ? (size_t)N * 13 + 64 : 0);
  SM_TRY(sm_scratch_get(ctx, SM_SCR_ARM0, (size_t)N * 16 + (size_t)(N + 2) * 4, &p0));
  SM_TRY(sm_scratch_get(ctx, SM_SCR_ARM1, s1, &p1));
  t.child = (int*)p0;
  t.level_start = t.child + (size_t)N * 4;
  uint8_t* q = (uint8_t*)p1;
  t.sync = (nl_sync*)q; q += 64;
  if (own_arrays) {
    t.parent = (int*)q; q += (size_t)N * 4;
    t.rank = (int*)q; q += (size_t)N * 4;
    t.order = (int*)q; q += (size_t)N * 4;
    t.wpar = q; q += (size_t)N;
    q += (64 - ((size_t)N & 63)) & 63;
  }
  t.nchild = q;
  // SM_SCR_NLREC: rcp[N] int4 | rcw[N] | rpp[N] | pos[N] | rnc[N] | rw[N]
  void* p2;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_NLREC, (size_t)N * 70 + 128, &p2));
  uint8_t* r = (uint8_t*)p2;
  t.rup = (uint4*)r; r += (size_t)N * 32;    // 16-byte aligned arrays first (N may be odd)
  t.rcp = (int4*)r; r += (size_t)N * 16;
  t.rdn = (int2*)r; r += (size_t)N * 8;
  t.rcw = (uint32_t*)r; r += (size_t)N * 4;
  t.rpp = (int*)r; r += (size_t)N * 4;
  t.pos = (int*)r; r += (size_t)N * 4;
  t.rnc = r; r += (size_t)N;
  t.rw = r;
  return SM_OK;
}

// d_vol == nullptr: d_A already holds the float64 cost [N][D] and receives the result (qx_tree_filter::filter).
static int nl_filter(sm_ctx* ctx, float* d_vol, double* d_A, int H, int W, int D, bool ones_plane, const nl_tree& t,
                     double sigma) {
  const size_t N = (size_t)H * W;
  const int Dp = ones_plane ? D + 1 : D;
  // weight table, built on the host with the same libm call as update_table
  double h_tab[256];
  sigma = fmax(0.01, sigma);
  for (int i = 0; i <= 255; i++) h_tab[i] = exp(-double(i) / (255 * sigma));
  void* p_tab;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_TAB, 8192, &p_tab));   // [0, 4096): exp tables of the cost kernel
  double* d_tab = (double*)((uint8_t*)p_tab + 4096);
  SM_CUDA(cudaMemcpyAsync(d_tab, h_tab, sizeof(h_tab), cudaMemcpyHostToDevice, ctx->stream));
  SM_CUDA(cudaStreamSynchronize(ctx->stream));   // h_tab is a stack array
  const int TB = 256, g = (int)min((size_t)ctx->num_sms * 16, (N * Dp + TB - 1) / TB);
  const int n = (int)N, gr = min(sm_div_up(n, TB), ctx->num_sms * 8);
  SM_LAUNCH(ctx, k_tf_positions, gr, TB, 0, n, t.order, t.pos);
  SM_LAUNCH(ctx, k_tf_records, gr, TB, 0, n, t.order, t.pos, t.rank, t.level_start, t.parent, t.wpar, t.child, t.nchild,
            t.rcp, t.rcw, t.rnc, t.rpp, t.rw, t.rup, t.rdn);
  const int gx = (int)min((size_t)ctx->num_sms * 8, (N + TFX_Q - 1) / TFX_Q);   // tiles of 32 positions
  if (d_vol) SM_LAUNCH(ctx, k_tf_load, gx, 256, 0, d_vol, d_A, N, D, Dp, t.order);
  const size_t tfwFixed = 2 * (size_t)TFW_CAP * 8 + (size_t)TFW_S * TFW_C * 40;
  if (d_vol) {
    // one CTA per plane, records and values streamed (k_tf_cta); level bounds in shared memory as far as they fit
    int h_nlev = 0;
    SM_CUDA(cudaMemcpyAsync(&h_nlev, &t.sync->nlevels, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SM_CUDA(cudaStreamSynchronize(ctx->stream));
    int lsCap = (int)min((size_t)h_nlev + 1, (227 * 1024 - tfwFixed - 2048 - 64) / 4);
    const size_t smem = tfwFixed + (size_t)lsCap * 4 + 16;   // + two pad entries on either side of the bounds (k_tf_cta_fast)
    double* a = d_A;
    int n_ = (int)N, dp = Dp;
    void* args[] = {(void*)&a, (void*)&n_, (void*)&dp, (void*)&lsCap, (void*)&t.level_start, (void*)&t.rup, (void*)&t.rdn,
                    (void*)&d_tab, (void*)&t.sync};
    // the lean level loop needs every level bound in shared memory and an even N; otherwise the general code at every level
    const void* fn = (h_nlev < lsCap && N % 2 == 0) ? (const void*)k_tf_cta_fast : (const void*)k_tf_cta;
    SM_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SM_CUDA(cudaLaunchKernel(fn, dim3(Dp), dim3(TFW_T), args, smem, ctx->stream));
    ctx->launches++;
  } else {
    // one CTA per plane while there are SMs for them, else the same number of planes for every CTA
    const int ppc = sm_div_up(Dp, min(Dp, ctx->num_sms));
    const int grid = sm_div_up(Dp, ppc);
    const int cap = NL_TF_CAP / ppc;
    const size_t smem = 2 * (size_t)cap * ppc * sizeof(double) + (size_t)NL_TF_LS * sizeof(int);
    const void* fn = d_vol ? (ppc == 1 ? (const void*)k_tf_sweeps<true, true> : (const void*)k_tf_sweeps<true, false>)
                           : (ppc == 1 ? (const void*)k_tf_sweeps<false, true> : (const void*)k_tf_sweeps<false, false>);
    double* a = d_A;
    size_t sn = d_vol ? 1 : (size_t)D, sd = d_vol ? N : 1;
    int dp = Dp, pp = ppc, cp = cap;
    void* args[] = {(void*)&a, (void*)&sn, (void*)&sd, (void*)&dp, (void*)&pp, (void*)&cp, (void*)&t.order, (void*)&t.level_start,
                    (void*)&t.rcp, (void*)&t.rcw, (void*)&t.rnc, (void*)&t.rpp, (void*)&t.rw, (void*)&d_tab, (void*)&t.sync};
    SM_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // 256 threads: measured at 640x480 D=64: 512 -> 8.12 ms, 256 -> 7.78, 128 -> 8.31, 64 -> 11.07
    SM_CUDA(cudaLaunchKernel(fn, dim3(grid), dim3(256), args, smem, ctx->stream));
    ctx->launches++;
  }
  if (d_vol) SM_LAUNCH(ctx, k_tf_store, gx, 256, 0, d_A, d_vol, N, D, Dp, t.order);
  return SM_OK;
}

extern "C" int sm_mst_build(sm_ctx* ctx, const uint8_t* d_bgr, int H, int W, int cn, int32_t* d_parent,
                            uint8_t* d_weight, int32_t* d_rank, int32_t* d_order) {
  SM_CHECK_ARG(ctx && d_bgr && d_parent && d_weight && d_rank && H > 0 && W > 0 && cn >= 1 && cn <= 4);
  SM_CHECK_ARG((long long)H * W < (1ll << 30));
  const int N = H * W;
  SM_CUDA(cudaSetDevice(ctx->device));
  // 1. ctmf(r = 1) on the guidance (NL/qx_mst_kruskals_image.cpp:174)
  void* p_img;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_IMG0, (size_t)N * cn, &p_img));
  SM_TRY(sm_median_u8(ctx, d_bgr, (uint8_t*)p_img, H, W, 1, cn));
  nl_tree t;
  SM_TRY(nl_tree_scratch(ctx, N, t, d_order == nullptr));
  t.parent = d_parent; t.wpar = d_weight; t.rank = d_rank;
  if (d_order) t.order = d_order;
  SM_TRY(nl_build_tree(ctx, (const uint8_t*)p_img, H, W, cn, t));
  return SM_OK;
}

extern "C" int sm_tree_filter(sm_ctx* ctx, float* d_vol, double* d_work, int H, int W, int D, const int32_t* d_parent,
                              const uint8_t* d_weight, const int32_t* d_rank, const int32_t* d_order, double sigma) {
  SM_CHECK_ARG(ctx && d_vol && d_work && d_parent && d_weight && d_rank && d_order && H > 0 && W > 0 && D > 0);
  const int N = H * W;
  SM_CUDA(cudaSetDevice(ctx->device));
  nl_tree t;
  SM_TRY(nl_tree_scratch(ctx, N, t, false));
  t.parent = (int*)d_parent; t.wpar = (uint8_t*)d_weight; t.rank = (int*)d_rank; t.order = (int*)d_order;
  const int TB = 256;
  SM_CUDA(cudaMemsetAsync(t.sync, 0, sizeof(nl_sync), ctx->stream));
  SM_LAUNCH(ctx, k_children_from_parent, sm_div_up(N, TB), TB, 0, H, W, t.parent, t.wpar, t.child, t.nchild);
  SM_LAUNCH(ctx, k_level_bounds, min(sm_div_up(N, TB), ctx->num_sms * 8), TB, 0, N, t.rank, t.order, t.level_start,
            t.sync);
  return nl_filter(ctx, d_vol, d_work, H, W, D, false, t, sigma);
}

extern "C" int sm_tree_filter_f64(sm_ctx* ctx, double* d_cost, int H, int W, int D, const int32_t* d_parent,
                                  const uint8_t* d_weight, const int32_t* d_rank, const int32_t* d_order, double sigma) {
  SM_CHECK_ARG(ctx && d_cost && d_parent && d_weight && d_rank && d_order && H > 0 && W > 0 && D > 0);
  const int N = H * W;
  SM_CUDA(cudaSetDevice(ctx->device));
  nl_tree t;
  SM_TRY(nl_tree_scratch(ctx, N, t, false));
  t.parent = (int*)d_parent; t.wpar = (uint8_t*)d_weight; t.rank = (int*)d_rank; t.order = (int*)d_order;
  const int TB = 256;
  SM_CUDA(cudaMemsetAsync(t.sync, 0, sizeof(nl_sync), ctx->stream));
  SM_LAUNCH(ctx, k_children_from_parent, sm_div_up(N, TB), TB, 0, H, W, t.parent, t.wpar, t.child, t.nchild);
  SM_LAUNCH(ctx, k_level_bounds, min(sm_div_up(N, TB), ctx->num_sms * 8), TB, 0, N, t.rank, t.order, t.level_start,
            t.sync);
  return nl_filter(ctx, nullptr, d_cost, H, W, D, false, t, sigma);
}

// StereoMatching::NL in two halves, so that the pipeline can build the tree (which needs the image alone) on a side stream
// while the cost kernels run: the tree lives in the ctx scratch slots between the two calls.
int smi_nl_tree(sm_ctx* ctx, const uint8_t* d_bgrL, int H, int W) {
  const int N = H * W;
  void* p_img;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_IMG0, (size_t)N * 3, &p_img));
  SM_TRY(sm_median_u8(ctx, d_bgrL, (uint8_t*)p_img, H, W, 1, 3));
  nl_tree t;
  SM_TRY(nl_tree_scratch(ctx, N, t, true));
  SM_TRY(nl_build_tree(ctx, (const uint8_t*)p_img, H, W, 3, t));
  const int TB = 256;
  SM_LAUNCH(ctx, k_children_from_parent, sm_div_up(N, TB), TB, 0, H, W, t.parent, t.wpar, t.child, t.nchild);
  return SM_OK;
}
int smi_nl_filter(sm_ctx* ctx, float* d_vol, double* d_work, int H, int W, int D) {
  // d_work: H*W*(D+1) doubles
  nl_tree t;
  SM_TRY(nl_tree_scratch(ctx, H * W, t, true));   // (same sizes as in smi_nl_tree: the same buffers)
  return nl_filter(ctx, d_vol, d_work, H, W, D, true, t, 0.1);   // sigma: NL/NLCCA.cpp:33
}
int smi_nl(sm_ctx* ctx, const uint8_t* d_bgrL, float* d_vol, double* d_work, int H, int W, int D) {
  SM_TRY(smi_nl_tree(ctx, d_bgrL, H, W));
  return smi_nl_filter(ctx, d_vol, d_work, H, W, D);
}

extern "C" int sm_nl(sm_ctx* ctx, const uint8_t* d_bgrL, float* d_vol, int H, int W, int D) {
  SM_CHECK_ARG(ctx && d_bgrL && d_vol && H > 0 && W > 0 && D > 0);
  SM_CHECK_ARG((long long)H * W < (1ll << 30));
  SM_CUDA(cudaSetDevice(ctx->device));
  void* work = nullptr;   // grow-only scratch: a stream of frames allocates once
  SM_TRY(sm_scratch_get(ctx, SM_SCR_NLWORK, (size_t)H * W * (D + 1) * sizeof(double), &work));
  return smi_nl(ctx, d_bgrL, d_vol, (double*)work, H, W, D);
}
