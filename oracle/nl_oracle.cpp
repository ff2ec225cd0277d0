// ============================================================================
// nl_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's NL/ non-local aggregation path:
// constant-time median (ctmf) semantics, the Kruskal MST over the 4-connected
// image grid, BFS rooting, and the two-pass tree filter, plus the glue in
// NLCCA::aggreCV and StereoMatching::NL.  Every function cites the reference
// file:line it follows.
//
// Pinning status: PINNED.  oracle/build_ref.sh compiles the reference's own
// NL/ctmf.c, NL/qx_mst_kruskals_image.cpp and NL/qx_tree_filter.cpp (where
// they lie under /root/reference) into oracle/_ref/libqxref.so;
// tests/golden/make_nl_golden.py ran that library on seeded inputs and the
// outputs are committed as tests/golden/nl_*.npz; tests/test_oracle_nl.py
// checks this restatement against them (and against libqxref.so directly when
// it is present).
// ============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

typedef unsigned char u8;
typedef short i16;

extern "C" {

// ---------------------------------------------------------------------------
// ctmf(src,dst,w,h,sstep,dstep,r,cn,memsize) (NL/ctmf.c:378-433, helper
// :193-339): (2r+1)^2 median of an 8-bit image with cn interleaved channels.
// Border handling is EDGE REPLICATION (first-row initialisation adds the top
// row r+1 times, :230; column indices are clamped with MIN/MAX, :244-252,
// :285, :310-311) -- the doc comment's "zero padded" is wrong.  The median is
// the ((2r+1)^2/2)-th order statistic counted from zero (:301 "t = 2*r*r+2*r",
// first bin whose running sum exceeds t).
// ---------------------------------------------------------------------------
void orc_ctmf(const u8* src, u8* dst, int W, int H, int src_step, int dst_step, int r, int cn) {
  const int t = 2 * r * r + 2 * r;
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++)
      for (int c = 0; c < cn; c++) {
        int hist[256] = {0};
        for (int dy = -r; dy <= r; dy++) {
          int yy = std::min(std::max(y + dy, 0), H - 1);
          for (int dx = -r; dx <= r; dx++) {
            int xx = std::min(std::max(x + dx, 0), W - 1);
            hist[src[(long)yy * src_step + xx * cn + c]]++;
          }
        }
        int sum = 0, k = 0;
        for (; k < 256; k++) { sum += hist[k]; if (sum > t) break; }
        dst[(long)y * dst_step + x * cn + c] = (u8)k;
      }
}

// ---------------------------------------------------------------------------
// MST of the 4-connected grid: qx_mst_kruskals_image::mst
// (NL/qx_mst_kruskals_image.cpp:167-186).
//  1. guidance <- ctmf(r=1) per channel (:174)
//  2. edges: all horizontal (y,x)-(y,x+1) row-major, then all vertical
//     (y,x)-(y+1,x) COLUMN-major (:46-69); weight = max_c |delta| (:22-25)
//  3. stable counting sort by weight (NL/qx_basic.h:76-99)
//  4. Kruskal with union-find (:188-228); each accepted edge is appended to the
//     adjacency list of both endpoints
//  5. BFS from vertex 0 over the adjacency lists (:229-277) -> parent, weight
//     (edge to parent), rank (depth), children (<=3, in adjacency order),
//     BFS order (node_id_from_parent_to_child).
// Outputs (all length H*W): parent (root: itself), weight (root: 0), rank,
// nr_child, children[3*i+j] (-1 padded), order.
// ---------------------------------------------------------------------------
static int uf_find(std::vector<int>& p, int x) {
  int r = x;
  while (p[r] != r) r = p[r];
  while (p[x] != r) { int nx = p[x]; p[x] = r; x = nx; }  // full path compression, iterative
  return r;
}

void orc_mst(const u8* image, int H, int W, int cn, int* parent, u8* weight, int* rank,
             int* nr_child, int* children, int* order, u8* filtered_out /*nullable*/) {
  const int N = H * W;
  std::vector<u8> img((size_t)N * cn);
  orc_ctmf(image, img.data(), W, H, W * cn, W * cn, 1, cn);
  if (filtered_out) std::memcpy(filtered_out, img.data(), img.size());
  const int E = (H - 1) * W + (W - 1) * H;
  std::vector<int> ea(E), eb(E);
  std::vector<u8> dist(E);
  int ne = 0;
  auto add = [&](int y0, int x0, int yt, int xt) {
    int a = y0 * W + x0, b = yt * W + xt;
    ea[ne] = a; eb[ne] = b;
    int m = 0;
    for (int c = 0; c < cn; c++) m = std::max(m, std::abs((int)img[(size_t)b * cn + c] - (int)img[(size_t)a * cn + c]));
    dist[ne] = (u8)m;
    ne++;
  };
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W - 1; x++) add(y, x, y, x + 1);
  for (int x = 0; x < W; x++)
    for (int y = 0; y < H - 1; y++) add(y, x, y + 1, x);
  // stable counting sort
  std::vector<int> id(E);
  {
    int start[257] = {0};
    for (int i = 0; i < E; i++) start[dist[i] + 1]++;
    for (int k = 0; k < 256; k++) start[k + 1] += start[k];
    for (int i = 0; i < E; i++) id[start[dist[i]]++] = i;
  }
  std::vector<int> uf(N);
  for (int i = 0; i < N; i++) uf[i] = i;
  std::vector<int> adj((size_t)N * 4, -1);
  std::vector<u8> adjw((size_t)N * 4, 0);
  std::vector<int> nadj(N, 0);
  for (int j = 0; j < E; j++) {
    int e = id[j], a = ea[e], b = eb[e];
    int pa = uf_find(uf, a), pb = uf_find(uf, b);
    if (pa != pb) {
      adj[(size_t)a * 4 + nadj[a]] = b; adjw[(size_t)a * 4 + nadj[a]] = dist[e]; nadj[a]++;
      adj[(size_t)b * 4 + nadj[b]] = a; adjw[(size_t)b * 4 + nadj[b]] = dist[e]; nadj[b]++;
      uf[pa] = pb;
    }
  }
  for (int i = 0; i < N; i++) { parent[i] = -1; nr_child[i] = 0; rank[i] = 0; }
  for (int i = 0; i < 3 * N; i++) children[i] = -1;
  parent[0] = 0; weight[0] = 0; order[0] = 0;
  int len = 1;
  for (int head = 0; head < len; head++) {
    int p = order[head];
    for (int i = 0; i < nadj[p]; i++) {
      int c = adj[(size_t)p * 4 + i];
      if (parent[c] == -1) {
        parent[c] = p;
        rank[c] = rank[p] + 1;
        weight[c] = adjw[(size_t)p * 4 + i];
        children[3 * p + nr_child[p]++] = c;
        order[len++] = c;
      }
    }
  }
}

// qx_tree_filter::update_table (NL/qx_tree_filter.cpp:21-25).
void orc_tree_table(double sigma_range, double* table /*256*/) {
  sigma_range = std::max(0.01, sigma_range);
  for (int i = 0; i <= 255; i++) table[i] = std::exp(-double(i) / (255 * sigma_range));
}

// qx_tree_filter::filter (NL/qx_tree_filter.cpp:61-117): backup <- cost; leaf-
// to-root in reverse BFS order backup[p] += sum_children w(c)*backup[c] (children
// in adjacency order); root copies; root-to-leaf in BFS order
// cost[i] = w*(cost[parent] - w*backup[i]) + backup[i].
void orc_tree_filter(double* cost, double* backup, int N, int D, const int* parent,
                     const u8* weight, const int* nr_child, const int* children, const int* order,
                     const double* table) {
  std::memcpy(backup, cost, sizeof(double) * (size_t)N * D);
  for (int i = N - 1; i >= 0; i--) {
    int id = order[i];
    double* vs = backup + (size_t)id * D;
    for (int j = 0; j < nr_child[id]; j++) {
      int c = children[3 * id + j];
      double w = table[weight[c]];
      const double* vc = backup + (size_t)c * D;
      for (int k = 0; k < D; k++) vs[k] += vc[k] * w;
    }
  }
  int root = order[0];
  std::memcpy(cost + (size_t)root * D, backup + (size_t)root * D, sizeof(double) * D);
  for (int i = 1; i < N; i++) {
    int id = order[i];
    const double* vp = cost + (size_t)parent[id] * D;
    const double* vc = backup + (size_t)id * D;
    double* vf = cost + (size_t)id * D;
    double w = table[weight[id]];
    for (int k = 0; k < D; k++) vf[k] = w * (vp[k] - w * vc[k]) + vc[k];
  }
}

// NLCCA::aggreCV (NL/NLCCA.cpp:27-95): sigma 0.1, f32 -> f64, build tree on the
// LEFT image bytes as given, filter, f64 -> f32.
void orc_nl_aggre(const u8* bgrL, int H, int W, int D, float* vol) {
  const int N = H * W;
  std::vector<int> parent(N), rank(N), nch(N), ch(3 * (size_t)N), order(N);
  std::vector<u8> wt(N);
  orc_mst(bgrL, H, W, 3, parent.data(), wt.data(), rank.data(), nch.data(), ch.data(), order.data(),
          nullptr);
  double table[256];
  orc_tree_table(0.1, table);
  std::vector<double> cost((size_t)N * D), tmp((size_t)N * D);
  for (size_t i = 0; i < cost.size(); i++) cost[i] = (double)vol[i];
  orc_tree_filter(cost.data(), tmp.data(), N, D, parent.data(), wt.data(), nch.data(), ch.data(),
                  order.data(), table);
  for (size_t i = 0; i < cost.size(); i++) vol[i] = (float)cost[i];
}

// StereoMatching::NL (stereoMatching.cpp:4892-4917): aggregate vm[0], aggregate
// an all-ones volume, divide (float / float), WTA (gen_dispFromVm, see
// orc_wta in stereo_oracle.cpp -- repeated here as the first-minimum scan).
void orc_nl(const u8* bgrL, int H, int W, int D, float* vol, i16* disp) {
  const size_t n = (size_t)H * W * D;
  orc_nl_aggre(bgrL, H, W, D, vol);
  std::vector<float> ones(n, 1.0f);
  orc_nl_aggre(bgrL, H, W, D, ones.data());
  for (size_t i = 0; i < n; i++) vol[i] /= ones[i];
  if (disp)
    for (long i = 0; i < (long)H * W; i++) {
      float m = 3.402823466e+38f;
      int best = -1;
      for (int d = 0; d < D; d++)
        if (m > vol[i * D + d]) { m = vol[i * D + d]; best = d; }
      disp[i] = (i16)best;
    }
}


// ---------------------------------------------------------------------------
// Yang's own driver, qx_nonlocal_cost_aggregation (API surface named by the north star; pinned against the
// reference's compiled class through oracle/_ref: qxref_nlca).
// ---------------------------------------------------------------------------
// rgb_2_gray (NL/qx_basic.h:72): (unsigned char)(0.299*in[0] + 0.587*in[1] + 0.114*in[2] + 0.5), double arithmetic.
static inline u8 qx_gray(const u8* in) { return (u8)(0.299 * in[0] + 0.587 * in[1] + 0.114 * in[2] + 0.5); }

// compute_gradient (NL/qx_nonlocal_cost_aggregation.cpp:219-236): central x-difference of the gray image + 127.5,
// one-sided (un-halved) at both borders.
void orc_nlca_gradient(const u8* img, int H, int W, float* grad) {
  for (int y = 0; y < H; y++) {
    const u8* row = img + (size_t)y * W * 3;
    float* g = grad + (size_t)y * W;
    float gray, gray_minus, gray_plus;
    gray_minus = qx_gray(row);
    gray = gray_plus = qx_gray(row + 3);
    g[0] = gray_plus - gray_minus + 127.5;
    for (int x = 1; x < W - 1; x++) {
      gray_plus = qx_gray(row + 3 * (x + 1));
      g[x] = 0.5 * (gray_plus - gray_minus) + 127.5;
      gray_minus = gray;
      gray = gray_plus;
    }
    g[W - 1] = gray_plus - gray_minus + 127.5;
  }
}

// matching_cost_from_color_and_gradient (NL/qx_nonlocal_cost_aggregation.cpp:190-218): for plane i the right image
// and gradient are shifted by i pixels (columns < i replicate column 0); cost = w*min(mean|dRGB|, maxc) +
// (1-w)*min(|dgrad|, maxg), double.
void orc_nlca_cost(const u8* left, const u8* right, int H, int W, int D, double maxc, double maxg, double wc,
                   double* vol) {
  std::vector<float> gl((size_t)H * W), gr((size_t)H * W);
  orc_nlca_gradient(left, H, W, gl.data());
  orc_nlca_gradient(right, H, W, gr.data());
  const double wci = 1 - wc;
  for (int i = 0; i < D; i++)
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++) {
        const int xs = x >= i ? x - i : 0;
        const u8* l = left + ((size_t)y * W + x) * 3;
        const u8* r = right + ((size_t)y * W + xs) * 3;
        double cost = 0;
        for (int c = 0; c < 3; c++) cost += std::abs((int)l[c] - (int)r[c]);
        cost = std::min(cost / 3, maxc);
        double cg = std::min((double)std::abs(gl[(size_t)y * W + x] - gr[(size_t)y * W + xs]), maxg);
        vol[((size_t)y * W + x) * D + i] = wc * cost + wci * cg;
      }
}

// qx_stereo_flip_corr_vol (NL/qx_basic.cpp:577-588): right-view volume from the left one along the diagonal; where
// x+d leaves the image the previous plane's value is repeated.
void orc_flip_vol(const double* vol, int H, int W, int D, double* volR) {
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++)
      for (int d = 0; d < D; d++) {
        size_t o = ((size_t)y * W + x) * D + d;
        if (x + d < W) volR[o] = vol[((size_t)y * W + x + d) * D + d];
        else volR[o] = volR[o - 1];
      }
}

// depth_best_cost / vec_min_pos (NL/qx_basic.cpp:589-602): first minimum, as unsigned char.
void orc_depth_best_cost(const double* vol, int H, int W, int D, u8* depth) {
  for (size_t p = 0; p < (size_t)H * W; p++) {
    const double* in = vol + p * D;
    double mv = in[0];
    int mp = 0;
    for (int i = 1; i < D; i++)
      if (in[i] < mv) { mv = in[i]; mp = i; }
    depth[p] = (u8)mp;
  }
}

// qx_detect_occlusion_left_right (NL/qx_basic.cpp:603-624): 255 where the left disparity is 0, maps outside the
// image, or differs from the right map at the matched column.
void orc_detect_occlusion(const u8* dl, const u8* dr, int H, int W, u8* mask) {
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      int d = dl[(size_t)y * W + x], xr = x - d;
      u8 m = 0;
      if (xr >= 0) { if (d == 0 || std::abs(d - (int)dr[(size_t)y * W + xr]) >= 1) m = 255; }
      else m = 255;
      mask[(size_t)y * W + x] = m;
    }
}

static void nlca_filter(double* vol, int N, int D, const u8* img, int H, int W, double sigma) {
  std::vector<int> parent(N), rank(N), nch(N), ch(3 * (size_t)N), order(N);
  std::vector<u8> wt(N);
  orc_mst(img, H, W, 3, parent.data(), wt.data(), rank.data(), nch.data(), ch.data(), order.data(), nullptr);
  double table[256];
  orc_tree_table(sigma, table);
  std::vector<double> tmp((size_t)N * D);
  orc_tree_filter(vol, tmp.data(), N, D, parent.data(), wt.data(), nch.data(), ch.data(), order.data(), table);
}

// init + matching_cost + disparity (NL/qx_nonlocal_cost_aggregation.cpp:22-109) with the class defaults
// (max colour difference 7, max gradient difference 2, weight on colour 0.11).  post = the optional non-local
// refinement: right disparity, occlusion mask, |d - disp| volume on the stable pixels, sigma/2, filter, WTA, median.
void orc_nlca_disparity(const u8* left, const u8* right, int H, int W, int D, double sigma, int post, u8* disp) {
  const int N = H * W;
  std::vector<double> cost((size_t)N * D), vol((size_t)N * D), costR((size_t)N * D);
  orc_nlca_cost(left, right, H, W, D, 7, 2, 0.11, cost.data());
  orc_flip_vol(cost.data(), H, W, D, costR.data());
  std::vector<u8> d0(N), dr(N), mask(N);
  vol = cost;
  nlca_filter(vol.data(), N, D, left, H, W, sigma);
  orc_depth_best_cost(vol.data(), H, W, D, d0.data());
  orc_ctmf(d0.data(), disp, W, H, W, W, 2, 1);
  if (!post) return;
  vol = costR;
  nlca_filter(vol.data(), N, D, right, H, W, sigma);
  orc_depth_best_cost(vol.data(), H, W, D, d0.data());
  orc_ctmf(d0.data(), dr.data(), W, H, W, W, 2, 1);
  orc_detect_occlusion(disp, dr.data(), H, W, mask.data());
  std::fill(vol.begin(), vol.end(), 0.0);
  for (int p = 0; p < N; p++)
    if (!mask[p])
      for (int d = 0; d < D; d++) vol[(size_t)p * D + d] = std::abs((int)disp[p] - d);
  nlca_filter(vol.data(), N, D, left, H, W, sigma / 2);
  orc_depth_best_cost(vol.data(), H, W, D, d0.data());
  orc_ctmf(d0.data(), disp, W, H, W, W, 2, 1);
}

}  // extern "C"
