// Host wrappers of the NL/ public surface (ctmf, qx_tree_filter, NLCCA) over the sm_b200 C ABI.
#include <cstring>
#include <stdexcept>
#include <string>

#include "NLCCA.h"
#include "ctmf.h"
#include "qx_tree_filter.h"

namespace {
sm_ctx* shared_ctx() {   // ctmf() is a free C function: it runs on one process-wide context
  static sm_ctx* c = nullptr;
  if (!c && sm_ctx_create(&c, 0, nullptr) != SM_OK) throw std::runtime_error(std::string("sm_ctx_create: ") + sm_last_error());
  return c;
}
void ok(int rc, const char* what) {
  if (rc != SM_OK) throw std::runtime_error(std::string(what) + ": " + sm_last_error());
}
struct Dev {
  sm_ctx* c; void* p = nullptr;
  Dev(sm_ctx* ctx, size_t n) : c(ctx) { ok(sm_dev_alloc(c, &p, n), "sm_dev_alloc"); }
  ~Dev() { sm_dev_free(c, p); }
};
void h2d(sm_ctx* c, void* d, const void* h, size_t n) { ok(sm_memcpy_h2d(c, d, h, n), "h2d"); ok(sm_ctx_sync(c), "sync"); }
void d2h(sm_ctx* c, void* h, const void* d, size_t n) { ok(sm_memcpy_d2h(c, h, d, n), "d2h"); ok(sm_ctx_sync(c), "sync"); }
}  // namespace

extern "C" void ctmf(const unsigned char* src, unsigned char* dst, int width, int height, int src_step_row, int dst_step_row,
                     int r, int channels, unsigned long memsize) {
  (void)memsize;
  sm_ctx* c = shared_ctx();
  const size_t row = (size_t)width * channels, n = row * height;
  Dev a(c, n), b(c, n);
  if ((size_t)src_step_row == row) h2d(c, a.p, src, n);
  else for (int y = 0; y < height; y++) h2d(c, (unsigned char*)a.p + y * row, src + (size_t)y * src_step_row, row);
  ok(sm_median_u8(c, (const uint8_t*)a.p, (uint8_t*)b.p, height, width, r, channels), "sm_median_u8");
  if ((size_t)dst_step_row == row) d2h(c, dst, b.p, n);
  else for (int y = 0; y < height; y++) d2h(c, dst + (size_t)y * dst_step_row, (unsigned char*)b.p + y * row, row);
}

qx_tree_filter::qx_tree_filter() {}
qx_tree_filter::~qx_tree_filter() { clean(); }
void qx_tree_filter::clean() {
  if (!ctx_) return;
  sm_dev_free(ctx_, d_parent_); sm_dev_free(ctx_, d_rank_); sm_dev_free(ctx_, d_order_);
  sm_dev_free(ctx_, d_weight_); sm_dev_free(ctx_, d_img_);
  sm_ctx_destroy(ctx_);
  ctx_ = nullptr;
}
int qx_tree_filter::init(int h, int w, int nr_channel, double sigma_range, int nr_neighbor) {
  if (nr_neighbor != QX_DEF_MST_KI_4NR_NEIGHBOR) throw std::runtime_error("qx_tree_filter: only the 4-connected grid is on the path");
  clean();
  m_h = h; m_w = w; m_nr_channel = nr_channel; m_nr_pixel = h * w;
  ok(sm_ctx_create(&ctx_, 0, nullptr), "sm_ctx_create");
  ok(sm_dev_alloc(ctx_, (void**)&d_parent_, (size_t)m_nr_pixel * 4), "alloc");
  ok(sm_dev_alloc(ctx_, (void**)&d_rank_, (size_t)m_nr_pixel * 4), "alloc");
  ok(sm_dev_alloc(ctx_, (void**)&d_order_, (size_t)m_nr_pixel * 4), "alloc");
  ok(sm_dev_alloc(ctx_, (void**)&d_weight_, (size_t)m_nr_pixel), "alloc");
  ok(sm_dev_alloc(ctx_, (void**)&d_img_, (size_t)m_nr_pixel * nr_channel), "alloc");
  update_table(sigma_range);
  return 0;
}
void qx_tree_filter::update_table(double sigma_range) { m_sigma = sigma_range; }   // the table is built per filter() call
int qx_tree_filter::build_tree(unsigned char* texture) {
  h2d(ctx_, d_img_, texture, (size_t)m_nr_pixel * m_nr_channel);
  ok(sm_mst_build(ctx_, d_img_, m_h, m_w, m_nr_channel, d_parent_, d_weight_, d_rank_, d_order_), "sm_mst_build");
  fetched_ = false;
  return 0;
}
int qx_tree_filter::filter(double* cost, double* cost_backup, int nr_plane) {
  (void)cost_backup;   // the reference's scratch volume; the GPU path needs none on the host
  const size_t n = (size_t)m_nr_pixel * nr_plane * sizeof(double);
  Dev a(ctx_, n);
  h2d(ctx_, a.p, cost, n);
  ok(sm_tree_filter_f64(ctx_, (double*)a.p, m_h, m_w, nr_plane, d_parent_, d_weight_, d_rank_, d_order_, m_sigma),
     "sm_tree_filter_f64");
  d2h(ctx_, cost, a.p, n);
  return 0;
}
void qx_tree_filter::fetch() {
  if (fetched_) return;
  h_parent_.resize(m_nr_pixel); h_rank_.resize(m_nr_pixel); h_weight_.resize(m_nr_pixel);
  d2h(ctx_, h_parent_.data(), d_parent_, (size_t)m_nr_pixel * 4);
  d2h(ctx_, h_rank_.data(), d_rank_, (size_t)m_nr_pixel * 4);
  d2h(ctx_, h_weight_.data(), d_weight_, (size_t)m_nr_pixel);
  fetched_ = true;
}
int* qx_tree_filter::get_rank() { fetch(); return h_rank_.data(); }
const std::vector<int>& qx_tree_filter::parent() { fetch(); return h_parent_; }
const std::vector<unsigned char>& qx_tree_filter::weight() { fetch(); return h_weight_; }

void NLCCA::aggreCV(const Mat& lImg, const Mat& rImg, const int maxDis, Mat& costVol) {
  (void)rImg;
  CV_Assert(lImg.type() == CV_8UC3 && costVol.depth() == CV_32F);
  const int h = lImg.rows, w = lImg.cols;
  CV_Assert(costVol.total() * costVol.channels() == (size_t)h * w * maxDis);
  sm_ctx* c = shared_ctx();
  const size_t npix = (size_t)h * w, vb = npix * maxDis * 4;
  Dev img(c, npix * 3), vol(c, vb), work(c, npix * maxDis * 8), par(c, npix * 4), rk(c, npix * 4), ord(c, npix * 4), wt(c, npix);
  h2d(c, img.p, lImg.data, npix * 3);
  h2d(c, vol.p, costVol.data, vb);
  ok(sm_mst_build(c, (const uint8_t*)img.p, h, w, 3, (int32_t*)par.p, (uint8_t*)wt.p, (int32_t*)rk.p, (int32_t*)ord.p), "sm_mst_build");
  ok(sm_tree_filter(c, (float*)vol.p, (double*)work.p, h, w, maxDis, (const int32_t*)par.p, (const uint8_t*)wt.p,
                    (const int32_t*)rk.p, (const int32_t*)ord.p, 0.1), "sm_tree_filter");   // sigma: NL/NLCCA.cpp:33
  d2h(c, costVol.data, vol.p, vb);
}

// ------------------------------------------------------------------ qx_nonlocal_cost_aggregation
#include "qx_nonlocal_cost_aggregation.h"

qx_nonlocal_cost_aggregation::qx_nonlocal_cost_aggregation() {}
qx_nonlocal_cost_aggregation::~qx_nonlocal_cost_aggregation() { clean(); }
void qx_nonlocal_cost_aggregation::clean() {
  if (!ctx_) return;
  void* ptrs[] = {d_left_, d_right_, d_disp_, d_disp2_, d_dispR_, d_mask_, d_wt_[0], d_wt_[1], d_cost_, d_costR_, d_vol_,
                  d_parent_[0], d_parent_[1], d_rank_[0], d_rank_[1], d_order_[0], d_order_[1]};
  for (void* p : ptrs) sm_dev_free(ctx_, p);
  sm_ctx_destroy(ctx_);
  ctx_ = nullptr;
}
int qx_nonlocal_cost_aggregation::init(int h, int w, int nr_plane, double sigma_range, double max_color_difference,
                                       double max_gradient_difference, double weight_on_color) {
  clean();
  if (nr_plane > 256) throw std::runtime_error("qx_nonlocal_cost_aggregation: disparities are unsigned char (nr_plane <= 256)");
  m_h = h; m_w = w; m_nr_plane = nr_plane; m_sigma_range = sigma_range;
  m_max_color_difference = max_color_difference; m_max_gradient_difference = max_gradient_difference;
  m_weight_on_color = weight_on_color;
  ok(sm_ctx_create(&ctx_, 0, nullptr), "sm_ctx_create");
  const size_t np = (size_t)h * w, nv = np * nr_plane * sizeof(double);
  auto A = [&](void** p, size_t n) { ok(sm_dev_alloc(ctx_, p, n), "sm_dev_alloc"); };
  A((void**)&d_left_, np * 3); A((void**)&d_right_, np * 3);
  A((void**)&d_disp_, np); A((void**)&d_disp2_, np); A((void**)&d_dispR_, np); A((void**)&d_mask_, np);
  A((void**)&d_cost_, nv); A((void**)&d_costR_, nv); A((void**)&d_vol_, nv);
  for (int i = 0; i < 2; i++) {
    A((void**)&d_wt_[i], np); A((void**)&d_parent_[i], np * 4); A((void**)&d_rank_[i], np * 4); A((void**)&d_order_[i], np * 4);
  }
  return 0;
}
int qx_nonlocal_cost_aggregation::matching_cost(unsigned char*** left, unsigned char*** right) {
  const size_t np = (size_t)m_h * m_w;
  h2d(ctx_, d_left_, left[0][0], np * 3);
  h2d(ctx_, d_right_, right[0][0], np * 3);
  ok(sm_nlca_cost(ctx_, d_left_, d_right_, m_h, m_w, m_nr_plane, m_max_color_difference, m_max_gradient_difference,
                  m_weight_on_color, d_cost_), "sm_nlca_cost");
  ok(sm_nlca_flip(ctx_, d_cost_, m_h, m_w, m_nr_plane, d_costR_), "sm_nlca_flip");
  ok(sm_mst_build(ctx_, d_left_, m_h, m_w, 3, d_parent_[0], d_wt_[0], d_rank_[0], d_order_[0]), "sm_mst_build");
  ok(sm_mst_build(ctx_, d_right_, m_h, m_w, 3, d_parent_[1], d_wt_[1], d_rank_[1], d_order_[1]), "sm_mst_build");
  return 0;
}
void qx_nonlocal_cost_aggregation::filter(double* d_vol, bool right_tree, double sigma) {
  const int t = right_tree ? 1 : 0;
  ok(sm_tree_filter_f64(ctx_, d_vol, m_h, m_w, m_nr_plane, d_parent_[t], d_wt_[t], d_rank_[t], d_order_[t], sigma),
     "sm_tree_filter_f64");
}
int qx_nonlocal_cost_aggregation::disparity(unsigned char** disparity, bool use_nonlocal_post_processing) {
  const int radius = 2;
  const size_t np = (size_t)m_h * m_w, nv = np * m_nr_plane * sizeof(double);
  auto copy = [&](double* dst, const double* src) {   // image_copy(m_cost_vol, ...): device to device
    ok(sm_memcpy_d2d(ctx_, dst, src, nv), "sm_memcpy_d2d");
  };
  copy(d_vol_, d_cost_);
  filter(d_vol_, false, m_sigma_range);
  ok(sm_depth_best_cost(ctx_, d_vol_, m_h, m_w, m_nr_plane, d_disp2_), "sm_depth_best_cost");
  ok(sm_median_u8(ctx_, d_disp2_, d_disp_, m_h, m_w, radius, 1), "sm_median_u8");
  if (use_nonlocal_post_processing) {
    copy(d_vol_, d_costR_);
    filter(d_vol_, true, m_sigma_range);
    ok(sm_depth_best_cost(ctx_, d_vol_, m_h, m_w, m_nr_plane, d_disp2_), "sm_depth_best_cost");
    ok(sm_median_u8(ctx_, d_disp2_, d_dispR_, m_h, m_w, radius, 1), "sm_median_u8");
    ok(sm_nlca_occlusion(ctx_, d_disp_, d_dispR_, m_h, m_w, d_mask_), "sm_nlca_occlusion");
    ok(sm_nlca_refine_cost(ctx_, d_disp_, d_mask_, m_h, m_w, m_nr_plane, d_vol_), "sm_nlca_refine_cost");
    filter(d_vol_, false, m_sigma_range / 2);   // m_tf.update_table(m_sigma_range / 2)
    ok(sm_depth_best_cost(ctx_, d_vol_, m_h, m_w, m_nr_plane, d_disp2_), "sm_depth_best_cost");
    ok(sm_median_u8(ctx_, d_disp2_, d_disp_, m_h, m_w, radius, 1), "sm_median_u8");
  }
  d2h(ctx_, disparity[0], d_disp_, np);
  return 0;
}
void qx_nonlocal_cost_aggregation::get_cost_volume(double* out) { d2h(ctx_, out, d_cost_, (size_t)m_h * m_w * m_nr_plane * 8); }
void qx_nonlocal_cost_aggregation::get_cost_volume_right(double* out) { d2h(ctx_, out, d_costR_, (size_t)m_h * m_w * m_nr_plane * 8); }
