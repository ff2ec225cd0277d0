// qx_nonlocal_cost_aggregation.h -- Yang's driver class (reference NL/qx_nonlocal_cost_aggregation.h:18-60) over the
// sm_b200 C ABI: same public methods, defaults and argument conventions (qx-style pointer-table images whose
// [0][0] / [0] is the flat buffer).  Volumes live on the GPU as float64, like the reference's double arrays.
#pragma once
#ifndef QX_NONLOCAL_COST_AGGREGATION_H
#define QX_NONLOCAL_COST_AGGREGATION_H
#include "../../../include/sm_b200.h"
#include "qx_basic.h"

class qx_nonlocal_cost_aggregation {
 public:
  qx_nonlocal_cost_aggregation();
  ~qx_nonlocal_cost_aggregation();
  void clean();
  int init(int h, int w, int nr_plane, double sigma_range = 0.1, double max_color_difference = 7,
           double max_gradient_difference = 2, double weight_on_color = 0.11);          // :22-56
  int matching_cost(unsigned char*** left, unsigned char*** right);                     // :57-71
  int disparity(unsigned char** disparity, bool use_nonlocal_post_processing = false);  // :72-109
  // additions: host copies of the device volumes (h*w*nr_plane doubles each)
  void get_cost_volume(double* out);
  void get_cost_volume_right(double* out);

 private:
  void filter(double* d_vol, bool right_tree, double sigma);
  sm_ctx* ctx_ = nullptr;
  int m_h = 0, m_w = 0, m_nr_plane = 0;
  double m_sigma_range = 0.1, m_max_color_difference = 7, m_max_gradient_difference = 2, m_weight_on_color = 0.11;
  unsigned char *d_left_ = nullptr, *d_right_ = nullptr, *d_disp_ = nullptr, *d_disp2_ = nullptr, *d_dispR_ = nullptr,
                *d_mask_ = nullptr, *d_wt_[2] = {nullptr, nullptr};
  double *d_cost_ = nullptr, *d_costR_ = nullptr, *d_vol_ = nullptr;
  int *d_parent_[2] = {nullptr, nullptr}, *d_rank_[2] = {nullptr, nullptr}, *d_order_[2] = {nullptr, nullptr};
};
#endif
