// qx_tree_filter.h -- the reference's non-local tree filter class (NL/qx_tree_filter.h:13-39) over the sm_b200 C ABI.
// Same public methods and argument meaning: init / build_tree / filter / update_table / get_rank.  Host pointers in,
// host pointers out (the reference's convention: caller-owned raw arrays); the work runs on the GPU.
#pragma once
#ifndef QX_TREE_FILTER_H
#define QX_TREE_FILTER_H
#include <vector>

#include "../../../include/sm_b200.h"

#define QX_DEF_MST_KI_SIGMA_RANGE 0.1
#define QX_DEF_MST_KI_4NR_NEIGHBOR 4

class qx_tree_filter {
 public:
  qx_tree_filter();
  ~qx_tree_filter();
  void clean();
  int init(int h, int w, int nr_channel, double sigma_range = QX_DEF_MST_KI_SIGMA_RANGE,
           int nr_neighbor = QX_DEF_MST_KI_4NR_NEIGHBOR);          // NL/qx_tree_filter.cpp:14-20
  int build_tree(unsigned char* texture);                          // NL/qx_tree_filter.cpp:26-37 -> mst()
  int filter(double* cost, double* cost_backup, int nr_plane);     // NL/qx_tree_filter.cpp:61-117
  int* get_rank();                                                 // depth of every node (host copy)
  void update_table(double sigma_range);                           // NL/qx_tree_filter.cpp:21-25
  // additions: the rooted tree on the host
  const std::vector<int>& parent();
  const std::vector<unsigned char>& weight();

 private:
  void fetch();
  sm_ctx* ctx_ = nullptr;
  int m_h = 0, m_w = 0, m_nr_channel = 0, m_nr_pixel = 0;
  double m_sigma = QX_DEF_MST_KI_SIGMA_RANGE;
  int *d_parent_ = nullptr, *d_rank_ = nullptr, *d_order_ = nullptr;
  unsigned char *d_weight_ = nullptr, *d_img_ = nullptr;
  bool fetched_ = false;
  std::vector<int> h_parent_, h_rank_;
  std::vector<unsigned char> h_weight_;
};
#endif
