// TEST INFRASTRUCTURE.  extern "C" entry points over the reference's own NL/
// classes, compiled TOGETHER WITH the reference sources by oracle/build_ref.sh
// into oracle/_ref/libqxref.so (never committed).  Nothing here restates the
// algorithm: it only calls qx_mst_kruskals_image / qx_tree_filter / ctmf and
// copies their outputs out.
#include "qx_basic.h"
#include "qx_mst_kruskals_image.h"
#include "qx_tree_filter.h"
#include "ctmf.h"
#include "qx_nonlocal_cost_aggregation.h"

// qx_timer lives in NL/qx_basic.cpp, which needs <windows.h>; it is only used
// for (disabled) prints, so a do-nothing definition stands in for that file.
void qx_timer::start() {}
double qx_timer::stop() { return 0; }
void qx_timer::time_display(char*, int) {}
void qx_timer::fps_display(char*, int) {}

extern "C" {

void qxref_ctmf(const unsigned char* src, unsigned char* dst, int w, int h, int sstep, int dstep,
                int r, int cn, unsigned long memsize) {
  ctmf(src, dst, w, h, sstep, dstep, r, cn, memsize);
}

// Runs qx_mst_kruskals_image::mst exactly as qx_tree_filter::build_tree does
// (NL/qx_tree_filter.cpp:26-37) and exports the tree arrays.
void qxref_mst(unsigned char* image, int h, int w, int cn, int* parent, unsigned char* weight,
               int* rank, int* nr_child, int* children /*3 per node, -1 padded*/, int* order) {
  qx_mst_kruskals_image mst;
  mst.init(h, w, cn);
  mst.mst(image);
  int n = h * w;
  int** ch = mst.get_children();
  for (int i = 0; i < n; i++) {
    parent[i] = mst.get_parent()[i];
    weight[i] = mst.get_weight()[i];
    rank[i] = mst.get_rank()[i];
    nr_child[i] = mst.get_nr_child()[i];
    order[i] = mst.get_node_id()[i];
    for (int j = 0; j < 3; j++) children[3 * i + j] = j < nr_child[i] ? ch[i][j] : -1;
  }
}

// qx_tree_filter::init + build_tree + filter on a caller-supplied f64 volume
// (what NLCCA::aggreCV does between its two conversion loops, NL/NLCCA.cpp:70-78).
void qxref_tree_filter(unsigned char* image, int h, int w, int nr_plane, double sigma, double* cost,
                       double* tmp) {
  qx_tree_filter tf;
  tf.init(h, w, 3, sigma, 4);
  tf.build_tree(image);
  tf.filter(cost, tmp, nr_plane);
}

// Yang's own driver class, qx_nonlocal_cost_aggregation (NL/qx_nonlocal_cost_aggregation.cpp:22-109, 190-236):
// init + matching_cost [+ disparity].  left/right: h*w*3 bytes.  Outputs (nullable): the raw left cost volume and
// its right-view flip (double [h][w][nr_plane]), the x-gradients (float [h][w]) and the final disparity map.
void qxref_nlca(unsigned char* left, unsigned char* right, int h, int w, int nr_plane, double sigma, int post,
                double* cost_out, double* cost_right_out, float* grad_left_out, unsigned char* disp_out) {
  unsigned char*** l = qx_allocu_3(h, w, 3);
  unsigned char*** r = qx_allocu_3(h, w, 3);
  memcpy(l[0][0], left, (size_t)h * w * 3);
  memcpy(r[0][0], right, (size_t)h * w * 3);
  qx_nonlocal_cost_aggregation nl;
  nl.init(h, w, nr_plane, sigma);
  nl.matching_cost(l, r);
  if (cost_out) memcpy(cost_out, nl.m_cost_vol_backup[0][0], sizeof(double) * h * w * nr_plane);
  if (cost_right_out) memcpy(cost_right_out, nl.m_cost_vol_right[0][0], sizeof(double) * h * w * nr_plane);
  if (grad_left_out) memcpy(grad_left_out, nl.m_gradient_left[0], sizeof(float) * h * w);
  if (disp_out) {
    unsigned char** d = qx_allocu(h, w);
    nl.disparity(d, post != 0);
    memcpy(disp_out, d[0], (size_t)h * w);
    qx_freeu(d);
  }
  qx_freeu_3(l);
  qx_freeu_3(r);
}
}
