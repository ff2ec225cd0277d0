// TEST INFRASTRUCTURE.  extern "C" entry points over the reference's own NL/
// classes, compiled TOGETHER WITH the reference sources by oracle/build_ref.sh
// into oracle/_ref/libqxref.so (never committed).  Nothing here restates the
// algorithm: it only calls qx_mst_kruskals_image / qx_tree_filter / ctmf and
// copies their outputs out.
#include "qx_basic.h"
#include "qx_mst_kruskals_image.h"
#include "qx_tree_filter.h"
#include "ctmf.h"

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
}
