#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Compiles the reference's own NL/ sources (ctmf.c,
# qx_mst_kruskals_image.cpp, qx_tree_filter.cpp) into oracle/_ref/libqxref.so.
#
# The sources are read from $REF (default /root/reference).  They are MSVC
# code, so four mechanical fixes are applied to TRANSIENT copies in a mktemp
# directory that is deleted afterwards (nothing of the reference is written
# into the repo; only the .so lands in oracle/_ref/, which is git-ignored):
#   * empty process.h / direct.h / io.h stand-ins (NL/qx_basic.h:13-15)
#   * -D__int64="long long"                          (NL/qx_basic.h:40)
#   * a prelude with <algorithm> <cstring> <cmath> <iostream> + using namespace
#     std (qx_basic.h only says so under _MSC_VER, :22-24)
#   * `unsigned char(x)` -> `(unsigned char)(x)`     (NL/qx_basic.h:72)
#   * the two dead member templates that name non-existent members are dropped
#     (NL/qx_tree_filter.cpp:38-60)
#   * NL/qx_basic.cpp (needs <windows.h>, fopen_s, ...) is not compiled as a whole;
#     qx_timer is a no-op in qxref_shim.cpp, and the four plain functions Yang's
#     driver needs (qx_stereo_flip_corr_vol, depth_best_cost, vec_min_pos,
#     qx_detect_occlusion_left_right, NL/qx_basic.cpp:577-624) are taken by line range
#   * NL/qx_nonlocal_cost_aggregation.cpp includes <opencv2/opencv.hpp> for ONE
#     function that is behind USE_CENCUS and never called; a compile-only stand-in
#     header (cv::Mat / cvtColor that abort()) lets the file build unchanged
# Sources are GBK/CRLF; iconv is not needed (comments only).
set -euo pipefail
REF="${REF:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/NL" ]; then
  echo "build_ref: $REF/NL not present; keeping any prebuilt $OUT/libqxref.so" >&2
  exit 0
fi
mkdir -p "$OUT"
T="$(mktemp -d)"
trap 'rm -rf "$T"' EXIT
mkdir -p "$T/shim"
: > "$T/shim/process.h"; : > "$T/shim/direct.h"; : > "$T/shim/io.h"
cat > "$T/prelude.h" <<'EOF'
#ifdef __cplusplus
#include <algorithm>
#include <cstring>
#include <cmath>
#include <cstdlib>
#include <iostream>
using namespace std;
#endif
EOF
mkdir -p "$T/shim/opencv2"
cat > "$T/shim/opencv2/opencv.hpp" <<'EOF'
#pragma once
#include <cstdlib>
typedef unsigned char uchar;
#define CV_8UC3 16
#define CV_RGB2GRAY 7
namespace cv {
struct Mat { int rows, cols; Mat() : rows(0), cols(0) {} static Mat zeros(int, int, int) { abort(); return Mat(); }
  template <typename T> T* ptr(int) { abort(); return 0; } };
inline void cvtColor(const Mat&, Mat&, int) { abort(); }
}
EOF
for f in qx_basic.h qx_mst_kruskals_image.h qx_mst_kruskals_image.cpp qx_tree_filter.h ctmf.h ctmf.c qx_nonlocal_cost_aggregation.h qx_nonlocal_cost_aggregation.cpp; do
  cp "$REF/NL/$f" "$T/$f"
done
sed -i 's/return(unsigned char(\(.*\)));}/return((unsigned char)(\1));}/' "$T/qx_basic.h"
sed '38,60d' "$REF/NL/qx_tree_filter.cpp" > "$T/qx_tree_filter.cpp"
{ echo '#include "qx_basic.h"'; sed -n '577,624p' "$REF/NL/qx_basic.cpp"; } > "$T/qx_basic_subset.cpp"
cp "$HERE/qxref_shim.cpp" "$T/qxref_shim.cpp"
CXXFLAGS="-O2 -fPIC -w -fpermissive -ffp-contract=off -D__int64=long\ long -include $T/prelude.h -I$T/shim -I$T"
${CC:-gcc} -O2 -fPIC -w -c "$T/ctmf.c" -o "$T/ctmf.o"
eval ${CXX:-g++} $CXXFLAGS -c "$T/qx_mst_kruskals_image.cpp" -o "$T/mst.o"
eval ${CXX:-g++} $CXXFLAGS -c "$T/qx_tree_filter.cpp" -o "$T/tf.o"
eval ${CXX:-g++} $CXXFLAGS -c "$T/qx_nonlocal_cost_aggregation.cpp" -o "$T/nlca.o"
eval ${CXX:-g++} $CXXFLAGS -c "$T/qx_basic_subset.cpp" -o "$T/basic.o"
eval ${CXX:-g++} $CXXFLAGS -c "$T/qxref_shim.cpp" -o "$T/shim.o"
${CXX:-g++} -shared -o "$OUT/libqxref.so" "$T/ctmf.o" "$T/mst.o" "$T/tf.o" "$T/nlca.o" "$T/basic.o" "$T/shim.o"
echo "build_ref: wrote $OUT/libqxref.so"
