#!/usr/bin/env python3
"""TEST INFRASTRUCTURE.  Compiles the reference's OWN function bodies for the dense-stereo hot path
(stereoMatching.h / stereoMatching.cpp) into oracle/_ref/libsmref.so, so that the CPU restatement in
stereo_oracle.cpp and the CUDA path can be checked against what the reference's code computes, not against a
reading of it.

stereoMatching.{h,cpp} cannot be built as files: they need OpenCV C++ (core/imgproc/highgui + opencv_contrib
ximgproc, stereoMatching.h:15-18), a `util.h` that is not in the repository (stereoMatching.h:2) and MSVC-only
constructs (stereoMatching.h:1256, stereoMatching.cpp:1896).  The functions on the hot path, however, use only a
thin slice of cv::Mat (ptr<T>, create, clone, copyTo, depth/type/channels, size[], `= scalar`) plus one
copyMakeBorder call.  So this script

  1. reads the two files from $REF (default /root/reference) WHERE THEY LIE,
  2. cuts the functions listed in CPP_FUNCS / H_FUNCS out of them by signature + brace matching (no line of the
     reference is edited; the only substitution is the compile-time switch `Do_refine = 0` -> `1`,
     stereoMatching.h:70, which the reference's author flips by hand to run the two-view / refine path),
  3. writes them, TRANSIENTLY (mkdtemp, deleted afterwards), between a class skeleton and the extern "C" entry
     points of smref_shim.inc (both ours), and
  4. compiles that against cv_standin.h (ours: the product's cvmat_lite.h + copyMakeBorder(REFLECT_101), stubs that
     abort for guidedFilter / imwrite paths the default parameters never reach).

Only the .so lands in oracle/_ref/ (git-ignored, travels to the GPU box).  Nothing of the reference is written
into the repository.  What is NOT the reference's code in the result: cv::Mat itself (stand-in; the three OpenCV
semantics the path relies on are pinned by cv2 golden vectors, tests/golden/make_opencv_golden.py), the 3x3 median
of the last refine step (cv::medianBlur in the reference, restated in the shim), and the orchestration of
dispOptimize()/refine() (those two reference functions reference a dozen unrelated subsystems; the shim calls the
same stage functions in the same order, stereoMatching.cpp:1046-1136, 1364-1510)."""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REF", "/root/reference")
OUT = os.path.join(HERE, "_ref")

# (regex that must match the first line of the definition, human name)
CPP_FUNCS = [
    (r"^StereoMatching::StereoMatching\(", "ctor"),
    (r"^void StereoMatching::asdCal\(", "asdCal"),
    (r"^void StereoMatching::censusCal\(", "censusCal"),
    (r"^void StereoMatching::ADCensusCal\(", "ADCensusCal"),
    (r"^int StereoMatching::HammingDistance\(", "HammingDistance"),
    (r"^void StereoMatching::gen_ad_sd_vm\(", "gen_ad_sd_vm"),
    (r"^void StereoMatching::gen_vm_from2vm_exp\(", "gen_vm_from2vm_exp"),
    (r"^static bool judgeColorDif\(", "judgeColorDif"),
    (r"^void StereoMatching::calHorVerDis\(Mat& I, Mat& cross, int L, int L_out,", "calHorVerDis"),
    (r"^void StereoMatching::genTrueHorVerArms\(", "genTrueHorVerArms"),
    (r"^void StereoMatching::adCensus\(", "adCensus"),
    (r"^void StereoMatching::calArms\(.*int L_out", "calArms"),
    (r"^void StereoMatching::initArm\(", "initArm"),
    (r"^void StereoMatching::cbca_core\(", "cbca_core"),
    (r"^void StereoMatching::cbca_aggregate\(", "cbca_aggregate"),
    (r"^void StereoMatching::gen1DCumu\(", "gen1DCumu"),
    (r"^void StereoMatching::gen_dispFromVm\(", "gen_dispFromVm"),
    (r"^void StereoMatching::genfinalVm_cbca\(", "genfinalVm_cbca"),
    (r"^void StereoMatching::sgm\(", "sgm"),
    (r"^void StereoMatching::costScan\(", "costScan"),
    (r"^void StereoMatching::gen_sgm_vm\(", "gen_sgm_vm"),
    (r"^void StereoMatching::wta_Co\(", "wta_Co"),
    (r"^void StereoMatching::LRConsistencyCheck_normal\(", "LRConsistencyCheck_normal"),
    (r"^void StereoMatching::LRConsistencyCheck\(", "LRConsistencyCheck"),
    (r"^void StereoMatching::LRConsistencyCheck_new\(", "LRConsistencyCheck_new"),
    (r"^void StereoMatching::regionVote_my\(", "regionVote_my"),
    (r"^void StereoMatching::properIpol\(", "properIpol"),
    # default-off refiner
    (r"^void StereoMatching::subpixelEnhancement\(", "subpixelEnhancement"),
    (r"^void StereoMatching::WM\(", "WM"),
    (r"^void StereoMatching::discontinuityAdjust\(", "discontinuityAdjust"),
    # aggregation == "NL": the caller of NLCCA::aggreCV (NL/NLCCA.cpp is compiled as its own translation unit below)
    (r"^void StereoMatching::NL\(\)", "NL"),
    # the sequential half of vmTop
    (r"^void StereoMatching::genDispFromTopCostVm2\(", "genDispFromTopCostVm2"),
    # the gradient cost family: "censusGrad" is the selector main_.cpp:15 compiles in
    (r"^void StereoMatching::censusGrad\(", "censusGrad"),
    (r"^void StereoMatching::grad\(vector<Mat>& vm_grad, float Trunc\)", "grad"),
    (r"^void StereoMatching::calGrad\(", "calGrad"),
    (r"^void StereoMatching::calGrad_y\(", "calGrad_y"),
    (r"^void StereoMatching::calgradvm\(", "calgradvm"),
    (r"^void StereoMatching::calgradvm_1d\(", "calgradvm_1d"),
    # the caller's cross-scale step: free functions of stereoMatching.cpp
    (r"^void PrintMat\(const Mat& mat\)", "PrintMat"),
    (r"^void SolveAll\(StereoMatching\*\*& smPyr", "SolveAll"),
]
# member functions defined inside the class body of stereoMatching.h
H_FUNCS = [
    (r"^\tstruct Parameters\s*$", "Parameters"),
    (r"^\tvoid genCensusCode\(vector<Mat>& I, vector<Mat>& census, int R_V, int R_U\)", "genCensusCode"),
    (r"^\tvoid genCensusCode_NC_Sur\(", "genCensusCode_NC_Sur"),
    (r"^\tvoid gen_cenVM_XOR\(", "gen_cenVM_XOR"),
    (r"^\tvoid cal1DCost\(", "cal1DCost"),
    (r"^\tstatic float min4\(", "min4"),
    (r"^\tvoid updateCost\(", "updateCost"),
    (r"^\tvoid selectTopCostFromVolumn\(", "selectTopCostFromVolumn"),
    (r"^\tvoid genDispFromTopCostVm\(Mat& topDisp, Mat& disp\)", "genDispFromTopCostVm"),
]


def read_lines(path):
    with open(path, "rb") as f:
        return f.read().decode("latin-1").replace("\r\n", "\n").split("\n")


def strip_code(line, in_block):
    """Return (code, in_block): the line without // and /* */ comments, string and char literals."""
    out, i, n = [], 0, len(line)
    while i < n:
        if in_block:
            k = line.find("*/", i)
            if k < 0:
                return "".join(out), True
            i, in_block = k + 2, False
            continue
        c = line[i]
        if c == "/" and i + 1 < n and line[i + 1] == "/":
            break
        if c == "/" and i + 1 < n and line[i + 1] == "*":
            in_block, i = True, i + 2
            continue
        if c in "\"'":
            q = c
            i += 1
            while i < n and line[i] != q:
                i += 2 if line[i] == "\\" else 1
            i += 1
            continue
        out.append(c)
        i += 1
    return "".join(out), in_block


def cut(lines, pattern, name, src):
    rx = re.compile(pattern)
    hits = [i for i, l in enumerate(lines) if rx.search(l)]
    if not hits:
        raise SystemExit(f"build_ref_sm: cannot find {name} in {src}")
    start = hits[0]
    first = start
    if first > 0 and lines[first - 1].strip().startswith("template"):
        first -= 1
    depth, seen, in_block = 0, False, False
    for j in range(start, len(lines)):
        s, in_block = strip_code(lines[j], in_block)
        for c in s:
            if c == "{":
                depth += 1
                seen = True
            elif c == "}":
                depth -= 1
        if seen and depth == 0:
            body = lines[first:j + 1]
            # `struct X { ... };` keeps its semicolon
            if name == "Parameters" and not body[-1].rstrip().endswith(";"):
                body[-1] += ";"
            return f"// ---- {src}:{first + 1}-{j + 1} ({name})\n" + "\n".join(body) + "\n"
    raise SystemExit(f"build_ref_sm: unbalanced braces cutting {name} from {src}")


def main():
    h_path = os.path.join(REF, "stereoMatching.h")
    c_path = os.path.join(REF, "stereoMatching.cpp")
    if not (os.path.isfile(h_path) and os.path.isfile(c_path)):
        print(f"build_ref_sm: {REF} not present; keeping any prebuilt {OUT}/libsmref.so", file=sys.stderr)
        return 0
    H = read_lines(h_path)
    Cc = read_lines(c_path)
    switches = [l for l in H[:120] if re.match(r"^\tstatic const bool \w+ = [01];", l)]
    if len(switches) < 20:
        raise SystemExit("build_ref_sm: compile-time switch block of stereoMatching.h not found")
    switches = [re.sub(r"\bDo_refine = 0;", "Do_refine = 1;", l) for l in switches]
    h_parts = {n: cut(H, p, n, "stereoMatching.h") for p, n in H_FUNCS}
    c_named = [(n, cut(Cc, p, n, "stereoMatching.cpp")) for p, n in CPP_FUNCS]

    tmp = tempfile.mkdtemp(prefix="smref_")
    try:
        tu = os.path.join(tmp, "smref_tu.cpp")
        # NL/NLCCA.cpp (the glue between StereoMatching::NL and Yang's classes) compiles unchanged next to transient
        # copies of the NL headers with the same mechanical MSVC -> gcc fixes as build_ref.sh; <opencv2/core.hpp> is
        # the stand-in.  The qx classes themselves come from oracle/_ref/libqxref.so (build_ref.sh ran first).
        nl_ok = os.path.isfile(os.path.join(OUT, "libqxref.so")) and os.path.isfile(os.path.join(REF, "NL", "NLCCA.cpp"))
        nl_objs = []
        if nl_ok:
            nld = os.path.join(tmp, "nl")
            os.makedirs(os.path.join(nld, "opencv2"))
            for h in ("process.h", "direct.h", "io.h"):
                open(os.path.join(nld, h), "w").close()
            for hpp in ("core.hpp", "opencv.hpp"):
                with open(os.path.join(nld, "opencv2", hpp), "w") as f:
                    f.write('#pragma once\n#include "cv_standin.h"\n')
            with open(os.path.join(nld, "prelude.h"), "w") as f:
                f.write("#include <algorithm>\n#include <cstring>\n#include <cmath>\n#include <cstdlib>\n#include <iostream>\nusing namespace std;\n")
            for h in ("qx_basic.h", "qx_mst_kruskals_image.h", "qx_tree_filter.h", "ctmf.h", "qx_nonlocal_cost_aggregation.h",
                      "NLCCA.h", "NLCCA.cpp"):
                txt = open(os.path.join(REF, "NL", h), "rb").read().decode("latin-1").replace("\r\n", "\n")
                if h == "qx_basic.h":
                    txt = re.sub(r"return\(unsigned char\((.*)\)\);}", r"return((unsigned char)(\1));}", txt)
                with open(os.path.join(nld, h), "w", encoding="latin-1") as f:
                    f.write(txt)
            nl_o = os.path.join(tmp, "nlcca.o")
            cxx0 = os.environ.get("CXX", "g++")
            r0 = subprocess.run([cxx0, "-O2", "-fPIC", "-std=c++17", "-w", "-fpermissive", "-ffp-contract=off", "-D__int64=long long",
                                 "-include", os.path.join(nld, "prelude.h"), "-I", nld, "-I", HERE,
                                 "-I", os.path.join(HERE, "..", "mystereomatching_b200", "host"),
                                 "-c", os.path.join(nld, "NLCCA.cpp"), "-o", nl_o], capture_output=True, text=True)
            if r0.returncode != 0:
                sys.stderr.write(r0.stderr[-4000:])
                raise SystemExit("build_ref_sm: NLCCA.cpp compile failed")
            nl_objs = [nl_o, "-L", OUT, "-lqxref", "-Wl,-rpath,$ORIGIN"]
        with open(tu, "w", encoding="latin-1") as f:
            f.write('#include "cv_standin.h"\n')
            if nl_ok:
                f.write('#define SMREF_HAVE_NL 1\n#include "%s"\n' % os.path.join(nld, "NLCCA.h"))
            f.write('#include "smref_class_head.inc"\n')      # ours: `class StereoMatching { public:` + statics
            f.write("\n".join(switches) + "\n")
            f.write(h_parts["Parameters"])
            f.write('#include "smref_class_decls.inc"\n')     # ours: declarations + data members
            for n in ("genCensusCode", "genCensusCode_NC_Sur", "gen_cenVM_XOR", "cal1DCost", "min4", "updateCost",
                      "selectTopCostFromVolumn", "genDispFromTopCostVm"):
                f.write(h_parts[n])
            f.write("};\n")
            f.write('#include "smref_class_tail.inc"\n')      # ours: static member definitions, stubs
            for n, part in c_named:
                if n == "NL" and not nl_ok:
                    continue
                f.write(part)
            f.write('#include "smref_shim.inc"\n')            # ours: extern "C" entry points
        os.makedirs(OUT, exist_ok=True)
        cxx = os.environ.get("CXX", "g++")
        cmd = [cxx, "-O2", "-fPIC", "-shared", "-std=c++17", "-w", "-fpermissive", "-ffp-contract=off",
               "-I", HERE, "-I", os.path.join(HERE, "..", "mystereomatching_b200", "host"),
               tu] + (["-D__int64=long long", "-I", os.path.join(tmp, "nl")] if nl_ok else []) + nl_objs + ["-o", os.path.join(OUT, "libsmref.so")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stderr[-6000:])
            if os.environ.get("SMREF_KEEP"):
                shutil.copy(tu, "/tmp/smref_tu_failed.cpp")
            raise SystemExit("build_ref_sm: compile failed")
        print(f"build_ref_sm: wrote {OUT}/libsmref.so")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
