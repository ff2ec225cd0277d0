"""ctypes binding of the CPU oracle (oracle/liboracle.so) and, when present, of the
compiled reference NL library (oracle/_ref/libqxref.so).

TEST INFRASTRUCTURE.  Import only from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never from the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


class OrcParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("D", "censusFunc", "paths", "iters", "L", "L_out", "tau",
                                       "tau_out", "minL", "corDifThres", "reduCoeffi1")] + \
               [(n, C.c_float) for n in ("adTrunc", "lamAD", "lamCen", "LRmaxDiff", "voteRatio")] + \
               [(n, C.c_int) for n in ("voteS", "voteNums", "DISP_OCC", "do_refine", "aggregation", "costcalc")] + \
               [(n, C.c_float) for n in ("cgLamCen", "cgLamG", "gradTrunc")] + \
               [("pyrLevels", C.c_int), ("crossLambda", C.c_float)]


def default_params(D, paths=4, census_func=3, do_refine=1, aggregation=1, costcalc=0, pyr_levels=1, cross_lambda=-1.0):
    """Reference defaults: stereoMatching.h:204-350, stereoMatching.cpp:905, 5270."""
    return OrcParams(D=D, censusFunc=census_func, paths=paths, iters=2, L=17, L_out=34, tau=20,
                     tau_out=6, minL=1, corDifThres=15, reduCoeffi1=4, adTrunc=1000.0, lamAD=10.0,
                     lamCen=30.0, LRmaxDiff=0.0, voteRatio=0.4, voteS=20, voteNums=2, DISP_OCC=-32,
                     do_refine=do_refine, aggregation=aggregation, costcalc=costcalc, cgLamCen=13.0, cgLamG=1.0,
                     gradTrunc=500.0, pyrLevels=pyr_levels, crossLambda=cross_lambda)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("stereo_oracle.cpp", "nl_oracle.cpp")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    I, F, P = C.c_int, C.c_float, C.c_void_p
    sig = {
        "orc_set_threads": ([I], None), "orc_get_threads": ([], I), "orc_has_openmp": ([], I),
        "orc_bgr2gray": ([u8p, I, I, u8p], None),
        "orc_census_code_length": ([I, I, I], I),
        "orc_census": ([u8p, I, I, I, I, I, u64p], None),
        "orc_hamming_vol": ([u64p, u64p, I, I, I, I, I, F, I, f32p], None),
        "orc_ad_vol": ([u8p, u8p, I, I, I, I, F, f32p], None),
        "orc_combine_exp": ([f32p, f32p, C.c_long, F, F, f32p], None),
        "orc_exp_tables": ([F, F, F, I, f32p, f32p], None),
        "orc_grad_xy": ([u8p, I, I, f32p, f32p], None),
        "orc_grad_vol": ([f32p, f32p, f32p, f32p, u16p, I, I, I, I, F, f32p], None),
        "orc_arms": ([u8p, I, I, I, I, I, I, I, I, u16p], None),
        "orc_arms_intersect": ([u16p, u16p, I, I, I, I, u16p], None),
        "orc_cbca": ([f32p, u16p, u16p, I, I, I, I, I, P], None),
        "orc_sgm_path": ([f32p, u8p, I, I, I, I, I, I, I, f32p], None),
        "orc_sgm": ([f32p, u8p, I, I, I, I, I, I], None),
        "orc_wta": ([f32p, I, I, I, i16p], None),
        "orc_wta_co": ([f32p, I, I, I, I, i16p, i16p], None),
        "orc_select_top": ([f32p, I, I, I, I, F, f32p], None),
        "orc_disp_from_top": ([f32p, u8p, I, I, I, I, I, I, I, I, i16p], None),
        "orc_subpixel": ([i16p, f32p, I, I, I, f32p], None),
        "orc_lrc_normal": ([i16p, i16p, I, I, F], None),
        "orc_lrc_label": ([i16p, i16p, I, I, I, F, I, I, u8p], None),
        "orc_lrc_label_right": ([i16p, i16p, I, I, I, F, I, I, u8p], None),
        "orc_region_vote": ([i16p, u16p, I, I, I, F, I], None),
        "orc_proper_ipol": ([i16p, u8p, I, I, I], None),
        "orc_median3_i16": ([i16p, I, I, i16p], None),
        "orc_median3_f32": ([f32p, I, I, f32p], None),
        "orc_da_edges": ([i16p, I, I, u8p], None),
        "orc_equalize_hist": ([u8p, I, I, u8p], None),
        "orc_gauss3_sigma4": ([u8p, I, I, u8p], None),
        "orc_canny3_l1": ([u8p, I, I, I, I, u8p], None),
        "orc_disc_adjust": ([i16p, f32p, I, I, I, u8p], I),
        "orc_wm": ([i16p, u8p, u8p, I, I, I], I),
        "orc_wm_lenient": ([i16p, u8p, u8p, I, I, I, i32p], I),
        "orc_solve_all_1level": ([f32p, C.c_long, F], None),
        "orc_cal_err": ([i16p, f32p, u8p, I, I, I, f32p, np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")], None),
        "orc_pyr_down_u8": ([u8p, I, I, I, u8p], None),
        "orc_cross_scale_weights": ([I, F, f32p], None),
        "orc_solve_all": ([C.POINTER(C.c_void_p), i32p, i32p, i32p, I, F], None),
        "orc_pipeline": ([u8p, u8p, u8p, u8p, I, I, C.POINTER(OrcParams), i16p, i16p, P, P], None),
        "orc_pipeline_ex": ([u8p, u8p, u8p, u8p, I, I, C.POINTER(OrcParams), i16p, i16p, P, P, P], None),
        "orc_ctmf": ([u8p, u8p, I, I, I, I, I, I], None),
        "orc_mst": ([u8p, I, I, I, i32p, u8p, i32p, i32p, i32p, i32p, P], None),
        "orc_tree_table": ([C.c_double, f64p], None),
        "orc_tree_filter": ([f64p, f64p, I, I, i32p, u8p, i32p, i32p, i32p, f64p], None),
        "orc_nl_aggre": ([u8p, I, I, I, f32p], None),
        "orc_nl": ([u8p, I, I, I, f32p, P], None),
        "orc_nlca_gradient": ([u8p, I, I, f32p], None),
        "orc_nlca_cost": ([u8p, u8p, I, I, I, C.c_double, C.c_double, C.c_double, f64p], None),
        "orc_flip_vol": ([f64p, I, I, I, f64p], None),
        "orc_depth_best_cost": ([f64p, I, I, I, u8p], None),
        "orc_detect_occlusion": ([u8p, u8p, I, I, u8p], None),
        "orc_nlca_disparity": ([u8p, u8p, I, I, I, C.c_double, I, u8p], None),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes, fn.restype = args, res
    _LIB = L
    return L


def ref_lib():
    """The reference's own NL sources compiled by oracle/build_ref.sh, or None."""
    global _REF
    if _REF is not None:
        return _REF
    so = os.path.join(_HERE, "_ref", "libqxref.so")
    if not os.path.exists(so):
        return None
    L = C.CDLL(so)
    I = C.c_int
    L.qxref_ctmf.argtypes = [u8p, u8p, I, I, I, I, I, I, C.c_ulong]
    L.qxref_mst.argtypes = [u8p, I, I, I, i32p, u8p, i32p, i32p, i32p, i32p]
    L.qxref_tree_filter.argtypes = [u8p, I, I, I, C.c_double, f64p, f64p]
    L.qxref_nlca.argtypes = [u8p, u8p, I, I, I, C.c_double, I, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    for f in (L.qxref_ctmf, L.qxref_mst, L.qxref_tree_filter, L.qxref_nlca):
        f.restype = None
    _REF = L
    return L


# ----------------------------------------------------------------------------- numpy-level helpers
def bgr2gray(bgr):
    H, W, _ = bgr.shape
    out = np.empty((H, W), np.uint8)
    lib().orc_bgr2gray(np.ascontiguousarray(bgr), H, W, out)
    return out


def census(gray, func=3, RV=3, RU=4):
    H, W = gray.shape
    nw = (lib().orc_census_code_length(func, RV, RU) + 63) // 64
    out = np.empty((H, W, nw), np.uint64)
    lib().orc_census(np.ascontiguousarray(gray), H, W, func, RV, RU, out)
    return out


def hamming_vol(cL, cR, D, func=3, LOR=0, trunc_rat=1.0):
    H, W, nw = cL.shape
    out = np.empty((H, W, D), np.float32)
    lib().orc_hamming_vol(cL, cR, H, W, D, nw, lib().orc_census_code_length(func, 3, 4), trunc_rat, LOR, out)
    return out


def ad_vol(bgrL, bgrR, D, LOR=0, trunc=1000.0):
    H, W, _ = bgrL.shape
    out = np.empty((H, W, D), np.float32)
    lib().orc_ad_vol(np.ascontiguousarray(bgrL), np.ascontiguousarray(bgrR), H, W, D, LOR, trunc, out)
    return out


def combine_exp(a, b, l0=10.0, l1=30.0):
    out = np.empty_like(a)
    lib().orc_combine_exp(a, b, a.size, l0, l1, out)
    return out


def adcensus_vol(bgrL, bgrR, grayL, grayR, D, LOR=0, func=3, trunc=1000.0, lamAD=10.0, lamCen=30.0):
    ad = ad_vol(bgrL, bgrR, D, LOR, trunc)
    cen = hamming_vol(census(grayL, func), census(grayR, func), D, func, LOR)
    return combine_exp(ad, cen, lamAD, lamCen)


def grad_xy(gray):
    """calGrad / calGrad_y on a gray image (stereoMatching.cpp:271-368)."""
    H, W = gray.shape
    gx, gy = np.empty((H, W), np.float32), np.empty((H, W), np.float32)
    lib().orc_grad_xy(np.ascontiguousarray(gray), H, W, gx, gy)
    return gx, gy


def grad_vol(grayL, grayR, arms_view, D, view=0, trunc=500.0):
    """grad() -> calgradvm (stereoMatching.cpp:603-656, 388-455); arms_view = HVL[view]."""
    H, W = grayL.shape
    gx0, gy0 = grad_xy(grayL)
    gx1, gy1 = grad_xy(grayR)
    out = np.empty((H, W, D), np.float32)
    lib().orc_grad_vol(gx0, gx1, gy0, gy1, np.ascontiguousarray(arms_view), H, W, D, view, trunc, out)
    return out


def censusgrad_vol(bgrL, bgrR, grayL, grayR, D, view=0, func=3, lamCen=13.0, lamG=1.0, trunc=500.0):
    """censusGrad (stereoMatching.cpp:25-48): 2 - exp(-census/lamCen) - exp(-grad/lamG)."""
    cl, cr = census(grayL, func), census(grayR, func)
    ham = hamming_vol(cl, cr, D, func, view)
    a = arms(bgrL if view == 0 else bgrR)
    return combine_exp(ham, grad_vol(grayL, grayR, a, D, view, trunc), lamCen, lamG)


def cal_err(dp, gt, mask, thres=1):
    """calErr<short> (stereoMatching.h:1748-1825): (PBM, RMS, sumNum, errorNum) over mask == 255."""
    H, W = dp.shape
    out, cnt = np.empty(2, np.float32), np.empty(2, np.int64)
    lib().orc_cal_err(np.ascontiguousarray(dp, np.int16), np.ascontiguousarray(gt, np.float32),
                      np.ascontiguousarray(mask, np.uint8), H, W, thres, out, cnt)
    return float(out[0]), float(out[1]), int(cnt[0]), int(cnt[1])


def pyr_down(img):
    """cv::pyrDown on an 8-bit image (1 or 3 channels)."""
    H, W = img.shape[:2]
    cn = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty(((H + 1) // 2, (W + 1) // 2) + ((cn,) if img.ndim == 3 else ()), np.uint8)
    lib().orc_pyr_down_u8(np.ascontiguousarray(img), H, W, cn, out)
    return out


def cross_scale_weights(levels, lam=0.3):
    """Row 0 of regMat.inv() in SolveAll (stereoMatching.cpp:2147-2170)."""
    w = np.empty(levels, np.float32)
    lib().orc_cross_scale_weights(levels, lam, w)
    return w


def solve_all(vols, lam=0.3):
    """SolveAll (stereoMatching.cpp:2142-2208) on a list of per-level volumes; returns the new level-0 volume."""
    vols = [np.ascontiguousarray(v, np.float32) for v in vols]
    out = vols[0].copy()
    ptrs = (C.c_void_p * len(vols))(out.ctypes.data, *[v.ctypes.data for v in vols[1:]])
    Hs = np.array([v.shape[0] for v in vols], np.int32)
    Ws = np.array([v.shape[1] for v in vols], np.int32)
    Ds = np.array([v.shape[2] for v in vols], np.int32)
    lib().orc_solve_all(ptrs, Hs, Ws, Ds, len(vols), lam)
    return out


def arms(img, L=17, L_out=34, tau=20, tau_out=6, minL=1):
    H, W, Cn = img.shape
    out = np.empty((H, W, 5), np.uint16)
    lib().orc_arms(np.ascontiguousarray(img), H, W, Cn, L, L_out, tau, tau_out, minL, out)
    return out


def arms_intersect(aL, aR, D, view=0):
    H, W, _ = aL.shape
    out = np.empty((H, W, D, 5), np.uint16)
    lib().orc_arms_intersect(aL, aR, H, W, D, view, out)
    return out


def cbca(vol, aL, aR, iters=2, view=0, want_area=False):
    H, W, D = vol.shape
    out = np.ascontiguousarray(vol.copy())
    area = np.empty((H, W, D), np.int32) if want_area else None
    lib().orc_cbca(out, aL, aR, H, W, D, iters, view, area.ctypes.data if want_area else None)
    return (out, area) if want_area else out


def sgm_path(vol, bgr, path, thr=15, redu=4):
    H, W, D = vol.shape
    rv = (+1, -1, 0, 0, +1, +1, -1, -1)[path]
    ru = (0, 0, +1, -1, -1, +1, +1, -1)[path]
    out = np.empty_like(vol)
    lib().orc_sgm_path(np.ascontiguousarray(vol), np.ascontiguousarray(bgr), H, W, D, rv, ru, thr, redu, out)
    return out


def sgm(vol, bgr, paths=4, thr=15, redu=4):
    H, W, D = vol.shape
    out = np.ascontiguousarray(vol.copy())
    lib().orc_sgm(out, np.ascontiguousarray(bgr), H, W, D, paths, thr, redu)
    return out


def wta(vol):
    H, W, D = vol.shape
    out = np.empty((H, W), np.int16)
    lib().orc_wta(np.ascontiguousarray(vol), H, W, D, out)
    return out


def select_top(vol, num, thres):
    """selectTopCostFromVolumn (stereoMatching.h:2405-2461): float [H][W][num+1][2]."""
    H, W, D = vol.shape
    out = np.empty((H, W, num + 1, 2), np.float32)
    lib().orc_select_top(np.ascontiguousarray(vol, np.float32), H, W, D, num, thres, out)
    return out


def disp_from_top(top, bgr, version=2, method=0, ts=10, has_cir2=True, color_limit=False, init=None):
    """genDispFromTopCostVm (version 1, stereoMatching.h:2466-2545) / genDispFromTopCostVm2 (version 2,
    stereoMatching.cpp:1514-1886).  top: float [H][W][num+1][2]; init: the map's content on entry (zeros)."""
    H, W, n1, _ = top.shape
    out = np.zeros((H, W), np.int16) if init is None else np.ascontiguousarray(init, np.int16).copy()
    lib().orc_disp_from_top(np.ascontiguousarray(top, np.float32), np.ascontiguousarray(bgr, np.uint8), H, W, n1 - 1, version,
                            method, ts, int(has_cir2), int(color_limit), out)
    return out


def subpixel(disp, vol):
    """subpixelEnhancement (stereoMatching.cpp:6138-6166): float [H][W]."""
    H, W, D = vol.shape
    out = np.empty((H, W), np.float32)
    lib().orc_subpixel(np.ascontiguousarray(disp, np.int16), np.ascontiguousarray(vol, np.float32), H, W, D, out)
    return out


def wta_co(vol, scale=16):
    H, W, D = vol.shape
    d1 = np.empty((H, W), np.int16)
    d2 = np.empty((H, W), np.int16)
    lib().orc_wta_co(np.ascontiguousarray(vol), H, W, D, scale, d1, d2)
    return d1, d2


def lrc_normal(d1, d2, max_diff=0.0):
    out = np.ascontiguousarray(d1.copy())
    lib().orc_lrc_normal(out, np.ascontiguousarray(d2), d1.shape[0], d1.shape[1], max_diff)
    return out


def lrc_label(d1, d2, D, max_diff=0.0, occ=-32, mis=-48):
    out = np.ascontiguousarray(d1.copy())
    mask = np.empty(d1.shape, np.uint8)
    lib().orc_lrc_label(out, np.ascontiguousarray(d2), d1.shape[0], d1.shape[1], D, max_diff, occ, mis, mask)
    return out, mask


def lrc_label_right(d1, d2, D, max_diff=0.0, occ=-32, mis=-48):
    """LRConsistencyCheck, LOR = 1 (stereoMatching.cpp:2336-2364): returns (labelled right map, errMask1)."""
    out = np.ascontiguousarray(d2.copy())
    mask1 = np.empty(d2.shape, np.uint8)
    lib().orc_lrc_label_right(np.ascontiguousarray(d1), out, d1.shape[0], d1.shape[1], D, max_diff, occ, mis, mask1)
    return out, mask1


def region_vote(dp, arms_l, D, ratio=0.4, S=20):
    out = np.ascontiguousarray(dp.copy())
    lib().orc_region_vote(out, arms_l, dp.shape[0], dp.shape[1], D, ratio, S)
    return out


def proper_ipol(dp, bgr, occ=-32):
    out = np.ascontiguousarray(dp.copy())
    lib().orc_proper_ipol(out, np.ascontiguousarray(bgr), dp.shape[0], dp.shape[1], occ)
    return out


def equalize_hist(img):
    out = np.empty_like(np.ascontiguousarray(img, np.uint8))
    lib().orc_equalize_hist(np.ascontiguousarray(img, np.uint8), img.shape[0], img.shape[1], out)
    return out


def gauss3_sigma4(img):
    out = np.empty_like(np.ascontiguousarray(img, np.uint8))
    lib().orc_gauss3_sigma4(np.ascontiguousarray(img, np.uint8), img.shape[0], img.shape[1], out)
    return out


def canny3_l1(img, low, high):
    out = np.empty_like(np.ascontiguousarray(img, np.uint8))
    lib().orc_canny3_l1(np.ascontiguousarray(img, np.uint8), img.shape[0], img.shape[1], int(low), int(high), out)
    return out


def da_edges(disp):
    """The edge map discontinuityAdjust works on (stereoMatching.cpp:6059-6064): convertTo(8U) -> equalizeHist ->
    GaussianBlur(3x3, 4) -> Canny(20, 60, 3)."""
    d = np.ascontiguousarray(disp, np.int16)
    out = np.empty(d.shape, np.uint8)
    lib().orc_da_edges(d, d.shape[0], d.shape[1], out)
    return out


def disc_adjust(disp, vol):
    """discontinuityAdjust (stereoMatching.cpp:6057-6135): (adjusted map, edge map, number of labels >= D met)."""
    H, W, D = vol.shape
    out = np.ascontiguousarray(disp, np.int16).copy()
    edge = np.empty((H, W), np.uint8)
    bad = lib().orc_disc_adjust(out, np.ascontiguousarray(vol, np.float32), H, W, D, edge)
    return out, edge, bad


def wm(disp, mask, bgr, D):
    """WM (stereoMatching.cpp:7340-7393): bilateral weighted median on the mask > 0 pixels; labels must lie in [0, D)."""
    out = np.ascontiguousarray(disp, np.int16).copy()
    H, W = out.shape
    n = lib().orc_wm(out, np.ascontiguousarray(mask, np.uint8), np.ascontiguousarray(bgr, np.uint8), H, W, D)
    if n < 0:
        raise ValueError("WM: a label outside [0, D) inside a window: undefined in the reference (stereoMatching.cpp:7371)")
    return out


def wm_lenient(disp, mask, bgr, D):
    """WM with the defined extension of sm_wm for labels outside [0, D): (map, number of such window neighbours)."""
    out = np.ascontiguousarray(disp, np.int16).copy()
    H, W = out.shape
    bad = np.zeros(1, np.int32)
    lib().orc_wm_lenient(out, np.ascontiguousarray(mask, np.uint8), np.ascontiguousarray(bgr, np.uint8), H, W, D, bad)
    return out, int(bad[0])


def median3_f32(dp):
    """cv::medianBlur(CV_32F, 3) (stereoMatching.cpp:1490)."""
    dp = np.ascontiguousarray(dp, np.float32)
    out = np.empty_like(dp)
    lib().orc_median3_f32(dp, dp.shape[0], dp.shape[1], out)
    return out


def median3_i16(dp):
    out = np.empty_like(dp)
    lib().orc_median3_i16(np.ascontiguousarray(dp), dp.shape[0], dp.shape[1], out)
    return out


def pipeline(bgrL, bgrR, grayL, grayR, params, want_vol=False, want_agg=False):
    """The whole chain.  Returns (dispL, dispR, vm[0] after sgm or None, stage ms); with want_agg the dict also holds
    "agg" = vm[0] as costCalculate() leaves it (aggregated, before sgm)."""
    H, W, _ = bgrL.shape
    dl = np.empty((H, W), np.int16)
    dr = np.empty((H, W), np.int16)
    vol = np.empty((H, W, params.D), np.float32) if want_vol else None
    agg = np.empty((H, W, params.D), np.float32) if want_agg else None
    ms = np.zeros(8, np.float32)
    lib().orc_pipeline_ex(np.ascontiguousarray(bgrL), np.ascontiguousarray(bgrR),
                          np.ascontiguousarray(grayL), np.ascontiguousarray(grayR), H, W,
                          C.byref(params), dl, dr, vol.ctypes.data if want_vol else None, ms.ctypes.data,
                          agg.ctypes.data if want_agg else None)
    names = ("census", "cost", "arms", "cbca", "sgm", "wta", "refine", "total")
    info = dict(zip(names, ms.tolist()))
    if want_agg:
        info["agg"] = agg
    return dl, dr, vol, info


def ctmf(img, r):
    img = np.ascontiguousarray(img)
    H, W = img.shape[:2]
    cn = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty_like(img)
    lib().orc_ctmf(img, out, W, H, W * cn, W * cn, r, cn)
    return out


def mst(img):
    img = np.ascontiguousarray(img)
    H, W = img.shape[:2]
    cn = 1 if img.ndim == 2 else img.shape[2]
    N = H * W
    r = dict(parent=np.empty(N, np.int32), weight=np.empty(N, np.uint8), rank=np.empty(N, np.int32),
             nr_child=np.empty(N, np.int32), children=np.empty(3 * N, np.int32), order=np.empty(N, np.int32))
    lib().orc_mst(img, H, W, cn, r["parent"], r["weight"], r["rank"], r["nr_child"], r["children"], r["order"], None)
    return r


def nl_aggre(bgr, vol):
    H, W, D = vol.shape
    out = np.ascontiguousarray(vol.copy())
    lib().orc_nl_aggre(np.ascontiguousarray(bgr), H, W, D, out)
    return out


def nl(bgr, vol):
    H, W, D = vol.shape
    out = np.ascontiguousarray(vol.copy())
    disp = np.empty((H, W), np.int16)
    lib().orc_nl(np.ascontiguousarray(bgr), H, W, D, out, disp.ctypes.data)
    return out, disp


def ref_ctmf(img, r):
    L = ref_lib()
    img = np.ascontiguousarray(img)
    H, W = img.shape[:2]
    cn = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty_like(img)
    L.qxref_ctmf(img, out, W, H, W * cn, W * cn, r, cn, H * W * cn)
    return out


def ref_mst(img):
    L = ref_lib()
    img = np.ascontiguousarray(img)
    H, W = img.shape[:2]
    cn = 1 if img.ndim == 2 else img.shape[2]
    N = H * W
    r = dict(parent=np.empty(N, np.int32), weight=np.empty(N, np.uint8), rank=np.empty(N, np.int32),
             nr_child=np.empty(N, np.int32), children=np.empty(3 * N, np.int32), order=np.empty(N, np.int32))
    L.qxref_mst(img, H, W, cn, r["parent"], r["weight"], r["rank"], r["nr_child"], r["children"], r["order"])
    return r


def ref_tree_filter(bgr, vol64, sigma=0.1):
    L = ref_lib()
    H, W, D = vol64.shape
    cost = np.ascontiguousarray(vol64.copy())
    tmp = np.empty_like(cost)
    L.qxref_tree_filter(np.ascontiguousarray(bgr), H, W, D, sigma, cost, tmp)
    return cost


# ----------------------------------------------------------------------------- Yang's driver (qx_nonlocal_cost_aggregation)
def nlca_gradient(img):
    H, W, _ = img.shape
    out = np.empty((H, W), np.float32)
    lib().orc_nlca_gradient(np.ascontiguousarray(img), H, W, out)
    return out


def nlca_cost(left, right, D, maxc=7.0, maxg=2.0, wc=0.11):
    H, W, _ = left.shape
    out = np.empty((H, W, D), np.float64)
    lib().orc_nlca_cost(np.ascontiguousarray(left), np.ascontiguousarray(right), H, W, D, maxc, maxg, wc, out)
    return out


def flip_vol(vol):
    H, W, D = vol.shape
    out = np.empty_like(vol)
    lib().orc_flip_vol(np.ascontiguousarray(vol), H, W, D, out)
    return out


def depth_best_cost(vol):
    H, W, D = vol.shape
    out = np.empty((H, W), np.uint8)
    lib().orc_depth_best_cost(np.ascontiguousarray(vol), H, W, D, out)
    return out


def detect_occlusion(dl, dr):
    H, W = dl.shape
    out = np.empty((H, W), np.uint8)
    lib().orc_detect_occlusion(np.ascontiguousarray(dl), np.ascontiguousarray(dr), H, W, out)
    return out


def nlca_disparity(left, right, D, sigma=0.1, post=False):
    H, W, _ = left.shape
    out = np.empty((H, W), np.uint8)
    lib().orc_nlca_disparity(np.ascontiguousarray(left), np.ascontiguousarray(right), H, W, D, sigma, int(post), out)
    return out


def ref_nlca(left, right, D, sigma=0.1, post=False, want_disp=True):
    """The reference's own qx_nonlocal_cost_aggregation (compiled into oracle/_ref)."""
    L = ref_lib()
    H, W, _ = left.shape
    cost = np.empty((H, W, D), np.float64)
    costR = np.empty((H, W, D), np.float64)
    grad = np.empty((H, W), np.float32)
    disp = np.empty((H, W), np.uint8) if want_disp else None
    L.qxref_nlca(np.ascontiguousarray(left), np.ascontiguousarray(right), H, W, D, sigma, int(post), cost.ctypes.data,
                 costR.ctypes.data, grad.ctypes.data, disp.ctypes.data if want_disp else None)
    return dict(cost=cost, cost_right=costR, grad_left=grad, disp=disp)


# ----------------------------------------------------------------------------- compiled StereoMatching reference
_SMREF = None


def smref_lib():
    """The reference's own stereoMatching.{h,cpp} hot-path function bodies compiled by oracle/build_ref_sm.py
    (oracle/_ref/libsmref.so), or None when it has not been built."""
    global _SMREF
    if _SMREF is not None:
        return _SMREF
    so = os.path.join(_HERE, "_ref", "libsmref.so")
    if not os.path.exists(so):
        return None
    L = C.CDLL(so)
    I, F, P, D_ = C.c_int, C.c_float, C.c_void_p, C.c_double
    sig = {
        "smref_create": ([u8p, u8p, u8p, u8p, I, I, I], P), "smref_destroy": ([P], None),
        "smref_do_refine": ([], I),
        "smref_set_param": ([P, C.c_char_p, D_], I), "smref_get_param": ([P, C.c_char_p], D_),
        "smref_census": ([P, I, u64p], I), "smref_census_cal": ([P, I, f32p, f32p], None),
        "smref_ad": ([P, I, F, f32p], None), "smref_combine_exp": ([P, f32p, f32p, F, F, f32p], None),
        "smref_adcensus": ([P], None), "smref_grad_xy": ([P, I, f32p, f32p], None),
        "smref_grad_vm": ([P, F, f32p, f32p], None), "smref_censusgrad": ([P], None), "smref_get_vm": ([P, I, f32p], None), "smref_set_vm": ([P, I, f32p], None),
        "smref_arms": ([P, u16p, u16p], None), "smref_arms_intersection": ([P, I, u16p], None),
        "smref_cbca": ([P, I], None), "smref_gen1dcumu": ([P, f32p, i32p, I, I], None),
        "smref_cal1dcost": ([P, I, f32p, i32p, I, I, I], None), "smref_genfinal": ([P, f32p, i32p], None),
        "smref_update_cost": ([P, I, f32p, I, I, I, I, I], None), "smref_cost_scan": ([P, I, I, f32p], None), "smref_sgm": ([P, I, I], None),
        "smref_wta": ([P, I, i16p], None), "smref_wta_co": ([P, I, i16p, i16p], None),
        "smref_select_top": ([P, I, I, F, f32p], None),
        "smref_subpixel": ([P, i16p, f32p], None),
        "smref_disp_from_top": ([P, f32p, I, I, I, I, I, I, i16p], None),
        "smref_wm": ([P, i16p, u8p], None),
        "smref_disc_adjust": ([P, i16p], None),
        "smref_lrc_normal": ([P, i16p, i16p], None), "smref_lrc_label": ([P, i16p, i16p, I, P], None),
        "smref_lrc_new": ([P, i16p, i16p, u8p], None),
        "smref_region_vote": ([P, i16p, F, I], None), "smref_proper_ipol": ([P, i16p], None),
        "smref_median3_i16": ([i16p, I, I, i16p], None),
        "smref_pipeline": ([P, I, I, P, P, P, P], None), "smref_pipeline2": ([P, I, I, I, P, P, P, P], None),
        "smref_pipeline_pyr": ([P, I, I, I, I, F, P, P, P], None),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes, fn.restype = args, res
    if hasattr(L, "smref_pipeline_nl"):   # present when NL/NLCCA.cpp + libqxref.so were available at build time
        L.smref_pipeline_nl.argtypes, L.smref_pipeline_nl.restype = [P, I, P, P, P, P], None
    _SMREF = L
    return L


class SmRef:
    """One compiled-reference `StereoMatching` instance (constructed as main_.cpp:60-64,138 does) on one stereo
    pair.  Methods are the reference's own functions; see oracle/smref_shim.inc."""

    def __init__(self, bgrL, bgrR, grayL, grayR, D):
        L = smref_lib()
        if L is None:
            raise RuntimeError("oracle/_ref/libsmref.so is not built (python oracle/build_ref_sm.py)")
        self.L = L
        self.H, self.W, _ = bgrL.shape
        self.D = D
        self.h = L.smref_create(np.ascontiguousarray(bgrL), np.ascontiguousarray(bgrR),
                                np.ascontiguousarray(grayL), np.ascontiguousarray(grayR), self.H, self.W, D - 1)

    def close(self):
        if self.h:
            self.L.smref_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set(self, name, v):
        if self.L.smref_set_param(self.h, name.encode(), float(v)) != 0:
            raise KeyError(name)

    def get(self, name):
        v = self.L.smref_get_param(self.h, name.encode())
        if v < -1e299:
            raise KeyError(name)
        return v

    def _vol(self):
        return np.empty((self.H, self.W, self.D), np.float32)

    def census(self, func=3):
        nw = 2 if func == 3 else 1
        out = np.empty((2, self.H, self.W, nw), np.uint64)
        assert self.L.smref_census(self.h, func, out) == nw
        return out[0], out[1]

    def census_cal(self, func=3):
        a, b = self._vol(), self._vol()
        self.L.smref_census_cal(self.h, func, a, b)
        return a, b

    def ad(self, LOR=0, trunc=1000.0):
        a = self._vol()
        self.L.smref_ad(self.h, LOR, trunc, a)
        return a

    def combine_exp(self, a, b, l0=10.0, l1=30.0):
        o = self._vol()
        self.L.smref_combine_exp(self.h, np.ascontiguousarray(a), np.ascontiguousarray(b), l0, l1, o)
        return o

    def adcensus(self):
        self.L.smref_adcensus(self.h)
        return self.vm(0), self.vm(1)

    def grad_xy(self, img):
        gx, gy = np.empty((self.H, self.W), np.float32), np.empty((self.H, self.W), np.float32)
        self.L.smref_grad_xy(self.h, img, gx, gy)
        return gx, gy

    def grad_vm(self, trunc=500.0):
        a, b = self._vol(), self._vol()
        self.L.smref_grad_vm(self.h, trunc, a, b)
        return a, b

    def censusgrad(self):
        self.L.smref_censusgrad(self.h)
        return self.vm(0), self.vm(1)

    def vm(self, view):
        a = self._vol()
        self.L.smref_get_vm(self.h, view, a)
        return a

    def set_vm(self, view, vol):
        assert vol.shape == (self.H, self.W, self.D)
        self.L.smref_set_vm(self.h, view, np.ascontiguousarray(vol, np.float32))

    def arms(self):
        a = np.empty((self.H, self.W, 5), np.uint16)
        b = np.empty((self.H, self.W, 5), np.uint16)
        self.L.smref_arms(self.h, a, b)
        return a, b

    def arms_intersection(self, view):
        o = np.empty((self.H, self.W, self.D, 5), np.uint16)
        self.L.smref_arms_intersection(self.h, view, o)
        return o

    def gen1dcumu(self, vol, area, dv, du):
        v, a = np.ascontiguousarray(vol, np.float32).copy(), np.ascontiguousarray(area, np.int32).copy()
        self.L.smref_gen1dcumu(self.h, v, a, dv, du)
        return v, a

    def cal1dcost(self, view, vol, area, dv, du, direc):
        v, a = np.ascontiguousarray(vol, np.float32).copy(), np.ascontiguousarray(area, np.int32).copy()
        self.L.smref_cal1dcost(self.h, view, v, a, dv, du, direc)
        return v, a

    def genfinal(self, vol, area):
        v = np.ascontiguousarray(vol, np.float32).copy()
        self.L.smref_genfinal(self.h, v, np.ascontiguousarray(area, np.int32))
        return v

    def update_cost(self, view, Lr, v, u, rv, ru, pre_is_inner):
        l = np.ascontiguousarray(Lr, np.float32).copy()
        self.L.smref_update_cost(self.h, view, l, v, u, rv, ru, int(pre_is_inner))
        return l

    def cbca(self, iters=2):
        self.L.smref_cbca(self.h, iters)
        return self.vm(0), self.vm(1)

    def cost_scan(self, view, path):
        o = self._vol()
        self.L.smref_cost_scan(self.h, view, path, o)
        return o

    def sgm(self, view, paths):
        self.L.smref_sgm(self.h, view, paths)
        return self.vm(view)

    def wta(self, view):
        o = np.empty((self.H, self.W), np.int16)
        self.L.smref_wta(self.h, view, o)
        return o

    def select_top(self, view, num, thres):
        """The reference's own selectTopCostFromVolumn on a clone of vm[view] (topDisp zero-initialised)."""
        o = np.empty((self.H, self.W, num + 1, 2), np.float32)
        self.L.smref_select_top(self.h, view, num, thres, o)
        return o

    def disp_from_top(self, top, version=2, method=0, ts=10, has_cir2=True, color_limit=False, init=None):
        """The reference's own genDispFromTopCostVm (version 1) / genDispFromTopCostVm2 (version 2) on `top`."""
        out = np.zeros((self.H, self.W), np.int16) if init is None else np.ascontiguousarray(init, np.int16).copy()
        self.L.smref_disp_from_top(self.h, np.ascontiguousarray(top, np.float32), top.shape[2] - 1, version, method, ts,
                                   int(has_cir2), int(color_limit), out)
        return out

    def subpixel(self, disp):
        """The reference's own subpixelEnhancement on the given disparity map and vm[0]."""
        o = np.empty((self.H, self.W), np.float32)
        self.L.smref_subpixel(self.h, np.ascontiguousarray(disp, np.int16), o)
        return o

    def disc_adjust(self, disp):
        """The reference's own discontinuityAdjust on the given map and vm[0] (its three OpenCV calls go to the restated,
        cv2-pinned functions of oracle/opencv_restated.h)."""
        o = np.ascontiguousarray(disp, np.int16).copy()
        self.L.smref_disc_adjust(self.h, o)
        return o

    def wm(self, disp, mask):
        """The reference's own WM on the given map and mask with I_c[0] as guidance."""
        o = np.ascontiguousarray(disp, np.int16).copy()
        self.L.smref_wm(self.h, o, np.ascontiguousarray(mask, np.uint8))
        return o

    def wta_co(self, view=0):
        a = np.empty((self.H, self.W), np.int16)
        b = np.empty((self.H, self.W), np.int16)
        self.L.smref_wta_co(self.h, view, a, b)
        return a, b

    def lrc_normal(self, d1, d2):
        a = np.ascontiguousarray(d1, np.int16).copy()
        self.L.smref_lrc_normal(self.h, a, np.ascontiguousarray(d2, np.int16))
        return a

    def lrc_label(self, d1, d2, LOR=0):
        a = np.ascontiguousarray(d1, np.int16).copy()
        b = np.ascontiguousarray(d2, np.int16).copy()
        m = np.zeros((self.H, self.W), np.uint8)
        self.L.smref_lrc_label(self.h, a, b, LOR, m.ctypes.data if LOR == 0 else None)
        return a, b, m

    def lrc_new(self, d1, d2, mask):
        m = np.ascontiguousarray(mask, np.uint8).copy()
        self.L.smref_lrc_new(self.h, np.ascontiguousarray(d1, np.int16), np.ascontiguousarray(d2, np.int16), m)
        return m

    def region_vote(self, dp, ratio=0.4, S=20):
        a = np.ascontiguousarray(dp, np.int16).copy()
        self.L.smref_region_vote(self.h, a, ratio, S)
        return a

    def proper_ipol(self, dp):
        a = np.ascontiguousarray(dp, np.int16).copy()
        self.L.smref_proper_ipol(self.h, a)
        return a

    def pipeline_pyr(self, levels, lam=0.3, paths=8, iters=2, costcalc=0):
        """main_.cpp:131-166 with PY_LEV = levels: returns (refined left map, vm[0], vm[1] right after SolveAll)."""
        rf = np.empty((self.H, self.W), np.int16)
        v0, v1 = self._vol(), self._vol()
        self.L.smref_pipeline_pyr(self.h, costcalc, paths, iters, levels, lam, rf.ctypes.data, v0.ctypes.data, v1.ctypes.data)
        return rf, v0, v1

    def has_nl(self):
        return hasattr(self.L, "smref_pipeline_nl")

    def pipeline_nl(self, paths=4, want_vol=False):
        """aggregation == "NL": ADCensusCal, the reference's own StereoMatching::NL() over its own NLCCA / qx_tree_filter,
        dispOptimize, refine.  Returns (wtaL, wtaR, refined, vm[0] right after NL() or None)."""
        wl = np.empty((self.H, self.W), np.int16)
        wr = np.empty((self.H, self.W), np.int16)
        rf = np.empty((self.H, self.W), np.int16)
        vol = self._vol() if want_vol else None
        self.L.smref_pipeline_nl(self.h, paths, wl.ctypes.data, wr.ctypes.data, rf.ctypes.data,
                                 vol.ctypes.data if want_vol else None)
        return wl, wr, rf, vol

    def pipeline(self, paths=4, iters=2, want_vol=False, costcalc=0):
        wl = np.empty((self.H, self.W), np.int16)
        wr = np.empty((self.H, self.W), np.int16)
        rf = np.empty((self.H, self.W), np.int16)
        vol = self._vol() if want_vol else None
        self.L.smref_pipeline2(self.h, costcalc, paths, iters, wl.ctypes.data, wr.ctypes.data, rf.ctypes.data,
                               vol.ctypes.data if want_vol else None)
        return wl, wr, rf, vol
