"""ctypes binding of the sm_b200 C ABI (include/sm_b200.h).

Plumbing only: device memory comes from torch tensors (``.data_ptr()``), the work
runs on the torch current stream, and every call goes straight into
``libsm_b200.so``.  There is no fallback of any kind: if the library is missing
or a call fails, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SM_B200_LIB") or os.path.join(_HERE, "libsm_b200.so")   # override: tuning builds
_LIB = None


class SmError(RuntimeError):
    pass


class SmParams(C.Structure):
    """Mirror of `struct sm_params` (include/sm_b200.h)."""
    _fields_ = [
        ("numDisparities", C.c_int), ("censusFunc", C.c_int), ("adTrunc", C.c_float),
        ("lamAD", C.c_float), ("lamCen", C.c_float),
        ("cbca_crossL", C.c_int), ("cbca_crossL_out", C.c_int), ("cbca_cTresh", C.c_int),
        ("cbca_cTresh_out", C.c_int), ("cbca_minArmL", C.c_int), ("cbca_iterationNum", C.c_int),
        ("sgm_paths", C.c_int), ("sgm_corDifThres", C.c_int), ("sgm_reduCoeffi1", C.c_int),
        ("LRmaxDiff", C.c_float), ("region_vote_nums", C.c_int), ("regVote_SThres", C.c_int),
        ("regVote_hratioThres", C.c_float), ("DISP_OCC", C.c_int), ("DISP_MIS", C.c_int),
        ("aggregation", C.c_int), ("Do_refine", C.c_int), ("Do_LRConsis", C.c_int),
        ("Do_regionVote", C.c_int), ("Do_properIpol", C.c_int), ("Do_lastMedianBlur", C.c_int),
        ("crossScaleLambda", C.c_float), ("sgm_grouped", C.c_int),
        ("costcalculation", C.c_int), ("cg_lamCen", C.c_float), ("cg_lamG", C.c_float), ("gradTrunc", C.c_float),
        ("pyramidLevels", C.c_int),
        ("Do_vmTop", C.c_int), ("vmTop_method", C.c_int), ("vmTop_Num", C.c_int), ("vmTop_thres", C.c_float),
        ("vmTop_ts", C.c_int), ("vmTop_hasCir2", C.c_int), ("vmTop_cir3_doColorLimit", C.c_int),
        ("keep_right_volume", C.c_int),
    ]


# name -> (argtypes, restype); the not-gpu test checks this table against the header.
_P, _I, _F, _D, _Z, _LL = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t, C.c_longlong
SIGNATURES = {
    "sm_params_default": ([C.POINTER(SmParams), _I], None),
    "sm_ctx_create": ([C.POINTER(_P), _I, _P], _I),
    "sm_ctx_destroy": ([_P], _I),
    "sm_ctx_sync": ([_P], _I),
    "sm_ctx_stream": ([_P], _P),
    "sm_last_error": ([], C.c_char_p),
    "sm_device_count": ([], _I),
    "sm_dev_alloc": ([_P, C.POINTER(_P), _Z], _I),
    "sm_dev_free": ([_P, _P], _I),
    "sm_host_alloc_pinned": ([C.POINTER(_P), _Z], _I),
    "sm_host_free_pinned": ([_P], _I),
    "sm_memcpy_h2d": ([_P, _P, _P, _Z], _I),
    "sm_memcpy_d2h": ([_P, _P, _P, _Z], _I),
    "sm_memset": ([_P, _P, _I, _Z], _I),
    "sm_memcpy_d2d": ([_P, _P, _P, _Z], _I),
    "sm_ctx_launch_count": ([_P], _LL),
    "sm_bgr2gray": ([_P, _P, _I, _I, _P], _I),
    "sm_census": ([_P, _P, _I, _I, _I, _P], _I),
    "sm_census_words": ([_I], _I),
    "sm_census_code_length": ([_I], _I),
    "sm_cost_hamming": ([_P, _P, _P, _I, _I, _I, _I, _I, _P], _I),
    "sm_cost_hamming_u16": ([_P, _P, _P, _I, _I, _I, _I, _I, _P], _I),
    "sm_cost_ad": ([_P, _P, _P, _I, _I, _I, _I, _F, _P], _I),
    "sm_cost_adcensus": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _I, _P], _I),
    "sm_combine_exp": ([_P, _P, _P, _Z, _F, _F, _P], _I),
    "sm_cumsum_1d": ([_P, _P, _P, _I, _I, _I, _I, _I], _I),
    "sm_span_1d": ([_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I], _I),
    "sm_div_area": ([_P, _P, _P, _Z], _I),
    "sm_update_cost": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I], _I),
    "sm_lrc_mask": ([_P, _P, _P, _I, _I, _P], _I),
    "sm_vol_to_f32": ([_P, _P, _I, _Z, _P], _I),
    "sm_cal_err": ([_P, _P, _P, _P, _I, _I, _I, C.POINTER(_LL), C.POINTER(_LL), C.POINTER(_D)], _I),
    "sm_pyr_down_u8": ([_P, _P, _I, _I, _I, _P], _I),
    "sm_cross_scale_weights": ([_I, _F, _P], _I),
    "sm_cross_scale": ([_P, _P, _P, _P, _P, _I, _F], _I),
    "sm_grad_xy": ([_P, _P, _I, _I, _P, _P], _I),
    "sm_cost_grad": ([_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _I, _P], _I),
    "sm_cost_censusgrad": ([_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _I, _P], _I),
    "sm_arms": ([_P, _P, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "sm_arms_intersect": ([_P, _P, _P, _I, _I, _I, _I, _P], _I),
    "sm_cbca": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _I], _I),
    "sm_median_u8": ([_P, _P, _P, _I, _I, _I, _I], _I),
    "sm_mst_build": ([_P, _P, _I, _I, _I, _P, _P, _P, _P], _I),
    "sm_tree_filter": ([_P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _D], _I),
    "sm_tree_filter_f64": ([_P, _P, _I, _I, _I, _P, _P, _P, _P, _D], _I),
    "sm_nl": ([_P, _P, _P, _I, _I, _I], _I),
    "sm_nlca_gradient": ([_P, _P, _I, _I, _P], _I),
    "sm_nlca_cost": ([_P, _P, _P, _I, _I, _I, _D, _D, _D, _P], _I),
    "sm_nlca_flip": ([_P, _P, _I, _I, _I, _P], _I),
    "sm_depth_best_cost": ([_P, _P, _I, _I, _I, _P], _I),
    "sm_nlca_occlusion": ([_P, _P, _P, _I, _I, _P], _I),
    "sm_nlca_refine_cost": ([_P, _P, _P, _I, _I, _I, _P], _I),
    "sm_sgm_path": ([_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "sm_sgm": ([_P, _P, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "sm_sgm_grouped": ([_P, _P, _P, _I, _I, _I, _I, _I, _P], _I),
    "sm_sgm_u16": ([_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P], _I),
    "sm_wta_u16": ([_P, _P, _I, _I, _I, _P], _I),
    "sm_sgm_grouped2": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P], _I),
    "sm_vol_accumulate": ([_P, _P, _P, _Z], _I),
    "sm_wta": ([_P, _P, _I, _I, _I, _P], _I),
    "sm_wta_co": ([_P, _P, _I, _I, _I, _I, _P, _P], _I),
    "sm_select_top_cost": ([_P, _P, _I, _I, _I, _I, _F, _P], _I),
    "sm_subpixel_enhancement": ([_P, _P, _P, _I, _I, _I, _P], _I),
    "sm_disp_from_top": ([_P, _P, _I, _I, _I, _P], _I),
    "sm_disp_from_top2": ([_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "sm_lrc": ([_P, _P, _P, _I, _I, _F], _I),
    "sm_lrc_label": ([_P, _P, _P, _I, _I, _I, _F, _I, _I, _P], _I),
    "sm_lrc_label_lor": ([_P, _P, _P, _I, _I, _I, _F, _I, _I, _I, _P, _P], _I),
    "sm_region_vote": ([_P, _P, _P, _P, _I, _I, _I, _F, _I], _I),
    "sm_proper_ipol": ([_P, _P, _P, _P, _I, _I, _I], _I),
    "sm_wm": ([_P, _P, _P, _P, _P, _I, _I, _I, _P], _I),
    "sm_discontinuity_adjust": ([_P, _P, _P, _I, _I, _I, _P], _I),
    "sm_median3_i16": ([_P, _P, _P, _I, _I], _I),
    "sm_median3_f32": ([_P, _P, _P, _I, _I], _I),
    "sm_cross_scale_1level": ([_P, _P, _Z, _F], _I),
    "sm_pipeline_create": ([_P, _I, _I, C.POINTER(SmParams), C.POINTER(_P)], _I),
    "sm_pipeline_destroy": ([_P], _I),
    "sm_pipeline_upload": ([_P, _P, _P, _P, _P], _I),
    "sm_pipeline_run_device": ([_P], _I),
    "sm_pipeline_download": ([_P, _P, _P], _I),
    "sm_pipeline_run": ([_P, _P, _P, _P, _P, _P, _P], _I),
    "sm_pipeline_buffer": ([_P, _I], _P),
    "sm_pipeline_bind_inputs": ([_P, _P, _P, _P, _P], _I),
    "sm_stream_create": ([C.POINTER(C.c_int), _I, _I, _I, C.POINTER(SmParams), _I, C.POINTER(_P)], _I),
    "sm_stream_submit": ([_P, _P, _P, _P, _P, _P, C.POINTER(_LL)], _I),
    "sm_stream_wait": ([_P, _LL], _I),
    "sm_stream_drain": ([_P], _I),
    "sm_stream_destroy": ([_P], _I),
    "sm_stream_device_count": ([_P], _I),
    "sm_stream_frames_done": ([_P, _I], _LL),
    "sm_stream_launch_count": ([_P], _LL),
    "sm_pipeline_enable_timing": ([_P, _I], _I),
    "sm_pipeline_stage_ms": ([_P, C.POINTER(C.c_float)], _I),
    "sm_pipeline_sgm_split_ms": ([_P, C.POINTER(C.c_float)], _I),
}


def lib():
    """Load libsm_b200.so (built in-tree by __graft_entry__.build()).  Loud failure if absent."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise SmError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU or PyTorch fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (args, res) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes, fn.restype = args, res
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise SmError(f"sm_b200 error {rc}: {lib().sm_last_error().decode(errors='replace')}")


def default_params(max_disp, **over):
    p = SmParams()
    lib().sm_params_default(C.byref(p), max_disp)
    for k, v in over.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def cross_scale_weights(levels, lam=0.3):
    w = (C.c_float * levels)()
    check(lib().sm_cross_scale_weights(levels, lam, w))
    return list(w)


def _ptr(t):
    """Device pointer of a torch CUDA tensor (must be contiguous) or None."""
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise SmError("expected a contiguous CUDA tensor")
    return t.data_ptr()


class Ctx:
    """One sm_ctx on `device`, launching on torch's current stream for that device."""

    def __init__(self, device=0, use_torch_stream=True):
        import torch
        self.torch = torch
        self.device = device
        L = lib()
        if L.sm_device_count() <= 0:
            raise SmError("no CUDA device visible: sm_b200 has no CPU fallback")
        torch.cuda.set_device(device)
        # torch's default stream has handle 0, which the C ABI reads as "create your own stream"; the
        # explicit handle of the legacy default stream is cudaStreamLegacy == 0x1.
        stream = (torch.cuda.current_stream(device).cuda_stream or 1) if use_torch_stream else None
        h = C.c_void_p()
        check(L.sm_ctx_create(C.byref(h), device, stream))
        self.h = h
        self.L = L

    def close(self):
        if getattr(self, "h", None):
            self.L.sm_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.L.sm_ctx_sync(self.h))

    def launches(self):
        return int(self.L.sm_ctx_launch_count(self.h))

    # -- helpers ---------------------------------------------------------------------------
    def dev(self, arr):
        """numpy array -> CUDA tensor on this ctx's device."""
        return self.torch.from_numpy(arr).to(f"cuda:{self.device}")

    def empty(self, shape, dtype):
        return self.torch.empty(shape, dtype=dtype, device=f"cuda:{self.device}")

    # -- stages (device tensors in / out) ---------------------------------------------------
    def bgr2gray(self, bgr):
        H, W, _ = bgr.shape
        out = self.empty((H, W), self.torch.uint8)
        check(self.L.sm_bgr2gray(self.h, _ptr(bgr), H, W, _ptr(out)))
        return out

    def census(self, gray, func=3):
        H, W = gray.shape
        nw = self.L.sm_census_words(func)
        out = self.empty((H, W, nw), self.torch.int64)  # uint64 bit patterns
        check(self.L.sm_census(self.h, _ptr(gray), H, W, func, _ptr(out)))
        return out

    def cost_hamming(self, cL, cR, D, func=3, LOR=0, u16=False):
        H, W, _ = cL.shape
        if u16:
            out = self.empty((H, W, D), self.torch.int16)
            check(self.L.sm_cost_hamming_u16(self.h, _ptr(cL), _ptr(cR), H, W, D, func, LOR, _ptr(out)))
        else:
            out = self.empty((H, W, D), self.torch.float32)
            check(self.L.sm_cost_hamming(self.h, _ptr(cL), _ptr(cR), H, W, D, func, LOR, _ptr(out)))
        return out

    def cost_ad(self, bgrL, bgrR, D, LOR=0, trunc=1000.0):
        H, W, _ = bgrL.shape
        out = self.empty((H, W, D), self.torch.float32)
        check(self.L.sm_cost_ad(self.h, _ptr(bgrL), _ptr(bgrR), H, W, D, LOR, trunc, _ptr(out)))
        return out

    def cost_adcensus(self, bgrL, bgrR, cL, cR, D, func=3, LOR=0, trunc=1000.0, lamAD=10.0, lamCen=30.0):
        H, W, _ = bgrL.shape
        out = self.empty((H, W, D), self.torch.float32)
        check(self.L.sm_cost_adcensus(self.h, _ptr(bgrL), _ptr(bgrR), _ptr(cL), _ptr(cR), H, W, D, func,
                                      trunc, lamAD, lamCen, LOR, _ptr(out)))
        return out

    def combine_exp(self, a, b, l0=10.0, l1=30.0):
        out = self.torch.empty_like(a)
        check(self.L.sm_combine_exp(self.h, _ptr(a), _ptr(b), a.numel(), l0, l1, _ptr(out)))
        return out

    def cumsum_1d(self, vol, area, dv, du):
        H, W, D = vol.shape
        check(self.L.sm_cumsum_1d(self.h, _ptr(vol), _ptr(area), H, W, D, dv, du))
        return vol, area

    def span_1d(self, vol, area, hvl_is, dv, du, direc):
        H, W, D = vol.shape
        tv = self.torch.empty_like(vol)
        ta = self.torch.empty_like(area) if area is not None else None
        check(self.L.sm_span_1d(self.h, _ptr(vol), _ptr(area), _ptr(hvl_is), _ptr(tv), _ptr(ta), H, W, D, dv, du, direc))
        return vol, area

    def div_area(self, vol, area):
        check(self.L.sm_div_area(self.h, _ptr(vol), _ptr(area), vol.numel()))
        return vol

    def update_cost(self, Lr, vm, bgr, v, u, rv, ru, pre_is_inner, thr=15, redu=4):
        H, W, n = vm.shape
        check(self.L.sm_update_cost(self.h, _ptr(Lr), _ptr(vm), _ptr(bgr), H, W, n, v, u, rv, ru, int(pre_is_inner), thr, redu))
        return Lr

    def vol_to_f32(self, vol):
        out = self.empty(tuple(vol.shape), self.torch.float32)
        check(self.L.sm_vol_to_f32(self.h, _ptr(vol), vol.element_size(), vol.numel(), _ptr(out)))
        return out

    def cal_err(self, disp, gt, mask, thres=1):
        """calErr: returns (PBM, RMS, sumNum, errorNum) over mask == 255."""
        H, W = disp.shape
        a, b, e = _LL(0), _LL(0), _D(0.0)
        check(self.L.sm_cal_err(self.h, _ptr(disp), _ptr(gt), _ptr(mask), H, W, thres, C.byref(a), C.byref(b), C.byref(e)))
        n = max(a.value, 1)
        return b.value / n, (e.value / n) ** 0.5, a.value, b.value

    def lrc_mask(self, d1, d2, mask):
        H, W = d1.shape
        check(self.L.sm_lrc_mask(self.h, _ptr(d1), _ptr(d2), H, W, _ptr(mask)))
        return mask

    def pyr_down(self, img):
        H, W = img.shape[:2]
        cn = 1 if img.dim() == 2 else img.shape[2]
        out = self.empty(((H + 1) // 2, (W + 1) // 2) + ((cn,) if img.dim() == 3 else ()), self.torch.uint8)
        check(self.L.sm_pyr_down_u8(self.h, _ptr(img), H, W, cn, _ptr(out)))
        return out

    def cross_scale(self, vols, lam=0.3):
        """SolveAll on a list of per-level device volumes; vols[0] is updated in place and returned."""
        n = len(vols)
        ptrs = (C.c_void_p * n)(*[_ptr(v) for v in vols])
        Hs = (C.c_int * n)(*[v.shape[0] for v in vols])
        Ws = (C.c_int * n)(*[v.shape[1] for v in vols])
        Ds = (C.c_int * n)(*[v.shape[2] for v in vols])
        check(self.L.sm_cross_scale(self.h, ptrs, Hs, Ws, Ds, n, lam))
        return vols[0]

    def grad_xy(self, gray):
        H, W = gray.shape
        gx = self.empty((H, W), self.torch.float32)
        gy = self.empty((H, W), self.torch.float32)
        check(self.L.sm_grad_xy(self.h, _ptr(gray), H, W, _ptr(gx), _ptr(gy)))
        return gx, gy

    def cost_grad(self, gL, gR, arms_view, D, LOR=0, trunc=500.0):
        """gL / gR = (gx, gy) of the left / right gray image; arms_view = HVL[LOR]."""
        H, W = gL[0].shape
        out = self.empty((H, W, D), self.torch.float32)
        check(self.L.sm_cost_grad(self.h, _ptr(gL[0]), _ptr(gL[1]), _ptr(gR[0]), _ptr(gR[1]), _ptr(arms_view), H, W, D,
                                  trunc, LOR, _ptr(out)))
        return out

    def cost_censusgrad(self, cL, cR, gL, gR, arms_view, D, func=3, LOR=0, lamCen=13.0, lamG=1.0, trunc=500.0):
        H, W = gL[0].shape
        out = self.empty((H, W, D), self.torch.float32)
        check(self.L.sm_cost_censusgrad(self.h, _ptr(cL), _ptr(cR), _ptr(gL[0]), _ptr(gL[1]), _ptr(gR[0]), _ptr(gR[1]),
                                        _ptr(arms_view), H, W, D, func, lamCen, lamG, trunc, LOR, _ptr(out)))
        return out

    def arms(self, bgr, L=17, L_out=34, tau=20, tau_out=6, minL=1):
        H, W, _ = bgr.shape
        out = self.empty((H, W, 5), self.torch.int16)  # uint16 bit patterns
        check(self.L.sm_arms(self.h, _ptr(bgr), H, W, L, L_out, tau, tau_out, minL, _ptr(out)))
        return out

    def arms_intersect(self, aL, aR, D, view=0):
        H, W, _ = aL.shape
        out = self.empty((H, W, D, 5), self.torch.int16)
        check(self.L.sm_arms_intersect(self.h, _ptr(aL), _ptr(aR), H, W, D, view, _ptr(out)))
        return out

    def cbca(self, vol, aL, aR, iters=2, view=0):
        """In place on `vol`."""
        H, W, D = vol.shape
        tmp = self.torch.empty_like(vol)
        check(self.L.sm_cbca(self.h, _ptr(vol), _ptr(tmp), _ptr(aL), _ptr(aR), H, W, D, iters, view))
        return vol

    def sgm_path(self, vol, bgr, path, thr=15, redu=4, accumulate_into=None):
        H, W, D = vol.shape
        if accumulate_into is None:
            out = self.torch.empty_like(vol)
            mode = 0
        else:
            out, mode = accumulate_into, 1
        check(self.L.sm_sgm_path(self.h, _ptr(vol), _ptr(bgr), H, W, D, path, thr, redu, mode, _ptr(out)))
        return out

    def sgm(self, vol, bgr, paths=4, thr=15, redu=4):
        H, W, D = vol.shape
        out = self.torch.empty_like(vol)
        check(self.L.sm_sgm(self.h, _ptr(vol), _ptr(bgr), H, W, D, paths, thr, redu, _ptr(out)))
        return out

    def sgm_u16(self, vol, bgr, paths=4, thr=15, redu=4, max_cost=71, want_disp=True):
        """sgm() on a uint16 volume (int16 tensor holding the bit patterns): (sum volume in fixed point = redu x the float
        sum, disparity map or None)."""
        H, W, D = vol.shape
        out = self.torch.empty_like(vol)
        disp = self.empty((H, W), self.torch.int16) if want_disp else None
        check(self.L.sm_sgm_u16(self.h, _ptr(vol), _ptr(bgr), H, W, D, paths, thr, redu, max_cost, _ptr(out), _ptr(disp)))
        return out, disp

    def wta_u16(self, vol):
        H, W, D = vol.shape
        out = self.empty((H, W), self.torch.int16)
        check(self.L.sm_wta_u16(self.h, _ptr(vol), H, W, D, _ptr(out)))
        return out

    def sgm_grouped(self, vol, bgr, thr=15, redu=4):
        H, W, D = vol.shape
        out = self.torch.empty_like(vol)
        check(self.L.sm_sgm_grouped(self.h, _ptr(vol), _ptr(bgr), H, W, D, thr, redu, _ptr(out)))
        return out

    def sgm_grouped2(self, volL, volR, bgrL, bgrR, thr=15, redu=4):
        H, W, D = volL.shape
        outL, outR = self.torch.empty_like(volL), self.torch.empty_like(volR)
        check(self.L.sm_sgm_grouped2(self.h, _ptr(volL), _ptr(volR), _ptr(bgrL), _ptr(bgrR), H, W, D, thr, redu,
                                     _ptr(outL), _ptr(outR)))
        return outL, outR

    def wta(self, vol):
        H, W, D = vol.shape
        out = self.empty((H, W), self.torch.int16)
        check(self.L.sm_wta(self.h, _ptr(vol), H, W, D, _ptr(out)))
        return out

    def subpixel_enhancement(self, disp, vol):
        """subpixelEnhancement (stereoMatching.cpp:6138-6166): float [H][W] from the short map and the view-0 volume."""
        H, W, D = vol.shape
        out = self.empty((H, W), self.torch.float32)
        check(self.L.sm_subpixel_enhancement(self.h, _ptr(disp), _ptr(vol), H, W, D, _ptr(out)))
        return out

    def select_top_cost(self, vol, num, thres):
        """selectTopCostFromVolumn (stereoMatching.h:2405-2461): float [H][W][num+1][2] = {d, cost} per candidate,
        [num][0] = candidate count."""
        H, W, D = vol.shape
        out = self.empty((H, W, num + 1, 2), self.torch.float32)
        check(self.L.sm_select_top_cost(self.h, _ptr(vol), H, W, D, num, thres, _ptr(out)))
        return out

    def disp_from_top(self, top, bgr=None, version=2, method=0, ts=10, has_cir2=True, color_limit=False, init=None):
        """genDispFromTopCostVm (version 1) / genDispFromTopCostVm2 (version 2) on a [H][W][num+1][2] candidate tensor."""
        H, W, n1, _ = top.shape
        out = self.torch.zeros((H, W), dtype=self.torch.int16, device=top.device) if init is None else init.clone()
        if version == 1:
            check(self.L.sm_disp_from_top(self.h, _ptr(top), H, W, n1 - 1, _ptr(out)))
        else:
            check(self.L.sm_disp_from_top2(self.h, _ptr(top), _ptr(bgr), H, W, n1 - 1, method, ts, int(has_cir2),
                                           int(color_limit), _ptr(out)))
        return out

    def wta_co(self, vol, scale=16):
        H, W, D = vol.shape
        d1 = self.empty((H, W), self.torch.int16)
        d2 = self.empty((H, W), self.torch.int16)
        check(self.L.sm_wta_co(self.h, _ptr(vol), H, W, D, scale, _ptr(d1), _ptr(d2)))
        return d1, d2

    def lrc(self, d1, d2, max_diff=0.0):
        H, W = d1.shape
        check(self.L.sm_lrc(self.h, _ptr(d1), _ptr(d2), H, W, max_diff))
        return d1

    def lrc_label(self, d1, d2, D, max_diff=0.0, occ=-32, mis=-48):
        H, W = d1.shape
        mask = self.empty((H, W), self.torch.uint8)
        check(self.L.sm_lrc_label(self.h, _ptr(d1), _ptr(d2), H, W, D, max_diff, occ, mis, _ptr(mask)))
        return d1, mask

    def lrc_label_lor(self, d1, d2, D, LOR, max_diff=0.0, occ=-32, mis=-48):
        """LRConsistencyCheck(D1, D2, errMask, LOR): returns (errMask, errMask1); d1 (LOR 0) or d2 (LOR 1) is labelled in place."""
        H, W = d1.shape
        mask = self.empty((H, W), self.torch.uint8)
        mask1 = self.empty((H, W), self.torch.uint8)
        check(self.L.sm_lrc_label_lor(self.h, _ptr(d1), _ptr(d2), H, W, D, max_diff, occ, mis, LOR, _ptr(mask), _ptr(mask1)))
        return mask, mask1

    def region_vote(self, disp, arms_l, D, ratio=0.4, S=20):
        H, W = disp.shape
        tmp = self.torch.empty_like(disp)
        check(self.L.sm_region_vote(self.h, _ptr(disp), _ptr(tmp), _ptr(arms_l), H, W, D, ratio, S))
        return disp

    def proper_ipol(self, disp, bgr, occ=-32):
        H, W = disp.shape
        tmp = self.torch.empty_like(disp)
        check(self.L.sm_proper_ipol(self.h, _ptr(disp), _ptr(tmp), _ptr(bgr), H, W, occ))
        return disp

    def wm(self, disp, mask, bgr, D):
        """WM (stereoMatching.cpp:7340-7393), in place on `disp`; returns (disp, number of out-of-range window labels)."""
        H, W = disp.shape
        tmp = self.torch.empty_like(disp)
        bad = self.torch.zeros((1,), dtype=self.torch.int32, device=disp.device)
        check(self.L.sm_wm(self.h, _ptr(disp), _ptr(tmp), _ptr(mask), _ptr(bgr), H, W, D, _ptr(bad)))
        return disp, int(bad.item())

    def discontinuity_adjust(self, disp, vol, want_edge=False):
        """discontinuityAdjust (stereoMatching.cpp:6057-6135), in place on `disp`; returns disp or (disp, edge map)."""
        H, W, D = vol.shape
        edge = self.empty((H, W), self.torch.uint8) if want_edge else None
        check(self.L.sm_discontinuity_adjust(self.h, _ptr(disp), _ptr(vol), H, W, D, _ptr(edge) if want_edge else None))
        return (disp, edge) if want_edge else disp

    def median3_f32(self, disp):
        """cv::medianBlur(CV_32F, 3) (stereoMatching.cpp:1490)."""
        H, W = disp.shape
        out = self.empty((H, W), self.torch.float32)
        check(self.L.sm_median3_f32(self.h, _ptr(disp), _ptr(out), H, W))
        return out

    def median3_i16(self, disp):
        H, W = disp.shape
        out = self.torch.empty_like(disp)
        check(self.L.sm_median3_i16(self.h, _ptr(disp), _ptr(out), H, W))
        return out

    def median_u8(self, img, r):
        H, W = img.shape[:2]
        cn = 1 if img.dim() == 2 else img.shape[2]
        out = self.torch.empty_like(img)
        check(self.L.sm_median_u8(self.h, _ptr(img), _ptr(out), H, W, r, cn))
        return out

    def cross_scale_1level(self, vol, lam=0.3):
        check(self.L.sm_cross_scale_1level(self.h, _ptr(vol), vol.numel(), lam))
        return vol

    def mst_build(self, img):
        H, W = img.shape[:2]
        cn = 1 if img.dim() == 2 else img.shape[2]
        N = H * W
        t = self.torch
        parent, rank, order = (self.empty((N,), t.int32) for _ in range(3))
        weight = self.empty((N,), t.uint8)
        check(self.L.sm_mst_build(self.h, _ptr(img), H, W, cn, _ptr(parent), _ptr(weight), _ptr(rank), _ptr(order)))
        return dict(parent=parent, weight=weight, rank=rank, order=order)

    def tree_filter(self, vol, tree, sigma=0.1):
        H, W, D = vol.shape
        work = self.empty((H * W * D,), self.torch.float64)
        check(self.L.sm_tree_filter(self.h, _ptr(vol), _ptr(work), H, W, D, _ptr(tree["parent"]),
                                    _ptr(tree["weight"]), _ptr(tree["rank"]), _ptr(tree["order"]), sigma))
        return vol

    def tree_filter_f64(self, cost, tree, H, W, sigma=0.1):
        """In place on a float64 [H*W, D] tensor."""
        D = cost.shape[-1]
        check(self.L.sm_tree_filter_f64(self.h, _ptr(cost), H, W, D, _ptr(tree["parent"]), _ptr(tree["weight"]),
                                        _ptr(tree["rank"]), _ptr(tree["order"]), sigma))
        return cost

    def vol_accumulate(self, acc, x):
        check(self.L.sm_vol_accumulate(self.h, _ptr(acc), _ptr(x), acc.numel()))
        return acc

    # -- Yang's driver (float64 volumes)
    def nlca_gradient(self, img):
        H, W, _ = img.shape
        out = self.empty((H, W), self.torch.float32)
        check(self.L.sm_nlca_gradient(self.h, _ptr(img), H, W, _ptr(out)))
        return out

    def nlca_cost(self, left, right, D, maxc=7.0, maxg=2.0, wc=0.11):
        H, W, _ = left.shape
        out = self.empty((H, W, D), self.torch.float64)
        check(self.L.sm_nlca_cost(self.h, _ptr(left), _ptr(right), H, W, D, maxc, maxg, wc, _ptr(out)))
        return out

    def nlca_flip(self, vol):
        H, W, D = vol.shape
        out = self.torch.empty_like(vol)
        check(self.L.sm_nlca_flip(self.h, _ptr(vol), H, W, D, _ptr(out)))
        return out

    def depth_best_cost(self, vol):
        H, W, D = vol.shape
        out = self.empty((H, W), self.torch.uint8)
        check(self.L.sm_depth_best_cost(self.h, _ptr(vol), H, W, D, _ptr(out)))
        return out

    def nlca_occlusion(self, dl, dr):
        H, W = dl.shape
        out = self.torch.empty_like(dl)
        check(self.L.sm_nlca_occlusion(self.h, _ptr(dl), _ptr(dr), H, W, _ptr(out)))
        return out

    def nlca_refine_cost(self, disp, mask, D):
        H, W = disp.shape
        out = self.empty((H, W, D), self.torch.float64)
        check(self.L.sm_nlca_refine_cost(self.h, _ptr(disp), _ptr(mask), H, W, D, _ptr(out)))
        return out

    def nlca_disparity(self, left, right, D, sigma=0.1, post=False):
        """qx_nonlocal_cost_aggregation::matching_cost + disparity composed from the C entry points."""
        H, W, _ = left.shape
        cost = self.nlca_cost(left, right, D)
        treeL = self.mst_build(left)
        vol = self.tree_filter_f64(cost.clone().view(H * W, D), treeL, H, W, sigma).view(H, W, D)
        disp = self.median_u8(self.depth_best_cost(vol), 2)
        if not post:
            return disp
        treeR = self.mst_build(right)
        volR = self.tree_filter_f64(self.nlca_flip(cost).view(H * W, D), treeR, H, W, sigma).view(H, W, D)
        dr = self.median_u8(self.depth_best_cost(volR), 2)
        mask = self.nlca_occlusion(disp, dr)
        ref = self.nlca_refine_cost(disp, mask, D)
        ref = self.tree_filter_f64(ref.view(H * W, D), treeL, H, W, sigma / 2).view(H, W, D)
        return self.median_u8(self.depth_best_cost(ref), 2)

    def nl(self, bgrL, vol):
        H, W, D = vol.shape
        check(self.L.sm_nl(self.h, _ptr(bgrL), _ptr(vol), H, W, D))
        return vol


class Pipeline:
    """Frame pipeline (sm_pipeline): host images in, host disparity out."""

    def __init__(self, ctx, H, W, params):
        self.ctx, self.H, self.W, self.params = ctx, H, W, params
        h = C.c_void_p()
        check(ctx.L.sm_pipeline_create(ctx.h, H, W, C.byref(params), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.ctx.L.sm_pipeline_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _np(a):
        return None if a is None else a.ctypes.data

    def upload(self, bgrL, bgrR, grayL=None, grayR=None):
        check(self.ctx.L.sm_pipeline_upload(self.h, self._np(bgrL), self._np(bgrR), self._np(grayL), self._np(grayR)))

    def run_device(self):
        check(self.ctx.L.sm_pipeline_run_device(self.h))

    def download(self, want_right=False):
        import numpy as np
        dl = np.empty((self.H, self.W), np.int16)
        dr = np.empty((self.H, self.W), np.int16) if want_right else None
        check(self.ctx.L.sm_pipeline_download(self.h, dl.ctypes.data, self._np(dr)))
        return (dl, dr) if want_right else dl

    def run(self, bgrL, bgrR, grayL=None, grayR=None, out=None):
        """The call a user makes: one stereo pair in (host), the left disparity map out (host)."""
        import numpy as np
        dl = out if out is not None else np.empty((self.H, self.W), np.int16)
        check(self.ctx.L.sm_pipeline_run(self.h, self._np(bgrL), self._np(bgrR), self._np(grayL), self._np(grayR),
                                         dl.ctypes.data, None))
        return dl

    def buffer(self, which, shape, dtype):
        """Zero-copy torch view of one of the pipeline's device buffers (tests)."""
        import torch
        ptr = self.ctx.L.sm_pipeline_buffer(self.h, which)
        n = 1
        for s in shape:
            n *= s
        itemsize = torch.empty((), dtype=dtype).element_size()

        class _Holder:
            pass
        holder = _Holder()
        holder.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": {torch.float32: "<f4", torch.int16: "<i2", torch.int64: "<i8"}[dtype],
            "data": (ptr, False), "version": 2, "strides": None,
        }
        return torch.as_tensor(holder, device=f"cuda:{self.ctx.device}")

    def enable_timing(self, on=True):
        check(self.ctx.L.sm_pipeline_enable_timing(self.h, 1 if on else 0))

    def stage_ms(self):
        arr = (C.c_float * 8)()
        check(self.ctx.L.sm_pipeline_stage_ms(self.h, arr))
        names = ("census", "cost", "arms", "aggregation", "sgm", "wta", "refine", "total")
        out = dict(zip(names, list(arr)))
        sp = (C.c_float * 2)()
        check(self.ctx.L.sm_pipeline_sgm_split_ms(self.h, sp))
        out["sgm_group"], out["sgm_path"] = float(sp[0]), float(sp[1])
        return out


class Stream:
    """sm_stream: frames submitted in order, frame i on device devices[i % n]; host buffers in, host disparity out."""

    def __init__(self, devices, H, W, params, queue_depth=4):
        self.L = lib()
        if self.L.sm_device_count() <= 0:
            raise SmError("no CUDA device visible: sm_b200 has no CPU fallback")
        self.H, self.W = H, W
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        check(self.L.sm_stream_create(arr, len(devices), H, W, C.byref(params), queue_depth, C.byref(h)))
        self.h = h
        self._keep = {}

    def submit(self, bgrL, bgrR, grayL, grayR, out):
        """numpy host arrays (pinned or not); `out`: int16 [H][W] that receives the left map.  Returns the ticket."""
        t = _LL(0)
        g = (None, None) if grayL is None else (grayL.ctypes.data, grayR.ctypes.data)
        check(self.L.sm_stream_submit(self.h, bgrL.ctypes.data, bgrR.ctypes.data, g[0], g[1], out.ctypes.data, C.byref(t)))
        self._keep[t.value] = (bgrL, bgrR, grayL, grayR, out)      # the buffers must outlive the frame
        return t.value

    def wait(self, ticket):
        check(self.L.sm_stream_wait(self.h, ticket))
        self._keep.pop(ticket, None)

    def drain(self):
        check(self.L.sm_stream_drain(self.h))
        self._keep.clear()

    def frames_done(self, worker):
        return int(self.L.sm_stream_frames_done(self.h, worker))

    def launches(self):
        return int(self.L.sm_stream_launch_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.sm_stream_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
