"""GPU suite: every CUDA stage, called through the C ABI (libsm_b200.so via ctypes), against the CPU oracle on
the same seeded inputs.  Bit-exact for integer/byte/index work; float volumes: see each test's tolerance
(north_star: 1e-4 relative; most stages are in fact bit-exact because the kernels keep the reference's
float operation order)."""
import os

import numpy as np
import pytest
import torch

from mystereomatching_b200 import capi, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

SHAPES = [(37, 53, 19), (75, 113, 24), (64, 160, 64), (40, 150, 130), (33, 70, 300)]


def _pair(H, W, D, kind="texture_warped", seed=11):
    return synth.make_pair(H, W, D, kind, seed)


def _u64(t):
    return t.cpu().numpy().view(np.uint64)


def _u16(t):
    return t.cpu().numpy().view(np.uint16)


def _bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


# ---------------------------------------------------------------- cost
@pytest.mark.parametrize("func", [0, 3])
@pytest.mark.parametrize("shape", [(1, 1), (2, 3), (7, 9), (37, 53), (64, 160), (130, 257)])
def test_census_bit_exact(ctx, func, shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    gray = rng.integers(0, 256, shape, dtype=np.uint8)
    gray[: shape[0] // 2] //= 16   # many ties
    got = _u64(ctx.census(ctx.dev(gray), func))
    assert np.array_equal(got, po.census(gray, func))


def test_bgr2gray_bit_exact(ctx):
    bgr = np.random.default_rng(0).integers(0, 256, (50, 71, 3), dtype=np.uint8)
    assert np.array_equal(ctx.bgr2gray(ctx.dev(bgr)).cpu().numpy(), po.bgr2gray(bgr))


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("LOR", [0, 1])
def test_cost_volumes_bit_exact(ctx, shape, LOR):
    H, W, D = shape
    p = _pair(H, W, D, "random_dot")
    for func in (0, 3):
        cL, cR = po.census(p["grayL"], func), po.census(p["grayR"], func)
        dL, dR = ctx.dev(cL.view(np.int64)), ctx.dev(cR.view(np.int64))
        ham = po.hamming_vol(cL, cR, D, func, LOR)
        assert np.array_equal(ctx.cost_hamming(dL, dR, D, func, LOR).cpu().numpy(), ham)
        got16 = ctx.cost_hamming(dL, dR, D, func, LOR, u16=True).cpu().numpy().view(np.uint16)
        assert np.array_equal(got16, ham.astype(np.uint16))
        bL, bR = ctx.dev(p["bgrL"]), ctx.dev(p["bgrR"])
        ref = po.adcensus_vol(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], D, LOR, func)
        got = ctx.cost_adcensus(bL, bR, dL, dR, D, func, LOR).cpu().numpy()
        assert _bits_equal(got, ref)      # table form: bit-exact, stronger than the 1e-4 the spec asks
    ad = po.ad_vol(p["bgrL"], p["bgrR"], D, LOR)
    assert _bits_equal(ctx.cost_ad(bL, bR, D, LOR).cpu().numpy(), ad)


def test_combine_exp_bit_exact(ctx):
    """gen_vm_from2vm_exp on two materialised volumes: expf as the host libm evaluates it (smd_expf_host) -> identical bits."""
    rng = np.random.default_rng(4)
    a = (rng.random(200000) * 255).astype(np.float32)
    b = rng.integers(0, 72, 200000).astype(np.float32)
    a[:7] = [0.0, 1e-30, 1000.0, 3e4, -2.5, -900.0, 1e38]        # underflow, overflow (negative cost), huge
    got = ctx.combine_exp(ctx.dev(a), ctx.dev(b)).cpu().numpy()
    ref = po.combine_exp(a, b)
    assert _bits_equal(got, ref)
    for l0, l1 in ((13.0, 1.0), (0.7, 250.0)):
        assert _bits_equal(ctx.combine_exp(ctx.dev(a), ctx.dev(b), l0, l1).cpu().numpy(), po.combine_exp(a, b, l0, l1))


# ---------------------------------------------------------------- arms / CBCA
@pytest.mark.parametrize("shape", [(1, 1), (1, 40), (40, 1), (37, 53), (80, 120)])
def test_arms_bit_exact(ctx, shape):
    H, W = shape
    p = _pair(H, W, 8, "texture_warped", seed=3) if min(H, W) > 4 else None
    img = p["bgrL"] if p else np.random.default_rng(5).integers(100, 130, (H, W, 3), dtype=np.uint8)
    got = _u16(ctx.arms(ctx.dev(img)))
    assert np.array_equal(got, po.arms(img))
    got = _u16(ctx.arms(ctx.dev(img), 5, 9, 30, 10, 0))
    assert np.array_equal(got, po.arms(img, 5, 9, 30, 10, 0))


def test_arms_intersect_bit_exact(ctx):
    p = _pair(30, 47, 12)
    aL, aR = po.arms(p["bgrL"]), po.arms(p["bgrR"])
    for view in (0, 1):
        got = _u16(ctx.arms_intersect(ctx.dev(aL.view(np.int16)), ctx.dev(aR.view(np.int16)), 12, view))
        assert np.array_equal(got, po.arms_intersect(aL, aR, 12, view))


@pytest.mark.parametrize("shape", SHAPES + [(48, 96, 40), (100, 132, 96), (44, 200, 36), (120, 64, 256)])   # last 4: wide-staging path
@pytest.mark.parametrize("view", [0, 1])
def test_cbca_bit_exact(ctx, shape, view):
    H, W, D = shape
    p = _pair(H, W, D)
    aL, aR = po.arms(p["bgrL"]), po.arms(p["bgrR"])
    vol = po.adcensus_vol(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], D, view)
    for iters in (1, 2, 3):
        ref = po.cbca(vol, aL, aR, iters, view)
        got = ctx.cbca(ctx.dev(vol.copy()), ctx.dev(aL.view(np.int16)), ctx.dev(aR.view(np.int16)), iters,
                       view).cpu().numpy()
        # the kernel keeps the reference's sequential prefix-sum order -> identical floats
        assert _bits_equal(got, ref), (iters, float(np.abs(got - ref).max()))


# ---------------------------------------------------------------- SGM
@pytest.mark.parametrize("shape", SHAPES + [(1, 30, 16), (30, 1, 16), (3, 3, 1)])
def test_sgm_paths_bit_exact_float(ctx, shape):
    H, W, D = shape
    rng = np.random.default_rng(H * W + D)
    bgr = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    bgr[:, : W // 2] //= 32         # flat half: P1/P2 not reduced there
    vol = (rng.random((H, W, D)) * 2).astype(np.float32)
    dv, db = ctx.dev(vol), ctx.dev(bgr)
    for path in range(8):
        ref = po.sgm_path(vol, bgr, path)
        got = ctx.sgm_path(dv, db, path).cpu().numpy()
        assert _bits_equal(got, ref), path
    for P in (4, 8):
        assert _bits_equal(ctx.sgm(dv, db, P).cpu().numpy(), po.sgm(vol, bgr, P)), P


def test_sgm_integer_cost_bit_exact(ctx):
    """Integer (Hamming) costs in a float volume: the 'integer-cost SGM' of costScan."""
    H, W, D = 50, 90, 40
    p = _pair(H, W, D, "random_dot")
    cL, cR = po.census(p["grayL"]), po.census(p["grayR"])
    vol = po.hamming_vol(cL, cR, D)
    ref = po.sgm(vol, p["bgrL"], 8)
    got = ctx.sgm(ctx.dev(vol), ctx.dev(p["bgrL"]), 8).cpu().numpy()
    assert np.array_equal(got, ref)


def test_sgm_u8_u16_cost_entries_bit_exact(ctx):
    """costScan's CV_8U / CV_16U entries (updateCost<uchar> / <ushort>): the u16 Hamming volume written by
    sm_cost_hamming_u16, converted on the device, gives exactly the float volume's SGM."""
    H, W, D = 40, 70, 24
    p = _pair(H, W, D, "random_dot")
    cL, cR = po.census(p["grayL"]), po.census(p["grayR"])
    dL, dR = ctx.dev(cL.view(np.int64)), ctx.dev(cR.view(np.int64))
    v16 = ctx.cost_hamming(dL, dR, D, 3, 0, u16=True)
    vf = ctx.vol_to_f32(v16)
    ham = po.hamming_vol(cL, cR, D)
    assert np.array_equal(vf.cpu().numpy(), ham)
    v8 = ctx.dev(ham.astype(np.uint8))
    assert np.array_equal(ctx.vol_to_f32(v8).cpu().numpy(), ham)
    assert np.array_equal(ctx.sgm(vf, ctx.dev(p["bgrL"]), 8).cpu().numpy(), po.sgm(ham, p["bgrL"], 8))


# ---------------------------------------------------------------- WTA / refine
@pytest.mark.parametrize("shape", SHAPES)
def test_wta_bit_exact(ctx, shape):
    H, W, D = shape
    rng = np.random.default_rng(9)
    vol = rng.integers(0, 6, (H, W, D)).astype(np.float32)    # many exact ties
    vol[0, 0, :] = np.finfo(np.float32).max                   # nothing beats FLT_MAX -> -1
    assert np.array_equal(ctx.wta(ctx.dev(vol)).cpu().numpy(), po.wta(vol))
    d1, d2 = ctx.wta_co(ctx.dev(vol))
    r1, r2 = po.wta_co(vol)
    assert np.array_equal(d1.cpu().numpy(), r1) and np.array_equal(d2.cpu().numpy(), r2)


def _noisy_disp(H, W, D, seed):
    rng = np.random.default_rng(seed)
    p = _pair(H, W, D, seed=seed)
    d = p["gt"].astype(np.int16)
    d[rng.random((H, W)) < 0.15] = -1
    d[rng.random((H, W)) < 0.03] = -32
    return p, d


def test_lrc_bit_exact(ctx):
    rng = np.random.default_rng(12)
    H, W, D = 45, 77, 20
    d1 = rng.integers(-1, D, (H, W)).astype(np.int16)
    d2 = rng.integers(0, D, (H, W)).astype(np.int16)
    for md in (0.0, 1.0):
        assert np.array_equal(ctx.lrc(ctx.dev(d1.copy()), ctx.dev(d2), md).cpu().numpy(), po.lrc_normal(d1, d2, md))
    g, m = ctx.lrc_label(ctx.dev(d1.copy()), ctx.dev(d2), D)
    r, rm = po.lrc_label(d1, d2, D)
    assert np.array_equal(g.cpu().numpy(), r) and np.array_equal(m.cpu().numpy(), rm)


def test_region_vote_and_ipol_and_median_bit_exact(ctx):
    H, W, D = 70, 110, 32
    p, d = _noisy_disp(H, W, D, 21)
    arms = po.arms(p["bgrL"])
    ref = po.region_vote(d, arms, D)
    got = ctx.region_vote(ctx.dev(d.copy()), ctx.dev(arms.view(np.int16)), D).cpu().numpy()
    assert np.array_equal(got, ref)
    assert (ref != d).any()    # the test input must actually exercise the vote
    for ratio, S in ((0.0, 5), (1.0, 5), (1.5, 5), (0.4, 0)):   # histogram path (mode always wins / never) and the unanimity path
        ref_r = po.region_vote(d, arms, D, ratio, S)
        got_r = ctx.region_vote(ctx.dev(d.copy()), ctx.dev(arms.view(np.int16)), D, ratio, S).cpu().numpy()
        assert np.array_equal(got_r, ref_r), (ratio, S)
    ref = po.proper_ipol(d, p["bgrL"])
    got = ctx.proper_ipol(ctx.dev(d.copy()), ctx.dev(p["bgrL"])).cpu().numpy()
    assert np.array_equal(got, ref)
    assert np.array_equal(ctx.median3_i16(ctx.dev(d)).cpu().numpy(), po.median3_i16(d))


@pytest.mark.parametrize("r,cn", [(1, 3), (1, 1), (2, 1), (3, 1)])
def test_median_u8_bit_exact(ctx, r, cn):
    rng = np.random.default_rng(r * 10 + cn)
    shape = (37, 59, cn) if cn > 1 else (37, 59)
    img = rng.integers(0, 64, shape, dtype=np.uint8)
    assert np.array_equal(ctx.median_u8(ctx.dev(img), r).cpu().numpy(), po.ctmf(img, r))


def test_cross_scale_bit_exact(ctx):
    v = np.random.default_rng(1).random(10000).astype(np.float32)
    ref = v.copy()
    po.lib().orc_solve_all_1level(ref, ref.size, 0.3)
    assert _bits_equal(ctx.cross_scale_1level(ctx.dev(v.copy()), 0.3).cpu().numpy(), ref)


# ---------------------------------------------------------------- whole pipeline
@pytest.mark.parametrize("cfg", [
    dict(H=90, W=140, D=32, kind="random_dot", paths=4),
    dict(H=96, W=150, D=48, kind="texture_warped", paths=8),
    dict(H=60, W=100, D=130, kind="texture_warped", paths=8),
])
def test_pipeline_matches_oracle(ctx, cfg):
    H, W, D = cfg["H"], cfg["W"], cfg["D"]
    p = _pair(H, W, D, cfg["kind"], seed=31)
    params = capi.default_params(D - 1, sgm_paths=cfg["paths"])
    pl = capi.Pipeline(ctx, H, W, params)
    dl, dr = None, None
    pl.upload(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"])
    pl.run_device()
    dl, dr = pl.download(want_right=True)
    vol = pl.buffer(0, (H, W, D), torch.float32).cpu().numpy()
    op = po.default_params(D, paths=cfg["paths"])
    rl, rr, rvol, _ = po.pipeline(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], op, want_vol=True)
    assert np.all(np.abs(vol - rvol) <= 1e-4 * np.abs(rvol))        # north_star: float volumes within 1e-4 rel
    assert (dl == rl).mean() >= 0.995                               # north_star: >= 99.5 % identical pixels
    b_gpu = synth.bad_k(dl, p["gt"], p["nonocc"], 2)
    b_cpu = synth.bad_k(rl, p["gt"], p["nonocc"], 2)
    assert abs(b_gpu - b_cpu) <= 0.1                                # bad-2 within 0.1 pp
    # the one-call host API gives the same map; gray computed on the device when not supplied
    again = pl.run(p["bgrL"], p["bgrR"])
    assert np.array_equal(again, dl)
    pl.close()


@pytest.mark.parametrize("paths", [4, 8])
def test_pipeline_one_level_solve_all_folded_into_cbca_is_bit_exact(ctx, paths):
    """main() as the reference compiles it (PY_LEV = 1, REG_LAMBDA = 0.3): the scale 1/(1+lambda) rides in the last
    CBCA pass; the volume after SGM (reference path order) and the map must equal the oracle's bit for bit."""
    H, W, D = 70, 120, 40
    p = _pair(H, W, D, "texture_warped", seed=41)
    params = capi.default_params(D - 1, sgm_paths=paths, sgm_grouped=0, crossScaleLambda=0.3, pyramidLevels=1)
    pl = capi.Pipeline(ctx, H, W, params)
    pl.upload(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"])
    pl.run_device()
    dl, _ = pl.download(want_right=True)
    vol = pl.buffer(0, (H, W, D), torch.float32).cpu().numpy()
    pl.close()
    op = po.default_params(D, paths=paths, pyr_levels=1, cross_lambda=0.3)
    rl, _, rvol, _ = po.pipeline(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], op, want_vol=True)
    assert _bits_equal(vol, rvol)
    assert np.array_equal(dl, rl)


def test_cal_err_matches_oracle(ctx):
    """calErr<short>: the two counts are exact, the RMS agrees to float accumulation accuracy."""
    H, W, D = 90, 140, 32
    p, d = _noisy_disp(H, W, D, 33)
    d[5:9, 7:30] += 3                                   # some errors above the threshold
    gt = p["gt"].astype(np.float32) + 0.25
    for mask_key, thres in (("nonocc", 1), ("all", 1), ("nonocc", 2)):
        mask = (p[mask_key].astype(np.uint8) * 255) if p[mask_key].dtype == np.bool_ else p[mask_key].astype(np.uint8)
        if mask.max() == 1:
            mask = mask * 255
        ref = po.cal_err(d, gt, mask, thres)
        got = ctx.cal_err(ctx.dev(d), ctx.dev(gt), ctx.dev(mask), thres)
        assert got[2] == ref[2] and got[3] == ref[3] and ref[2] > 0 and 0 < ref[3] < ref[2]
        assert abs(got[0] - ref[0]) < 1e-6 and abs(got[1] - ref[1]) <= 1e-5 * ref[1]


def test_pipeline_region_of_validity_errors(ctx):
    with pytest.raises(capi.SmError):
        capi.Pipeline(ctx, 10, 10, capi.default_params(600))       # D > 512 (CV_CN_MAX)
    with pytest.raises(capi.SmError):
        ctx.sgm_path(ctx.dev(np.zeros((2, 2, 2), np.float32)), ctx.dev(np.zeros((2, 2, 3), np.uint8)), 9)


def test_subpixel_and_float_median_argument_errors(ctx):
    L = capi.lib()
    m = ctx.dev(np.zeros((4, 5), np.float32))
    d = ctx.dev(np.zeros((4, 5), np.int16))
    v = ctx.dev(np.zeros((4, 5, 3), np.float32))
    p = capi._ptr
    assert L.sm_median3_f32(ctx.h, p(m), p(m), 4, 5) != 0            # in place is not supported (reads neighbours)
    assert L.sm_median3_f32(ctx.h, p(m), None, 4, 5) != 0
    assert L.sm_subpixel_enhancement(ctx.h, p(d), None, 4, 5, 3, p(m)) != 0
    assert L.sm_subpixel_enhancement(ctx.h, p(d), p(v), 4, 0, 3, p(m)) != 0
    assert L.sm_subpixel_enhancement(ctx.h, p(d), p(v), 4, 5, 3, p(m)) == 0
    assert b"" != L.sm_last_error()


# ---------------------------------------------------------------- NL: MST + tree filter
def _nl_images():
    rng = np.random.default_rng(5)
    noisy = rng.integers(0, 256, (45, 61, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:40, 0:52]
    smooth = np.stack([(xx * 3 + yy) // 4, (xx + yy * 2) // 5, (xx * yy) // 40], -1).astype(np.uint8)
    flat = np.full((17, 23, 3), 9, np.uint8)                       # every weight 0: pure index tie-breaking
    tex = _pair(70, 96, 8, "texture_warped", seed=2)["bgrL"]
    return {"noisy": noisy, "smooth": smooth, "flat": flat, "tex": tex, "row": noisy[:1].copy(),
            "col": noisy[:, :1].copy(), "one": noisy[:1, :1].copy()}


@pytest.mark.timeout(120)
@pytest.mark.parametrize("name", ["noisy", "smooth", "flat", "tex", "row", "col", "one"])
def test_mst_identical_to_reference_tree(ctx, name):
    img = _nl_images()[name]
    H, W = img.shape[:2]
    t = ctx.mst_build(ctx.dev(img))
    ref = po.mst(img)
    parent, weight, rank, order = (t[k].cpu().numpy() for k in ("parent", "weight", "rank", "order"))
    assert np.array_equal(parent[1:], ref["parent"][1:]) and parent[0] == 0     # edge-for-edge the same tree
    assert np.array_equal(weight[1:], ref["weight"][1:])
    assert np.array_equal(rank, ref["rank"])
    assert np.array_equal(np.sort(order), np.arange(H * W))                     # a permutation ...
    assert np.all(np.diff(rank[order]) >= 0)                                    # ... grouped by depth
    assert np.array_equal(ctx.median_u8(ctx.dev(img), 1).cpu().numpy(), po.ctmf(img, 1))


@pytest.mark.timeout(300)
@pytest.mark.parametrize("shape", [(1080, 1920), (480, 640), (9, 2049)])
def test_mst_rooting_large_frames(ctx, shape):
    """The rooting's own scan / radix sort across many tiles (1080p: 2 M nodes, a 4 M-edge tour = 2025 scan tiles, more
    than one round of the tile-sum kernel; depths need three 8-bit passes): same tree, depth and a level-grouped order."""
    H, W = shape
    rng = np.random.default_rng(H + W)
    yy, xx = np.mgrid[0:H, 0:W]
    base = ((xx // 7 + yy // 5) % 256).astype(np.uint8)
    img = np.stack([base, base // 2, rng.integers(0, 4, (H, W), dtype=np.uint8)], -1).astype(np.uint8)
    t = ctx.mst_build(ctx.dev(img))
    ref = po.mst(img)
    parent, weight, rank, order = (t[k].cpu().numpy() for k in ("parent", "weight", "rank", "order"))
    assert np.array_equal(parent[1:], ref["parent"][1:]) and parent[0] == 0
    assert np.array_equal(weight[1:], ref["weight"][1:])
    assert np.array_equal(rank, ref["rank"])
    assert np.array_equal(np.sort(order), np.arange(H * W))
    assert np.all(np.diff(rank[order]) >= 0)
    assert order[0] == 0


@pytest.mark.timeout(120)
@pytest.mark.parametrize("name,D", [("noisy", 6), ("smooth", 5), ("flat", 3), ("tex", 33)])
def test_tree_filter_bit_exact(ctx, name, D):
    img = _nl_images()[name]
    H, W = img.shape[:2]
    vol = np.random.default_rng(3).random((H, W, D)).astype(np.float32)
    tree = ctx.mst_build(ctx.dev(img))
    for sigma in (0.1, 0.05):
        got = ctx.tree_filter(ctx.dev(vol.copy()), tree, sigma).cpu().numpy()
        t = po.mst(img)
        table = np.empty(256, np.float64)
        po.lib().orc_tree_table(sigma, table)
        cost = vol.astype(np.float64).copy()
        tmp = np.empty_like(cost)
        po.lib().orc_tree_filter(cost, tmp, H * W, D, t["parent"], t["weight"], t["nr_child"], t["children"],
                                 t["order"], table)
        assert _bits_equal(got, cost.astype(np.float32))     # same f64 operation order -> identical after the cast


@pytest.mark.timeout(180)
@pytest.mark.parametrize("shape", [(200, 300), (201, 301), (96, 1024)])
def test_tree_filter_bit_exact_many_chunks(ctx, shape):
    """Enough nodes for dozens of 1024-position chunks (the streamed filter's ring wraps many times, levels straddle
    chunk boundaries); odd N takes the general kernel, even N the lean level loop -- both must equal the oracle."""
    H, W = shape
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    img[: H // 2] = (img[: H // 2] // 64) * 64          # a flatter half: wider levels
    D = 3
    vol = rng.random((H, W, D)).astype(np.float32) - 0.25   # negative costs too (the -0.0 identity must hold)
    tree = ctx.mst_build(ctx.dev(img))
    got = ctx.tree_filter(ctx.dev(vol.copy()), tree, 0.1).cpu().numpy()
    t = po.mst(img)
    table = np.empty(256, np.float64)
    po.lib().orc_tree_table(0.1, table)
    cost = vol.astype(np.float64).copy()
    tmp = np.empty_like(cost)
    po.lib().orc_tree_filter(cost, tmp, H * W, D, t["parent"], t["weight"], t["nr_child"], t["children"], t["order"], table)
    assert _bits_equal(got, cost.astype(np.float32))


@pytest.mark.timeout(120)
def test_tree_filter_matches_reference_golden(ctx, golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "nl_ref.npz"))       # outputs of the reference's own qx_tree_filter
    img, vol = g["mst_noisy_img"], g["tf_vol"]
    tree = ctx.mst_build(ctx.dev(img))
    for key, sigma in (("tf_out_sigma0p1", 0.1), ("tf_out_sigma0p05", 0.05)):
        got = ctx.tree_filter(ctx.dev(vol.copy()), tree, sigma).cpu().numpy()
        assert _bits_equal(got, g[key].astype(np.float32))
    for k in ("parent", "weight", "rank"):
        assert np.array_equal(tree[k].cpu().numpy()[1:], g["mst_noisy_" + k][1:])


@pytest.mark.timeout(180)
@pytest.mark.parametrize("shape", [(48, 64, 16), (60, 90, 33)])
def test_nl_aggregation_matches_oracle(ctx, shape):
    H, W, D = shape
    p = _pair(H, W, D, "texture_warped", seed=8)
    vol = po.adcensus_vol(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], D, 0)
    ref, rdisp = po.nl(p["bgrL"], vol)
    got = ctx.nl(ctx.dev(p["bgrL"]), ctx.dev(vol.copy()))
    assert _bits_equal(got.cpu().numpy(), ref)
    assert np.array_equal(ctx.wta(got).cpu().numpy(), rdisp)


@pytest.mark.timeout(180)
def test_pipeline_nl_matches_oracle_composition(ctx):
    """aggregation = "NL": StereoMatching::NL aggregates vm[0] only; SGM/WTA/LRC/refine follow as usual."""
    H, W, D = 64, 96, 32
    p = _pair(H, W, D, "texture_warped", seed=13)
    params = capi.default_params(D - 1, sgm_paths=4, aggregation=2)
    pl = capi.Pipeline(ctx, H, W, params)
    got = pl.run(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"])
    pl.close()
    whole, _, _, _ = po.pipeline(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"],
                                 po.default_params(D, paths=4, aggregation=2))
    v0 = po.adcensus_vol(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], D, 0)
    v1 = po.adcensus_vol(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], D, 1)
    v0, _ = po.nl(p["bgrL"], v0)
    d0 = po.wta(po.sgm(v0, p["bgrL"], 4))
    d1 = po.wta(po.sgm(v1, p["bgrR"], 4))
    d = po.lrc_normal(d0, d1)
    arms = po.arms(p["bgrL"])
    for _ in range(2):
        d = po.region_vote(d, arms, D)
    for _ in range(2):
        d = po.proper_ipol(d, p["bgrL"])
    d = po.median3_i16(d)
    assert np.array_equal(whole, d)            # the oracle's own NL chain is this composition
    assert (got == d).mean() >= 0.995


@pytest.mark.timeout(180)
def test_pipeline_nl_stream_of_frames_replays_the_mst_graph(ctx):
    """A stream of NL frames through one pipeline: the tree build runs on the pipeline's side stream, where the fixed
    launch chains of the MST (Boruvka rounds, Euler-tour rooting) are captured into CUDA graphs on the second frame and
    replayed from the third on (smi_graphed).  Every frame must equal what a fresh pipeline (eager launches) gives for
    it, different images included, and the launch counter must advance as it does without graphs."""
    H, W, D = 72, 104, 32
    params = capi.default_params(D - 1, sgm_paths=4, aggregation=2)
    pairs = [_pair(H, W, D, "texture_warped", seed=40 + i) for i in range(3)]
    fresh = []
    for p in pairs:
        pl = capi.Pipeline(ctx, H, W, params)
        fresh.append(pl.run(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"]).copy())
        pl.close()
    pl = capi.Pipeline(ctx, H, W, params)
    per_frame = []
    for k in range(7):                       # eager, capture, then replays; the images change under the graph
        p = pairs[k % 3]
        l0 = ctx.launches()
        got = pl.run(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"])
        per_frame.append(ctx.launches() - l0)
        assert np.array_equal(got, fresh[k % 3]), f"frame {k}"
    pl.close()
    # a replayed frame counts the launches the same frame made eagerly (the number of Boruvka rounds depends on the image)
    assert all(per_frame[k] == per_frame[k - 3] for k in range(3, 7)), per_frame


# ---------------------------------------------------------------- Yang's driver (qx_nonlocal_cost_aggregation)
@pytest.mark.timeout(180)
def test_nlca_stages_bit_exact(ctx, golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "nlca_ref.npz"))      # outputs of the reference's own compiled class
    L, R, D = g["left"], g["right"], int(g["D"])
    dL, dR = ctx.dev(L), ctx.dev(R)
    assert np.array_equal(ctx.nlca_gradient(dL).cpu().numpy(), g["grad_left"])
    cost = ctx.nlca_cost(dL, dR, D)
    assert np.array_equal(cost.cpu().numpy(), g["cost"])
    assert np.array_equal(ctx.nlca_flip(cost).cpu().numpy(), g["cost_right"])
    assert np.array_equal(ctx.nlca_disparity(dL, dR, D).cpu().numpy(), g["disp"])
    assert np.array_equal(ctx.nlca_disparity(dL, dR, D, post=True).cpu().numpy(), g["disp_post"])
    # helpers against the oracle on random data
    rng = np.random.default_rng(6)
    vol = rng.integers(0, 5, (33, 47, 9)).astype(np.float64)
    assert np.array_equal(ctx.depth_best_cost(ctx.dev(vol)).cpu().numpy(), po.depth_best_cost(vol))
    assert np.array_equal(ctx.nlca_flip(ctx.dev(vol)).cpu().numpy(), po.flip_vol(vol))
    dl = rng.integers(0, 9, (33, 47)).astype(np.uint8)
    dr = rng.integers(0, 9, (33, 47)).astype(np.uint8)
    m = po.detect_occlusion(dl, dr)
    assert np.array_equal(ctx.nlca_occlusion(ctx.dev(dl), ctx.dev(dr)).cpu().numpy(), m)
    ref = np.where(m[..., None] != 0, 0.0, np.abs(dl[..., None].astype(np.int64) - np.arange(9)).astype(np.float64))
    assert np.array_equal(ctx.nlca_refine_cost(ctx.dev(dl), ctx.dev(m), 9).cpu().numpy(), ref)


# ---------------------------------------------------------------- grouped SGM sweeps (3 paths per pass over the volume)
@pytest.mark.timeout(120)
@pytest.mark.parametrize("shape", [(40, 64, 128), (37, 53, 132), (50, 200, 256), (64, 96, 68), (9, 8, 72), (120, 330, 96)])
def test_sgm_grouped_matches_reference_sum(ctx, shape):
    """Every path volume is exact; only the order of the eight additions differs from gen_sgm_vm's."""
    H, W, D = shape
    rng = np.random.default_rng(H + W + D)
    bgr = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    bgr[:, : W // 2] //= 32
    vol = (rng.random((H, W, D)) * 2).astype(np.float32)
    ref = po.sgm(vol, bgr, 8)
    got = ctx.sgm_grouped(ctx.dev(vol), ctx.dev(bgr)).cpu().numpy()
    rel = np.abs(got - ref) / np.abs(ref)
    assert rel.max() <= 1e-6, float(rel.max())          # north_star: float SGM volumes within 1e-4 relative
    # the same eight path volumes added in the grouped order reproduce the result bit for bit
    L = [po.sgm_path(vol, bgr, p) for p in range(8)]
    grouped = ((L[0] + L[4]) + L[5])
    grouped = ((grouped + L[1]) + L[6]) + L[7]
    grouped = (grouped + L[2]) + L[3]
    if D % 4 == 0 and D > 64:
        assert _bits_equal(got, grouped)
    # integer-valued costs: exact in any order
    ivol = rng.integers(0, 72, (H, W, D)).astype(np.float32)
    assert np.array_equal(ctx.sgm_grouped(ctx.dev(ivol), ctx.dev(bgr)).cpu().numpy(), po.sgm(ivol, bgr, 8))


@pytest.mark.timeout(120)
@pytest.mark.parametrize("shape", [(40, 64, 128), (37, 53, 132), (50, 200, 256), (64, 96, 68), (33, 600, 96)])
def test_sgm_grouped_two_views_in_one_launch_equals_one_view_at_a_time(ctx, shape):
    H, W, D = shape
    rng = np.random.default_rng(H * W + D)
    bgr = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(2)]
    vol = [(rng.random((H, W, D)) * 2).astype(np.float32) for _ in range(2)]
    one = [ctx.sgm_grouped(ctx.dev(vol[i]), ctx.dev(bgr[i])).cpu().numpy() for i in range(2)]
    gl, gr = ctx.sgm_grouped2(ctx.dev(vol[0]), ctx.dev(vol[1]), ctx.dev(bgr[0]), ctx.dev(bgr[1]))
    assert _bits_equal(gl.cpu().numpy(), one[0]) and _bits_equal(gr.cpu().numpy(), one[1])


@pytest.mark.timeout(180)
def test_pipeline_reference_order_is_bit_exact_and_grouped_agrees(ctx):
    H, W, D = 80, 144, 96
    p = _pair(H, W, D, "texture_warped", seed=41)
    op = po.default_params(D, paths=8)
    rl, _, rvol, _ = po.pipeline(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], op, want_vol=True)
    out = {}
    for grouped in (0, 1):
        pl = capi.Pipeline(ctx, H, W, capi.default_params(D - 1, sgm_paths=8, sgm_grouped=grouped))
        pl.upload(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"])
        pl.run_device()
        out[grouped] = (pl.download(), pl.buffer(0, (H, W, D), torch.float32).cpu().numpy())
        pl.close()
    assert _bits_equal(out[0][1], rvol) and np.array_equal(out[0][0], rl)      # reference order: identical
    assert np.all(np.abs(out[1][1] - rvol) <= 1e-6 * np.abs(rvol))
    assert (out[1][0] == rl).mean() >= 0.995


# ---------------------------------------------------------------- BASELINE's full size, through size-independent properties
@pytest.mark.timeout(300)
def test_full_size_grouped_sweeps_equal_sum_of_single_path_kernels(ctx):
    """1920x1080 D=256 (all 148 CTAs, uneven column split): the grouped sweeps must reproduce, bit for bit, the eight
    single-path volumes (each bit-exact against the oracle at small sizes) added in the grouped order."""
    H, W, D = 1080, 1920, 256
    p = _pair(H, W, D, "texture_warped", seed=5)
    bgr = ctx.dev(p["bgrL"])
    g = torch.Generator(device="cuda").manual_seed(11)
    vol = torch.rand((H, W, D), device="cuda", generator=g) * 2
    got = ctx.sgm_grouped(vol, bgr)
    acc = ctx.sgm_path(vol, bgr, 0)
    for k in (4, 5, 1, 6, 7, 2, 3):
        acc = acc + ctx.sgm_path(vol, bgr, k)       # IEEE float add, left to right: the grouped order
    assert torch.equal(got.view(torch.int32), acc.view(torch.int32))
    # and a second call gives the same bits (edge hand-off buffers are re-armed per launch)
    assert torch.equal(ctx.sgm_grouped(vol, bgr).view(torch.int32), got.view(torch.int32))
    # both views in one launch (two CTAs per SM): the same bits for each
    bgr2 = ctx.dev(p["bgrR"])
    vol2 = torch.rand((H, W, D), device="cuda", generator=g) * 2
    gl, gr = ctx.sgm_grouped2(vol, vol2, bgr, bgr2)
    assert torch.equal(gl.view(torch.int32), got.view(torch.int32))
    assert torch.equal(gr.view(torch.int32), ctx.sgm_grouped(vol2, bgr2).view(torch.int32))


@pytest.mark.timeout(300)
def test_full_size_pipeline_is_deterministic_and_order_insensitive(ctx):
    """1920x1080 D=256, 8 paths: same frame twice -> identical maps; reference path order vs grouped sweeps agree on
    >= 99.5 % of the pixels; every disparity lies in [DISP_OCC.., D)."""
    H, W, D = 1080, 1920, 256
    p = _pair(H, W, D, "texture_warped", seed=6)
    maps = {}
    for grouped in (1, 0):
        pl = capi.Pipeline(ctx, H, W, capi.default_params(D - 1, sgm_paths=8, sgm_grouped=grouped))
        a = pl.run(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"]).copy()
        b = pl.run(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"]).copy()
        pl.close()
        assert np.array_equal(a, b)
        maps[grouped] = a
    assert (maps[0] == maps[1]).mean() >= 0.995
    assert maps[1].max() < D and maps[1].min() >= -48        # DISP_MIS, the lowest label refine can leave
    assert abs(synth.bad_k(maps[1], p["gt"], p["nonocc"], 2) - synth.bad_k(maps[0], p["gt"], p["nonocc"], 2)) <= 0.1


# ---------------------------------------------------------------- candidate disparities (first half)
@pytest.mark.timeout(120)
def test_select_top_cost_matches_reference_golden(ctx, golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "top_ref.npz"))    # outputs of the reference's own selectTopCostFromVolumn
    for k in g.files:
        if "_top_" not in k:
            continue
        vol = g[k.split("_top_")[0] + "_vol"]
        num, thres = int(k.split("_n")[1].split("_")[0]), float(k.split("_t")[-1])
        dv = ctx.dev(vol.copy())
        got = ctx.select_top_cost(dv, num, thres).cpu().numpy()
        assert _bits_equal(got, g[k]), k
        assert _bits_equal(dv.cpu().numpy(), vol)             # the volume itself is left alone (the reference clones it)


# ---------------------------------------------------------------- sub-pixel step (default-off refiner)
@pytest.mark.timeout(120)
def test_subpixel_enhancement_matches_reference_golden(ctx, golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "subpixel_ref.npz"))    # outputs of the reference's own subpixelEnhancement
    n = 0
    for k in g.files:
        if "_se_" not in k:
            continue
        pre, tag = k.split("_se_")
        got = ctx.subpixel_enhancement(ctx.dev(g[f"{pre}_disp_{tag}"].copy()), ctx.dev(g[pre + "_vol"].copy())).cpu().numpy()
        assert _bits_equal(got, g[k]), k
        n += 1
    assert n == 3


@pytest.mark.timeout(120)
def test_median3_f32_matches_cv2_golden_and_oracle(ctx, golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "subpixel_ref.npz"))    # cv2.medianBlur(CV_32F, 3) outputs
    pairs = [(k.replace("_semed_", "_se_"), k) for k in g.files if "_semed_" in k]
    pairs += [(k[:-4] + "_map", k) for k in g.files if k.endswith("_med")]
    assert len(pairs) == 8
    for src, want in pairs:
        assert _bits_equal(ctx.median3_f32(ctx.dev(g[src].copy())).cpu().numpy(), g[want]), want
    m = (np.random.default_rng(3).random((270, 481)) * 64).astype(np.float32)
    m[::2] = np.floor(m[::2])
    assert _bits_equal(ctx.median3_f32(ctx.dev(m)).cpu().numpy(), po.median3_f32(m))


@pytest.mark.timeout(180)
@pytest.mark.parametrize("shape", [(37, 53, 33), (24, 40, 256), (1, 1, 3), (9, 11, 2), (270, 480, 64)])
def test_subpixel_enhancement_matches_oracle(ctx, shape):
    H, W, D = shape
    rng = np.random.default_rng(23)
    vol = rng.random((H, W, D)).astype(np.float32) * 50
    vol[::3] = np.round(vol[::3])                             # flat parabolas: denom == 0
    vol[::4, ::2] = np.finfo(np.float32).max                  # out-of-range padding values: inf - inf -> NaN, unchanged
    disp = rng.integers(-1, D, (H, W)).astype(np.int16)
    disp[::7, ::5] = -32
    got = ctx.subpixel_enhancement(ctx.dev(disp), ctx.dev(vol)).cpu().numpy()
    assert _bits_equal(got, po.subpixel(disp, vol))


@pytest.mark.timeout(180)
@pytest.mark.parametrize("shape,num,thres", [((37, 53, 33), 6, 1.08), ((24, 40, 256), 6, 1.3), ((30, 31, 64), 2, 1.01),
                                             ((9, 11, 5), 8, 100.0)])
def test_select_top_cost_matches_oracle(ctx, shape, num, thres):
    H, W, D = shape
    rng = np.random.default_rng(17)
    vol = rng.random((H, W, D)).astype(np.float32) + 0.5
    vol[::3] = np.round(vol[::3] * 4) / 4                     # equal minima: the lowest d must win
    vol[1, 1] = 0.0                                           # first cost 0: cost < 0 * thres never holds
    vol[2, 2, 0] = -0.0                                       # the reference's '>' does not tell -0 from +0
    vol[2, 2, 1:] = np.where(np.arange(1, D) % 2 == 0, 0.0, 1.0)
    got = ctx.select_top_cost(ctx.dev(vol.copy()), num, thres).cpu().numpy()
    assert _bits_equal(got, po.select_top(vol, num, thres))


# ---------------------------------------------------------------- LRConsistencyCheck, LOR = 1 (VERDICT r01 missing #8)
def test_lrc_label_right_view_matches_reference_golden_and_oracle(ctx, golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "r02_ref.npz"))     # outputs of the reference's own LRConsistencyCheck(.., LOR = 1)
    for t in "abc":
        D = int(g[f"lor1_{t}_D"])
        d1, d2 = ctx.dev(g[f"lor1_{t}_d1"].copy()), ctx.dev(g[f"lor1_{t}_d2"].copy())
        mask, mask1 = ctx.lrc_label_lor(d1, d2, D, 1)
        assert np.array_equal(d2.cpu().numpy(), g[f"lor1_{t}_d2_after"]), t
        assert np.array_equal(d1.cpu().numpy(), g[f"lor1_{t}_d1_after"]), t        # the left map is only read
        assert np.array_equal(mask.cpu().numpy(), g[f"lor1_{t}_errmask"]), t        # errMask stays zero on this branch
        assert np.array_equal(mask1.cpu().numpy(), po.lrc_label_right(g[f"lor1_{t}_d1"], g[f"lor1_{t}_d2"], D)[1]), t
    rng = np.random.default_rng(31)
    H, W, D = 61, 133, 40
    d1 = rng.integers(-1, D, (H, W)).astype(np.int16)
    d2 = rng.integers(-1, D, (H, W)).astype(np.int16)
    for md in (0.0, 2.0):
        a, b = ctx.dev(d1.copy()), ctx.dev(d2.copy())
        _, m1 = ctx.lrc_label_lor(a, b, D, 1, md)
        want, wm = po.lrc_label_right(d1, d2, D, md)
        assert np.array_equal(b.cpu().numpy(), want) and np.array_equal(m1.cpu().numpy(), wm)
        a, b = ctx.dev(d1.copy()), ctx.dev(d2.copy())                              # LOR = 0 through the same entry point
        m0, m1 = ctx.lrc_label_lor(a, b, D, 0, md)
        want, wm = po.lrc_label(d1, d2, D, md)
        assert np.array_equal(a.cpu().numpy(), want) and np.array_equal(m0.cpu().numpy(), wm) and int(m1.max()) == 0


# ---------------------------------------------------------------- WM (VERDICT r01 missing #2)
@pytest.mark.timeout(300)
def test_wm_matches_reference_golden(ctx, golden_dir):
    """sm_wm against the outputs of the reference's OWN WM (tests/golden/wm_ref.npz): identical labels -- the weights come
    from expf as the host libm computes it (smd_expf_host) and every float sum keeps the reference's order."""
    import os
    g = np.load(os.path.join(golden_dir, "wm_ref.npz"))
    n = 0
    for k in g.files:
        if not k.endswith("_out"):
            continue
        tag = k.split("_")[0]
        disp, mask = g[k[:-4] + "_in"], g[k[:-4] + "_mask"]
        got, bad = ctx.wm(ctx.dev(disp.copy()), ctx.dev(mask), ctx.dev(g[tag + "_bgr"]), int(g[tag + "_D"]))
        assert np.array_equal(got.cpu().numpy(), g[k]), k
        assert bad == 0
        n += 1
    assert n == 9


@pytest.mark.timeout(300)
@pytest.mark.parametrize("shape", [(1, 1, 4), (3, 40, 9), (40, 3, 70), (96, 128, 32), (120, 160, 300)])
def test_wm_matches_oracle(ctx, shape):
    H, W, D = shape
    rng = np.random.default_rng(H * 7 + W)
    bgr = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    bgr[:, : W // 2] //= 8                                   # flat half: colour weights near 1
    disp = np.clip((np.arange(W)[None, :] * (D - 1) // max(W - 1, 1)) + rng.integers(-1, 2, (H, W)), 0, D - 1).astype(np.int16)
    disp[rng.random((H, W)) < 0.1] = rng.integers(0, D)
    mask = (rng.random((H, W)) < 0.35).astype(np.uint8) * 255
    got, bad = ctx.wm(ctx.dev(disp.copy()), ctx.dev(mask), ctx.dev(bgr), D)
    assert np.array_equal(got.cpu().numpy(), po.wm(disp, mask, bgr, D)) and bad == 0
    # labels outside [0, D) (DISP_OCC, -1 ...): undefined in the reference; sm_wm's defined extension = the oracle's lenient form
    disp2 = disp.copy()
    disp2[rng.random((H, W)) < 0.2] = -32
    got, bad = ctx.wm(ctx.dev(disp2.copy()), ctx.dev(mask), ctx.dev(bgr), D)
    want, wbad = po.wm_lenient(disp2, mask, bgr, D)
    assert np.array_equal(got.cpu().numpy(), want) and bad == wbad


# ---------------------------------------------------------------- vmTop, second half (VERDICT r01 missing #1)
def _top_variants(g):
    for k in g.files:
        if k.startswith("top_") and "_out_" in k:
            tag = k.split("_")[1]
            f = k.split("_out_")[1].split("_")
            yield k, tag, int(f[0][1:]), int(f[1][1:]), int(f[2][2:]), int(f[3][1:]), int(f[4][1:])


@pytest.mark.timeout(300)
def test_disp_from_top_cost_matches_reference_golden(ctx, golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "r02_ref.npz"))     # outputs of the reference's own genDispFromTopCostVm{,2}
    n = 0
    for k, tag, ver, m, ts, c2, cl in _top_variants(g):
        got = ctx.disp_from_top(ctx.dev(g[f"top_{tag}_in"]), ctx.dev(g[f"top_{tag}_bgr"]), ver, m, ts, c2, cl)
        assert np.array_equal(got.cpu().numpy(), g[k]), k
        n += 1
    assert n == 56


def _random_top(rng, H, W, D, num, spread):
    """Candidate lists with distinct disparities per pixel, ascending costs on a coarse grid (many ties)."""
    n = rng.integers(1, num + 1, (H, W))
    base = rng.integers(0, D, (H, W, 1))
    d = np.clip(base + rng.integers(-spread, spread + 1, (H, W, num)), 0, D - 1)
    far = rng.random((H, W, num)) < 0.3
    d = np.where(far, rng.integers(0, D, (H, W, num)), d)
    for k in range(1, num):                      # make the disparities of a pixel distinct
        for _ in range(4):
            dup = (d[..., k:k + 1] == d[..., :k]).any(-1)
            d[..., k] = np.where(dup, (d[..., k] + 1 + rng.integers(0, 3, (H, W))) % D, d[..., k])
        dup = (d[..., k:k + 1] == d[..., :k]).any(-1)
        n = np.where(dup & (n > k), k, n)         # still a duplicate: cut the list there
    c = np.sort(np.round(rng.random((H, W, num)) * 8) / 8 + 1.0, axis=-1)
    top = np.zeros((H, W, num + 1, 2), np.float32)
    keep = np.arange(num)[None, None, :] < n[..., None]
    top[..., :num, 0] = np.where(keep, d, 0)
    top[..., :num, 1] = np.where(keep, c, 0)
    top[..., num, 0] = n
    return top


@pytest.mark.timeout(600)
@pytest.mark.parametrize("shape", [(97, 131, 64, 6, 2), (60, 333, 256, 2, 40), (200, 65, 32, 16, 1), (33, 1000, 128, 4, 60)])
def test_disp_from_top_cost_matches_oracle(ctx, shape):
    H, W, D, num, spread = shape
    rng = np.random.default_rng(H + W)
    top = _random_top(rng, H, W, D, num, spread)
    bgr = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    bgr[:, : W // 2] //= 32
    dt, db = ctx.dev(top), ctx.dev(bgr)
    for ver, m, ts, c2, cl in [(1, 0, 10, 1, 0), (2, 0, 10, 1, 0), (2, 0, 2, 1, 0), (2, 0, 10, 0, 0), (2, 0, 1, 1, 1), (2, 1, 10, 1, 0),
                               (2, 2, 10, 1, 0)]:
        got = ctx.disp_from_top(dt, db, ver, m, ts, c2, cl).cpu().numpy()
        want = po.disp_from_top(top, bgr, ver, m, ts, c2, cl)
        assert np.array_equal(got, want), (ver, m, ts, c2, cl, int((got != want).sum()))
    # the raster-dependent case on nearly every pixel (ts = 0: no candidate ever has a partner), twice: same answer
    a = ctx.disp_from_top(dt, db, 2, 0, 0, 1, 0).cpu().numpy()
    assert np.array_equal(a, po.disp_from_top(top, bgr, 2, 0, 0, 1, 0))
    assert np.array_equal(a, ctx.disp_from_top(dt, db, 2, 0, 0, 1, 0).cpu().numpy())


@pytest.mark.timeout(600)
@pytest.mark.parametrize("method", [0, 1, 2])
def test_pipeline_with_vmtop_matches_oracle_chain(ctx, method):
    """dispOptimize with param_.Do_vmTop (stereoMatching.cpp:1111-1121) inside sm_pipeline: the left map after the whole
    refinement equals the oracle's stage functions composed the same way (M = 2, lamc = 109, ts = 10 as main_.cpp:62-64)."""
    H, W, D = 60, 90, 24
    p = _pair(H, W, D, "texture_warped", seed=17)
    aL, aR = po.arms(p["bgrL"]), po.arms(p["bgrR"])
    maps = []
    for view, bgr in ((0, p["bgrL"]), (1, p["bgrR"])):
        vol = po.adcensus_vol(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], D, view)
        vol = po.sgm(po.cbca(vol, aL, aR, 2, view), bgr, 4)
        top = po.select_top(vol, 2, np.float32(1.09))
        maps.append(po.disp_from_top(top, p["bgrL"], 2, method, 10, True, False))
    d = po.lrc_normal(maps[0], maps[1], 0.0)
    for _ in range(2):
        d = po.region_vote(d, aL, D)
    for _ in range(2):
        d = po.proper_ipol(d, p["bgrL"])
    want = po.median3_i16(d)
    pl = capi.Pipeline(ctx, H, W, capi.default_params(D - 1, sgm_paths=4, Do_vmTop=1, vmTop_method=method))
    got = pl.run(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"]).copy()
    pl.close()
    assert np.array_equal(got, want), float((got == want).mean())


@pytest.mark.timeout(300)
def test_pipeline_right_volume_kept_or_dropped_gives_the_same_maps(ctx):
    """sm_params.keep_right_volume: the last sgm path of view 1 either stores its sum (vm[1] as the reference leaves it) or
    only feeds the WTA -- identical maps, and with keep = 1 the stored volume equals the oracle's vm[1]."""
    H, W, D = 80, 200, 160
    p = _pair(H, W, D, "texture_warped", seed=23)
    aL, aR = po.arms(p["bgrL"]), po.arms(p["bgrR"])
    v1 = po.sgm(po.cbca(po.adcensus_vol(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], D, 1), aL, aR, 2, 1), p["bgrR"], 8)
    out = {}
    for paths, grouped in ((8, 1), (8, 0), (4, 0)):
        for keep in (0, 1):
            pl = capi.Pipeline(ctx, H, W, capi.default_params(D - 1, sgm_paths=paths, sgm_grouped=grouped, keep_right_volume=keep))
            pl.upload(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"])
            pl.run_device()
            out[keep] = pl.download(want_right=True)
            if keep and paths == 8 and not grouped:
                assert _bits_equal(pl.buffer(1, (H, W, D), torch.float32).cpu().numpy(), v1)
            pl.close()
        assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]), (paths, grouped)


# ---------------------------------------------------------------- native uint16 integer-cost path (VERDICT r01 missing #3)
def _u16(t):
    return t.cpu().numpy().view(np.uint16)


@pytest.mark.timeout(600)
@pytest.mark.parametrize("shape", [(37, 53, 19), (64, 160, 64), (40, 150, 130), (48, 96, 256), (33, 70, 300), (1, 30, 16), (30, 1, 16),
                                   (40, 700, 128), (36, 1930, 256), (2, 40, 256)])   # D = 128 / 256: the grouped row sweeps (k_sgm_group_u16)
def test_sgm_u16_is_the_float_reference_exactly(ctx, shape):
    """costScan's integer entry (stereoMatching.cpp:2007-2014) as a native uint16 path: for every path count and power-of-two
    reduCoeffi1 the fixed-point sum equals reduCoeffi1 x the float volume the reference's sgm() leaves, and the fused WTA
    equals gen_dispFromVm of it."""
    H, W, D = shape
    rng = np.random.default_rng(H * W + D)
    bgr = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    bgr[:, : W // 2] //= 32
    if min(H, W) > 8:
        p = _pair(H, W, D, "random_dot", seed=9)
        cL, cR = po.census(p["grayL"], 3), po.census(p["grayR"], 3)
        ham = po.hamming_vol(cL, cR, D, 3, 0)                 # a real 71-bit Hamming volume (out-of-range = 71)
        bgr = p["bgrL"]
    else:
        ham = rng.integers(0, 72, (H, W, D)).astype(np.float32)
    vol16 = ctx.dev(ham.astype(np.uint16).view(np.int16))
    db = ctx.dev(bgr)
    for paths, redu in ((4, 4), (8, 4), (8, 1), (5, 2), (1, 4)):
        ref = po.sgm(ham, bgr, paths, 15, redu)
        got, disp = ctx.sgm_u16(vol16, db, paths, 15, redu)
        assert np.array_equal(_u16(got).astype(np.float32) / redu, ref), (paths, redu)
        assert np.array_equal(disp.cpu().numpy(), po.wta(ref)), (paths, redu)
    with pytest.raises(capi.SmError):
        ctx.sgm_u16(vol16, db, 8, 15, 3)                      # 1/3 is not exact in the reference's floats either
    with pytest.raises(capi.SmError):
        ctx.sgm_u16(vol16, db, 8, 15, 4, max_cost=3000)       # 8 * 3003 * 4 does not fit 16 bits
    assert np.array_equal(ctx.wta_u16(vol16).cpu().numpy(), po.wta(ham))


@pytest.mark.timeout(600)
@pytest.mark.parametrize("agg,redu", [(0, 4), (0, 3), (1, 4)])
def test_pipeline_census_cost_matches_oracle(ctx, agg, redu):
    """costcalculation = "Census" (stereoMatching.cpp:975-976) through sm_pipeline: uint16 volumes when there is no aggregation
    and reduCoeffi1 is a power of two, float32 otherwise -- the refined map equals the oracle's either way."""
    H, W, D = 90, 140, 48
    p = _pair(H, W, D, "texture_warped", seed=29)
    op = po.default_params(D, paths=8, aggregation=agg, costcalc=2)
    op.reduCoeffi1 = redu
    rl, rr, rvol, _ = po.pipeline(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], op, want_vol=True)
    pl = capi.Pipeline(ctx, H, W, capi.default_params(D - 1, sgm_paths=8, sgm_grouped=0, aggregation=agg, costcalculation=2,
                                                     sgm_reduCoeffi1=redu))
    got = pl.run(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"]).copy()
    if agg == 0 and redu == 4:
        v = pl.buffer(0, (H, W, D), torch.int16)
        assert np.array_equal(_u16(v).astype(np.float32) / redu, rvol)       # fixed point = redu x the float sum
    elif redu == 4:
        assert _bits_equal(pl.buffer(0, (H, W, D), torch.float32).cpu().numpy(), rvol)
    pl.close()
    assert np.array_equal(got, rl)      # (redu = 3: float32 volumes, the same float operations in the same order)


# ---------------------------------------------------------------- discontinuityAdjust
def _da_case(H, W, D, seed, kind="piecewise"):
    rng = np.random.default_rng(seed)
    if kind == "piecewise":
        m = np.full((H, W), D // 3, np.int16)
        for _ in range(max(4, H * W // 600)):
            y0, x0 = rng.integers(0, max(1, H - 1)), rng.integers(0, max(1, W - 1))
            y1, x1 = rng.integers(y0 + 1, H + 1), rng.integers(x0 + 1, W + 1)
            m[y0:y1, x0:x1] = rng.integers(0, D)
        m[:, W // 2:] += (np.arange(W - W // 2) // 5).astype(np.int16)[None, :]
        m = np.clip(m, 0, D - 1).astype(np.int16)
        m[rng.random((H, W)) < 0.02] = -32
        m[rng.random((H, W)) < 0.01] = -48
    elif kind == "noise":
        m = rng.integers(0, D, (H, W)).astype(np.int16)
    else:
        m = np.full((H, W), 5 % D, np.int16)
    vol = (rng.random((H, W, D)) * 3).astype(np.float32)
    return m, vol


@pytest.mark.gpu
def test_disc_adjust_matches_reference_golden(ctx, golden_dir):
    """sm_discontinuity_adjust against the reference's own discontinuityAdjust outputs and cv2's edge maps (da_ref.npz)."""
    g = np.load(os.path.join(golden_dir, "da_ref.npz"))
    for pre in ("a", "b0", "b1", "b2"):
        d = ctx.dev(g[pre + "_disp"].copy())
        out, edge = ctx.discontinuity_adjust(d, ctx.dev(g[pre + "_vol"].copy()), want_edge=True)
        assert np.array_equal(edge.cpu().numpy(), g[pre + "_edge_cv2"]), pre
        assert np.array_equal(out.cpu().numpy(), g[pre + "_out"]), pre


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 1, 4), (2, 9, 4), (3, 3, 8), (5, 31, 8), (33, 32, 16), (34, 65, 16), (60, 97, 40),
                                   (270, 480, 64), (135, 1920, 32)])
@pytest.mark.parametrize("kind", ["piecewise", "noise", "const"])
def test_disc_adjust_matches_oracle(ctx, shape, kind):
    H, W, D = shape
    m, vol = _da_case(H, W, D, seed=H * 7 + W, kind=kind)
    want, edge_want, bad = po.disc_adjust(m, vol)
    assert bad == 0
    out, edge = ctx.discontinuity_adjust(ctx.dev(m.copy()), ctx.dev(vol), want_edge=True)
    assert np.array_equal(edge.cpu().numpy(), edge_want)
    assert np.array_equal(out.cpu().numpy(), want)
    if kind == "piecewise" and H * W > 2000:
        assert (want != m).sum() > 10
    # deterministic under the row pipeline: a second run lands on the same map
    out2 = ctx.discontinuity_adjust(ctx.dev(m.copy()), ctx.dev(vol))
    assert np.array_equal(out2.cpu().numpy(), want)


@pytest.mark.gpu
def test_disc_adjust_out_of_range_labels_are_not_indices(ctx):
    """Labels >= D (undefined in the reference) follow the oracle's defined extension; argument errors are refused."""
    H, W, D = 40, 70, 16
    m, vol = _da_case(H, W, D, seed=3)
    m[::7, ::5] = D + 3
    want, _, bad = po.disc_adjust(m, vol)
    assert bad > 0
    out = ctx.discontinuity_adjust(ctx.dev(m.copy()), ctx.dev(vol))
    assert np.array_equal(out.cpu().numpy(), want)
    L = ctx.L
    assert L.sm_discontinuity_adjust(ctx.h, None, None, H, W, D, None) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("ratio", [0.4, 1.0, 0.0])
def test_region_vote_unanimous_regions_and_out_of_range_votes(ctx, ratio):
    """regionVote_my's integer test (stereoMatching.cpp:7263: maxCount / validCount is an int division): for 0 < ratio <= 1 a
    pixel changes iff every valid vote of its region agrees -- the early-exit kernel -- incl. regions holding a vote >= D
    (counted as valid, in no bin: never unanimous); ratio 0 keeps the histogram kernel.  Both against the oracle."""
    H, W, D = 90, 150, 24
    rng = np.random.default_rng(77)
    p = _pair(H, W, D, "texture_warped", seed=31)
    arms = po.arms(p["bgrL"])
    d = np.full((H, W), 7, np.int16)
    d[:, W // 2:] = 11                                   # two constant halves: unanimous regions away from the seam
    d[40:50, 20:60] = rng.integers(0, D, (10, 40))       # a noisy patch: mixed regions
    d[rng.random((H, W)) < 0.3] = -32
    d[10, 10:14] = D + 5                                 # votes >= D
    ref = po.region_vote(d, arms, D, ratio, 20)
    got = ctx.region_vote(ctx.dev(d.copy()), ctx.dev(arms.view(np.int16)), D, ratio, 20).cpu().numpy()
    assert np.array_equal(got, ref)
    assert ((ref != d) & (d < 0)).sum() > 500            # the unanimous regions did fill their holes
