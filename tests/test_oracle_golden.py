"""CPU suite: pins the oracle (oracle/) against (i) cv2 golden vectors for the three OpenCV semantics on the
path, (ii) outputs of the reference's own NL/ sources (tests/golden/nl_ref.npz, made by
tests/golden/make_nl_golden.py from oracle/_ref/libqxref.so), (iii) hand-derived known-answer cases."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po


@pytest.fixture(scope="module")
def cvg(golden_dir):
    return np.load(os.path.join(golden_dir, "opencv_semantics.npz"))


@pytest.fixture(scope="module")
def nlg(golden_dir):
    return np.load(os.path.join(golden_dir, "nl_ref.npz"))


# ---------------------------------------------------------------- OpenCV semantics
def test_bgr2gray_matches_cv2(cvg):
    assert np.array_equal(po.bgr2gray(cvg["bgr"]), cvg["bgr2gray"])


def test_median3_i16_matches_cv2(cvg):
    assert np.array_equal(po.median3_i16(cvg["disp"]), cvg["median3"])


def _census_via_padded(padded, H, W, RV, RU, func):
    """Census evaluated on cv2's own BORDER_REFLECT_101 padded image (pure Python, small)."""
    dv_sur = [-1, -1, -1, 0, 1, 1, 1, 0, -1]
    du_sur = [-1, 0, 1, 1, 1, 0, -1, -1, -1]
    nbits = (2 * RV + 1) * (2 * RU + 1) + (8 if func == 3 else 0)
    nw = (nbits + 63) // 64
    out = np.zeros((H, W, nw), np.uint64)
    for v in range(H):
        for u in range(W):
            bits = []
            c = int(padded[v + RV, u + RU])
            for dv in range(-RV, RV + 1):
                for du in range(-RU, RU + 1):
                    bits.append(1 if c - int(padded[v + RV + dv, u + RU + du]) < 0 else 0)
            if func == 3:
                for i in range(8):
                    a = int(padded[v + RV + dv_sur[i], u + RU + du_sur[i]])
                    b = int(padded[v + RV + dv_sur[i + 1], u + RU + du_sur[i + 1]])
                    bits.append(1 if a - b < 0 else 0)
            for w in range(nw):
                chunk = bits[64 * w:64 * (w + 1)]
                val = 0
                for b in chunk:
                    val = (val << 1) | b
                out[v, u, w] = val
    return out


@pytest.mark.parametrize("func", [0, 3])
def test_census_uses_reflect101(cvg, func):
    gray, padded = cvg["gray"], cvg["reflect"]  # padded by cv2 with (3,3,4,4)
    H, W = gray.shape
    assert np.array_equal(po.census(gray, func), _census_via_padded(padded, H, W, 3, 4, func))


def test_census_known_answer_ramp():
    # strictly increasing along u, constant along v: neighbour > centre  <=>  du > 0 (interior pixel)
    gray = np.tile(np.arange(40, dtype=np.uint8) * 3, (12, 1))
    c0 = po.census(gray, 0)
    row = int("000001111" * 7, 2)
    assert int(c0[6, 20, 0]) == row
    c3 = po.census(gray, 3)
    # ring clockwise from top-left: tl<t, t<tr, tr=r, r=br, br>b, b>bl, bl=l, l=tl -> bits 1,1,0,0,0,0,0,0
    assert int(c3[6, 20, 0]) == (row << 1) | 1
    assert int(c3[6, 20, 1]) == 0b1000000


# ---------------------------------------------------------------- cost
def test_hamming_and_ad_known_answers():
    rng = np.random.default_rng(1)
    H, W, D = 5, 12, 6
    cL = rng.integers(0, 2**63, (H, W, 2), dtype=np.uint64)
    cR = rng.integers(0, 2**63, (H, W, 2), dtype=np.uint64)
    cL[..., 1] &= np.uint64(0x7F)
    cR[..., 1] &= np.uint64(0x7F)
    for LOR in (0, 1):
        vol = po.hamming_vol(cL, cR, D, 3, LOR)
        for v, u, d in [(0, 0, 0), (2, 3, 3), (4, 11, 5), (1, 2, 4), (3, 9, 4)]:
            ul, ur = (u, u - d) if LOR == 0 else (u + d, u)
            if ur < 0 or ul >= W:
                exp = 71
            else:
                exp = min(71, sum(bin(int(cL[v, ul, k]) ^ int(cR[v, ur, k])).count("1") for k in range(2)))
            assert vol[v, u, d] == exp
    bL = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    bR = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    ad = po.ad_vol(bL, bR, D, 0)
    assert ad[2, 1, 3] == 1000.0  # u-d < 0
    exp = np.float32(np.abs(bL[2, 7].astype(np.float32) - bR[2, 5].astype(np.float32)).sum()) / np.float32(3)
    assert ad[2, 7, 2] == exp
    comb = po.combine_exp(ad, po.hamming_vol(cL, cR, D, 3, 0))
    oor = np.float32(2) - np.exp(np.float32(-100.0)) - np.exp(np.float32(-71.0) / np.float32(30))
    assert abs(float(comb[2, 1, 3]) - float(oor)) < 1e-6


def test_exp_tables_reproduce_combine_bitwise():
    """The table form the fused GPU kernel uses is bit-identical to the three-pass reference arithmetic."""
    rng = np.random.default_rng(2)
    H, W, D = 9, 40, 16
    bL = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    bR = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    gL, gR = po.bgr2gray(bL), po.bgr2gray(bR)
    ref = po.adcensus_vol(bL, bR, gL, gR, D, LOR=0)
    tabAD = np.empty(767, np.float32)
    tabCen = np.empty(72, np.float32)
    po.lib().orc_exp_tables(1000.0, 10.0, 30.0, 71, tabAD, tabCen)
    cen = po.hamming_vol(po.census(gL), po.census(gR), D).astype(np.int64)
    k = np.full((H, W, D), 766, np.int64)
    for d in range(D):
        k[:, d:, d] = np.abs(bL[:, d:].astype(np.int64) - bR[:, :W - d].astype(np.int64)).sum(-1)
    tab = (np.float32(2) - tabAD[k]) - tabCen[cen]
    assert np.array_equal(tab.view(np.uint32), ref.view(np.uint32))


# ---------------------------------------------------------------- arms / CBCA
def test_arms_known_answers():
    img = np.full((9, 11, 3), 100, np.uint8)
    a = po.arms(img, L=3, L_out=5, tau=20, tau_out=6, minL=1)
    assert list(a[4, 5]) == [5, 5, 4, 4, 18]       # limited by L_out (h) and the border (v)
    assert list(a[0, 0]) == [0, 5, 0, 5, 10]       # border forces 0
    img2 = img.copy()
    img2[4, 7] = (100, 100, 130)                   # one channel jumps by 30 > tau
    a2 = po.arms(img2, L=3, L_out=5, tau=20, tau_out=6, minL=1)
    assert a2[4, 5, 1] == 1                        # right arm stops before u=7
    assert a2[4, 7, 0] == 1 and a2[4, 7, 1] == 1   # minL keeps 1 although the neighbour differs
    img3 = img.copy()
    img3[4, 9:] = 110                              # within tau (20) but above tau_out (6)
    a3 = po.arms(img3, L=3, L_out=5, tau=20, tau_out=6, minL=1)
    assert a3[4, 5, 1] == 3                        # arm 4 > L needs tau_out: stops at 3


def test_cbca_matches_bruteforce_small():
    rng = np.random.default_rng(3)
    H, W, D = 7, 9, 4
    aL = po.arms(rng.integers(90, 110, (H, W, 3), dtype=np.uint8), 2, 3, 8, 4, 1)
    aR = po.arms(rng.integers(90, 110, (H, W, 3), dtype=np.uint8), 2, 3, 8, 4, 1)
    vol = rng.integers(0, 16, (H, W, D)).astype(np.float32)  # small ints: float sums are exact
    out, area = po.cbca(vol, aL, aR, iters=1, view=0, want_area=True)
    isect = po.arms_intersect(aL, aR, D, 0).astype(int)
    for v in range(H):
        for u in range(W):
            for d in range(D):
                s, n = 0.0, 0
                for vn in range(v - isect[v, u, d, 2], v + isect[v, u, d, 3] + 1):
                    for un in range(u - isect[vn, u, d, 0], u + isect[vn, u, d, 1] + 1):
                        s += vol[vn, un, d]
                        n += 1
                assert area[v, u, d] == n
                assert out[v, u, d] == np.float32(s) / np.float32(n)


# ---------------------------------------------------------------- SGM / WTA / refine
def test_sgm_known_answer_1x3():
    # 1 x 3 x 3 volume, path 3 (predecessor u-1): hand-evaluated recurrence, P1=1, P2=3, flat image
    vol = np.array([[[5, 1, 4], [2, 6, 0], [3, 3, 3]]], np.float32)
    bgr = np.zeros((1, 3, 3), np.uint8)
    lr = po.sgm_path(vol, bgr, 3)
    assert lr[0, 0].tolist() == [5, 1, 4]
    # u=1: prev=[5,1,4], minC=1, P1=1-1=0: d0: min(4, inf, 1+0, 3)=1 -> 3; d1: min(0,5,4,3)=0 -> 6; d2: min(3,1,inf,3)=1 -> 1
    assert lr[0, 1].tolist() == [3, 6, 1]
    # u=2: prev=[3,6,1], minC=1, P1=0: d0: min(2,inf,6,3)=2 -> 5; d1: min(5,3,1,3)=1 -> 4; d2: min(0,6,inf,3)=0 -> 3
    assert lr[0, 2].tolist() == [5, 4, 3]
    bgr[0, 1] = (0, 40, 0)  # colour step 40 > 15 at u=1 and u=2: P1=0.25, P2=0.75
    lr = po.sgm_path(vol, bgr, 3)
    # u=1: P1=0.25-1=-0.75: d0: min(4, inf, 0.25, 0.75) -> 2.25; d1: min(0, 4.25, 3.25, .75)=0 -> 6; d2: min(3, .25, inf, .75) -> .25
    assert lr[0, 1].tolist() == [2.25, 6.0, 0.25]
    total = po.sgm(vol, bgr, paths=4)
    parts = sum(po.sgm_path(vol, bgr, p) for p in range(4))
    assert np.array_equal(total, parts)


def test_wta_first_minimum_and_lrc():
    vol = np.array([[[3, 1, 1, 2], [0, 0, 5, 5], [9, 8, 7, 7]]], np.float32)
    assert po.wta(vol).tolist() == [[1, 0, 2]]
    big = np.full((1, 1, 4), np.finfo(np.float32).max, np.float32)
    assert po.wta(big).tolist() == [[-1]]
    d1 = np.array([[0, 1, 2, 1, 5]], np.int16)
    d2 = np.array([[0, 9, 1, 9, 9]], np.int16)
    # u=0 d=0: D2[0]=0 ok; u=1 d=1: D2[0]=0 !=1 -> -1; u=2 d=2: D2[0]=0 -> -1; u=3 d=1: D2[2]=1 ok; u=4: u-d<0 -> -1
    assert po.lrc_normal(d1, d2).tolist() == [[0, -1, -1, 1, -1]]
    lab, mask = po.lrc_label(d1, d2, 8)
    # u=1: is there dd with D2[1-dd]==dd? dd=0: D2[1]=9; dd=1: D2[0]=0 -> no -> OCC; u=2: dd=1: D2[1]=9, dd=2: D2[0]=0 no, dd=0: D2[2]=1 no -> OCC
    assert lab.tolist() == [[0, -32, -32, 1, -32]]
    assert mask.tolist() == [[0, 255, 255, 0, 255]]


def test_region_vote_integer_division_rule():
    H, W, D = 9, 9, 8
    arms = np.zeros((H, W, 5), np.uint16)
    arms[..., :4] = 2
    arms[:2, :, 2] = 0; arms[-2:, :, 3] = 0; arms[:, :2, 0] = 0; arms[:, -2:, 1] = 0
    dp = np.full((H, W), 3, np.int16)
    dp[4, 4] = -1
    out = po.region_vote(dp, arms, D, 0.4, 20)
    assert out[4, 4] == 3            # 24 valid votes, unanimous
    dp[4, 5] = 4                     # one dissenting vote: 23/24 -> integer division gives 0 < 0.4
    assert po.region_vote(dp, arms, D, 0.4, 20)[4, 4] == -1
    dp[4, 5] = 3
    assert po.region_vote(dp, arms, D, 0.4, 24)[4, 4] == -1   # validNum <= S


def test_proper_ipol_rules():
    H, W = 7, 9
    bgr = np.zeros((H, W, 3), np.uint8)
    dp = np.full((H, W), -1, np.int16)
    dp[3, 6] = 5; bgr[3, 6] = (10, 0, 0)     # to the right (direction 2: pw=2 -> steps of 1)
    dp[3, 1] = 2; bgr[3, 1] = (3, 0, 0)      # to the left, closer in colour
    out = po.proper_ipol(dp, bgr)
    assert out[3, 3] == 2
    dp2 = dp.copy(); dp2[3, 3] = -32          # DISP_OCC takes the smallest disparity
    assert po.proper_ipol(dp2, bgr)[3, 3] == 2
    bgr[3, 1] = (255, 0, 0); bgr[3, 6] = (255, 0, 0)   # distance 255 can never win
    assert po.proper_ipol(dp, bgr)[3, 3] == -1


def test_solve_all_one_level():
    v = np.array([1.3, 2.6, 0.0], np.float32)
    out = v.copy()
    po.lib().orc_solve_all_1level(out, out.size, 0.3)
    inv = np.float32(1.0) / (np.float32(1.0) + np.float32(0.3))
    assert np.array_equal(out, inv * v)


# ---------------------------------------------------------------- NL: pinned against the compiled reference
def test_ctmf_matches_reference(nlg):
    assert np.array_equal(po.ctmf(nlg["ctmf_in3"], 1), nlg["ctmf_r1_cn3"])
    assert np.array_equal(po.ctmf(nlg["ctmf_in1"], 2), nlg["ctmf_r2_cn1"])
    assert np.array_equal(po.ctmf(nlg["ctmf_in1"], 1), nlg["ctmf_r1_cn1"])


@pytest.mark.parametrize("name", ["ramp", "smooth", "noisy"])
def test_mst_matches_reference(nlg, name):
    got = po.mst(nlg[f"mst_{name}_img"])
    for k in ("parent", "weight", "rank", "nr_child", "children", "order"):
        ref = nlg[f"mst_{name}_{k}"]
        g = got[k]
        if k == "parent":
            ref = ref.copy(); g = g.copy()
            ref[0] = 0; g[0] = 0     # root's parent: implementation detail (-1 vs itself)
        if k == "weight":
            ref = ref.copy(); g = g.copy()
            ref[0] = 0; g[0] = 0
        assert np.array_equal(g, ref), k


def test_tree_filter_matches_reference(nlg):
    vol = nlg["tf_vol"]
    for img, key, sigma in ((nlg["mst_noisy_img"], "tf_out_sigma0p1", 0.1),
                            (nlg["mst_noisy_img"], "tf_out_sigma0p05", 0.05),
                            (np.ascontiguousarray(nlg["mst_smooth_img"][:24, :30]), "tf_smooth_out", 0.1)):
        t = po.mst(img)
        table = np.empty(256, np.float64)
        po.lib().orc_tree_table(sigma, table)
        cost = vol.astype(np.float64).copy()
        tmp = np.empty_like(cost)
        H, W, D = cost.shape
        po.lib().orc_tree_filter(cost, tmp, H * W, D, t["parent"], t["weight"], t["nr_child"], t["children"],
                                 t["order"], table)
        assert np.array_equal(cost, nlg[key]), key


@pytest.mark.skipif(po.ref_lib() is None, reason="oracle/_ref/libqxref.so not built (needs /root/reference)")
def test_nl_restatement_against_live_reference():
    rng = np.random.default_rng(77)
    img = rng.integers(0, 256, (33, 41, 3), dtype=np.uint8)
    img[5:20, 8:30] //= 8    # flat-ish region: many weight ties
    a, b = po.mst(img), po.ref_mst(img)
    for k in ("rank", "nr_child", "children", "order"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["parent"][1:], b["parent"][1:])
    assert np.array_equal(a["weight"][1:], b["weight"][1:])
    vol = rng.random((33, 41, 5)).astype(np.float32)
    ours = po.nl_aggre(img, vol)
    ref = po.ref_tree_filter(img, vol.astype(np.float64), 0.1).astype(np.float32)
    assert np.array_equal(ours, ref)
    assert np.array_equal(po.ctmf(img, 1), po.ref_ctmf(img, 1))


# ---------------------------------------------------------------- whole chain
def test_pipeline_recovers_synthetic_disparity():
    from mystereomatching_b200 import synth
    pair = synth.make_pair(60, 96, 24, "texture_warped", seed=5)
    p = po.default_params(24, paths=4)
    dl, dr, _, ms = po.pipeline(pair["bgrL"], pair["bgrR"], pair["grayL"], pair["grayR"], p)
    assert synth.bad_k(dl, pair["gt"], pair["nonocc"], 2) < 15.0
    po.lib().orc_set_threads(4)
    dl4, _, _, _ = po.pipeline(pair["bgrL"], pair["bgrR"], pair["grayL"], pair["grayR"], p)
    po.lib().orc_set_threads(1)
    assert np.array_equal(dl, dl4)   # thread count never changes a result


# ---------------------------------------------------------------- Yang's driver class: pinned against the compiled reference
@pytest.fixture(scope="module")
def nlcag(golden_dir):
    return np.load(os.path.join(golden_dir, "nlca_ref.npz"))


def test_nlca_cost_and_flip_match_reference(nlcag):
    L, R, D = nlcag["left"], nlcag["right"], int(nlcag["D"])
    assert np.array_equal(po.nlca_gradient(L), nlcag["grad_left"])
    c = po.nlca_cost(L, R, D)
    assert np.array_equal(c, nlcag["cost"])
    assert np.array_equal(po.flip_vol(c), nlcag["cost_right"])


def test_nlca_disparity_matches_reference(nlcag):
    L, R, D = nlcag["left"], nlcag["right"], int(nlcag["D"])
    assert np.array_equal(po.nlca_disparity(L, R, D, 0.1, post=False), nlcag["disp"])
    assert np.array_equal(po.nlca_disparity(L, R, D, 0.1, post=True), nlcag["disp_post"])


def test_nlca_helpers_known_answers():
    v = np.array([[[3.0, 1.0, 1.0], [0.5, 0.5, 0.2]]])
    assert po.depth_best_cost(v).tolist() == [[1, 2]]                     # first minimum
    dl = np.array([[0, 1, 2, 3]], np.uint8)
    dr = np.array([[1, 9, 9, 9]], np.uint8)
    # x=0: d=0 -> 255; x=1: xr=0, |1-1|=0 -> 0; x=2: xr=0, |2-1|>=1 -> 255; x=3: xr=0, |3-1| -> 255
    assert po.detect_occlusion(dl, dr).tolist() == [[255, 0, 255, 255]]
    vol = np.arange(2 * 3 * 2, dtype=np.float64).reshape(1, 6, 2) if False else np.arange(12, dtype=np.float64).reshape(1, 3, 4)
    f = po.flip_vol(vol)
    # right[x][d] = left[x+d][d] while x+d < W, else the previous plane's value
    assert f[0, 0].tolist() == [vol[0, 0, 0], vol[0, 1, 1], vol[0, 2, 2], vol[0, 2, 2]]
    assert f[0, 2].tolist() == [vol[0, 2, 0]] * 4


# ---------------------------------------------------------------- cross-scale step
def test_pyr_down_matches_cv2(cvg):
    for k in range(int(cvg["n_pyr"])):
        assert np.array_equal(po.pyr_down(cvg[f"pyr_in{k}"]), cvg[f"pyr_out{k}"]), k


def test_cross_scale_weights_match_cv2_invert(cvg):
    for a, lam in enumerate(cvg["inv_lams"]):
        for n in range(1, 8):
            got = po.cross_scale_weights(n, float(lam))
            assert np.array_equal(got.view(np.uint32), cvg["inv_rows"][a, n - 1, :n].view(np.uint32)), (lam, n)


def test_solve_all_known_answer():
    """SolveAll by hand: 2 levels, weights w; out[y,x,d] = w0*v0[y,x,d] + w1*v1[y//2,x//2,(d+1)//2] in float order."""
    rng = np.random.default_rng(3)
    v0 = rng.random((5, 6, 7)).astype(np.float32)
    v1 = rng.random((3, 3, 5)).astype(np.float32)
    w = po.cross_scale_weights(2, 0.3)
    want = np.empty_like(v0)
    for y in range(5):
        for x in range(6):
            for d in range(7):
                s = np.float32(0)
                s = np.float32(s + np.float32(w[0] * v0[y, x, d]))
                s = np.float32(s + np.float32(w[1] * v1[y // 2, x // 2, (d + 1) // 2]))
                want[y, x, d] = s
    assert np.array_equal(po.solve_all([v0, v1], 0.3).view(np.uint32), want.view(np.uint32))
    one = po.solve_all([v0], 0.3)
    ref = v0.copy().reshape(-1)
    po.lib().orc_solve_all_1level(ref, ref.size, 0.3)
    assert np.allclose(one.reshape(-1), ref, rtol=2e-7)      # 1/(1+l): float division vs (float)(1./double) of cv


def test_cal_err_known_answer():
    """calErr by hand: 5 masked pixels; errors = one invalid (-1), one |7.5 - 5| > 1, one invalid (-48) -> PBM 3/5;
    squared error sum 0 + 2 + 6.25 + 0.25 + 2 = 10.5 -> RMS sqrt(2.1)."""
    dp = np.array([[3, -1, 5], [2, 2, -48]], np.int16)
    gt = np.array([[3, 4, 7.5], [0, 2.5, 1]], np.float32)
    mask = np.array([[255, 255, 255], [0, 255, 255]], np.uint8)
    pbm, rms, n, e = po.cal_err(dp, gt, mask, 1)
    assert (n, e) == (5, 3) and abs(pbm - 0.6) < 1e-7 and abs(rms - np.sqrt(2.1)) < 1e-6


def test_select_top_equals_reference(golden_dir):
    """orc_select_top against the outputs of the reference's own selectTopCostFromVolumn (tests/golden/top_ref.npz)."""
    import os
    g = np.load(os.path.join(golden_dir, "top_ref.npz"))
    n = 0
    for k in g.files:
        if "_top_" not in k:
            continue
        vol = g[k.split("_top_")[0] + "_vol"]
        num, thres = int(k.split("_n")[1].split("_")[0]), float(k.split("_t")[-1])
        got = po.select_top(vol, num, thres)
        assert np.array_equal(got.view(np.uint32), g[k].view(np.uint32)), k
        n += 1
    assert n == 5
    # known answer by hand: costs [5, 2, 2, 7, 2.1], thres 1.08 -> d=1 (first of the equal minima), then d=2 (2 < 2.16),
    # then d=4 (2.1 < 2.16), then 5 fails
    v = np.array([[[5, 2, 2, 7, 2.1]]], np.float32)
    t = po.select_top(v, 4, 1.08)[0, 0]
    assert t[:, 0].tolist() == [1, 2, 4, 0, 3] and t[:3, 1].tolist() == [2, 2, np.float32(2.1)] and t[3].tolist() == [0, 0]


def test_subpixel_equals_reference(golden_dir):
    """orc_subpixel against the outputs of the reference's own subpixelEnhancement (tests/golden/subpixel_ref.npz)."""
    import os
    g = np.load(os.path.join(golden_dir, "subpixel_ref.npz"))
    n = changed = 0
    for k in g.files:
        if "_se_" not in k:
            continue
        pre, tag = k.split("_se_")
        disp = g[f"{pre}_disp_{tag}"]
        got = po.subpixel(disp, g[pre + "_vol"])
        assert np.array_equal(got.view(np.uint32), g[k].view(np.uint32)), k
        assert np.array_equal(got, np.trunc(got))           # the reference truncates on the short: integer-valued
        changed += int((got != disp).sum())
        n += 1
    assert n == 3 and changed > 0
    # known answers by hand (D=5): costs 4,1,2 around d=2 -> diff = (2-4)/(2*(2+4-2)) = -0.25 -> (short)2.25 = 2;
    # costs 2,1,4 -> diff = +0.25 -> (short)1.75 = 1; flat 3,3,3 -> denom 0, unchanged; d = 0, D-1, negative: unchanged
    v = np.array([[[9, 4, 1, 2, 9], [9, 2, 1, 4, 9], [9, 3, 3, 3, 9], [0, 5, 5, 5, 5], [5, 5, 5, 5, 0], [1, 2, 3, 4, 5]]],
                 np.float32)
    d = np.array([[2, 2, 2, 0, 4, -32]], np.int16)
    assert po.subpixel(d, v).tolist() == [[2.0, 1.0, 2.0, 0.0, 4.0, -32.0]]


def test_median3_f32_matches_cv2(golden_dir):
    """orc_median3_f32 against cv2.medianBlur(CV_32F, 3) on the reference's sub-pixel maps and a fractional map."""
    import os
    g = np.load(os.path.join(golden_dir, "subpixel_ref.npz"))
    pairs = [(k.replace("_semed_", "_se_"), k) for k in g.files if "_semed_" in k]
    pairs += [(k[:-4] + "_map", k) for k in g.files if k.endswith("_med")]
    assert len(pairs) == 8
    for src, want in pairs:
        assert np.array_equal(po.median3_f32(g[src]).view(np.uint32), g[want].view(np.uint32)), want


def test_wm_equals_reference(golden_dir):
    """orc_wm against the outputs of the reference's own WM (tests/golden/wm_ref.npz); labels outside [0, D) -- where the
    reference is undefined (stereoMatching.cpp:7371) -- are refused."""
    import os
    g = np.load(os.path.join(golden_dir, "wm_ref.npz"))
    n = changed = 0
    for k in g.files:
        if not k.endswith("_out"):
            continue
        tag = k.split("_")[0]
        disp, mask = g[k[:-4] + "_in"], g[k[:-4] + "_mask"]
        got = po.wm(disp, mask, g[tag + "_bgr"], int(g[tag + "_D"]))
        assert np.array_equal(got, g[k]), k
        assert np.array_equal(got[mask == 0], disp[mask == 0])      # only flagged pixels are rewritten
        changed += int((got != disp).sum())
        n += 1
    assert n == 9 and changed > 0
    bad = g["a_noisy_some_in"].copy()
    bad[3, 3] = -32                                                 # DISP_OCC inside somebody's window
    with pytest.raises(ValueError):
        po.wm(bad, g["a_noisy_some_mask"], g["a_bgr"], int(g["a_D"]))


def test_opencv_restatements_equal_cv2(golden_dir):
    """oracle/opencv_restated.h (equalizeHist, GaussianBlur(3x3, 4), Canny(20, 60, 3)) against cv2 4.13's outputs
    (tests/golden/da_ref.npz): every image incl. the constant one and the 1x1 ... 3x3 shapes, bit for bit."""
    from oracle import pyoracle as po
    g = np.load(os.path.join(golden_dir, "da_ref.npz"))
    names = [k[4:] for k in g.files if k.startswith("img_")]
    assert len(names) >= 9
    for k in names:
        im = g["img_" + k]
        assert np.array_equal(po.equalize_hist(im), g["eq_" + k]), k
        assert np.array_equal(po.gauss3_sigma4(im), g["gb_" + k]), k
        assert np.array_equal(po.canny3_l1(im, 20, 60), g["cn_" + k]), k
        assert np.array_equal(po.canny3_l1(g["gb_" + k], 20, 60), g["cnb_" + k]), k


def test_disc_adjust_equals_reference(golden_dir):
    """orc_disc_adjust against the reference's own discontinuityAdjust (da_ref.npz), and its edge chain against cv2's."""
    from oracle import pyoracle as po
    g = np.load(os.path.join(golden_dir, "da_ref.npz"))
    for pre in ("a", "b0", "b1", "b2"):
        disp, vol = g[pre + "_disp"], g[pre + "_vol"]
        assert np.array_equal(po.da_edges(disp), g[pre + "_edge_cv2"]), pre
        got, edge, bad = po.disc_adjust(disp, vol)
        assert bad == 0
        assert np.array_equal(edge, g[pre + "_edge_cv2"])
        assert np.array_equal(got, g[pre + "_out"]), pre
        assert (got != disp).sum() > 20          # the case exercises the update
        if po.smref_lib() is not None and pre == "b0":
            H, W, D = vol.shape
            z = np.zeros((H, W, 3), np.uint8)
            r = po.SmRef(z, z, z[..., 0].copy(), z[..., 0].copy(), D)
            r.adcensus()
            r.arms()
            r.cbca(1)                                # leaves vm[0] as the 3-D Mat the function indexes (after ADCensusCal
                                                     # alone it is a 2-D multi-channel header, on which at(h, w, d) is undefined)
            r.set_vm(0, vol)
            assert np.array_equal(r.disc_adjust(disp), got)
            r.close()
