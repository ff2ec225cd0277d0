"""The C++ host side: class StereoMatching / NLCCA / qx_tree_filter / ctmf with the reference's names and signatures
(mystereomatching_b200/host/), driven the way the reference's main() drives it, compared with the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

from mystereomatching_b200 import synth
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "mystereomatching_b200", "host")
EXE = os.path.join(HOST, "stage_api_test")
LIB = os.path.join(ROOT, "mystereomatching_b200", "libstereomatching_b200.so")


def test_host_library_exports_the_stage_api():
    assert os.path.exists(LIB) and os.path.exists(EXE), "run __graft_entry__.build()"
    out = subprocess.check_output(["nm", "-DC", "--defined-only", LIB], text=True)
    for sym in ("StereoMatching::pipeline()", "StereoMatching::costCalculate()", "StereoMatching::dispOptimize()",
                "StereoMatching::refine()", "StereoMatching::ADCensusCal()", "StereoMatching::CBCA()",
                "StereoMatching::NL()", "StereoMatching::sgm(cv::Mat&, bool)",
                "StereoMatching::costScan(cv::Mat&, cv::Mat&, int, int, bool)",
                "StereoMatching::gen_dispFromVm(cv::Mat&, cv::Mat&)", "StereoMatching::wta_Co(",
                "StereoMatching::selectTopCostFromVolumn(cv::Mat&, cv::Mat&, float)",
                "StereoMatching::subpixelEnhancement(cv::Mat&, cv::Mat&)", "StereoMatching::discontinuityAdjust(cv::Mat&)",
                "StereoMatching::regionVote_my(cv::Mat&, float, int)", "StereoMatching::properIpol(",
                "StereoMatching::LRConsistencyCheck_normal(", "StereoMatching::genCensusCode_NC_Sur(",
                "StereoMatching::gen_cenVM_XOR(", "StereoMatching::cbca_core(", "StereoMatching::genTrueHorVerArms(",
                "StereoMatching::gen1DCumu(", "StereoMatching::cal1DCost(", "StereoMatching::genfinalVm_cbca(",
                "void StereoMatching::updateCost<float>(", "StereoMatching::LRConsistencyCheck_new(",
                "void StereoMatching::calErr<short>(", "SolveAll(StereoMatching**&, int, float)", "pyrDown_u8(", "StereoMatching::censusGrad(", "StereoMatching::grad(", "StereoMatching::calGrad(",
                "StereoMatching::calGrad_y(", "StereoMatching::calgradvm(",
                "void StereoMatching::calHorVerDis<unsigned char>(", "void StereoMatching::calArms<unsigned char>(",
                "NLCCA::aggreCV(", "qx_tree_filter::filter(double*, double*, int)", "qx_tree_filter::build_tree(",
                "qx_nonlocal_cost_aggregation::matching_cost(unsigned char***, unsigned char***)",
                "qx_nonlocal_cost_aggregation::disparity(unsigned char**, bool)", "qx_nonlocal_cost_aggregation::init(",
                "ctmf"):
        assert sym in out, sym


def _run(tmp_path, pair, D, paths, mode):
    H, W = pair["grayL"].shape
    inp = tmp_path / "in.bin"
    with open(inp, "wb") as f:
        for k in ("bgrL", "bgrR", "grayL", "grayR"):
            f.write(np.ascontiguousarray(pair[k]).tobytes())
    prefix = str(tmp_path / "out")
    r = subprocess.run([EXE, str(inp), prefix, str(H), str(W), str(D), str(paths), mode], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return prefix, r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["pipeline", "stages"])
def test_cpp_class_matches_oracle(tmp_path, mode):
    H, W, D, P = 72, 110, 40, 4
    pair = synth.make_pair(H, W, D, "texture_warped", seed=17)
    prefix, _ = _run(tmp_path, pair, D, P, mode)
    ref, _, vol, _ = po.pipeline(pair["bgrL"], pair["bgrR"], pair["grayL"], pair["grayR"], po.default_params(D, paths=P),
                                 want_vol=True)
    dp = np.fromfile(prefix + ".dp0.i16", np.int16).reshape(H, W)
    assert (dp == ref).mean() >= 0.995
    hvl = np.fromfile(prefix + ".hvl0.u16", np.uint16).reshape(H, W, 5)
    assert np.array_equal(hvl, po.arms(pair["bgrL"]))
    if mode == "pipeline":
        gt = dp.astype(np.float32) + np.where(np.arange(W)[None, :] % 7 == 0, 2.0, 0.0).astype(np.float32)
        pbm, rms, _, _ = po.cal_err(dp, gt, np.full((H, W), 255, np.uint8), 1)
        e = np.fromfile(prefix + ".err_all.f32", np.float32)
        assert abs(e[0] - pbm) < 1e-6 and abs(e[1] - rms) <= 1e-5 * rms and pbm > 0
    if mode == "stages":
        aL, aR = po.arms(pair["bgrL"]), po.arms(pair["bgrR"])
        cost = po.adcensus_vol(pair["bgrL"], pair["bgrR"], pair["grayL"], pair["grayR"], D, 0)
        cb = po.cbca(cost, aL, aR, 2, 0)
        got = np.fromfile(prefix + ".vm0_cbca.f32", np.float32).reshape(H, W, D)
        assert np.array_equal(got.view(np.uint32), cb.view(np.uint32))
        # selectTopCostFromVolumn on a clone of the SGM volume: candidates = the oracle's, taken entries -> FLT_MAX
        sg = np.fromfile(prefix + ".vm0_sgm.f32", np.float32).reshape(H, W, D)
        top = np.fromfile(prefix + ".top0.f32", np.float32).reshape(H, W, 7, 2)
        want = po.select_top(sg, 6, 1.08)
        assert np.array_equal(top.view(np.uint32), want.view(np.uint32))
        # subpixelEnhancement on the WTA map of that volume
        dw = np.fromfile(prefix + ".dp0_wta.i16", np.int16).reshape(H, W)
        se = np.fromfile(prefix + ".se0.f32", np.float32).reshape(H, W)
        se_want = po.subpixel(dw, sg)
        assert np.array_equal(se.view(np.uint32), se_want.view(np.uint32)) and (se_want != dw).any()
        # discontinuityAdjust on a clone of the WTA map over vm[0]
        da = np.fromfile(prefix + ".da0.i16", np.int16).reshape(H, W)
        da_want, _, bad = po.disc_adjust(dw, sg)
        assert bad == 0 and np.array_equal(da, da_want)
        # refine() with Do_subpixelEnhancement: SE = median3(subpixel(DP[0] before the last median, vm[0]))
        pm = np.fromfile(prefix + ".dp0_premed.i16", np.int16).reshape(H, W)
        ser = np.fromfile(prefix + ".se_refine.f32", np.float32).reshape(H, W)
        assert np.array_equal(ser.view(np.uint32), po.median3_f32(po.subpixel(pm, sg)).view(np.uint32))
        assert np.array_equal(po.median3_i16(pm), dp)            # and the two-call refine ends where one call does
        marked = np.fromfile(prefix + ".top0_vm.f32", np.float32).reshape(H, W, D)
        exp = sg.copy()
        for k in range(6):
            sel = want[:, :, 6, 0] > k
            vv, uu = np.nonzero(sel)
            exp[vv, uu, want[vv, uu, k, 0].astype(np.int64)] = np.finfo(np.float32).max
        assert np.array_equal(marked.view(np.uint32), exp.view(np.uint32)) and (want[:, :, 6, 0] > 1).any()
        wta = np.fromfile(prefix + ".dp0_wta.i16", np.int16).reshape(H, W)
        assert np.array_equal(wta, po.wta(po.sgm(cb, pair["bgrL"], P)))
        ad = np.fromfile(prefix + ".ad0.f32", np.float32).reshape(H, W, D)
        assert np.array_equal(ad, po.ad_vol(pair["bgrL"], pair["bgrR"], D, 0))
        cen = np.fromfile(prefix + ".cenL.u64", np.uint64).reshape(H, W, 2)
        assert np.array_equal(cen, po.census(pair["grayL"], 3))
        cv0 = np.fromfile(prefix + ".cen0.f32", np.float32).reshape(H, W, D)
        ham = po.hamming_vol(po.census(pair["grayL"]), po.census(pair["grayR"]), D)
        assert np.array_equal(cv0, ham)
        lr = np.fromfile(prefix + ".lr3.f32", np.float32).reshape(H, W, D)
        assert np.array_equal(lr, po.sgm_path(ham, pair["bgrL"], 3))
        # gen1DCumu / cal1DCost / genfinalVm_cbca one by one == one fused CBCA iteration
        parts = np.fromfile(prefix + ".parts_cbca1.f32", np.float32).reshape(H, W, D)
        assert np.array_equal(parts.view(np.uint32), po.cbca(ad, aL, aR, 1, 0).view(np.uint32))
        # updateCost<float> at one pixel restores the wiped row of the path volume
        assert np.array_equal(np.fromfile(prefix + ".lr3_pixel.f32", np.float32).reshape(H, W, D), lr)
        assert np.array_equal(np.fromfile(prefix + ".lr3_u16.f32", np.float32).reshape(H, W, D), lr)   # CV_16U entry
        # LRConsistencyCheck_new on the final DP[0] / DP[1]
        m = np.fromfile(prefix + ".lrc_new.u8", np.uint8).reshape(H, W)
        assert set(np.unique(m)) <= {0, 255} and (m == 0).any()


@pytest.mark.gpu
def test_cpp_class_censusgrad_chain_matches_oracle(tmp_path):
    """costcalculation = "censusGrad" (the selector the reference's main_.cpp:15 compiles in) through the C++ class."""
    H, W, D, P = 60, 96, 24, 4
    pair = synth.make_pair(H, W, D, "texture_warped", seed=23)
    prefix, _ = _run(tmp_path, pair, D, P, "censusgrad")
    bl, br, gl, gr = pair["bgrL"], pair["bgrR"], pair["grayL"], pair["grayR"]
    gx0, _ = po.grad_xy(gl)
    _, gy1 = po.grad_xy(gr)
    assert np.array_equal(np.fromfile(prefix + ".gx0.f32", np.float32).reshape(H, W), gx0)
    assert np.array_equal(np.fromfile(prefix + ".gy1.f32", np.float32).reshape(H, W), gy1)
    gv = np.fromfile(prefix + ".gradvm1.f32", np.float32).reshape(H, W, D)
    assert np.array_equal(gv.view(np.uint32), po.grad_vol(gl, gr, po.arms(br), D, 1).view(np.uint32))
    cg = np.fromfile(prefix + ".cg0.f32", np.float32).reshape(H, W, D)
    ref = po.censusgrad_vol(bl, br, gl, gr, D, 0)
    assert np.all(np.abs(cg - ref) <= 1e-4 * np.abs(ref))
    res = po.pipeline(bl, br, gl, gr, po.default_params(D, paths=P, costcalc=1))
    dp = np.fromfile(prefix + ".dp0.i16", np.int16).reshape(H, W)
    assert (dp == res[0]).mean() >= 0.995


@pytest.mark.gpu
def test_cpp_pyramid_main_flow_matches_oracle(tmp_path):
    """The reference's main() loop with PY_LEV = 3 written against the C++ class: per-level costCalculate(),
    the free function SolveAll(smPyr, PY_LEV, 0.3), dispOptimize(), refine()."""
    H, W, D, P = 66, 100, 32, 4
    pair = synth.make_pair(H, W, D, "texture_warped", seed=29)
    prefix, _ = _run(tmp_path, pair, D, P, "pyramid")
    res = po.pipeline(pair["bgrL"], pair["bgrR"], pair["grayL"], pair["grayR"],
                      po.default_params(D, paths=P, pyr_levels=3, cross_lambda=0.3))
    dp = np.fromfile(prefix + ".dp0.i16", np.int16).reshape(H, W)
    assert np.array_equal(dp, res[0])       # 4 paths in the reference's order: identical map


@pytest.mark.gpu
def test_cpp_nl_surface_matches_oracle(tmp_path):
    H, W, D = 64, 80, 6          # ctmf r = 2 in the reference needs h*w >= 5*544
    pair = synth.make_pair(H, W, 8, "texture_warped", seed=4)
    prefix, _ = _run(tmp_path, pair, D, 4, "nl")
    img = pair["bgrL"]
    assert np.array_equal(np.fromfile(prefix + ".ctmf.u8", np.uint8).reshape(H, W, 3), po.ctmf(img, 1))
    t = po.mst(img)
    assert np.array_equal(np.fromfile(prefix + ".rank.i32", np.int32), t["rank"])
    assert np.array_equal(np.fromfile(prefix + ".parent.i32", np.int32)[1:], t["parent"][1:])
    i = np.arange(H * W * D, dtype=np.uint64)
    cost = ((i * np.uint64(2654435761)) % np.uint64(1000)).astype(np.float64) / 1000.0   # 64-bit product, as in C++
    table = np.empty(256, np.float64)
    po.lib().orc_tree_table(0.1, table)
    ref = cost.reshape(H * W, D).copy()
    tmp = np.empty_like(ref)
    po.lib().orc_tree_filter(ref, tmp, H * W, D, t["parent"], t["weight"], t["nr_child"], t["children"], t["order"], table)
    got = np.fromfile(prefix + ".tf.f64", np.float64).reshape(H * W, D)
    assert np.array_equal(got, ref)                                     # float64, bit for bit
    vol = ((i * np.uint64(2654435761)) % np.uint64(1000)).astype(np.float32) / np.float32(1000.0)
    ag = po.nl_aggre(img, vol.reshape(H, W, D))
    assert np.array_equal(np.fromfile(prefix + ".aggre.f32", np.float32).reshape(H, W, D), ag)
    # qx_nonlocal_cost_aggregation: init / matching_cost / disparity
    L, R = pair["bgrL"], pair["bgrR"]
    assert np.array_equal(np.fromfile(prefix + ".nlca_cost.f64", np.float64).reshape(H, W, D), po.nlca_cost(L, R, D))
    assert np.array_equal(np.fromfile(prefix + ".nlca_disp.u8", np.uint8).reshape(H, W), po.nlca_disparity(L, R, D))
    assert np.array_equal(np.fromfile(prefix + ".nlca_disp_post.u8", np.uint8).reshape(H, W),
                          po.nlca_disparity(L, R, D, post=True))


@pytest.mark.gpu
def test_cpp_error_convention(tmp_path):
    pair = synth.make_pair(24, 32, 8, "random_dot", seed=1)
    _, out = _run(tmp_path, pair, 8, 4, "errors")
    assert "caught 2" in out
