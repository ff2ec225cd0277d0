"""Generates tests/golden/sm_ref.npz by RUNNING THE REFERENCE'S OWN stereoMatching.{h,cpp} FUNCTION BODIES
(oracle/build_ref_sm.py cuts them from /root/reference and compiles them against a cv::Mat stand-in into
oracle/_ref/libsmref.so) on seeded synthetic stereo pairs.  Needs /root/reference; the committed .npz travels.
Every array is the output of a reference function at a stage boundary of the hot path.
Run:  python tests/golden/make_sm_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import pyoracle as po  # noqa: E402
from mystereomatching_b200 import synth  # noqa: E402

assert po.smref_lib() is not None, "run python oracle/build_ref_sm.py first (needs /root/reference)"
out = {}


def stages(tag, H, W, D, kind, seed, full):
    p = synth.make_pair(H, W, D, kind, seed=seed)
    bl, br, gl, gr = p["bgrL"], p["bgrR"], p["grayL"], p["grayR"]
    for k, v in (("bgrL", bl), ("bgrR", br), ("grayL", gl), ("grayR", gr)):
        out[f"{tag}_{k}"] = v
    out[f"{tag}_D"] = np.int32(D)
    r = po.SmRef(bl, br, gl, gr, D)
    for func in (0, 3):
        cl, cr = r.census(func)                       # genCensusCode<uchar> / genCensusCode_NC_Sur
        out[f"{tag}_census{func}_L"], out[f"{tag}_census{func}_R"] = cl, cr
    if full:
        h0, h1 = r.census_cal(3)                      # censusCal -> gen_cenVM_XOR
        out[f"{tag}_hamming3_v0"], out[f"{tag}_hamming3_v1"] = h0.astype(np.uint8), h1.astype(np.uint8)
        assert np.array_equal(h0, out[f"{tag}_hamming3_v0"].astype(np.float32))
        h0, _ = r.census_cal(0)
        out[f"{tag}_hamming0_v0"] = h0.astype(np.uint8)
        a0, a1 = r.ad(0), r.ad(1)                     # gen_ad_sd_vm (AD, trunc 1000)
        out[f"{tag}_ad_v0"], out[f"{tag}_ad_v1"] = a0, a1
    r.set("censusFunc", 3)
    v0, v1 = r.adcensus()                             # ADCensusCal
    out[f"{tag}_adcensus_v0"], out[f"{tag}_adcensus_v1"] = v0, v1
    a0, a1 = r.arms()                                 # initArm + calArms<uchar> (calHorVerDis)
    out[f"{tag}_arms_L"], out[f"{tag}_arms_R"] = a0, a1
    if full:
        out[f"{tag}_isect_v0"] = r.arms_intersection(0).astype(np.uint8)   # genTrueHorVerArms
        out[f"{tag}_isect_v1"] = r.arms_intersection(1).astype(np.uint8)
        r1 = po.SmRef(bl, br, gl, gr, D)
        r1.set_vm(0, v0)
        r1.set_vm(1, v1)
        c1, _ = r1.cbca(1)                            # one iteration only
        out[f"{tag}_cbca1_v0"] = c1
        r1.close()
    if full:
        # the CBCA sub-steps one by one on view 0 (first iteration: H cumsum, H span, V cumsum, V span, divide)
        ones = np.ones((H, W, D), np.int32)
        sv, sa = r.gen1dcumu(v0, ones, 0, -1)         # gen1DCumu along u
        out[f"{tag}_part_cumH_vol"], out[f"{tag}_part_cumH_area"] = sv, sa.astype(np.int16)
        sv, sa = r.cal1dcost(0, sv, sa, 0, -1, 0)     # cal1DCost, horizontal span
        out[f"{tag}_part_spanH_vol"], out[f"{tag}_part_spanH_area"] = sv, sa.astype(np.int16)
        sv, sa = r.gen1dcumu(sv, sa, -1, 0)           # gen1DCumu along v
        sv, sa = r.cal1dcost(0, sv, sa, -1, 0, 1)     # cal1DCost, vertical span
        out[f"{tag}_part_spanV_vol"], out[f"{tag}_part_spanV_area"] = sv, sa.astype(np.int16)
        out[f"{tag}_part_final_vol"] = r.genfinal(sv, sa)   # genfinalVm_cbca  (== cbca1_v0)
    c0, c1 = r.cbca(2)                                # cbca_aggregate -> cbca_core, 2 iterations, both views
    out[f"{tag}_cbca2_v0"], out[f"{tag}_cbca2_v1"] = c0, c1
    if full:
        for path in range(8):                         # costScan / updateCost<float>
            out[f"{tag}_path{path}_v0"] = r.cost_scan(0, path)
        out[f"{tag}_path5_v1"] = r.cost_scan(1, 5)
        r4 = po.SmRef(bl, br, gl, gr, D)
        r4.set_vm(0, c0)
        out[f"{tag}_sgm4_v0"] = r4.sgm(0, 4)           # the reference's own sgm() (4 paths) + gen_sgm_vm
        r4.close()
    s0, s1 = r.sgm(0, 8), r.sgm(1, 8)
    out[f"{tag}_sgm8_v0"], out[f"{tag}_sgm8_v1"] = s0, s1
    d0, d1 = r.wta(0), r.wta(1)                       # gen_dispFromVm
    out[f"{tag}_wta_v0"], out[f"{tag}_wta_v1"] = d0, d1
    w1, w2 = r.wta_co(0)                              # wta_Co
    out[f"{tag}_wtaco_D1"], out[f"{tag}_wtaco_D2"] = w1, w2
    lrc = r.lrc_normal(d0, d1)                        # LRConsistencyCheck_normal
    out[f"{tag}_lrc"] = lrc
    out[f"{tag}_lrc_new_mask"] = r.lrc_new(d0, d1, np.full((H, W), 255, np.uint8))   # LRConsistencyCheck_new
    la, lb, lm = r.lrc_label(d0, d1, 0)               # LRConsistencyCheck (labelling form)
    out[f"{tag}_lrc_label"], out[f"{tag}_lrc_mask"] = la, lm
    rv1 = r.region_vote(lrc, 0.4, 20)                 # regionVote_my
    out[f"{tag}_vote1"] = rv1
    out[f"{tag}_vote_s3"] = r.region_vote(lrc, 0.4, 3)
    rv2 = r.region_vote(rv1, 0.4, 20)
    ip1 = r.proper_ipol(rv2)                          # properIpol
    out[f"{tag}_ipol1"] = ip1
    out[f"{tag}_ipol_labelled"] = r.proper_ipol(la)   # DISP_OCC / DISP_MIS branches
    r.close()
    r = po.SmRef(bl, br, gl, gr, D)                   # gradient cost family, fresh instance
    for i in (0, 1):                                  # calGrad / calGrad_y
        gx, gy = r.grad_xy(i)
        out[f"{tag}_gx{i}"], out[f"{tag}_gy{i}"] = gx.astype(np.float16), gy.astype(np.float16)   # multiples of 0.5, |.| <= 255
        assert np.array_equal(gx, out[f"{tag}_gx{i}"].astype(np.float32))
    g0, g1 = r.grad_vm(500.0)                         # grad() -> calgradvm
    out[f"{tag}_gradvm_v0"], out[f"{tag}_gradvm_v1"] = g0, g1
    v0, v1 = r.censusgrad()                           # censusGrad(vm)
    out[f"{tag}_censusgrad_v0"], out[f"{tag}_censusgrad_v1"] = v0, v1
    r.close()
    rcg = po.SmRef(bl, br, gl, gr, D)
    _, _, rf, _ = rcg.pipeline(8, 2, costcalc=1)      # main_.cpp:15's own selector through the whole chain
    out[f"{tag}_pipe8_censusgrad_refined"] = rf
    rcg.close()
    for levels in ((2, 3) if full else (3,)):         # main_.cpp:131-166 with PY_LEV = levels (pyrDown pyramid + SolveAll)
        rp = po.SmRef(bl, br, gl, gr, D)
        rf, s0, s1 = rp.pipeline_pyr(levels, 0.3, 8, 2, 0)
        out[f"{tag}_pyr{levels}_refined"] = rf
        if full:
            out[f"{tag}_pyr{levels}_vm0"] = s0          # vm[0] right after SolveAll
            if levels == 2:
                out[f"{tag}_pyr{levels}_vm1"] = s1
        rp.close()
    for paths in (4, 8):                              # whole default chain, fresh instance
        r = po.SmRef(bl, br, gl, gr, D)
        wl, wr, rf, _ = r.pipeline(paths, 2)
        out[f"{tag}_pipe{paths}_wtaL"], out[f"{tag}_pipe{paths}_wtaR"], out[f"{tag}_pipe{paths}_refined"] = wl, wr, rf
        r.close()


stages("a", 24, 40, 12, "texture_warped", 11, full=True)
stages("b", 16, 22, 32, "random_dot", 12, full=False)      # D > W: most (u, d) out of range
stages("c", 30, 52, 16, "random_dot", 13, full=False)
path = os.path.join(os.path.dirname(__file__), "sm_ref.npz")
np.savez_compressed(path, **out)
print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")
