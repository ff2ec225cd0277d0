// Test driver for the C++ stage API (mystereomatching_b200/host/stereoMatching.h).  Reads a raw stereo pair written
// by tests/test_cpp_host.py, drives the class the way the reference's main() does (main_.cpp:138-166), and writes
// raw outputs that the Python test compares with the CPU oracle.
//   stage_api_test <in.bin> <out_prefix> <H> <W> <D> <paths> <mode>
// in.bin = bgrL | bgrR | grayL | grayR (u8).  mode: "pipeline" | "stages" | "censusgrad" | "nl" | "errors"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>

#include "NL/NLCCA.h"
#include "NL/ctmf.h"
#include "NL/qx_nonlocal_cost_aggregation.h"
#include "NL/qx_tree_filter.h"
#include "stereoMatching.h"

static void dump(const std::string& path, const void* p, size_t n) {
  std::ofstream f(path, std::ios::binary);
  f.write((const char*)p, (std::streamsize)n);
}

int main(int argc, char** argv) {
  if (argc < 8) { fprintf(stderr, "usage\n"); return 2; }
  const std::string in = argv[1], out = argv[2], mode = argv[7];
  const int H = atoi(argv[3]), W = atoi(argv[4]), D = atoi(argv[5]), paths = atoi(argv[6]);
  const size_t npix = (size_t)H * W;
  std::vector<unsigned char> buf(npix * 8);
  { std::ifstream f(in, std::ios::binary); f.read((char*)buf.data(), (std::streamsize)buf.size()); if (!f) { fprintf(stderr, "short input\n"); return 2; } }
  cv::Mat I1c(H, W, CV_8UC3, buf.data()), I2c(H, W, CV_8UC3, buf.data() + npix * 3);
  cv::Mat I1g(H, W, CV_8UC1, buf.data() + npix * 6), I2g(H, W, CV_8UC1, buf.data() + npix * 7);
  cv::Mat DT, m0, m1, m2;
  try {
    if (mode == "errors") {
      // CV_Assert-style failures must surface as cv::Exception
      int caught = 0;
      try { StereoMatching::Parameters p(600, H, W, 13, 1, 2, 109, 10, "", 1); StereoMatching s(I1c, I2c, I1g, I2g, DT, m0, m1, m2, p); }
      catch (const cv::Exception&) { caught++; }
      try {
        StereoMatching::Parameters p(D - 1, H, W, 13, 1, 2, 109, 10, "", 1);
        StereoMatching s(I1c, I2c, I1g, I2g, DT, m0, m1, m2, p);
        cv::Mat lr, v = s.vm[0];
        s.costScan(lr, s.vm[0], 2, 2, true);   // not a direction of the table
      } catch (const cv::Exception&) { caught++; }
      printf("caught %d\n", caught);
      return caught == 2 ? 0 : 1;
    }
    if (mode == "nl") {
      // the NL/ surface on its own: ctmf, qx_tree_filter, NLCCA::aggreCV
      std::vector<unsigned char> med(npix * 3);
      ctmf(I1c.data, med.data(), W, H, W * 3, W * 3, 1, 3, (unsigned long)npix * 3);
      dump(out + ".ctmf.u8", med.data(), med.size());
      qx_tree_filter tf;
      tf.init(H, W, 3, 0.1, 4);
      tf.build_tree(I1c.data);
      dump(out + ".rank.i32", tf.get_rank(), npix * 4);
      dump(out + ".parent.i32", tf.parent().data(), npix * 4);
      std::vector<double> cost(npix * D), backup(npix * D);
      for (size_t i = 0; i < cost.size(); i++) cost[i] = (double)((i * 2654435761u) % 1000) / 1000.0;
      tf.filter(cost.data(), backup.data(), D);
      dump(out + ".tf.f64", cost.data(), cost.size() * 8);
      int sz[3] = {H, W, D};
      cv::Mat vol(3, sz, CV_32FC1);
      for (size_t i = 0; i < npix * D; i++) vol.ptr<float>()[i] = (float)((i * 2654435761u) % 1000) / 1000.0f;
      NLCCA nl;
      nl.aggreCV(I1c, I2c, D, vol);
      dump(out + ".aggre.f32", vol.data, npix * D * 4);
      // Yang's driver class, used as in his own main(): init, matching_cost, disparity
      unsigned char*** l3 = qx_allocu_3(H, W, 3);
      unsigned char*** r3 = qx_allocu_3(H, W, 3);
      memcpy(l3[0][0], I1c.data, npix * 3);
      memcpy(r3[0][0], I2c.data, npix * 3);
      unsigned char** disp = qx_allocu(H, W);
      qx_nonlocal_cost_aggregation nlca;
      nlca.init(H, W, D);
      nlca.matching_cost(l3, r3);
      std::vector<double> cv(npix * D);
      nlca.get_cost_volume(cv.data());
      dump(out + ".nlca_cost.f64", cv.data(), cv.size() * 8);
      nlca.disparity(disp, false);
      dump(out + ".nlca_disp.u8", disp[0], npix);
      nlca.disparity(disp, true);
      dump(out + ".nlca_disp_post.u8", disp[0], npix);
      qx_freeu(disp); qx_freeu_3(l3); qx_freeu_3(r3);
      return 0;
    }
    if (mode == "pyramid") {
      // the reference's main() flow with PY_LEV = 3 (main_.cpp:131-166): pyramid loop, SolveAll, dispOptimize, refine
      StereoMatching::costcalculation = "ADCensus";
      StereoMatching::aggregation = "CBCA";
      StereoMatching::optimization = "sgm";
      const int PY_LEV = 3;
      int maxDisp = D - 1, disSc = 1;
      cv::Mat c1 = I1c, c2 = I2c, g1 = I1g, g2 = I2g;
      StereoMatching** smPsy = new StereoMatching*[PY_LEV];
      for (int p = 0; p < PY_LEV; p++) {
        StereoMatching::Parameters param(maxDisp, c1.rows, c1.cols, 13, 1, 2, 109, 10, "", disSc);
        smPsy[p] = new StereoMatching(c1, c2, g1, g2, DT, m0, m1, m2, param);
        smPsy[p]->setSgmPaths(paths);
        smPsy[p]->costCalculate();
        maxDisp = maxDisp / 2 + 1;
        disSc *= 2;
        cv::Mat t;
        pyrDown_u8(c1, t); c1 = t; pyrDown_u8(c2, t); c2 = t; pyrDown_u8(g1, t); g1 = t; pyrDown_u8(g2, t); g2 = t;
      }
      SolveAll(smPsy, PY_LEV, 0.3f);
      smPsy[0]->dispOptimize();
      smPsy[0]->refine();
      smPsy[0]->syncToHost(false);
      dump(out + ".dp0.i16", smPsy[0]->DP[0].data, npix * 2);
      for (int p = 0; p < PY_LEV; p++) delete smPsy[p];
      delete[] smPsy;
      return 0;
    }
    // "censusgrad": the selectors main_.cpp:15-17 compiles in (censusGrad + CBCA + sgm)
    StereoMatching::costcalculation = mode == "censusgrad" ? "censusGrad" : "ADCensus";
    StereoMatching::aggregation = "CBCA";
    StereoMatching::optimization = "sgm";
    StereoMatching::Parameters param(D - 1, H, W, 13, 1, 2, 109, 10, "", 1);   // main_.cpp:60-64, 138
    StereoMatching sm(I1c, I2c, I1g, I2g, DT, m0, m1, m2, param);
    sm.setSgmPaths(paths);
    if (mode == "pipeline") {
      sm.pipeline();
      // calErr on the refined map against a synthetic truth (= the map itself shifted by one on a stripe), "all" mask
      sm.syncToHost(false);
      cv::Mat gtm(H, W, CV_32FC1), all(H, W, CV_8UC1, cv::Scalar::all(255));
      for (int v = 0; v < H; v++)
        for (int u = 0; u < W; u++) gtm.ptr<float>(v)[u] = (float)sm.DP[0].ptr<short>(v)[u] + (u % 7 == 0 ? 2.f : 0.f);
      sm.I_mask[1] = all;
      sm.DT = gtm;
      sm.calErr<short>(sm.DP[0], sm.DT, "final");
      float e[2] = {sm.lastErr[1].PBM, sm.lastErr[1].RMS};
      dump(out + ".err_all.f32", e, sizeof(e));
    } else if (mode == "censusgrad") {
      // the gradient family through its explicit-argument forms, then the whole chain with the reference's own selectors
      std::vector<cv::Mat> g(2), gy(2);
      for (int i = 0; i < 2; i++) { sm.calGrad(g[i], sm.I_g[i]); sm.calGrad_y(gy[i], sm.I_g[i]); }
      dump(out + ".gx0.f32", g[0].data, npix * 4);
      dump(out + ".gy1.f32", gy[1].data, npix * 4);
      cv::Mat gv;
      sm.calgradvm(gv, g, gy, 1, 500);
      dump(out + ".gradvm1.f32", gv.data, npix * D * 4);
      std::vector<cv::Mat> cg(2);
      sm.censusGrad(cg);
      dump(out + ".cg0.f32", cg[0].data, npix * D * 4);
      sm.pipeline();
    } else {   // "stages": the same chain, one public stage method at a time, through host-visible Mats where the API has them
      sm.ADCensusCal();
      sm.initArm();
      sm.calArms<uchar>(sm.I_c, sm.HVL, sm.HVL_INTERSECTION, 17, 34, 20, 6);
      sm.cbca_core(sm.HVL, sm.HVL_INTERSECTION, sm.vm, 2);
      dump(out + ".vm0_cbca.f32", sm.hostVm(0).data, npix * D * 4);
      for (int i = 0; i < 2; i++) sm.sgm(sm.vm[i], i == 0);
      {  // the vmTop candidate selection as dispOptimize() calls it (stereoMatching.cpp:1114-1119): on a clone of vm
        dump(out + ".vm0_sgm.f32", sm.hostVm(0).data, npix * D * 4);
        int szTop[] = {H, W, 6 + 1, 2};
        cv::Mat topDisp(4, szTop, CV_32F);
        cv::Mat vm_copy = sm.hostVm(0).clone();
        sm.selectTopCostFromVolumn(vm_copy, topDisp, 1.08f);
        dump(out + ".top0.f32", topDisp.data, npix * 7 * 2 * 4);
        dump(out + ".top0_vm.f32", vm_copy.data, npix * D * 4);
      }
      sm.DP[0].create(H, W, CV_16SC1); sm.DP[1].create(H, W, CV_16SC1);
      for (int i = 0; i < 2; i++) sm.gen_dispFromVm(sm.vm[i], sm.DP[i]);
      dump(out + ".dp0_wta.i16", sm.hostDP(0).data, npix * 2);
      {  // the default-off sub-pixel step as refine() would call it (stereoMatching.cpp:1482-1485)
        cv::Mat dsp = sm.hostDP(0).clone(), SE(H, W, CV_32F);
        sm.subpixelEnhancement(dsp, SE);
        dump(out + ".se0.f32", SE.data, npix * 4);
      }
      {  // the default-off discontinuity step as refine() would call it (stereoMatching.cpp:1473-1475), on a host clone
        cv::Mat dsp = sm.hostDP(0).clone();
        sm.discontinuityAdjust(dsp);
        dump(out + ".da0.i16", dsp.data, npix * 2);
      }
      {  // refine() with the sub-pixel step on (stereoMatching.cpp:1482-1490), split in two calls so that the map the
         // step reads (DP[0] before the last median) can be dumped; the end state equals a single refine()
        StereoMatching::Do_subpixelEnhancement = true;
        StereoMatching::Do_lastMedianBlur = false;
        sm.refine();
        dump(out + ".dp0_premed.i16", sm.hostDP(0).data, npix * 2);
        dump(out + ".se_refine.f32", sm.SE.data, npix * 4);
        StereoMatching::Do_subpixelEnhancement = false;
        StereoMatching::Do_LRConsis = StereoMatching::Do_regionVote = StereoMatching::Do_properIpol = false;
        StereoMatching::Do_lastMedianBlur = true;
        sm.refine();
        StereoMatching::Do_LRConsis = StereoMatching::Do_regionVote = StereoMatching::Do_properIpol = true;
      }
      // explicit-argument forms: host Mats in, host Mats out
      cv::Mat ad, lr;
      sm.gen_ad_sd_vm(ad, 0, 0, 1000);
      dump(out + ".ad0.f32", ad.data, npix * D * 4);
      std::vector<cv::Mat> cen(2);
      sm.genCensusCode_NC_Sur(sm.I_g, cen, 3, 4);
      dump(out + ".cenL.u64", cen[0].data, npix * 16);
      cv::Mat cv0;
      sm.gen_cenVM_XOR(cen, cv0, 71, 1.0f, 0);
      dump(out + ".cen0.f32", cv0.data, npix * D * 4);
      sm.costScan(lr, cv0, 0, -1, true);
      dump(out + ".lr3.f32", lr.data, npix * D * 4);
      // the sub-steps cbca_core composes, one public method at a time (first iteration, view 0), on host Mats
      {
        cv::Mat v0;
        sm.gen_ad_sd_vm(v0, 0, 0, 1000);            // any float volume serves; AD is at hand
        int sz[3] = {H, W, D};
        cv::Mat areaIS(3, sz, CV_32SC1, cv::Scalar::all(1)), none;
        std::vector<cv::Mat> hvl = sm.HVL, is(2);
        sm.syncToHost(false);
        sm.genTrueHorVerArms(sm.HVL, is);
        sm.gen1DCumu(v0, none, areaIS, 0, -1);
        sm.cal1DCost(v0, sm.HVL[0], none, areaIS, is[0], 0, -1, 0);
        sm.gen1DCumu(v0, none, areaIS, -1, 0);
        sm.cal1DCost(v0, sm.HVL[0], none, areaIS, is[0], -1, 0, 1);
        sm.genfinalVm_cbca(v0, none, areaIS, 0);
        dump(out + ".parts_cbca1.f32", v0.data, npix * D * 4);
        // updateCost<float> at one pixel of path 3 (rv = 0, ru = -1): recompute pixel (H/2, W/2) of lr
        cv::Mat lr2 = lr.clone();
        for (int d = 0; d < D; d++) lr2.ptr<float>(H / 2, W / 2)[d] = -1.f;
        sm.updateCost<float>(lr2, cv0, H / 2, W / 2, D, 0, -1, true, true);
        dump(out + ".lr3_pixel.f32", lr2.data, npix * D * 4);
        // costScan's integer-cost entry: the Hamming volume as CV_16U gives the same Lr as the float volume
        cv::Mat cv16(3, sz, CV_16UC1), lr16;
        for (size_t i = 0; i < npix * D; i++) cv16.ptr<ushort>()[i] = (ushort)cv0.ptr<float>()[i];
        sm.costScan(lr16, cv16, 0, -1, true);
        dump(out + ".lr3_u16.f32", lr16.data, npix * D * 4);
        cv::Mat mask(H, W, CV_8UC1, cv::Scalar::all(255));
        sm.LRConsistencyCheck_new(mask);
        dump(out + ".lrc_new.u8", mask.data, npix);
      }
    }
    sm.syncToHost(false);
    dump(out + ".dp0.i16", sm.DP[0].data, npix * 2);
    dump(out + ".hvl0.u16", sm.HVL[0].data, npix * 10);
  } catch (const cv::Exception& e) {
    fprintf(stderr, "cv::Exception: %s\n", e.what());
    return 1;
  } catch (const std::exception& e) {
    fprintf(stderr, "exception: %s\n", e.what());
    return 1;
  }
  return 0;
}
