// sm_io.h -- the file side of the reference's driver: what main_.cpp:25-129 does with
// cv::imread / convertTo before it constructs a StereoMatching, and what saveDispMap (stereoMatching.h:2005-2110)
// writes afterwards, without OpenCV: PNG (zlib), PGM / PPM, PFM.
//
// Pinned against cv2 4.13 (tests/golden/io_ref.npz + the PNG / PPM fixtures beside it, tests/test_host_io.py):
//   imread(path, 1)  -> BGR, 3 x 8 bit (gray files replicated, alpha dropped, 16-bit samples >> 8)
//   imread(path, 0)  -> gray 8 bit.  The conversion of a COLOUR file depends on the codec, as in OpenCV:
//        PNG        libpng's rgb_to_gray with (0.299, 0.587):  (9797 R + 19234 G + 3737 B) >> 15   (truncation)
//        PGM / PPM  imgcodecs' own BGR2GRAY:                   (4899 R + 9617 G + 1868 B + 8192) >> 14
//     (cv::cvtColor(BGR2GRAY) is a third formula, (9798 R + 19235 G + 3735 B + 16384) >> 15; sm_bgr2gray_mode has all three)
//   DT.convertTo(CV_32F, 1 / disp_reduceCoeff)                 main_.cpp:126-129
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace smio {

struct Image {          // row-major, interleaved channels
  int h = 0, w = 0, c = 0;
  std::vector<uint8_t> data;
  bool empty() const { return data.empty(); }
};
struct ImageF {
  int h = 0, w = 0;
  std::vector<float> data;
};

// ---- readers (false + *err on failure; out is left empty, like an empty cv::Mat)
bool imread_color(const std::string& path, Image& out, std::string* err = nullptr);   // cv::imread(path, 1)
bool imread_gray(const std::string& path, Image& out, std::string* err = nullptr);    // cv::imread(path, 0)
bool read_pfm(const std::string& path, ImageF& out, std::string* err = nullptr);      // Middlebury 2014 ground truth (bottom-up rows)
// ---- writers
bool write_png(const std::string& path, const uint8_t* data, int h, int w, int c, std::string* err = nullptr);  // c = 1 | 3 (BGR in memory)
bool write_pnm(const std::string& path, const uint8_t* data, int h, int w, int c, std::string* err = nullptr);  // P5 / P6
bool write_pfm(const std::string& path, const float* data, int h, int w, std::string* err = nullptr);

// ---- the dataset table of main_.cpp:31-39 (33 Middlebury objects) and the pair loader of main_.cpp:85-129
struct MiddleburyEntry {
  const char* object;
  const char* left;
  const char* right;
  const char* disp;
  float disp_reduceCoeff;   // ground-truth scale: DT = png / coeff
  int maxdisp;
};
extern const MiddleburyEntry kMiddlebury[33];
const MiddleburyEntry* middlebury_find(const std::string& object);

struct StereoPair {
  Image I1_c, I2_c, I1_g, I2_g;          // imread(.., 1) / imread(.., 0) of the left / right view
  Image all_mask, nonocc_mask, disc_mask;   // may be empty ("can't read mask img" is not fatal in the reference)
  ImageF DT;                              // ground truth / disp_reduceCoeff; empty if the file is missing
  int maxdisp = 0;
};
// imgroot = root + object + "/"; file names from the table with the given extension (".png" in the reference).
bool load_middlebury(const std::string& root, const std::string& object, StereoPair& out, std::string* err = nullptr,
                     const std::string& ext = ".png");

// ---- saveDispMap<short> (stereoMatching.h:2005-2110): the disparity map as an 8-bit BGR picture -- valid pixels scaled to
// [0, 255] over [min valid, max], DISP_OCC blue, DISP_MIS red, DISP_PKR yellow, err_ip_dispV magenta, cor_ip_dispV cyan
// (Parameters defaults -50 / -100, stereoMatching.h:316-317); with dt != nullptr and all_mask the pixels whose error
// exceeds 1 are painted red (the "_err.png" variant).
void disp_to_bgr(const int16_t* disp, int h, int w, int DISP_OCC, int DISP_MIS, int DISP_PKR, std::vector<uint8_t>& bgr,
                 const float* dt = nullptr, const uint8_t* all_mask = nullptr, int err_ip_dispV = -50, int cor_ip_dispV = -100);

}  // namespace smio
