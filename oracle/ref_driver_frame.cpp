// ref_driver_frame.cpp — OUR driver around the reference's own Src/Frame.cpp (compiled UNMODIFIED from /root/reference by
// oracle/Makefile, target _ref, against the library stand-ins of ref_shim/): Frame::CreateDirPyrs (Src/Frame.cpp:104-181),
// run by the reference's Frame constructor, on a caller-supplied float image.  The feature front-end's image pyramids
// (CreateIndPyrs, OpenCV) are no-ops in the stand-in and not on the path.  TEST INFRASTRUCTURE: tests/test_oracle_vs_reference.py
// compares every level with the oracle's pyramid, which is what the GPU pyramid kernel is held to bit for bit.
#include "CalibData.h"
#include "Detector.h"
#include "EnergyFunctional.h"
#include "Frame.h"
#include "ImmaturePoint.h"
#include "MapPoint.h"
#include "Settings.h"

using namespace FSLAM;
#define REF_HAVE_FRAME_CPP
#include "ref_stubs.inc"

namespace FSLAM {
// Src/Detector.cpp (the ORB front-end, OpenCV) is not on the path: Frame::Extract is never called here
void FeatureDetector::ExtractFeatures(cv::Mat&, std::vector<Vec3f*>&, int, std::vector<float*>&, std::vector<cv::KeyPoint>&, cv::Mat&, int&,
                                      int, std::shared_ptr<IndexThreadReduce<Vec10>>) { abort(); }
}  // namespace FSLAM

extern "C" {

// dI_out: per level w_l*h_l*3 floats {I, dx, dy}, levels concatenated; abs_out: per level w_l*h_l floats.
int ref_frame_create_dir_pyrs(const float* img, int w, int h, int levels, float* dI_out, float* abs_out) {
  DirPyrLevels = levels;     // Src/Settings.cpp:28
  IndPyrLevels = 1;          // the (stand-in) feature pyramid: one level
  Mat33 K = Mat33::Identity();
  K(0, 0) = w; K(1, 1) = w; K(0, 2) = 0.5 * w; K(1, 2) = 0.5 * h;
  // CalibData with ONE direct level (its constructor writes out of bounds for more, CalibData.h:163-166); the level sizes
  // CreateDirPyrs reads are appended the way the constructor's loop forms them (:155)
  int dirPyr = 1;
  std::shared_ptr<CalibData> calib(new CalibData(w, h, K, 0.f, nullptr, nullptr, dirPyr, 1, 1.2f));
  for (int i = 1; i < levels; ++i) { calib->wpyr.push_back(calib->wpyr[i - 1] / 2); calib->hpyr.push_back(calib->hpyr[i - 1] / 2); }
  std::shared_ptr<ImageData> im(new ImageData(w, h));
  im->timestamp = 0; im->ExposureL = im->ExposureR = 1.f;
  im->fImgL.assign(img, img + (size_t)w * h);
  Frame f(im, 0, 1.f, nullptr, calib, nullptr, false);   // runs CreateIndPyrs (no-op here) and CreateDirPyrs
  size_t o3 = 0, o1 = 0;
  for (int l = 0; l < levels; ++l) {
    const size_t n = (size_t)calib->wpyr[l] * calib->hpyr[l];
    for (size_t i = 0; i < n; ++i) {
      for (int c = 0; c < 3; ++c) dI_out[o3 + 3 * i + c] = f.DirPyr[l][i][c];
      abs_out[o1 + i] = f.absSquaredGrad[l][i];
    }
    o3 += 3 * n; o1 += n;
  }
  return 0;
}

}  // extern "C"
