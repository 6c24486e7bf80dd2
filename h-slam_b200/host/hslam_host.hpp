// hslam_host.hpp — C++ host mirror of the reference's hot-path classes above the C ABI.
//
// The reference (AUBVRL/H-SLAM, namespace FSLAM) reaches the path through member functions of
// Frame, CoarseTracker, EnergyFunctional and Accumulated*Hessian.  This header keeps
// those class and method names and forwards to include/hslam_b200.h, so a maintainer can swap the
// bodies under `#ifdef HSLAM_B200`.  It is header-only, C++11, and does NOT need
// Eigen/Sophus: small POD types stand in for SE3 / AffLight / MatXX; when HSLAM_WITH_EIGEN is defined,
// overloads taking the reference's own Eigen/Sophus types are provided as well.
//
// Error behaviour mirrors the reference: no exceptions on the tracking/BA calls themselves
// (trackNewestCoarse returns false, outputs stay untouched on abort); a CUDA failure throws
// hslam_b200::Error from the constructors / calls, because there is NO CPU fallback to continue on.
#pragma once
#include <array>
#include <cmath>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/hslam_b200.h"

namespace hslam_b200 {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
inline void check(int rc) { if (rc != 0) throw Error(std::string("hslam_b200: ") + hslam_last_error()); }

// ---- stand-ins for Sophus::SE3d / FSLAM::AffLight / Eigen dynamic matrices -------------------
struct SE3 {                      // T = [R | t], R row-major; same layout as the C ABI's 12 doubles
  double m[12];
  SE3() { for (int i = 0; i < 12; ++i) m[i] = 0; m[0] = m[4] = m[8] = 1; }
  const double* data() const { return m; }
  double* data() { return m; }
  SE3 operator*(const SE3& o) const {
    SE3 r;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) r.m[3 * i + j] = m[3 * i] * o.m[j] + m[3 * i + 1] * o.m[3 + j] + m[3 * i + 2] * o.m[6 + j];
      r.m[9 + i] = m[3 * i] * o.m[9] + m[3 * i + 1] * o.m[10] + m[3 * i + 2] * o.m[11] + m[9 + i];
    }
    return r;
  }
  SE3 inverse() const {
    SE3 r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[3 * i + j] = m[3 * j + i];
    for (int i = 0; i < 3; ++i) r.m[9 + i] = -(r.m[3 * i] * m[9] + r.m[3 * i + 1] * m[10] + r.m[3 * i + 2] * m[11]);
    return r;
  }
  // Sophus SE3::exp (Thirdparty/Sophus/sophus/se3.hpp:407-427), tangent = [upsilon, omega]
  static SE3 exp(const double* xi) {
    SE3 T;
    const double* om = xi + 3;
    double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2], th = std::sqrt(th2);
    double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0}, O2[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) O2[3 * i + j] = O[3 * i] * O[j] + O[3 * i + 1] * O[3 + j] + O[3 * i + 2] * O[6 + j];
    double a, b, c;
    if (th < 1e-10) { a = 1.0; b = 0.5; c = 1.0 / 6.0; }
    else { a = std::sin(th) / th; b = (1 - std::cos(th)) / th2; c = (th - std::sin(th)) / (th2 * th); }
    double V[9];
    for (int i = 0; i < 9; ++i) { T.m[i] = (i % 4 == 0) + a * O[i] + b * O2[i]; V[i] = (i % 4 == 0) + b * O[i] + c * O2[i]; }
    for (int i = 0; i < 3; ++i) T.m[9 + i] = V[3 * i] * xi[0] + V[3 * i + 1] * xi[1] + V[3 * i + 2] * xi[2];
    return T;
  }
  // Sophus SE3::Adj (se3.hpp:131-139), row-major 6x6
  void Adj(double* Ad) const {
    const double* t = m + 9;
    double Ht[9] = {0, -t[2], t[1], t[2], 0, -t[0], -t[1], t[0], 0};
    for (int i = 0; i < 36; ++i) Ad[i] = 0;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
      Ad[6 * i + j] = m[3 * i + j];
      Ad[6 * (i + 3) + j + 3] = m[3 * i + j];
      Ad[6 * i + j + 3] = Ht[3 * i] * m[j] + Ht[3 * i + 1] * m[3 + j] + Ht[3 * i + 2] * m[6 + j];
    }
  }
};

struct AffLight {                 // Include/GlobalTypes.h:326-352
  double a, b;
  AffLight(double a_ = 0, double b_ = 0) : a(a_), b(b_) {}
  static void fromToVecExposure(float eF, float eT, AffLight g2F, AffLight g2T, double out[2]) {
    if (eF == 0 || eT == 0) eF = eT = 1;
    out[0] = std::exp(g2T.a - g2F.a) * eT / eF;
    out[1] = g2T.b - out[0] * g2F.b;
  }
};

struct MatXX {                    // column-major dynamic matrix (Eigen default layout)
  int rows = 0, cols = 0;
  std::vector<double> v;
  void resize(int r, int c) { rows = r; cols = c; v.assign((size_t)r * c, 0.0); }
  double& operator()(int r, int c) { return v[(size_t)r + (size_t)rows * c]; }
  double operator()(int r, int c) const { return v[(size_t)r + (size_t)rows * c]; }
};
typedef std::vector<double> VecX;

// HCalib accessors the path reads (Include/CalibData.h:93-98)
struct CalibData {
  int Width = 0, Height = 0;
  float fx = 0, fy = 0, cx = 0, cy = 0;
  float fxl() const { return fx; } float fyl() const { return fy; } float cxl() const { return cx; } float cyl() const { return cy; }
  float fxli() const { return 1.0f / fx; } float fyli() const { return 1.0f / fy; }
};

// ---- context -----------------------------------------------------------------------------------
class Context {                   // replaces the IndexThreadReduce pools of System (Src/System.cpp:33-34)
 public:
  explicit Context(int device = 0, const hslam_params* p = nullptr) { check(hslam_create(p, device, &h_)); }
  ~Context() { hslam_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  hslam_ctx* get() const { return h_; }
 private:
  hslam_ctx* h_ = nullptr;
};

// ---- Frame: DirPyr lives on the device with the Frame's lifetime (Src/Frame.cpp:104-192) ------
class Frame {
 public:
  Frame(Context& ctx, int w, int h, int levels) : ctx_(ctx), w_(w), h_(h), levels_(levels) {}
  ~Frame() { if (f_) hslam_frame_destroy(f_); }
  Frame(const Frame&) = delete;
  Frame& operator=(const Frame&) = delete;
  // void Frame::CreateDirPyrs(std::vector<float>& Img, std::vector<Vec3f*>& DirPyr)
  void CreateDirPyrs(const std::vector<float>& Img, const float* gammaBGrad256 = nullptr) {
    if ((int)Img.size() != w_ * h_) throw Error("CreateDirPyrs: image size mismatch");
    if (!f_) check(hslam_frame_create(ctx_.get(), Img.data(), w_, h_, levels_, gammaBGrad256, &f_));
    else check(hslam_frame_update(f_, Img.data(), 0));
  }
  // host copies for code that still reads DirPyr[lvl] (Vec3f AoS) / absSquaredGrad[lvl]
  void downloadDirPyr(int lvl, std::vector<float>& dI_xyz, std::vector<float>& absSquaredGrad) const {
    int w, h;
    check(hslam_frame_level_size(f_, lvl, &w, &h));
    dI_xyz.resize((size_t)w * h * 3); absSquaredGrad.resize((size_t)w * h);
    check(hslam_frame_download(f_, lvl, dI_xyz.data(), absSquaredGrad.data()));
  }
  hslam_frame* get() const { return f_; }
  float ab_exposure = 1.0f;
  float frameEnergyTH = 8 * 8 * HSLAM_PATTERN_NUM;   // Src/Frame.cpp:47
 private:
  Context& ctx_;
  int w_, h_, levels_;
  hslam_frame* f_ = nullptr;
};

// ---- CoarseTracker (Include/CoarseTracker.h:14-95) ---------------------------------------------
class CoarseTracker {
 public:
  CoarseTracker(Context& ctx, int w, int h) : ctx_(ctx) {
    check(hslam_tracker_create(ctx.get(), w, h, &t_));
    for (int i = 0; i < 5; ++i) lastResiduals[i] = NAN;
    for (int i = 0; i < 3; ++i) lastFlowIndicators[i] = 1000;
  }
  ~CoarseTracker() { hslam_tracker_destroy(t_); }
  CoarseTracker(const CoarseTracker&) = delete;
  CoarseTracker& operator=(const CoarseTracker&) = delete;

  // void makeK(std::shared_ptr<CalibData> HCalib)   (Src/CoarseTracker.cpp:72-101)
  void makeK(const CalibData& HCalib, int levels) {
    levels_ = levels;
    check(hslam_tracker_make_K(t_, levels, HCalib.Width, HCalib.Height, HCalib.fxl(), HCalib.fyl(), HCalib.cxl(), HCalib.cyl()));
  }
  // setCoarseTrackingRef: the per-level clouds are the output of makeCoarseDepthL0
  // (Src/CoarseTracker.cpp:105-263, still host code), plus lastRef's exposure and aff_g2l (:496-500)
  void setCoarseTrackingRef(int lvl, const float* pc_u, const float* pc_v, const float* pc_idepth, const float* pc_color, int pc_n) {
    check(hslam_tracker_set_ref(t_, lvl, pc_u, pc_v, pc_idepth, pc_color, pc_n));
  }
  void setRefFrame(float ab_exposure, AffLight aff_g2l, int frameID) {
    lastRef_aff_g2l = aff_g2l; refFrameID = frameID; firstCoarseRMSE = -1;
    check(hslam_tracker_set_ref_aff(t_, ab_exposure, aff_g2l.a, aff_g2l.b));
  }
  // bool trackNewestCoarse(shared_ptr<FrameShell>&, SE3&, AffLight&, int coarsestLvl, Vec5 minResForAbort)
  bool trackNewestCoarse(const Frame& newFrame, SE3& lastToNew_out, AffLight& aff_g2l_out, int coarsestLvl,
                         const double minResForAbort[5]) {
    double aff[2] = {aff_g2l_out.a, aff_g2l_out.b};
    int ok = 0;
    check(hslam_tracker_track(t_, newFrame.get(), newFrame.ab_exposure, lastToNew_out.data(), aff, coarsestLvl,
                              minResForAbort, lastResiduals, lastFlowIndicators, &ok, iterations));
    aff_g2l_out.a = aff[0]; aff_g2l_out.b = aff[1];
    return ok != 0;
  }
  // The motion hypotheses of System::trackNewCoarse (System.cpp:348-481) in ONE launch: every lastF_2_fh_tries[i] is
  // tracked by its own thread-block cluster; the caller then walks the results in try order with the reference's
  // acceptance rule.  ok / residuals / flow per hypothesis, each bit-identical to a single trackNewestCoarse call.
  void trackNewestCoarse_hypotheses(const Frame& newFrame, std::vector<SE3>& lastToNew_io, std::vector<AffLight>& aff_io, int coarsestLvl,
                                    const double minResForAbort[5], std::vector<int>& ok, std::vector<double>& residuals5,
                                    std::vector<double>& flow3) {
    const int n = (int)lastToNew_io.size();
    std::vector<const hslam_frame*> fr(n, newFrame.get());
    std::vector<float> ex(n, newFrame.ab_exposure);
    std::vector<double> T(12 * (size_t)n), a(2 * (size_t)n);
    for (int i = 0; i < n; ++i) { std::memcpy(&T[12 * i], lastToNew_io[i].data(), 12 * sizeof(double)); a[2 * i] = aff_io[i].a; a[2 * i + 1] = aff_io[i].b; }
    ok.assign(n, 0); residuals5.assign(5 * (size_t)n, 0.0); flow3.assign(3 * (size_t)n, 0.0);
    check(hslam_tracker_track_batch(t_, n, fr.data(), ex.data(), T.data(), a.data(), coarsestLvl, minResForAbort, residuals5.data(),
                                    flow3.data(), ok.data(), nullptr));
    for (int i = 0; i < n; ++i) { std::memcpy(lastToNew_io[i].data(), &T[12 * i], 12 * sizeof(double)); aff_io[i].a = a[2 * i]; aff_io[i].b = a[2 * i + 1]; }
  }
  // Vec6 calcRes(int lvl, const SE3& refToNew, AffLight aff_g2l, float cutoffTH)   (private in the reference)
  std::array<double, 6> calcRes(int lvl, const Frame& newFrame, const SE3& refToNew, AffLight aff_g2l, float cutoffTH) {
    std::array<double, 6> r;
    check(hslam_tracker_calc_res(t_, lvl, newFrame.get(), refToNew.data(), aff_g2l.a, aff_g2l.b, newFrame.ab_exposure, cutoffTH,
                                 r.data(), &buf_warped_n));
    lastCutoff_ = cutoffTH;
    return r;
  }
  // void calcGSSSE(int lvl, Mat88& H_out, Vec8& b_out, const SE3& refToNew, AffLight aff_g2l)
  void calcGSSSE(int lvl, const Frame& newFrame, double H_out[64], double b_out[8], const SE3& refToNew, AffLight aff_g2l) {
    check(hslam_tracker_calc_gs(t_, lvl, newFrame.get(), refToNew.data(), aff_g2l.a, aff_g2l.b, newFrame.ab_exposure, lastCutoff_,
                                H_out, b_out));
  }
  // pure outputs, as in the reference
  double lastResiduals[5];
  double lastFlowIndicators[3];
  double firstCoarseRMSE = -1;
  int iterations[5] = {0, 0, 0, 0, 0};
  AffLight lastRef_aff_g2l;
  int refFrameID = -1;
  int buf_warped_n = 0;
 private:
  Context& ctx_;
  hslam_tracker* t_ = nullptr;
  int levels_ = 0;
  float lastCutoff_ = 20.0f;
};

// ---- window BA: EnergyFunctional + residual linearisation --------------------------------------
// Flat records the host keeps per point / residual (MapPoint, PointFrameResidual of the reference).
struct PointRec { float u, v, idepth, idepth_zero, priorF; float color[8], weights[8]; int hostIdx; std::vector<int> residuals; };
struct ResidualRec { int point, targetIdx; int8_t state_state = HSLAM_RES_IN; bool isLinearized = false, isActive = false; };
struct FrameState {               // FrameOptimizationData, Include/Frame.h:116-275
  SE3 worldToCam_evalPT;
  double state[10] = {0}, state_zero[10] = {0};
  double prior[8] = {0}, delta_prior[8] = {0};
  const Frame* frame = nullptr;
};

class EnergyFunctional {
 public:
  explicit EnergyFunctional(Context& ctx) : ctx_(ctx) { check(hslam_ba_create(ctx.get(), &ba_)); hslam_params_default(&prm_); }
  ~EnergyFunctional() { hslam_ba_destroy(ba_); }
  EnergyFunctional(const EnergyFunctional&) = delete;
  EnergyFunctional& operator=(const EnergyFunctional&) = delete;

  std::vector<FrameState> frames;
  std::vector<PointRec> allPoints;
  std::vector<ResidualRec> residuals;
  int nFrames = 0, nPoints = 0, nResiduals = 0;
  int resInA = 0, resInL = 0;
  double cPrior[4] = {5e9, 5e9, 5e9, 5e9};   // setting_initialCalibHessian (EnergyFunctional.cpp:63)
  float cDeltaF[4] = {0, 0, 0, 0};
  std::vector<double> adHost, adTarget;       // n*n*64 col-major
  std::vector<float> adHTdeltaF, precalc;

  // setAdjointsF (EnergyFunctional.cpp:22-82) + setDeltaF (:128-152) + FrameFramePrecalc::set
  // (OptimizationClasses.cpp:13-39) for every ordered pair; uploads the tables.
  void setAdjointsF_setDeltaF_setPrecalc(const CalibData& HCalib) {
    const int n = nFrames = (int)frames.size();
    adHost.assign((size_t)n * n * 64, 0); adTarget.assign((size_t)n * n * 64, 0);
    adHTdeltaF.assign((size_t)n * n * 8, 0); precalc.assign((size_t)n * n * HSLAM_PRECALC_STRIDE, 0);
    std::vector<SE3> PRE_w2c(n);
    std::vector<AffLight> aff(n), aff0(n);
    for (int f = 0; f < n; ++f) {          // FrameOptimizationData::setState, Frame.h:184-197
      const double* s = frames[f].state;
      double sc[6] = {prm_.scale_xi_trans * s[0], prm_.scale_xi_trans * s[1], prm_.scale_xi_trans * s[2],
                      prm_.scale_xi_rot * s[3], prm_.scale_xi_rot * s[4], prm_.scale_xi_rot * s[5]};
      PRE_w2c[f] = SE3::exp(sc) * frames[f].worldToCam_evalPT;
      aff[f] = AffLight(prm_.scale_a * s[6], prm_.scale_b * s[7]);
      aff0[f] = AffLight(frames[f].state_zero[6] * prm_.scale_a, frames[f].state_zero[7] * prm_.scale_b);
    }
    const float K[9] = {HCalib.fx, 0, HCalib.cx, 0, HCalib.fy, HCalib.cy, 0, 0, 1};
    float Ki[9];
    inv3(K, Ki);
    const double rs[8] = {prm_.scale_xi_trans, prm_.scale_xi_trans, prm_.scale_xi_trans, prm_.scale_xi_rot,
                          prm_.scale_xi_rot, prm_.scale_xi_rot, prm_.scale_a, prm_.scale_b};
    for (int h = 0; h < n; ++h)
      for (int t = 0; t < n; ++t) {
        const int idx = h + t * n;
        SE3 ll0 = frames[t].worldToCam_evalPT * frames[h].worldToCam_evalPT.inverse();
        SE3 ll = PRE_w2c[t] * PRE_w2c[h].inverse();
        float R[9], tt[3], KR[9], KRKi[9];
        for (int i = 0; i < 9; ++i) R[i] = (float)ll.m[i];
        for (int i = 0; i < 3; ++i) tt[i] = (float)ll.m[9 + i];
        mul3(K, R, KR); mul3(KR, Ki, KRKi);
        float* pc = &precalc[(size_t)HSLAM_PRECALC_STRIDE * idx];
        for (int i = 0; i < 9; ++i) pc[i] = KRKi[i];
        for (int i = 0; i < 3; ++i) pc[9 + i] = dot3(K[3 * i], K[3 * i + 1], K[3 * i + 2], tt[0], tt[1], tt[2]);
        for (int i = 0; i < 9; ++i) pc[12 + i] = (float)ll0.m[i];
        for (int i = 0; i < 3; ++i) pc[21 + i] = (float)ll0.m[9 + i];
        double a2[2], a02[2];
        const float eh = frames[h].frame ? frames[h].frame->ab_exposure : 1.f, et = frames[t].frame ? frames[t].frame->ab_exposure : 1.f;
        AffLight::fromToVecExposure(eh, et, aff[h], aff[t], a2);
        pc[24] = (float)a2[0]; pc[25] = (float)a2[1]; pc[26] = (float)aff0[h].b;
        double Ad[36], AH[64] = {0}, AT[64] = {0};
        ll0.Adj(Ad);
        for (int i = 0; i < 8; ++i) AH[9 * i] = AT[9 * i] = 1;
        for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) AH[8 * r + c] = -Ad[6 * c + r];
        AffLight::fromToVecExposure(eh, et, aff0[h], aff0[t], a02);
        const float affLL0 = (float)a02[0];
        AT[8 * 6 + 6] = -affLL0; AH[8 * 6 + 6] = affLL0; AT[8 * 7 + 7] = -1; AH[8 * 7 + 7] = affLL0;
        for (int r = 0; r < 8; ++r) for (int c = 0; c < 8; ++c) {
          AH[8 * r + c] *= rs[r]; AT[8 * r + c] *= rs[r];
          adHost[64 * (size_t)idx + r + 8 * c] = AH[8 * r + c];
          adTarget[64 * (size_t)idx + r + 8 * c] = AT[8 * r + c];
        }
        for (int c = 0; c < 8; ++c) {
          float a = 0, b = 0;
          for (int r = 0; r < 8; ++r) {
            a += (float)(frames[h].state[r] - frames[h].state_zero[r]) * (float)AH[8 * r + c];
            b += (float)(frames[t].state[r] - frames[t].state_zero[r]) * (float)AT[8 * r + c];
          }
          adHTdeltaF[8 * (size_t)idx + c] = a + b;
        }
      }
    upload_window(HCalib);
  }

  // makeIDX (EnergyFunctional.cpp:819-840): re-pack the point/residual graph after a topology change
  void makeIDX() {
    const int P = nPoints = (int)allPoints.size(), R = nResiduals = (int)residuals.size();
    std::vector<float> u(P), v(P), id(P), idz(P), prior(P), delta(P), col(8 * (size_t)P), wgt(8 * (size_t)P);
    std::vector<int32_t> host(P), rb(P + 1), ri, rp(R), rt(R);
    std::vector<int8_t> st(R);
    std::vector<uint8_t> lin(R), act(R);
    for (int p = 0; p < P; ++p) {
      const PointRec& q = allPoints[p];
      u[p] = q.u; v[p] = q.v; id[p] = q.idepth; idz[p] = q.idepth_zero; prior[p] = q.priorF; delta[p] = q.idepth - q.idepth_zero;
      std::memcpy(&col[8 * (size_t)p], q.color, 32); std::memcpy(&wgt[8 * (size_t)p], q.weights, 32);
      host[p] = q.hostIdx; rb[p] = (int)ri.size();
      for (int r : q.residuals) ri.push_back(r);
    }
    rb[P] = (int)ri.size();
    for (int r = 0; r < R; ++r) {
      rp[r] = residuals[r].point; rt[r] = residuals[r].targetIdx; st[r] = residuals[r].state_state;
      lin[r] = residuals[r].isLinearized; act[r] = residuals[r].isActive;
    }
    hslam_point_soa ps = {P, u.data(), v.data(), id.data(), idz.data(), col.data(), wgt.data(), prior.data(), delta.data(),
                          host.data(), rb.data(), ri.data()};
    hslam_residual_soa rs = {R, rp.data(), rt.data(), st.data(), lin.data(), act.data(), nullptr};
    check(hslam_ba_set_topology(ba_, &ps, &rs));
  }

  // System::linearizeAll(fixLinearization) (FullSystemOptimize.cpp:103-165): returns lastEnergyP
  double linearizeAll(bool fixLinearization, std::vector<int8_t>* newState = nullptr, std::vector<float>* newEnergyWithOutlier = nullptr) {
    double e = 0;
    if (newState) newState->resize(nResiduals);
    if (newEnergyWithOutlier) newEnergyWithOutlier->resize(nResiduals);
    check(hslam_ba_linearize(ba_, fixLinearization, &e, newState ? newState->data() : nullptr, nullptr,
                             newEnergyWithOutlier ? newEnergyWithOutlier->data() : nullptr, nullptr, nullptr));
    return e;
  }
  void applyRes_Reductor() { check(hslam_ba_apply_res(ba_)); }
  // void System::setNewFrameEnergyTH() (FullSystemOptimize.cpp:60-101; System::linearizeAll calls it after every pass, :124):
  // exact nth_element on the device over the residuals into the newest keyframe; the frame's threshold is updated like
  // newFrame->frame->frameEnergyTH, the per-residual energies never reach the host
  float setNewFrameEnergyTH() {
    float th = 0.f;
    check(hslam_ba_new_frame_energy_th(ba_, nFrames - 1, &th));
    const_cast<Frame*>(frames[nFrames - 1].frame)->frameEnergyTH = th;   // the reference writes newFrame->frame->frameEnergyTH
    return th;
  }
  // GN-loop Jacobian mode (hslam_ba_set_jacobian_mode): per-pixel rows + projectedTo only when a residual is fixed
  void setJacobianMode(int mode) { check(hslam_ba_set_jacobian_mode(ba_, mode)); }

  // void accumulateAF_MT(MatXX& H, VecX& b, bool MT) etc. (EnergyFunctional.cpp:155-220); MT is ignored:
  // the CUDA grid replaces IndexThreadReduce.
  void accumulateAF_MT(MatXX& H, VecX& b, bool /*MT*/) { acc(HSLAM_ACC_A, H, b, &resInA); }
  void accumulateLF_MT(MatXX& H, VecX& b, bool /*MT*/) { acc(HSLAM_ACC_L, H, b, &resInL); }
  void accumulateSCF_MT(MatXX& H, VecX& b, bool /*MT*/) { acc(HSLAM_ACC_SC, H, b, nullptr); }
  // the three of them in one pass (what solveSystemF needs, EnergyFunctional.cpp:717-724)
  void accumulateALL(MatXX& HA, VecX& bA, MatXX& HL, VecX& bL, MatXX& Hsc, VecX& bsc) {
    const int N = 8 * nFrames + 4;
    std::vector<double> out(3 * ((size_t)N * N + N));
    int nres[2];
    check(hslam_ba_accumulate_all(ba_, out.data(), nres));
    resInA = nres[0]; resInL = nres[1];
    MatXX* Hs[3] = {&HA, &HL, &Hsc}; VecX* bs[3] = {&bA, &bL, &bsc};
    for (int s = 0; s < 3; ++s) {
      Hs[s]->resize(N, N); bs[s]->assign(N, 0.0);
      const double* src = out.data() + s * ((size_t)N * N + N);
      std::memcpy(Hs[s]->v.data(), src, sizeof(double) * N * N);
      std::memcpy(bs[s]->data(), src + (size_t)N * N, sizeof(double) * N);
    }
  }
  // void resubstituteF_MT(VecX x, shared_ptr<CalibData>& HCalib, bool MT)   (EnergyFunctional.cpp:222-247)
  void resubstituteF_MT(const VecX& x, std::vector<float>& point_steps) {
    point_steps.resize(nPoints);
    check(hslam_ba_resubstitute(ba_, x.data(), point_steps.data()));
  }
  // side effects of addPoint on efPoint (HdiF, bdSumF) and MapPoint::idepth_hessian
  void pointResults(std::vector<float>& HdiF, std::vector<float>& bdSumF, std::vector<float>& idepth_hessian) {
    HdiF.resize(nPoints); bdSumF.resize(nPoints); idepth_hessian.resize(nPoints);
    check(hslam_ba_point_results(ba_, HdiF.data(), bdSumF.data(), idepth_hessian.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
  }
  // void EnergyFunctional::marginalizePointsF()   (EnergyFunctional.cpp:545-609): the accumulate + stitch part for
  // the points whose stateFlag is toMarg; the caller keeps H = M - Msc, b = Mb - Mbsc, HM += ..., bM += ...
  void marginalizePointsF(const std::vector<int32_t>& pointsToMarg, float idepthFixPriorMargFac, MatXX& M, VecX& Mb, MatXX& Msc,
                          VecX& Mbsc, int& resInM_increment) {
    const int N = 8 * nFrames + 4;
    M.resize(N, N); Msc.resize(N, N); Mb.assign(N, 0.0); Mbsc.assign(N, 0.0);
    check(hslam_ba_marginalize_points(ba_, pointsToMarg.data(), (int)pointsToMarg.size(), idepthFixPriorMargFac, M.v.data(),
                                      Mb.data(), Msc.v.data(), Mbsc.data(), &resInM_increment));
  }
  // double EnergyFunctional::calcLEnergyF_MT()   (EnergyFunctional.cpp:349-364)
  double calcLEnergyF_MT() {
    double e = 0;
    check(hslam_ba_calc_l_energy(ba_, &e));
    return e;
  }
  hslam_ba* get() const { return ba_; }

 private:
  static float dot3(float m0, float m1, float m2, float x, float y, float z) {
    volatile float a = m0 * x, b = m1 * y, c = m2 * z; volatile float bc = b + c; return a + bc;   // a0 + (a1 + a2)
  }
  static void mul3(const float* A, const float* B, float* C) {
    float t[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) t[3 * i + j] = dot3(A[3 * i], A[3 * i + 1], A[3 * i + 2], B[j], B[3 + j], B[6 + j]);
    std::memcpy(C, t, sizeof(t));
  }
  static float cof(const float* m, int i, int j) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    volatile float a = m[3 * i1 + j1] * m[3 * i2 + j2], b = m[3 * i1 + j2] * m[3 * i2 + j1];
    return a - b;
  }
  static void inv3(const float* m, float* r) {   // Eigen compute_inverse_size3 (cofactors)
    float c0 = cof(m, 0, 0), c1 = cof(m, 1, 0), c2 = cof(m, 2, 0);
    float det = dot3(c0, c1, c2, m[0], m[3], m[6]);
    float inv = 1.0f / det;
    r[0] = c0 * inv; r[1] = c1 * inv; r[2] = c2 * inv;
    r[3] = cof(m, 0, 1) * inv; r[4] = cof(m, 1, 1) * inv; r[5] = cof(m, 2, 1) * inv;
    r[6] = cof(m, 0, 2) * inv; r[7] = cof(m, 1, 2) * inv; r[8] = cof(m, 2, 2) * inv;
  }
  void upload_window(const CalibData& HCalib) {
    const int n = nFrames;
    std::vector<const hslam_frame*> fr(n);
    std::vector<float> eth(n);
    std::vector<double> prior(8 * (size_t)n), dprior(8 * (size_t)n);
    for (int f = 0; f < n; ++f) {
      if (!frames[f].frame) throw Error("EnergyFunctional: frame without device pyramid");
      fr[f] = frames[f].frame->get(); eth[f] = frames[f].frame->frameEnergyTH;
      std::memcpy(&prior[8 * (size_t)f], frames[f].prior, 64); std::memcpy(&dprior[8 * (size_t)f], frames[f].delta_prior, 64);
    }
    hslam_window w;
    w.n_frames = n; w.w = HCalib.Width; w.h = HCalib.Height;
    w.fx = HCalib.fxl(); w.fy = HCalib.fyl(); w.cx = HCalib.cxl(); w.cy = HCalib.cyl(); w.fxi = HCalib.fxli(); w.fyi = HCalib.fyli();
    w.frames = fr.data(); w.frameEnergyTH = eth.data(); w.precalc = precalc.data(); w.adHost = adHost.data();
    w.adTarget = adTarget.data(); w.adHTdeltaF = adHTdeltaF.data();
    for (int i = 0; i < 4; ++i) { w.cDeltaF[i] = cDeltaF[i]; w.cPrior[i] = cPrior[i]; }
    w.framePrior = prior.data(); w.frameDeltaPrior = dprior.data();
    check(hslam_ba_set_window(ba_, &w));
  }
  void acc(int which, MatXX& H, VecX& b, int* nres) {
    const int N = 8 * nFrames + 4;
    H.resize(N, N); b.assign(N, 0.0);
    check(hslam_ba_accumulate(ba_, which, H.v.data(), b.data(), nres));
  }
  Context& ctx_;
  hslam_ba* ba_ = nullptr;
  hslam_params prm_;
};

// void System::traceNewCoarse(std::shared_ptr<FrameShell> fh)   (Src/Mapping.cpp:494-538): ImmaturePoint::traceOn for every
// immature point of every keyframe against the new frame.  `pts` holds the ImmaturePoint members as SoA; the per-host
// tables are K*R*K^-1, K*t and fromToVecExposure(host, fh) exactly as traceNewCoarse forms them.
inline void traceNewCoarse(Context& ctx, const Frame& fh, int n_hosts, const float* KRKi, const float* Kt, const float* aff,
                           const hslam_immature_soa& pts) {
  check(hslam_trace_immature(ctx.get(), fh.get(), nullptr, n_hosts, KRKi, Kt, aff, &pts));
}

// std::shared_ptr<MapPoint> System::optimizeImmaturePoint(...)   (Src/FullSystemOptPoint.cpp:24-175) for a list of candidates.
inline void optimizeImmaturePoints(Context& ctx, const std::vector<const Frame*>& frameHessians, const CalibData& HCalib,
                                   const float* RtAff, float minIdepthH_act, int GNItsOnPointActivation, int minObs,
                                   const hslam_activate_soa& pts) {
  std::vector<const hslam_frame*> fr;
  for (const Frame* f : frameHessians) fr.push_back(f->get());
  check(hslam_activate_immature(ctx.get(), (int)fr.size(), fr.data(), HCalib.fxl(), HCalib.fyl(), HCalib.cxl(), HCalib.cyl(), HCalib.fxli(),
                                HCalib.fyli(), RtAff, minIdepthH_act, GNItsOnPointActivation, minObs, &pts));
}

}  // namespace hslam_b200
