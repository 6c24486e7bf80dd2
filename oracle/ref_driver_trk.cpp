// ref_driver_trk.cpp — C interface around the UNMODIFIED reference translation units of Stage T and of the immature-point
// tracing / activation: Src/CoarseTracker.cpp (makeK, calcRes, calcGSSSE + Accumulator9, trackNewestCoarse), Src/ImmaturePoint.cpp
// (the constructor, traceOn, linearizeResidual) and Src/FullSystemOptPoint.cpp (System::optimizeImmaturePoint), compiled from /root/reference against the stand-ins of oracle/ref_shim/ into
// oracle/_ref/libhslam_ref_trk.so (oracle/Makefile).  TEST INFRASTRUCTURE, our code: it builds the objects those member
// functions work on from the same flat inputs the oracle takes and returns flat outputs; tests/test_oracle_vs_reference.py
// compares them with the oracle's float restatement.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>
#define private public      // our driver fills the tracker's reference clouds, calls calcRes / calcGSSSE directly and sets up
#include "CoarseTracker.h"  // a bare System object (its CalibData member is private) for System::optimizeImmaturePoint
#include "System.h"
#undef private
#include "CalibData.h"
#include "EnergyFunctional.h"
#include "Frame.h"
#include "ImmaturePoint.h"
#include "MapPoint.h"

#include "hslam_oracle.h"

using namespace FSLAM;
#define REF_HAVE_IMMATUREPOINT_CPP
#include "ref_stubs.inc"

struct ref_trk {
  std::shared_ptr<CalibData> calib;
  std::unique_ptr<CoarseTracker> trk;
  std::shared_ptr<FrameShell> ref, cur;
  int levels = 0, w = 0, h = 0;
};

static std::shared_ptr<FrameShell> make_shell(std::shared_ptr<CalibData>& calib, float exposure, int levels) {
  std::shared_ptr<ImageData> img(new ImageData(1, 1));
  img->timestamp = 0; img->ExposureL = exposure; img->ExposureR = exposure;
  std::shared_ptr<FrameShell> fs(new FrameShell(img, nullptr, calib, nullptr, false));
  fs->frame->DirPyr.assign(levels, nullptr);
  fs->frame->efFrame.reset(new FrameOptimizationData(0, exposure));
  return fs;
}

static SE3 se3_from(const double T[12]) {
  Eigen::Matrix<double, 3, 3> R; Eigen::Matrix<double, 3, 1> t;
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R(i, j) = T[3 * i + j]; t[i] = T[9 + i]; }
  return SE3(R, t);
}
static void se3_to(const SE3& S, double T[12]) {
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T[3 * i + j] = S.rotationMatrix()(i, j); T[9 + i] = S.translation()[i]; }
}

extern "C" {

ref_trk* ref_trk_create(int w, int h, int levels, float fx, float fy, float cx, float cy) {
  ref_trk* t = new ref_trk();
  t->w = w; t->h = h;
  DirPyrLevels = levels;                        // Src/Settings.cpp:28 (the constructor and makeK loop over it)
  Mat33 K = Mat33::Identity();
  K(0, 0) = fx; K(1, 1) = fy; K(0, 2) = cx; K(1, 2) = cy;
  // CalibData with ONE direct pyramid level: its constructor writes pyrfxi[i] .. pyrcyi[i] of one-element vectors for
  // i >= 1 (CalibData.h:163-166, an out-of-bounds write the survey notes); CoarseTracker::makeK derives the per-level
  // intrinsics itself from Width / Height / fxl() .. cyl() and the global DirPyrLevels
  int dirPyr = 1;
  t->calib.reset(new CalibData(w, h, K, 0.f, nullptr, nullptr, dirPyr, 1, 1.2f));
  t->levels = levels;
  DirPyrLevels = levels;
  t->trk.reset(new CoarseTracker(w, h));
  t->trk->makeK(t->calib);
  t->ref = make_shell(t->calib, 1.f, levels);
  t->cur = make_shell(t->calib, 1.f, levels);
  t->trk->lastRef = t->ref;
  t->trk->newFrame = t->cur;
  return t;
}
void ref_trk_destroy(ref_trk* t) { delete t; }
int ref_trk_levels(ref_trk* t) { return t->levels; }

int ref_trk_get_K(ref_trk* t, int lvl, float fxfycxcy[4], int wh[2], float K9[9], float Ki9[9]) {
  CoarseTracker& c = *t->trk;
  fxfycxcy[0] = c.fx[lvl]; fxfycxcy[1] = c.fy[lvl]; fxfycxcy[2] = c.cx[lvl]; fxfycxcy[3] = c.cy[lvl];
  wh[0] = c.w[lvl]; wh[1] = c.h[lvl];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { K9[3 * i + j] = c.K[lvl](i, j); Ki9[3 * i + j] = c.Ki[lvl](i, j); }
  return 0;
}

int ref_trk_set_ref(ref_trk* t, int lvl, const float* u, const float* v, const float* idepth, const float* color, int n) {
  CoarseTracker& c = *t->trk;
  memcpy(c.pc_u[lvl], u, sizeof(float) * n); memcpy(c.pc_v[lvl], v, sizeof(float) * n);
  memcpy(c.pc_idepth[lvl], idepth, sizeof(float) * n); memcpy(c.pc_color[lvl], color, sizeof(float) * n);
  c.pc_n[lvl] = n;
  return 0;
}

// dI_new: the whole pyramid in the oracle's layout (levels concatenated, Vec3f AoS)
int ref_trk_set_frames(ref_trk* t, float ref_exposure, float new_exposure, double ref_aff_a, double ref_aff_b, const float* dI_new) {
  t->ref->ab_exposure = ref_exposure; t->cur->ab_exposure = new_exposure;
  t->trk->lastRef_aff_g2l = AffLight(ref_aff_a, ref_aff_b);
  for (int l = 0; l < t->levels; ++l)
    t->cur->frame->DirPyr[l] = reinterpret_cast<Vec3f*>(const_cast<float*>(dI_new) + 3 * orc_pyramid_offset(t->w, t->h, l));
  return 0;
}

int ref_trk_calc_res(ref_trk* t, int lvl, const double T[12], double aff_a, double aff_b, float cutoffTH, double out6[6],
                     float* const* warped /*8 arrays, nullable*/, int* n_warped) {
  CoarseTracker& c = *t->trk;
  Vec6 r = c.calcRes(lvl, se3_from(T), AffLight(aff_a, aff_b), cutoffTH);
  for (int i = 0; i < 6; ++i) out6[i] = r[i];
  if (n_warped) *n_warped = c.buf_warped_n;     // padded to a multiple of 4 by calcRes
  if (warped) {
    const float* src[8] = {c.buf_warped_idepth, c.buf_warped_u, c.buf_warped_v, c.buf_warped_dx, c.buf_warped_dy,
                           c.buf_warped_residual, c.buf_warped_weight, c.buf_warped_refColor};
    for (int k = 0; k < 8; ++k) memcpy(warped[k], src[k], sizeof(float) * c.buf_warped_n);
  }
  return 0;
}

// on the warped buffers of the last ref_trk_calc_res at this level
int ref_trk_calc_gs(ref_trk* t, int lvl, const double T[12], double aff_a, double aff_b, double H[64], double b[8]) {
  Mat88 Hm; Vec8 bm;
  t->trk->calcGSSSE(lvl, Hm, bm, se3_from(T), AffLight(aff_a, aff_b));
  for (int j = 0; j < 8; ++j) { for (int i = 0; i < 8; ++i) H[i + 8 * j] = Hm(i, j); b[j] = bm[j]; }
  return 0;
}

int ref_trk_track(ref_trk* t, double T_io[12], double aff_io[2], int coarsestLvl, const double minResForAbort[5],
                  double lastResiduals[5], double lastFlow[3], int* ok) {
  SE3 T = se3_from(T_io);
  AffLight aff(aff_io[0], aff_io[1]);
  Vec5 mra;
  for (int i = 0; i < 5; ++i) mra[i] = minResForAbort[i];
  const bool good = t->trk->trackNewestCoarse(t->cur, T, aff, coarsestLvl, mra);
  se3_to(T, T_io);
  aff_io[0] = aff.a; aff_io[1] = aff.b;
  for (int i = 0; i < 5; ++i) lastResiduals[i] = t->trk->lastResiduals[i];
  for (int i = 0; i < 3; ++i) lastFlow[i] = t->trk->lastFlowIndicators[i];
  *ok = good ? 1 : 0;
  return 0;
}

// CoarseTracker::setCoarseTrackingRef -> makeCoarseDepthL0 (Src/CoarseTracker.cpp:105-263, 486-499): n active points whose
// residual into the newest keyframe is IN, given as centerProjectedTo (u, v, new_idepth) and efPoint->HdiF; dI_ref = the
// keyframe's pyramid (oracle layout).  The clouds are read back with ref_trk_get_ref.
int ref_trk_make_coarse_depth(ref_trk* t, int n, const float* cpt3, const float* hdif, const float* dI_ref) {
  for (int l = 0; l < t->levels; ++l)
    t->ref->frame->DirPyr[l] = reinterpret_cast<Vec3f*>(const_cast<float*>(dI_ref) + 3 * orc_pyramid_offset(t->w, t->h, l));
  std::shared_ptr<FrameShell> host = make_shell(t->calib, 1.f, t->levels);
  host->frame->DirPyr[0] = t->ref->frame->DirPyr[0];      // the ImmaturePoint constructor samples its host image
  host->frame->pointHessians.clear();
  for (int i = 0; i < n; ++i) {
    std::shared_ptr<ImmaturePoint> raw(new ImmaturePoint(8.f, 8.f, host, 0.f, t->calib));
    raw->idepth_min = raw->idepth_max = 1.f;
    std::shared_ptr<MapPoint> mp(new MapPoint(raw, t->calib));
    mp->efPoint->HdiF = hdif[i];
    std::shared_ptr<PointFrameResidual> r(new PointFrameResidual(mp, host, t->ref));
    r->isActiveAndIsGoodNEW = true;
    for (int k = 0; k < 3; ++k) r->centerProjectedTo[k] = cpt3[3 * i + k];
    mp->lastResiduals[0].first = r; mp->lastResiduals[0].second = ResState::IN;
    host->frame->pointHessians.push_back(mp);
  }
  std::vector<std::shared_ptr<FrameShell>> fhs;
  fhs.push_back(host); fhs.push_back(t->ref);
  t->trk->setCoarseTrackingRef(fhs);
  for (auto& mp : host->frame->pointHessians) mp->Clear();
  return 0;
}
int ref_trk_get_ref(ref_trk* t, int lvl, float* u, float* v, float* idepth, float* color, int* n_out) {
  CoarseTracker& c = *t->trk;
  const int n = c.pc_n[lvl];
  if (u) { memcpy(u, c.pc_u[lvl], sizeof(float) * n); memcpy(v, c.pc_v[lvl], sizeof(float) * n);
           memcpy(idepth, c.pc_idepth[lvl], sizeof(float) * n); memcpy(color, c.pc_color[lvl], sizeof(float) * n); }
  *n_out = n;
  return 0;
}

// ImmaturePoint::traceOn (Src/ImmaturePoint.cpp:40-351) for n points against one frame; same argument list as
// orc_trace_immature (the settings are the reference's own globals of Src/Settings.cpp).
int ref_trace_immature(const float* dI0, int w, int h, float fx, float fy, float cx, float cy, int n_hosts, const float* KRKi, const float* Kt,
                       const float* aff, int n, const int32_t* host_idx, const float* u, const float* v, const float* gradH,
                       const float* color, const float* weights, const float* energyTH, float* idepth_min_io, float* idepth_max_io,
                       float* quality_io, int8_t* status_io, float* lastTraceUV, float* lastTracePixelInterval) {
  Mat33 K = Mat33::Identity();
  K(0, 0) = fx; K(1, 1) = fy; K(0, 2) = cx; K(1, 2) = cy;
  int dirPyr = 1;
  std::shared_ptr<CalibData> calib(new CalibData(w, h, K, 0.f, nullptr, nullptr, dirPyr, 1, 1.2f));
  std::shared_ptr<FrameShell> frame = make_shell(calib, 1.f, 1);
  frame->frame->DirPyr[0] = reinterpret_cast<Vec3f*>(const_cast<float*>(dI0));
  // a host whose image is the frame itself: the constructor only samples it for colours we overwrite below
  std::shared_ptr<FrameShell> host = make_shell(calib, 1.f, 1);
  host->frame->DirPyr[0] = frame->frame->DirPyr[0];
  (void)n_hosts;
  for (int p = 0; p < n; ++p) {
    std::shared_ptr<ImmaturePoint> ip(new ImmaturePoint(8.f, 8.f, host, 0.f, calib));
    ip->u = u[p]; ip->v = v[p];
    for (int k = 0; k < 8; ++k) { ip->color[k] = color[8 * p + k]; ip->weights[k] = weights[8 * p + k]; }
    ip->gradH(0, 0) = gradH[4 * p]; ip->gradH(1, 0) = gradH[4 * p + 1]; ip->gradH(0, 1) = gradH[4 * p + 2]; ip->gradH(1, 1) = gradH[4 * p + 3];
    ip->energyTH = energyTH[p];
    ip->idepth_min = idepth_min_io[p]; ip->idepth_max = idepth_max_io[p];
    ip->quality = quality_io[p];
    ip->lastTraceStatus = (ImmaturePointStatus)status_io[p];
    ip->lastTraceUV[0] = lastTraceUV[2 * p]; ip->lastTraceUV[1] = lastTraceUV[2 * p + 1];
    ip->lastTracePixelInterval = lastTracePixelInterval[p];
    const int hI = host_idx[p];
    Mat33f M; Vec3f t3; Vec2f a2;
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) M(i, j) = KRKi[9 * hI + 3 * i + j]; t3[i] = Kt[3 * hI + i]; }
    a2[0] = aff[2 * hI]; a2[1] = aff[2 * hI + 1];
    ip->traceOn(frame, M, t3, a2, calib, false);
    idepth_min_io[p] = ip->idepth_min; idepth_max_io[p] = ip->idepth_max; quality_io[p] = ip->quality;
    status_io[p] = (int8_t)ip->lastTraceStatus;
    lastTraceUV[2 * p] = ip->lastTraceUV[0]; lastTraceUV[2 * p + 1] = ip->lastTraceUV[1];
    lastTracePixelInterval[p] = ip->lastTracePixelInterval;
  }
  return 0;
}

// System::optimizeImmaturePoint (Src/FullSystemOptPoint.cpp:24-175, compiled unmodified) with ImmaturePoint::linearizeResidual
// (Src/ImmaturePoint.cpp:389-451; projectPoint / derive_idepth of Include/DirectProjection.h) for every candidate.  The
// System object is a zero-filled block with frameHessians and Calib set by hand (never constructed, never destroyed: the
// constructor would start the whole SLAM system).  Same argument list as orc_activate_immature.
int ref_activate_immature(int n_frames, const float* const* dI0, int w, int h, float fx, float fy, float cx, float cy, const float* RtAff,
                          float huber, float minIdepthH_act, int GNIts, int minObs, int n, const int32_t* host_idx, const float* u,
                          const float* v, const float* color, const float* weights, const float* energyTH, const float* idepth_min,
                          const float* idepth_max, int8_t* ok_out, float* idepth_out, int8_t* res_state_out) {
  Mat33 K = Mat33::Identity();
  K(0, 0) = fx; K(1, 1) = fy; K(0, 2) = cx; K(1, 2) = cy;
  int dirPyr = 1;
  std::shared_ptr<CalibData> calib(new CalibData(w, h, K, 0.f, nullptr, nullptr, dirPyr, 1, 1.2f));
  setting_huberTH = huber;
  setting_minIdepthH_act = minIdepthH_act;
  setting_GNItsOnPointActivation = GNIts;
  std::vector<std::shared_ptr<FrameShell>> frames;
  for (int f = 0; f < n_frames; ++f) {
    frames.push_back(make_shell(calib, 1.f, 1));
    frames[f]->frame->DirPyr[0] = reinterpret_cast<Vec3f*>(const_cast<float*>(dI0[f]));
    frames[f]->frame->idx = f;
  }
  for (int hh = 0; hh < n_frames; ++hh) {   // the three members of FrameFramePrecalc that linearizeResidual reads
    frames[hh]->frame->targetPrecalc.resize(n_frames);
    for (int t = 0; t < n_frames; ++t) {
      const float* T = RtAff + 14 * (size_t)(hh + t * n_frames);
      FrameFramePrecalc& pc = frames[hh]->frame->targetPrecalc[t];
      for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) pc.PRE_RTll(i, j) = T[3 * i + j]; pc.PRE_tTll[i] = T[9 + i]; }
      pc.PRE_aff_mode[0] = T[12]; pc.PRE_aff_mode[1] = T[13];
    }
  }
  System* sys = static_cast<System*>(std::calloc(1, sizeof(System)));
  new (&sys->frameHessians) std::vector<std::shared_ptr<FrameShell>>(frames);
  new (&sys->Calib) std::shared_ptr<CalibData>(calib);
  std::vector<std::shared_ptr<ImmaturePointTemporaryResidual>> residuals;   // as System::activatePointsMT_Reductor provides them
  for (int f = 0; f < n_frames; ++f) residuals.push_back(std::shared_ptr<ImmaturePointTemporaryResidual>(new ImmaturePointTemporaryResidual()));
  for (int p = 0; p < n; ++p) {
    const int host = host_idx[p];
    if (host < 0 || host >= n_frames) return -1;
    ok_out[p] = 0; idepth_out[p] = 0;
    for (int f = 0; f < n_frames; ++f) res_state_out[(size_t)p * n_frames + f] = 1;
    std::shared_ptr<ImmaturePoint> point(new ImmaturePoint(8.f, 8.f, frames[host], 0.f, calib));
    point->u = u[p]; point->v = v[p];
    for (int k = 0; k < 8; ++k) { point->color[k] = color[8 * (size_t)p + k]; point->weights[k] = weights[8 * (size_t)p + k]; }
    point->energyTH = energyTH[p];
    point->idepth_min = idepth_min[p]; point->idepth_max = idepth_max[p];
    std::shared_ptr<MapPoint> mp = sys->optimizeImmaturePoint(point, minObs, residuals);
    if (!mp) continue;
    ok_out[p] = 1; idepth_out[p] = mp->idepth;
    // the final states of the temporary residuals, in frame order without the host (:28-38)
    int k = 0;
    for (int f = 0; f < n_frames; ++f)
      if (f != host) res_state_out[(size_t)p * n_frames + f] = (int8_t)residuals[k++]->state_state;
    for (auto& r : mp->residuals) r->Clear();
    mp->Clear();
  }
  for (auto& r : residuals) r->target.reset();
  sys->frameHessians.clear(); sys->Calib.reset();
  return 0;   // (the block itself is leaked on purpose: ~System must not run)
}

}  // extern "C"
