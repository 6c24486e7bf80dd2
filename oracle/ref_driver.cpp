// ref_driver.cpp — C interface around the UNMODIFIED reference translation units of the BA hot path.  TEST INFRASTRUCTURE.
//
// oracle/Makefile compiles, from where they lie under /root/reference (nothing is copied into this repository):
//     Src/OptimizationClasses.cpp      FrameFramePrecalc::set, PointFrameResidual::linearize / applyRes / fixLinearizationF
//     Src/AccumulatedTopHessian.cpp    AccumulatedTopHessianSSE::addPoint<mode>, stitchDouble, stitchDoubleInternal
//     Src/AccumulatedSCHessian.cpp     AccumulatedSCHessianSSE::addPoint, stitchDouble, stitchDoubleInternal
//     Src/EnergyFunctional.cpp         EnergyFunctional: constructor, setAdjointsF, setDeltaF, accumulate{A,L,SC}F_MT,
//                                      resubstituteF_MT / resubstituteFPt, calcLEnergyF_MT / calcLEnergyPt
//     Src/FullSystemOptimize.cpp       System::linearizeAll / linearizeAll_Reductor / setNewFrameEnergyTH
//     Src/Settings.cpp                 the setting_* globals
// together with their own headers (Include/GlobalTypes.h, MatrixAccumulators.h, OptimizationClasses.h, DirectProjection.h,
// CalibData.h, Frame.h, MapPoint.h, EnergyFunctional.h, Accumulated*Hessian.h, IndexThreadReduce.h, System.h, ...) against
// the small stand-ins of oracle/ref_shim/ for the libraries this image lacks (Eigen, Sophus-on-Eigen, Boost.Thread, the
// OpenCV / DBoW3 type names), into oracle/_ref/libhslam_ref.so.  This file is OUR code: it builds the reference's object
// graph (FrameShell / Frame / FrameOptimizationData / MapPoint / PointFrameResidual / EnergyFunctional / CalibData) from the
// same flat window description the oracle and the CUDA library take, calls the reference's member functions and hands the
// results back as flat arrays.  It also DEFINES the few out-of-line symbols whose own translation units are not part of
// the path (Frame / ImmaturePoint constructors, static members, System::setPrecalcValues): trivial bodies, they only have
// to satisfy the linker.
//
// tests/test_oracle_vs_reference.py feeds one window to this library and to the oracle's float restatement and compares
// every output: that is what pins the oracle to the reference's own code.
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
// System keeps its CalibData in a private member; the driver sets up a bare System object by hand (its constructor starts the
// whole SLAM system), so the reference headers are read with `private` opened up — the standard headers above are not
#define private public
#include "System.h"
#undef private
#include "AccumulatedSCHessian.h"
#include "AccumulatedTopHessian.h"
#include "CalibData.h"
#include "EnergyFunctional.h"
#include "Frame.h"
#include "MapPoint.h"
#include "OptimizationClasses.h"

#include "IndexThreadReduce.h"

#include "hslam_oracle.h"   // orc_window + the hslam_*_soa data-format structs

using namespace FSLAM;

#define REF_HAVE_ENERGYFUNCTIONAL_CPP   /* Src/EnergyFunctional.cpp is part of this library */
#include "ref_stubs.inc"

struct ref_ba {
  std::shared_ptr<CalibData> calib;
  std::shared_ptr<EnergyFunctional> ef;
  std::vector<std::shared_ptr<FrameShell>> frames;
  std::vector<std::shared_ptr<MapPoint>> points;              // caller order
  std::vector<std::shared_ptr<PointFrameResidual>> residuals; // caller order
  // adjoint tables and the three accumulators are the EnergyFunctional's own (its constructor news the accumulators, its
  // destructor delete[]s the tables: Src/EnergyFunctional.cpp:84-121)
  int n = 0;
};

namespace FSLAM {
void System::setPrecalcValues() { std::abort(); }   // Src/System.cpp (not compiled); only System::optimize calls it
}  // namespace FSLAM

extern "C" {

ref_ba* ref_ba_create(void) { return new ref_ba(); }
void ref_ba_destroy(ref_ba* b) {
  if (!b) return;
  for (auto& r : b->residuals) r->Clear();
  for (auto& p : b->points) p->Clear();
  delete b;
}

int ref_ba_set_window(ref_ba* b, const orc_window* w) {
  const int n = w->n_frames;
  const bool fresh = b->n != n || b->frames.empty();
  if (fresh) {
    b->n = n;
    Mat33 K = Mat33::Identity();
    K(0, 0) = w->fx; K(1, 1) = w->fy; K(0, 2) = w->cx; K(1, 2) = w->cy;
    int dirPyr = 1;
    b->calib.reset(new CalibData(w->w, w->h, K, 0.f, nullptr, nullptr, dirPyr, 1, 1.2f));
    b->ef.reset(new EnergyFunctional());
    b->frames.clear();
    std::shared_ptr<ImageData> img(new ImageData(1, 1));
    img->timestamp = 0; img->ExposureL = 1.f; img->ExposureR = 1.f;
    for (int f = 0; f < n; ++f) {
      std::shared_ptr<FrameShell> fs(new FrameShell(img, nullptr, b->calib, nullptr, false));
      fs->frame->idx = f;
      fs->frame->efFrame.reset(new FrameOptimizationData(f, 1.f));
      fs->frame->targetPrecalc.resize(n);
      fs->frame->DirPyr.resize(1);
      b->frames.push_back(fs);
    }
    b->ef->frames = b->frames;
    b->ef->nFrames = n;
    b->ef->adHost = new Mat88[n * n]; b->ef->adTarget = new Mat88[n * n];       // as setAdjointsF allocates them (:26-27, 75-78)
    b->ef->adHostF = new Mat88f[n * n]; b->ef->adTargetF = new Mat88f[n * n];
    b->ef->adHTdeltaF = new Mat18f[n * n];
  }
  for (int f = 0; f < n; ++f) {
    Frame& F = *b->frames[f]->frame;
    F.DirPyr[0] = reinterpret_cast<Vec3f*>(const_cast<float*>(w->dI[f]));
    F.frameEnergyTH = w->frameEnergyTH[f];
    for (int i = 0; i < 8; ++i) { F.efFrame->prior[i] = w->framePrior[8 * f + i]; F.efFrame->delta_prior[i] = w->frameDeltaPrior[8 * f + i]; }
  }
  for (int h = 0; h < n; ++h)
    for (int t = 0; t < n; ++t) {
      const int idx = h + t * n;
      const float* p = w->precalc + (size_t)HSLAM_PRECALC_STRIDE * idx;
      FrameFramePrecalc& P = b->frames[h]->frame->targetPrecalc[t];
      P.host = b->frames[h]; P.target = b->frames[t];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { P.PRE_KRKiTll(i, j) = p[3 * i + j]; P.PRE_RTll_0(i, j) = p[12 + 3 * i + j]; }
      for (int i = 0; i < 3; ++i) { P.PRE_KtTll[i] = p[9 + i]; P.PRE_tTll_0[i] = p[21 + i]; }
      P.PRE_aff_mode[0] = p[24]; P.PRE_aff_mode[1] = p[25]; P.PRE_b0_mode = p[26];
      for (int c = 0; c < 8; ++c)
        for (int r = 0; r < 8; ++r) {   // col-major tables
          b->ef->adHost[idx](r, c) = w->adHost[64 * (size_t)idx + r + 8 * c];
          b->ef->adTarget[idx](r, c) = w->adTarget[64 * (size_t)idx + r + 8 * c];
        }
      b->ef->adHostF[idx] = b->ef->adHost[idx].cast<float>();     // Src/EnergyFunctional.cpp:73-74
      b->ef->adTargetF[idx] = b->ef->adTarget[idx].cast<float>();
      for (int i = 0; i < 8; ++i) b->ef->adHTdeltaF[idx][i] = w->adHTdeltaF[8 * (size_t)idx + i];
    }
  for (int i = 0; i < 4; ++i) { b->ef->cDeltaF[i] = w->cDeltaF[i]; b->ef->cPrior[i] = w->cPrior[i]; }
  b->ef->cPriorF = b->ef->cPrior.cast<float>();
  return 0;
}

int ref_ba_set_topology(ref_ba* b, const hslam_point_soa* p, const hslam_residual_soa* r) {
  for (auto& q : b->residuals) q->Clear();
  for (auto& q : b->points) q->Clear();
  b->points.clear(); b->residuals.clear();
  for (auto& f : b->frames) f->frame->pointHessians.clear();
  const int P = p->n_points, R = r->n_residuals;
  for (int i = 0; i < P; ++i) {
    std::shared_ptr<FrameShell>& host = b->frames[p->host_idx[i]];
    std::shared_ptr<ImmaturePoint> raw(new ImmaturePoint(p->u[i], p->v[i], host, 0.f, b->calib));
    memcpy(raw->color, p->color + 8 * (size_t)i, sizeof(float) * 8);
    memcpy(raw->weights, p->weights + 8 * (size_t)i, sizeof(float) * 8);
    raw->idepth_min = raw->idepth_max = p->idepth[i];
    std::shared_ptr<MapPoint> mp(new MapPoint(raw, b->calib));
    mp->setIdepth(p->idepth[i]);
    mp->efPoint->setIdepthZero(p->idepth_zero[i]);
    mp->efPoint->priorF = p->priorF[i];
    mp->efPoint->deltaF = p->deltaF[i];
    mp->setPointStatus(ACTIVE);
    mp->idx = i;
    b->points.push_back(mp);
    host->frame->pointHessians.push_back(mp);
  }
  b->residuals.resize(R);
  for (int i = 0; i < P; ++i)
    for (int k = p->res_begin[i]; k < p->res_begin[i + 1]; ++k) {
      const int id = p->res_index[k];
      std::shared_ptr<MapPoint>& mp = b->points[i];
      std::shared_ptr<PointFrameResidual> res(new PointFrameResidual(mp, mp->host, b->frames[r->target_idx[id]]));
      res->hostIDX = p->host_idx[i]; res->targetIDX = r->target_idx[id];
      res->state_state = (ResState)r->state_state[id];
      res->state_energy = r->state_energy ? r->state_energy[id] : 0.0;
      res->isLinearized = r->is_linearized[id] != 0;
      res->isActiveAndIsGoodNEW = r->is_active[id] != 0;
      res->res_toZeroF.setZero();
      mp->residuals.push_back(res);     // the point's own order
      b->residuals[id] = res;
    }
  b->ef->allPoints = b->points;
  return 0;
}

int ref_ba_update_points(ref_ba* b, const float* idepth, const float* deltaF, const float* idepth_zero) {
  for (size_t i = 0; i < b->points.size(); ++i) {
    if (idepth) b->points[i]->setIdepth(idepth[i]);
    if (idepth_zero) b->points[i]->efPoint->setIdepthZero(idepth_zero[i]);
    if (deltaF) b->points[i]->efPoint->deltaF = deltaF[i];
  }
  return 0;
}

// System::linearizeAll_Reductor over activeResiduals = the non-linearised residuals (FullSystemOptimize.cpp:19-52,371-391)
int ref_ba_linearize(ref_ba* b, int fix, double* energy_sum, int8_t* newState, float* newEnergy, float* newEnergyWO, float* cpt,
                     float* pto) {
  double e = 0;
  for (auto& r : b->residuals) {
    if (r->isLinearized) continue;
    e += r->linearize(b->calib);
    if (fix) r->applyRes(true);
  }
  if (energy_sum) *energy_sum = e;
  for (size_t i = 0; i < b->residuals.size(); ++i) {
    PointFrameResidual& r = *b->residuals[i];
    if (newState) newState[i] = (int8_t)r.state_NewState;
    if (newEnergy) newEnergy[i] = (float)r.state_NewEnergy;
    if (newEnergyWO) newEnergyWO[i] = (float)r.state_NewEnergyWithOutlier;
    if (cpt) for (int k = 0; k < 3; ++k) cpt[3 * i + k] = r.centerProjectedTo[k];
    if (pto) for (int k = 0; k < 8; ++k) { pto[16 * i + 2 * k] = r.projectedTo[k][0]; pto[16 * i + 2 * k + 1] = r.projectedTo[k][1]; }
  }
  return 0;
}

int ref_ba_apply_res(ref_ba* b) {   // System::applyRes_Reductor (FullSystemOptimize.cpp:55-59)
  for (auto& r : b->residuals) if (!r->isLinearized) r->applyRes(true);
  return 0;
}

int ref_ba_fix_linearization(ref_ba* b, const int32_t* ids, int count) {
  for (int i = 0; i < count; ++i) {
    b->residuals[ids[i]]->fixLinearizationF(b->ef);
  }
  return 0;
}

// EnergyFunctional::accumulate{A,L,SC}F_MT (Src/EnergyFunctional.cpp:155-220): the reference's own member functions,
// single-thread branch (MT = false)
int ref_ba_accumulate(ref_ba* b, int which, double* H, double* bvec, int* nres) {
  MatXX Hm; VecX bm;
  EnergyFunctional* ef = b->ef.get();
  const int n = b->n;
  if (which == 0) {
    ef->accumulateAF_MT(Hm, bm, false);
    if (nres) *nres = ef->resInA;
  } else if (which == 1) {
    ef->accumulateLF_MT(Hm, bm, false);
    if (nres) *nres = ef->resInL;
  } else {
    ef->accumulateSCF_MT(Hm, bm, false);
  }
  const int N = 8 * n + 4;
  for (int j = 0; j < N; ++j) for (int i = 0; i < N; ++i) H[i + (size_t)N * j] = Hm(i, j);
  for (int i = 0; i < N; ++i) bvec[i] = bm[i];
  return 0;
}

// EnergyFunctional::marginalizePointsF (EnergyFunctional.cpp:545-609), the part on the path: priorF scaling (:564),
// addPoint<2> + SC addPoint(p, false) over the points to marginalise (:574-579), stitchDouble(usePrior=false) of both
// (:582-583).  The caller forms H = M - Msc, b = Mb - Mbsc (:587-588).
int ref_ba_marginalize_points(ref_ba* b, const int32_t* ids, int count, float prior_scale, double* M, double* Mb, double* Msc,
                              double* Mbsc, int* nres) {
  EnergyFunctional* ef = b->ef.get();
  const int n = b->n;
  b->ef->accSSE_bot->setZero(n);
  b->ef->accSSE_top_A->setZero(n);
  for (int i = 0; i < count; ++i) {
    if (ids[i] < 0 || ids[i] >= (int)b->points.size()) return -1;
    std::shared_ptr<MapPoint>& p = b->points[ids[i]];
    p->efPoint->priorF *= prior_scale;
    b->ef->accSSE_top_A->addPoint<2>(p, ef);
    b->ef->accSSE_bot->addPoint(p, false);
  }
  MatXX Mm, Mscm; VecX Mbm, Mbscm;
  b->ef->accSSE_top_A->stitchDouble(Mm, Mbm, ef, false, false);
  b->ef->accSSE_bot->stitchDouble(Mscm, Mbscm, ef);
  if (nres) *nres = b->ef->accSSE_top_A->nres[0];
  const int N = 8 * n + 4;
  for (int j = 0; j < N; ++j) for (int i = 0; i < N; ++i) { M[i + (size_t)N * j] = Mm(i, j); Msc[i + (size_t)N * j] = Mscm(i, j); }
  for (int i = 0; i < N; ++i) { Mb[i] = Mbm[i]; Mbsc[i] = Mbscm[i]; }
  return 0;
}

int ref_ba_point_results(ref_ba* b, double* HdiF, double* bdSumF, double* idepth_hessian, double* Hdd_accAF, double* bd_accAF,
                         double* Hcd_accAF, double* Hdd_accLF, double* bd_accLF, double* Hcd_accLF) {
  for (size_t i = 0; i < b->points.size(); ++i) {
    const MapPoint& p = *b->points[i];
    const MapPointOptimizationData& e = *p.efPoint;
    if (HdiF) HdiF[i] = e.HdiF;
    if (bdSumF) bdSumF[i] = e.bdSumF;
    if (idepth_hessian) idepth_hessian[i] = p.idepth_hessian;
    if (Hdd_accAF) Hdd_accAF[i] = e.Hdd_accAF;
    if (bd_accAF) bd_accAF[i] = e.bd_accAF;
    if (Hdd_accLF) Hdd_accLF[i] = e.Hdd_accLF;
    if (bd_accLF) bd_accLF[i] = e.bd_accLF;
    for (int k = 0; k < 4; ++k) {
      if (Hcd_accAF) Hcd_accAF[4 * i + k] = e.Hcd_accAF[k];
      if (Hcd_accLF) Hcd_accLF[4 * i + k] = e.Hcd_accLF[k];
    }
  }
  return 0;
}

// EnergyFunctional::resubstituteF_MT / resubstituteFPt (Src/EnergyFunctional.cpp:222-274), single-thread branch: the
// reference's own member function on the driver's window.  x: (8n+4) solution; step_out: P floats (efPoint->step).
int ref_ba_resubstitute(ref_ba* b, const double* x, float* step_out) {
  const int N = 8 * b->n + 4;
  VecX xv(N);
  for (int i = 0; i < N; ++i) xv[i] = x[i];
  b->ef->resubstituteF_MT(xv, b->calib, false);
  for (size_t i = 0; i < b->points.size(); ++i) step_out[i] = b->points[i]->efPoint->step;
  return 0;
}

// EnergyFunctional::calcLEnergyF_MT / calcLEnergyPt (Src/EnergyFunctional.cpp:289-364) with the reference's own
// IndexThreadReduce (one worker thread, chunks of 50 points).
int ref_ba_calc_l_energy(ref_ba* b, double* E_out) {
  static IndexThreadReduce<Vec10> red;
  b->ef->red = &red;
  EFDeltaValid = EFAdjointsValid = EFIndicesValid = true;   // the tables were set by ref_ba_set_window / ref_ba_set_topology
  *E_out = b->ef->calcLEnergyF_MT();
  b->ef->red = nullptr;
  return 0;
}

// EnergyFunctional::setAdjointsF + setDeltaF (Src/EnergyFunctional.cpp:22-82, 128-152) from the frames' poses and states:
// w2c_evalPT (12: R row-major, t), state / state_zero (10) per frame, exposures.  Out: adHost / adTarget (n*n*64, column-
// major 8x8, index host + target*n) and adHTdeltaF (n*n*8).
int ref_ef_adjoints_deltas(int n, const double* w2c_evalPT, const double* state, const double* state_zero, const float* ab_exposure,
                           double* adHost_out, double* adTarget_out, float* adHTdeltaF_out) {
  Mat33 K = Mat33::Identity();
  K(0, 0) = 300; K(1, 1) = 300; K(0, 2) = 160; K(1, 2) = 120;
  int dirPyr = 1;
  std::shared_ptr<CalibData> calib(new CalibData(320, 240, K, 0.f, nullptr, nullptr, dirPyr, 1, 1.2f));
  std::shared_ptr<ImageData> img(new ImageData(1, 1));
  img->timestamp = 0; img->ExposureL = 1.f; img->ExposureR = 1.f;
  EnergyFunctional ef;
  for (int f = 0; f < n; ++f) {
    std::shared_ptr<FrameShell> fs(new FrameShell(img, nullptr, calib, nullptr, false));
    fs->frame->idx = f;
    fs->ab_exposure = ab_exposure[f];
    fs->frame->efFrame.reset(new FrameOptimizationData(f, ab_exposure[f]));
    FrameOptimizationData& e = *fs->frame->efFrame;
    Eigen::Matrix<double, 3, 3> R; Eigen::Matrix<double, 3, 1> t;
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R(i, j) = w2c_evalPT[12 * f + 3 * i + j]; t[i] = w2c_evalPT[12 * f + 9 + i]; }
    e.worldToCam_evalPT = SE3(R, t);
    Vec10 s, s0;
    for (int i = 0; i < 10; ++i) { s[i] = state[10 * f + i]; s0[i] = state_zero[10 * f + i]; }
    e.state_zero = s0;   // (setStateZero also builds the nullspaces, which the path does not use)
    e.setState(s);
    ef.frames.push_back(fs);
  }
  ef.nFrames = n;
  ef.setAdjointsF(calib);
  ef.setDeltaF(calib);
  for (int idx = 0; idx < n * n; ++idx) {
    for (int c = 0; c < 8; ++c)
      for (int r = 0; r < 8; ++r) {
        adHost_out[64 * (size_t)idx + r + 8 * c] = ef.adHost[idx](r, c);
        adTarget_out[64 * (size_t)idx + r + 8 * c] = ef.adTarget[idx](r, c);
      }
    for (int i = 0; i < 8; ++i) adHTdeltaF_out[8 * (size_t)idx + i] = ef.adHTdeltaF[idx][i];
  }
  return 0;
}

// System::linearizeAll(false) (Src/FullSystemOptimize.cpp:103-165): linearizeAll_Reductor over activeResiduals (:19-52) and
// setNewFrameEnergyTH (:60-101), the reference's own member functions, single-thread branch.  The System object is a
// zero-filled block with the five members the functions touch set by hand (never constructed, never destroyed: the
// constructor would start the whole SLAM system).  activeResiduals in the order System::optimize builds it (:371-391):
// frames -> points -> the point's residuals, the non-linearised ones.  Out: the summed energy and the new
// frameEnergyTH of the newest frame (frameHessians.back()).
int ref_ba_system_linearize_all(ref_ba* b, double* energy_out, float* frameEnergyTH_out) {
  System* sys = static_cast<System*>(std::calloc(1, sizeof(System)));
  new (&sys->frameHessians) std::vector<std::shared_ptr<FrameShell>>(b->frames);
  new (&sys->activeResiduals) std::vector<std::shared_ptr<PointFrameResidual>>();
  new (&sys->allResVec) std::vector<float>();
  new (&sys->Calib) std::shared_ptr<CalibData>(b->calib);
  new (&sys->ef) std::shared_ptr<EnergyFunctional>(b->ef);
  for (auto& fh : b->frames)
    for (auto& ph : fh->frame->pointHessians)
      for (auto& r : ph->residuals)
        if (!r->isLinearized) sys->activeResiduals.push_back(r);
  multiThreading = false;
  const Vec3 res = sys->linearizeAll(false);
  *energy_out = res[0];
  *frameEnergyTH_out = b->frames.back()->frame->frameEnergyTH;
  sys->frameHessians.clear(); sys->activeResiduals.clear(); sys->Calib.reset(); sys->ef.reset();
  return 0;   // (the block itself is leaked on purpose: ~System must not run)
}

// RawResidualJacobian in member order (Include/OptimizationClasses.h:48-77): resF[8], Jpdxi[2][6], Jpdc[2][4], Jpdd[2],
// JIdx[2][8], JabF[2][8], JIdx2[4], JabJIdx[4], Jab2[4] (2x2 blocks column-major)
int ref_ba_residual_results(ref_ba* b, int8_t* state_state, uint8_t* is_active, double* JpJdF, double* J74, double* rtz) {
  for (size_t i = 0; i < b->residuals.size(); ++i) {
    const PointFrameResidual& r = *b->residuals[i];
    if (state_state) state_state[i] = (int8_t)r.state_state;
    if (is_active) is_active[i] = r.isActiveAndIsGoodNEW ? 1 : 0;
    if (JpJdF) for (int k = 0; k < 8; ++k) JpJdF[8 * i + k] = r.JpJdF[k];
    if (rtz) for (int k = 0; k < 8; ++k) rtz[8 * i + k] = r.res_toZeroF[k];
    if (J74) {
      double* o = J74 + 74 * i;
      const RawResidualJacobian& J = *r.J;
      int q = 0;
      for (int k = 0; k < 8; ++k) o[q++] = J.resF[k];
      for (int a = 0; a < 2; ++a) for (int k = 0; k < 6; ++k) o[q++] = J.Jpdxi[a][k];
      for (int a = 0; a < 2; ++a) for (int k = 0; k < 4; ++k) o[q++] = J.Jpdc[a][k];
      for (int k = 0; k < 2; ++k) o[q++] = J.Jpdd[k];
      for (int a = 0; a < 2; ++a) for (int k = 0; k < 8; ++k) o[q++] = J.JIdx[a][k];
      for (int a = 0; a < 2; ++a) for (int k = 0; k < 8; ++k) o[q++] = J.JabF[a][k];
      for (int c = 0; c < 2; ++c) for (int rr = 0; rr < 2; ++rr) o[q++] = J.JIdx2(rr, c);
      for (int c = 0; c < 2; ++c) for (int rr = 0; rr < 2; ++rr) o[q++] = J.JabJIdx(rr, c);
      for (int c = 0; c < 2; ++c) for (int rr = 0; rr < 2; ++rr) o[q++] = J.Jab2(rr, c);
    }
  }
  return 0;
}

// FrameFramePrecalc::set (Src/OptimizationClasses.cpp:13-39) for one ordered pair from the frames' poses and states:
// w2c_evalPT (12: R row-major, t), state / state_zero (10) per frame, exposures.  Out: the 28 floats of the precalc record.
int ref_precalc_set(int n, const double* w2c_evalPT, const double* state, const double* state_zero, const float* ab_exposure, float fx,
                    float fy, float cx, float cy, int w, int h, float* precalc_out) {
  Mat33 K = Mat33::Identity();
  K(0, 0) = fx; K(1, 1) = fy; K(0, 2) = cx; K(1, 2) = cy;
  int dirPyr = 1;
  std::shared_ptr<CalibData> calib(new CalibData(w, h, K, 0.f, nullptr, nullptr, dirPyr, 1, 1.2f));
  std::shared_ptr<ImageData> img(new ImageData(1, 1));
  img->timestamp = 0; img->ExposureL = 1.f; img->ExposureR = 1.f;
  std::vector<std::shared_ptr<FrameShell>> fr;
  for (int f = 0; f < n; ++f) {
    std::shared_ptr<FrameShell> fs(new FrameShell(img, nullptr, calib, nullptr, false));
    fs->frame->idx = f;
    fs->ab_exposure = ab_exposure[f];
    fs->frame->efFrame.reset(new FrameOptimizationData(f, ab_exposure[f]));
    FrameOptimizationData& e = *fs->frame->efFrame;
    Eigen::Matrix<double, 3, 3> R; Eigen::Matrix<double, 3, 1> t;
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R(i, j) = w2c_evalPT[12 * f + 3 * i + j]; t[i] = w2c_evalPT[12 * f + 9 + i]; }
    e.worldToCam_evalPT = SE3(R, t);
    Vec10 s, s0;
    for (int i = 0; i < 10; ++i) { s[i] = state[10 * f + i]; s0[i] = state_zero[10 * f + i]; }
    e.state_zero = s0;   // (setStateZero also builds the nullspaces, which the path does not use)
    e.setState(s);
    fs->frame->targetPrecalc.resize(n);
    fr.push_back(fs);
  }
  for (int hI = 0; hI < n; ++hI)
    for (int tI = 0; tI < n; ++tI) {
      FrameFramePrecalc& P = fr[hI]->frame->targetPrecalc[tI];
      P.set(fr[hI], fr[tI], calib);
      float* o = precalc_out + (size_t)HSLAM_PRECALC_STRIDE * (hI + tI * n);
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { o[3 * i + j] = P.PRE_KRKiTll(i, j); o[12 + 3 * i + j] = P.PRE_RTll_0(i, j); }
      for (int i = 0; i < 3; ++i) { o[9 + i] = P.PRE_KtTll[i]; o[21 + i] = P.PRE_tTll_0[i]; }
      o[24] = P.PRE_aff_mode[0]; o[25] = P.PRE_aff_mode[1]; o[26] = P.PRE_b0_mode; o[27] = 0.f;
      P.host.reset(); P.target.reset();
    }
  return 0;
}

}  // extern "C"
