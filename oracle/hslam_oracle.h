/*
 * hslam_oracle.h — C interface of the CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A CPU restatement of the reference's (AUBVRL/H-SLAM, namespace FSLAM)
 * photometric Gauss-Newton hot path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (h-slam_b200/) never does.
 *
 * PARITY PINNED AGAINST THE COMPILED REFERENCE (round 2): the reference ships no
 * tests, golden vectors or fixtures for this path and its
 * build needs Eigen/Boost/OpenCV, which this image lacks.  oracle/Makefile
 * therefore compiles the reference's own hot-path translation units UNMODIFIED,
 * from where they lie under /root/reference (OptimizationClasses.cpp,
 * AccumulatedTopHessian.cpp, AccumulatedSCHessian.cpp, EnergyFunctional.cpp,
 * FullSystemOptimize.cpp, FullSystemOptPoint.cpp, CoarseTracker.cpp, ImmaturePoint.cpp, Frame.cpp, Settings.cpp), against the minimal library stand-ins of
 * oracle/ref_shim/ (fixed-size Eigen subset with Eigen 3.3's evaluation order,
 * Sophus SE3, boost->std, type stubs) into oracle/_ref/libhslam_ref*.so;
 * tests/test_oracle_vs_reference.py drives both on the same inputs: linearize /
 * applyRes, System::linearizeAll with setNewFrameEnergyTH, the frame pyramid, the tracker's LM path, makeCoarseDepthL0, traceOn and the point
 * activation (linearizeResidual) agree BIT FOR BIT with the float restatement,
 * the stitched systems (incl. the marginalisation pass) to 1e-9..1e-12.  The
 * analytic known answers of tests/test_oracle_*.py (finite differences, dense
 * Schur elimination, zero motion, group identities) remain as a second anchor.
 *
 * Two restatements of every function:
 *   *_f : float-faithful — fp32 arithmetic in the operation order written in
 *         the reference source (Eigen fixed-size 3-vector reductions evaluated
 *         as a0+(a1+a2), the order of Eigen 3.3's redux_novec_unroller), the
 *         4-lane SSE layout and the 1/1k/1M shiftUp tiers of MatrixAccumulators;
 *         compiled with -ffp-contract=off.
 *   *_d : the same mathematics in fp64 on the same fp32 inputs.
 */
#ifndef HSLAM_ORACLE_H
#define HSLAM_ORACLE_H
#include <stdint.h>
#include "../include/hslam_b200.h" /* data-format structs only (hslam_params, *_soa) */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- SE3 (Thirdparty/Sophus/sophus/se3.hpp:131-139,407-427; so3.hpp:343-369) ---- */
void orc_se3_exp(const double xi[6] /*upsilon,omega*/, double T[12] /*R row-major, t*/);
void orc_se3_log(const double T[12], double xi[6]);
void orc_se3_mul(const double A[12], const double B[12], double C[12]);
void orc_se3_inv(const double A[12], double Ainv[12]);
void orc_se3_adj(const double T[12], double Adj[36] /*row-major 6x6*/);
/* AffLight::fromToVecExposure (Include/GlobalTypes.h:334-346) */
void orc_aff_from_to(float expF, float expT, double aF, double bF, double aT, double bT, double out2[2]);
/* Pivoted LDLT solve of an n x n (n<=16) symmetric system, col-major; stands in
 * for Eigen's ldlt().solve() (Src/CoarseTracker.cpp:563). */
void orc_ldlt_solve(const double* A, const double* rhs, int n, double* x);

/* ---- pyramid: Frame::CreateDirPyrs (Src/Frame.cpp:104-181) ----
 * dI_out: concatenated levels, Vec3f AoS (3 floats/px); absgrad_out likewise 1
 * float/px.  Rows 0 and h-1 (uninitialised in the reference) are written 0. */
int orc_pyramid(const float* img, int w, int h, int levels, const float* gammaBGrad256,
                float* dI_out, float* absgrad_out);
int64_t orc_pyramid_offset(int w, int h, int lvl); /* pixel offset of level lvl */

/* ---- tracker ---- */
typedef struct orc_trk_level {
  int w, h;
  float fx, fy, cx, cy;
  float K[9];   /* row-major */
  float Ki[9];  /* row-major */
  int n;
  const float* pc_u;
  const float* pc_v;
  const float* pc_idepth;
  const float* pc_color;
} orc_trk_level;

/* CoarseTracker::makeK (Src/CoarseTracker.cpp:72-101) */
void orc_make_K(int levels, int w0, int h0, float fx0, float fy0, float cx0, float cy0,
                orc_trk_level* out_levels);
/* CoarseTracker::makeCoarseDepthL0 (Src/CoarseTracker.cpp:105-263).  Input =
 * the (u,v,new_idepth,weight) splats of the IN residuals; dI_ref = reference
 * frame pyramid (orc_pyramid layout).  Output arrays per level are caller
 * allocated with w_l*h_l floats each; pc_n[l] receives the count. */
int orc_make_coarse_depth(int levels, int w0, int h0, int n_splat, const int* su, const int* sv,
                          const float* sidepth, const float* sweight, const float* dI_ref,
                          float* const* pc_u, float* const* pc_v, float* const* pc_idepth,
                          float* const* pc_color, int* pc_n);

typedef struct orc_trk_frame_pair {
  float ref_exposure, new_exposure;
  double ref_aff_a, ref_aff_b;
} orc_trk_frame_pair;

/* CoarseTracker::calcRes (Src/CoarseTracker.cpp:329-485). warped (nullable):
 * 8 arrays {idepth,u,v,dx,dy,residual,weight,refColor}, each >= n+4 floats
 * (doubles for _d); *n_warped = count BEFORE padding; mask (nullable, n bytes):
 * 0 = gated out, 1 = inlier, 2 = saturated. */
int orc_trk_calc_res_f(const orc_trk_level* L, int lvl, const float* dI_new, const orc_trk_frame_pair* fp,
                       const hslam_params* prm, const double refToNew[12], double aff_a, double aff_b,
                       float cutoffTH, double out6[6], float* const* warped, int* n_warped, uint8_t* mask);
int orc_trk_calc_res_d(const orc_trk_level* L, int lvl, const float* dI_new, const orc_trk_frame_pair* fp,
                       const hslam_params* prm, const double refToNew[12], double aff_a, double aff_b,
                       float cutoffTH, double out6[6], double* const* warped, int* n_warped, uint8_t* mask);
/* CoarseTracker::calcGSSSE (Src/CoarseTracker.cpp:267-324) + Accumulator9
 * (Include/MatrixAccumulators.h:982-1345) on the warped buffers of calc_res. */
int orc_trk_calc_gs_f(const orc_trk_level* L, const orc_trk_frame_pair* fp, const hslam_params* prm,
                      double aff_a, double aff_b, const float* const* warped, int n_warped_unpadded,
                      double H[64], double b[8]);
int orc_trk_calc_gs_d(const orc_trk_level* L, const orc_trk_frame_pair* fp, const hslam_params* prm,
                      double aff_a, double aff_b, const double* const* warped, int n_warped_unpadded,
                      double H[64], double b[8]);
/* CoarseTracker::trackNewestCoarse (Src/CoarseTracker.cpp:506-683).
 * dI_new = whole pyramid (orc_pyramid layout).  trace (nullable): per LM
 * evaluation 4 doubles {lvl, E/N, accepted, lambda} up to trace_cap rows,
 * *trace_n = rows written. */
int orc_trk_track_f(const orc_trk_level* levels, int n_levels, const float* dI_new, const orc_trk_frame_pair* fp,
                    const hslam_params* prm, double refToNew_io[12], double aff_io[2], int coarsestLvl,
                    const double minResForAbort[5], double lastResiduals[5], double lastFlow[3],
                    int* ok, int* iterations, double* trace, int trace_cap, int* trace_n);
int orc_trk_track_d(const orc_trk_level* levels, int n_levels, const float* dI_new, const orc_trk_frame_pair* fp,
                    const hslam_params* prm, double refToNew_io[12], double aff_io[2], int coarsestLvl,
                    const double minResForAbort[5], double lastResiduals[5], double lastFlow[3],
                    int* ok, int* iterations, double* trace, int trace_cap, int* trace_n);

/* ---- window BA ---- */
/* Host-side state preparation restated from FrameOptimizationData::setState
 * (Include/Frame.h:184-197), FrameFramePrecalc::set (Src/OptimizationClasses.cpp:13-39),
 * EnergyFunctional::setAdjointsF / setDeltaF (Src/EnergyFunctional.cpp:22-82,128-152).
 * Inputs per frame: worldToCam_evalPT (12 doubles), state[10], state_zero[10],
 * ab_exposure.  Outputs: precalc n*n*28 float, adHost/adTarget n*n*64 double
 * col-major, adHTdeltaF n*n*8 float, all indexed host + target*n. */
int orc_ba_prepare_window(int n, const double* w2c_evalPT, const double* state, const double* state_zero,
                          const float* ab_exposure, float fx, float fy, float cx, float cy,
                          const hslam_params* prm, float* precalc, double* adHost, double* adTarget,
                          float* adHTdeltaF);

typedef struct orc_window {
  int n_frames;
  int w, h;
  float fx, fy, cx, cy, fxi, fyi;
  const float* const* dI;       /* n_frames pointers to level-0 Vec3f AoS images */
  const float* frameEnergyTH;
  const float* precalc;
  const double* adHost;
  const double* adTarget;
  const float* adHTdeltaF;
  float cDeltaF[4];
  double cPrior[4];
  const double* framePrior;
  const double* frameDeltaPrior;
} orc_window;

typedef struct orc_ba orc_ba;  /* holds J, states, accumulators; scalar chosen at creation */
orc_ba* orc_ba_create(int use_double, const hslam_params* prm);
void orc_ba_destroy(orc_ba* ba);
int orc_ba_set_window(orc_ba* ba, const orc_window* w);
int orc_ba_set_topology(orc_ba* ba, const hslam_point_soa* pts, const hslam_residual_soa* res);
int orc_ba_update_points(orc_ba* ba, const float* idepth, const float* deltaF, const float* idepth_zero);
/* threads<=1: the reference's non-MT branch; threads>1: IndexThreadReduce
 * semantics (Include/IndexThreadReduce.h:70-207) on a std::thread pool. */
int orc_ba_set_threads(orc_ba* ba, int threads);
int orc_ba_linearize(orc_ba* ba, int fixLinearization, double* energy_sum, int8_t* newState,
                     float* newEnergy, float* newEnergyWithOutlier, float* centerProjectedTo,
                     float* projectedTo);
int orc_ba_apply_res(orc_ba* ba);
int orc_ba_fix_linearization(orc_ba* ba, const int32_t* ids, int count);
int orc_ba_accumulate(orc_ba* ba, int which, double* H, double* b, int* nres);
int orc_ba_point_results(orc_ba* ba, double* HdiF, double* bdSumF, double* idepth_hessian,
                         double* Hdd_accAF, double* bd_accAF, double* Hcd_accAF,
                         double* Hdd_accLF, double* bd_accLF, double* Hcd_accLF);
int orc_ba_residual_results(orc_ba* ba, int8_t* state_state, uint8_t* is_active, double* JpJdF,
                            double* J74, double* res_toZeroF);
/* Raw relative accumulators before stitching, for fine-grained parity:
 * top: n*n blocks of 13*13 doubles (col-major) for mode `which` (0/1);
 * sc : accD n*n*n*64, accE n*n*32, accEB n*n*8, accHcc 16, accbc 4. */
int orc_ba_raw_top(orc_ba* ba, int which, double* acc1313);
int orc_ba_raw_sc(orc_ba* ba, double* accD, double* accE, double* accEB, double* accHcc, double* accbc);
int orc_ba_resubstitute(orc_ba* ba, const double* x, float* step_out);
/* EnergyFunctional::marginalizePointsF (Src/EnergyFunctional.cpp:545-609), accumulate + stitch part, over the listed
 * points (caller order ids): priorF *= prior_scale, addPoint<2>, SC addPoint(p,false), stitchDouble(usePrior=false).
 * Outputs col-major (8n+4)^2 / (8n+4): M, Mb (top), Msc, Mbsc (Schur); nres = accSSE_top_A->nres[0]. */
int orc_ba_marginalize_points(orc_ba* ba, const int32_t* point_ids, int count, float prior_scale, double* M, double* Mb,
                              double* Msc, double* Mbsc, int* nres);
/* EnergyFunctional::calcLEnergyF_MT (Src/EnergyFunctional.cpp:349-364). */
int orc_ba_calc_l_energy(orc_ba* ba, double* out);
/* System::setNewFrameEnergyTH (Src/FullSystemOptimize.cpp:60-101) after a linearisation pass: std::nth_element at
 * (int)(setting_frameEnergyTHN * count) over state_NewEnergyWithOutlier of the non-linearised residuals into frame
 * `newest`; also stores the threshold as the window's frameEnergyTH[newest]. */
int orc_ba_new_frame_energy_th(orc_ba* ba, int newest, float* th_out);

/* ---- immature point tracing ---- */
#define ORC_IPS_GOOD 0
#define ORC_IPS_OOB 1
#define ORC_IPS_OUTLIER 2
#define ORC_IPS_SKIPPED 3
#define ORC_IPS_BADCONDITION 4
#define ORC_IPS_UNINITIALIZED 5
typedef struct orc_trace_settings {   /* Src/Settings.cpp:68,85-93 */
  float maxPixSearch, slackInterval, stepsize, minImprovementFactor, GNThreshold, extraSlackOnTH, huberTH;
  int minTraceTestRadius, GNIterations;
} orc_trace_settings;
/* ImmaturePoint::traceOn (Src/ImmaturePoint.cpp:40-351) for n points against one frame (level-0 Vec3f image). */
int orc_trace_immature(const float* dI0, int w, int h, const orc_trace_settings* S, int n_hosts, const float* KRKi,
                       const float* Kt, const float* aff, int n, const int32_t* host_idx, const float* u, const float* v,
                       const float* gradH, const float* color, const float* weights, const float* energyTH,
                       float* idepth_min_io, float* idepth_max_io, float* quality_io, int8_t* status_io,
                       float* lastTraceUV, float* lastTracePixelInterval);

/* System::optimizeImmaturePoint (Src/FullSystemOptPoint.cpp:24-175) + ImmaturePoint::linearizeResidual
 * (Src/ImmaturePoint.cpp:389-451).  RtAff: per ordered pair (host + target*n) 14 floats: PRE_RTll row-major, PRE_tTll,
 * PRE_aff_mode.  res_state_out: per point n_frames states (HSLAM_RES_*; the host's own slot is OOB). */
int orc_activate_immature(int n_frames, const float* const* dI0, int w, int h, float fx, float fy, float cx, float cy, float fxi,
                          float fyi, const float* RtAff, float huber, float minIdepthH_act, int GNIts, int minObs, int n,
                          const int32_t* host_idx, const float* u, const float* v, const float* color, const float* weights,
                          const float* energyTH, const float* idepth_min, const float* idepth_max, int8_t* ok_out,
                          float* idepth_out, int8_t* res_state_out);

#ifdef __cplusplus
}
#endif
#endif
