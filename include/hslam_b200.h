/*
 * hslam_b200.h — C ABI of the B200-native photometric Gauss-Newton hot path.
 *
 * This is the drop-in boundary for the DSO-style hot path of AUBVRL/H-SLAM
 * (namespace FSLAM in the reference).  The reference has no FFI of its own
 * (SURVEY.md §8b): the path is reached through C++ member functions.  Every
 * entry point below therefore cites the reference member function (file:line,
 * relative to the reference root) whose work it replaces; the C++ host shim in
 * h-slam_b200/host/ keeps those member-function names and forwards here.
 *
 * Conventions
 *  - POD only, no C++/torch types.  Every function returns 0 on success and a
 *    negative hslam_status on failure; hslam_last_error() (thread-local) holds
 *    the message.  Nothing throws across this boundary.
 *  - There is NO CPU fallback: if no CUDA device is usable every compute entry
 *    point returns HSLAM_ERR_CUDA.
 *  - A handle is single-threaded; different handles may be used concurrently
 *    from different host threads (the reference runs two CoarseTracker
 *    instances on two threads, Src/System.cpp:127-133).
 *  - Matrices crossing the boundary are COLUMN-MAJOR (Eigen default), except
 *    rigid transforms which are 12 doubles: R row-major (9) then t (3).
 *  - "host pointer" / "device pointer" is stated per argument.
 */
#ifndef HSLAM_B200_H
#define HSLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSLAM_ABI_VERSION 2
#define HSLAM_MAX_PYR_LEVELS 10   /* CoarseTracker::PYR_LEVELS, Include/CoarseTracker.h:32 */
#define HSLAM_PATTERN_NUM 8       /* ChosenPattern 8, Include/GlobalTypes.h:33,181 */
#define HSLAM_CPARS 4             /* Include/GlobalTypes.h:223 */
#define HSLAM_MAX_FRAMES 16       /* window size cap of this build (reference default 7, Src/Settings.cpp:58) */
#define HSLAM_PRECALC_STRIDE 28   /* floats per (host,target) record, see hslam_window.precalc */
#define HSLAM_J_FLOATS 74         /* sizeof(RawResidualJacobian)/4, Include/OptimizationClasses.h:48-77 */

typedef enum {
  HSLAM_OK = 0,
  HSLAM_ERR_ARG = -1,      /* bad argument (null, size, level out of range) */
  HSLAM_ERR_CUDA = -2,     /* CUDA runtime error or no device */
  HSLAM_ERR_STATE = -3,    /* call order violated (e.g. accumulate before set_window) */
  HSLAM_ERR_NOMEM = -4,
  HSLAM_ERR_NCCL = -5
} hslam_status;

/* ResState, Include/GlobalTypes.h (enum ResState {IN=0, OOB, OUT}) */
#define HSLAM_RES_IN 0
#define HSLAM_RES_OOB 1
#define HSLAM_RES_OUT 2

/* Settings consumed by the hot path (Src/Settings.cpp; SURVEY.md §5 "Config"). */
typedef struct hslam_params {
  float huberTH;                 /* setting_huberTH = 9                 (Settings.cpp:68)  */
  float coarseCutoffTH;          /* setting_coarseCutoffTH = 20         (Settings.cpp:77)  */
  float outlierTHSumComponent;   /* setting_outlierTHSumComponent=2500  (Settings.cpp:64)  */
  float affineOptModeA;          /* 1e12  (Settings.cpp:109); <0 fixes a */
  float affineOptModeB;          /* 1e8   (Settings.cpp:110); <0 fixes b */
  float scale_xi_rot;            /* SCALE_XI_ROT   = 1    (GlobalTypes.h:35) */
  float scale_xi_trans;          /* SCALE_XI_TRANS = 0.5  (GlobalTypes.h:36) */
  float scale_a;                 /* SCALE_A = 10          (GlobalTypes.h:40) */
  float scale_b;                 /* SCALE_B = 1000        (GlobalTypes.h:41) */
  float scale_f;                 /* SCALE_F = 50          (GlobalTypes.h:37) */
  float scale_c;                 /* SCALE_C = 50          (GlobalTypes.h:38) */
  float scale_idepth;            /* SCALE_IDEPTH = 1      (GlobalTypes.h:34) */
  int   max_tracker_iterations[5]; /* {10,20,50,50,50}    (CoarseTracker.cpp:518) */
  /* System::setNewFrameEnergyTH (Src/FullSystemOptimize.cpp:60-101) */
  float frameEnergyTHN;          /* setting_frameEnergyTHN = 0.7          (Settings.cpp:73) */
  float frameEnergyTHFacMedian;  /* setting_frameEnergyTHFacMedian = 1.5  (Settings.cpp:74) */
  float frameEnergyTHConstWeight;/* setting_frameEnergyTHConstWeight = 0.5(Settings.cpp:75) */
  float overallEnergyTHWeight;   /* setting_overallEnergyTHWeight = 1     (Settings.cpp:66) */
} hslam_params;

typedef struct hslam_ctx hslam_ctx;
typedef struct hslam_frame hslam_frame;
typedef struct hslam_tracker hslam_tracker;
typedef struct hslam_ba hslam_ba;

const char* hslam_last_error(void);
int  hslam_abi_version(void);
void hslam_params_default(hslam_params* p);

/* One context = one CUDA device + one stream.  Replaces the three
 * IndexThreadReduce pools of System (Src/System.cpp:33-34,51). */
int hslam_create(const hslam_params* params /*nullable => defaults*/, int device, hslam_ctx** out);
int hslam_destroy(hslam_ctx* ctx);
int hslam_synchronize(hslam_ctx* ctx);
/* cudaStream_t of the context, for callers that time or interoperate. */
int hslam_get_stream(hslam_ctx* ctx, void** cuda_stream_out);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t hslam_launch_count(hslam_ctx* ctx);
/* Page-locked host memory owned by the context.  Buffers that cross the boundary every iteration
 * (the Eigen storage behind MatXX H / VecX b, the per-residual state arrays) can be placed here:
 * copies to and from such memory are direct DMA, any other host pointer is staged through an
 * internal pinned block and memcpy'd.  Freed by hslam_host_free or hslam_destroy. */
int hslam_host_alloc(hslam_ctx* ctx, size_t bytes, void** out);
int hslam_host_free(hslam_ctx* ctx, void* p);
/* Per-kernel device timing with CUDA events on the context stream (measurement only).
 * enable=1 starts bracketing every launch with a pair of events; hslam_profile_read
 * synchronises, returns for each kernel class its summed duration (ms) and launch count
 * since the last read and clears the counters.  Classes: see HSLAM_K_* below. */
#define HSLAM_K_PYRAMID 0
#define HSLAM_K_TRK_EVAL 1
#define HSLAM_K_TRK_TRACK 2
#define HSLAM_K_BA_LINEARIZE 3
#define HSLAM_K_BA_APPLY 4
#define HSLAM_K_BA_TOP 5
#define HSLAM_K_BA_FOLD_TOP 6
#define HSLAM_K_BA_POINT_SUM 7
#define HSLAM_K_BA_SCHUR 8
#define HSLAM_K_BA_FOLD_SC 9
#define HSLAM_K_BA_STITCH 10
#define HSLAM_K_BA_RESUB 11
#define HSLAM_K_BA_FIXLIN 12
#define HSLAM_K_UNPACK 13
#define HSLAM_K_BA_ENERGY 14
#define HSLAM_K_BA_EXPORT 15
#define HSLAM_K_TRACE 16
#define HSLAM_K_COARSE_DEPTH 17
#define HSLAM_K_BA_ENERGY_TH 18
#define HSLAM_K_COUNT 19
int hslam_profile_enable(hslam_ctx* ctx, int enable);
int hslam_profile_read(hslam_ctx* ctx, float ms_out[HSLAM_K_COUNT], int count_out[HSLAM_K_COUNT]);
const char* hslam_kernel_name(int k);

/* ---------------------------------------------------------------- frames ---
 * Replaces Frame::CreateDirPyrs (Src/Frame.cpp:104-181): level 0 copies the
 * intensities, level l>0 is the 2x2 box mean of level l-1, dx/dy are central
 * differences over flat indices w <= i < w*(h-1) (row ends wrap exactly as in
 * the reference), absSquaredGrad = dx^2+dy^2 (x gammaBGrad[I]^2 if given).
 * Rows 0 and h-1, which the reference leaves uninitialised, hold dx=dy=0.
 * Device layout: one float4 {I, dx, dy, absSquaredGrad} per pixel per level.
 * img: w*h float32 intensities in [0,255] (photometrically corrected). */
int hslam_frame_create(hslam_ctx* ctx, const float* img_host, int w, int h, int levels,
                       const float* gammaBGrad256_host /*nullable*/, hslam_frame** out);
/* Same, input already resident in HBM (device pointer). */
int hslam_frame_create_from_device(hslam_ctx* ctx, const float* img_dev, int w, int h, int levels,
                                   const float* gammaBGrad256_host /*nullable*/, hslam_frame** out);
/* Re-use an existing frame's buffers for a new image of the same size. */
int hslam_frame_update(hslam_frame* f, const float* img, int img_is_device);
/* For host code that still reads DirPyr / absSquaredGrad: AoS Vec3f (12 B
 * stride) and float planes exactly as the reference stores them. */
int hslam_frame_download(hslam_frame* f, int lvl, float* dI_xyz_host /*w*h*3, nullable*/,
                         float* absSqGrad_host /*w*h, nullable*/);
int hslam_frame_level_size(hslam_frame* f, int lvl, int* w, int* h);
int hslam_frame_destroy(hslam_frame* f);

/* --------------------------------------------------------------- tracker ---
 * One per CoarseTracker instance (Include/CoarseTracker.h:14-95). */
int hslam_tracker_create(hslam_ctx* ctx, int w, int h, hslam_tracker** out);
int hslam_tracker_destroy(hslam_tracker* trk);
/* How hslam_tracker_track runs the LM loop: cluster_ctas = 8 or 16 -> ONE thread-block cluster runs the whole frame
 * (cluster barrier + distributed shared memory between evaluations, the other SMs stay free for more trackers);
 * 0 -> cooperative kernel over up to two blocks per SM with a grid barrier per evaluation.  Default: the largest cluster
 * the device can schedule (environment HSLAM_TRACKER_MODE=coop selects 0).  Same arithmetic per point in both; the
 * partial sums are grouped per block, so H,b agree to fp32 round-off between the modes. */
int hslam_tracker_set_mode(hslam_tracker* trk, int cluster_ctas);
int hslam_tracker_get_mode(hslam_tracker* trk, int* cluster_ctas);
/* Diagnostics of the last hslam_tracker_track: SM clock cycles block 0 spent in {solve + SE3 exp + constants,
 * evaluation, evaluation + barrier + fold, accept + normal equations}, and the number of evaluations. */
int hslam_tracker_last_profile(hslam_tracker* trk, long long phase_cycles[4], int* evals);
/* CoarseTracker::makeK (Src/CoarseTracker.cpp:72-101).  The caller passes the
 * level-0 intrinsics (HCalib->fxl() ...); levels 1.. are derived by the
 * reference rule.  Ki = K^-1 is derived with Eigen's 3x3 cofactor formula. */
int hslam_tracker_make_K(hslam_tracker* trk, int levels, int w0, int h0,
                         float fx0, float fy0, float cx0, float cy0);
/* Read back the per-level intrinsics (fx,fy,cx,cy,w,h,Ki row-major 9). */
int hslam_tracker_get_K(hslam_tracker* trk, int lvl, float out_fxfycxcy[4], int out_wh[2], float out_Ki[9]);
/* Upload the reference point cloud pc_u/pc_v/pc_idepth/pc_color of one level
 * (output of makeCoarseDepthL0, Src/CoarseTracker.cpp:216-261). Host pointers. */
int hslam_tracker_set_ref(hslam_tracker* trk, int lvl, const float* u, const float* v,
                          const float* idepth, const float* color, int n);
/* lastRef->ab_exposure and lastRef_aff_g2l (Src/CoarseTracker.cpp:496-500). */
int hslam_tracker_set_ref_aff(hslam_tracker* trk, float ref_exposure, double ref_aff_a, double ref_aff_b);
/* CoarseTracker::makeCoarseDepthL0 on the device (Src/CoarseTracker.cpp:105-263): builds the
 * per-level reference clouds pc_* of this tracker from the active points that are IN the newest keyframe `lastRef`.
 * Per contributing point (host arrays): centerProjectedTo (3 floats: u, v, new_idepth of its residual into lastRef,
 * CoarseTracker.cpp:119-121) and efPoint->HdiF (weight = sqrtf(1e-3/(HdiF+1e-12)), :122).  Replaces the host loop +
 * hslam_tracker_set_ref for every level; call hslam_tracker_set_ref_aff as before. */
int hslam_tracker_make_coarse_depth(hslam_tracker* trk, const hslam_frame* lastRef, int n_points,
                                    const float* centerProjectedTo /*n*3*/, const float* HdiF /*n*/);
/* The same, device-resident: the inputs are read from a window handle of the SAME context — for every listed
 * residual (caller ids: ph->lastResiduals[0] of each contributing point, state IN, target == lastRef) its
 * centerProjectedTo of the last linearisation and the HdiF of its point from the last Schur pass. */
int hslam_tracker_make_coarse_depth_from_ba(hslam_tracker* trk, const hslam_frame* lastRef, hslam_ba* ba,
                                            const int32_t* residual_ids, int count);
/* The reference cloud of a level as the tracker holds it (pc_u, pc_v, pc_idepth, pc_color, pc_n). */
int hslam_tracker_get_ref(hslam_tracker* trk, int lvl, int capacity, float* u, float* v, float* idepth, float* color, int* n_out);

/* CoarseTracker::calcRes (Src/CoarseTracker.cpp:329-485).
 * out6 = {E, numTermsInE, sumSqShiftT/(n+0.1), 0, sumSqShiftRT/(n+0.1), numSaturated/numTermsInE}.
 * The warped buffers are not materialised: the same launch also reduces the
 * normal equations (see calc_gs). n_warped_out (nullable) receives buf_warped_n
 * BEFORE padding to a multiple of 4. */
int hslam_tracker_calc_res(hslam_tracker* trk, int lvl, const hslam_frame* newFrame,
                           const double refToNew[12], double aff_a, double aff_b,
                           float new_exposure, float cutoffTH, double out6[6], int* n_warped_out);
/* CoarseTracker::calcGSSSE (Src/CoarseTracker.cpp:267-324) evaluated on the
 * inliers of calcRes at the same (pose, aff, cutoffTH): H (8x8 col-major) and
 * b (8) with the 1/n_padded normalisation and SCALE_* scaling applied. */
int hslam_tracker_calc_gs(hslam_tracker* trk, int lvl, const hslam_frame* newFrame,
                          const double refToNew[12], double aff_a, double aff_b,
                          float new_exposure, float cutoffTH, double H[64], double b[8]);
/* CoarseTracker::trackNewestCoarse (Src/CoarseTracker.cpp:506-683): the whole
 * coarse-to-fine LM loop in ONE cooperative launch (8x8 LDLT, SE3 exp and the
 * accept/reject logic run on the device).  refToNew_io / aff_io are
 * lastToNew_out / aff_g2l_out.  lastResiduals[5] is NaN-initialised,
 * lastFlow[3] = lastFlowIndicators.  *ok receives the bool return value.
 * iterations_out (nullable, 5 ints) receives the LM iterations per level. */
int hslam_tracker_track(hslam_tracker* trk, const hslam_frame* newFrame, float new_exposure,
                        double refToNew_io[12], double aff_io[2], int coarsestLvl,
                        const double minResForAbort[5], double lastResiduals[5],
                        double lastFlow[3], int* ok, int* iterations_out);
/* Asynchronous pair of the above for batched/pipelined callers: enqueue does
 * the (pinned) H2D of the 14 state doubles and the launch; finish waits and
 * fills the outputs. */
int hslam_tracker_track_enqueue(hslam_tracker* trk, const hslam_frame* newFrame, float new_exposure,
                                const double refToNew[12], const double aff[2], int coarsestLvl,
                                const double minResForAbort[5]);
int hslam_tracker_track_finish(hslam_tracker* trk, double refToNew_out[12], double aff_out[2],
                               double lastResiduals[5], double lastFlow[3], int* ok,
                               int* iterations_out);

/* Batched form: `n` members — several incoming frames, or several initial poses (the motion hypotheses
 * lastF_2_fh_tries of System::trackNewCoarse, Src/System.cpp:348-481; pass the same frame n times) — are tracked against
 * this tracker's reference in ONE launch, one thread-block cluster per member (needs the cluster mode).  All arrays are
 * member-major: T_io n*12, aff_io n*2, new_exposure n (nullable: 1), lastResiduals n*5, lastFlow n*3, ok n,
 * iterations_out n*5 (each output nullable).  Every member's result is bit-identical to hslam_tracker_track of that
 * member alone; the caller applies the reference's sequential acceptance rule (System.cpp:428-481) to the results. */
int hslam_tracker_track_batch(hslam_tracker* trk, int n, const hslam_frame* const* frames, const float* new_exposure,
                              double* T_io, double* aff_io, int coarsestLvl, const double minResForAbort[5],
                              double* lastResiduals, double* lastFlow, int* ok, int* iterations_out);

/* -------------------------------------------------------------- window BA ---
 * Device-resident mirror of EnergyFunctional + its points/residuals
 * (Include/EnergyFunctional.h:30-121).  Re-pack only when the reference
 * invalidates EFIndicesValid / EFAdjointsValid / EFDeltaValid
 * (Src/EnergyFunctional.cpp:17-19). */

/* Per-window state; all pointers are HOST pointers, copied on the call.
 * Pair index convention everywhere: idx = host + target*n_frames
 * (Src/EnergyFunctional.cpp:60, AccumulatedTopHessian.cpp:53). */
typedef struct hslam_window {
  int n_frames;
  int w, h;                      /* HCalib->Width/Height */
  float fx, fy, cx, cy;          /* HCalib->fxl(),fyl(),cxl(),cyl()   (Include/CalibData.h:93-96) */
  float fxi, fyi;                /* HCalib->fxli(),fyli()             (Include/CalibData.h:97-98) */
  const hslam_frame* const* frames;  /* n_frames, index = Frame::idx */
  const float* frameEnergyTH;    /* n_frames, Frame::frameEnergyTH */
  /* FrameFramePrecalc (Src/OptimizationClasses.cpp:13-39), HSLAM_PRECALC_STRIDE
   * floats per pair: [0..8] PRE_KRKiTll row-major, [9..11] PRE_KtTll,
   * [12..20] PRE_RTll_0 row-major, [21..23] PRE_tTll_0, [24..25] PRE_aff_mode,
   * [26] PRE_b0_mode, [27] unused. */
  const float* precalc;          /* n*n*28 */
  const double* adHost;          /* n*n*64 col-major, EnergyFunctional::adHost   (EnergyFunctional.cpp:60) */
  const double* adTarget;        /* n*n*64 col-major, EnergyFunctional::adTarget (EnergyFunctional.cpp:61) */
  const float* adHTdeltaF;       /* n*n*8,  EnergyFunctional::adHTdeltaF         (EnergyFunctional.cpp:136) */
  float  cDeltaF[4];             /* EnergyFunctional::cDeltaF (EnergyFunctional.cpp:140) */
  double cPrior[4];              /* EnergyFunctional::cPrior  (EnergyFunctional.cpp:63)  */
  const double* framePrior;      /* n*8, FrameOptimizationData::prior       (Include/Frame.h:265) */
  const double* frameDeltaPrior; /* n*8, FrameOptimizationData::delta_prior (Include/Frame.h:266) */
} hslam_window;

/* MapPoint + MapPointOptimizationData flattened (Include/MapPoint.h:16-162).
 * HOST pointers.  res_begin/res_index give each point's `residuals` vector in
 * the point's own order (the order the reference iterates, it fixes the float
 * summation order of Hdd/bd/Hcd, AccumulatedTopHessian.cpp:35-120). */
typedef struct hslam_point_soa {
  int n_points;
  const float* u;            /* P */
  const float* v;            /* P */
  const float* idepth;       /* P   MapPoint::idepth */
  const float* idepth_zero;  /* P   efPoint->idepth_zero */
  const float* color;        /* P*8 MapPoint::color,   point-major */
  const float* weights;      /* P*8 MapPoint::weights, point-major */
  const float* priorF;       /* P   efPoint->priorF */
  const float* deltaF;       /* P   efPoint->deltaF (EnergyFunctional.cpp:148) */
  const int32_t* host_idx;   /* P   host->frame->idx */
  const int32_t* res_begin;  /* P+1 CSR offsets into res_index */
  const int32_t* res_index;  /* residual ids (index into hslam_residual_soa) */
} hslam_point_soa;

/* PointFrameResidual flattened (Include/OptimizationClasses.h:79-163). HOST pointers. */
typedef struct hslam_residual_soa {
  int n_residuals;
  const int32_t* point_idx;    /* R */
  const int32_t* target_idx;   /* R  targetIDX (hostIDX comes from the point) */
  const int8_t*  state_state;  /* R  HSLAM_RES_* */
  const uint8_t* is_linearized;/* R  PointFrameResidual::isLinearized */
  const uint8_t* is_active;    /* R  isActiveAndIsGoodNEW */
  const double*  state_energy; /* R  nullable (zeros) */
} hslam_residual_soa;

int hslam_ba_create(hslam_ctx* ctx, hslam_ba** out);
int hslam_ba_destroy(hslam_ba* ba);
/* setAdjointsF/setDeltaF/setPrecalcValues results (EnergyFunctional.cpp:22-82,
 * 128-152; System.cpp:321-331).  Cheap: O(n^2) floats. */
int hslam_ba_set_window(hslam_ba* ba, const hslam_window* win);
/* Topology: call after makeIDX / insertResidual / dropResidual / removePoint
 * (EnergyFunctional.cpp:819-840,371-380,428-454).  Residuals are re-ordered on
 * the device by (host,target) pair; all per-residual outputs are returned in
 * the caller's order. */
int hslam_ba_set_topology(hslam_ba* ba, const hslam_point_soa* pts, const hslam_residual_soa* res);
/* Per-iteration point state only (idepth, deltaF change in doStepFromBackup /
 * setDeltaF; FullSystemOptimize.cpp:171-264). Host pointers, P floats each. */
int hslam_ba_update_points(hslam_ba* ba, const float* idepth, const float* deltaF,
                           const float* idepth_zero /*nullable: unchanged*/);

/* System::linearizeAll_Reductor -> PointFrameResidual::linearize
 * (FullSystemOptimize.cpp:19-52,114; OptimizationClasses.cpp:43-233) over every
 * residual that is not linearised.  If fixLinearization != 0 applyRes(true) is
 * done in the same pass (FullSystemOptimize.cpp:26-28).  Outputs (host, caller
 * order, each nullable): energy_sum = stats[0]; newState/newEnergy/
 * newEnergyWithOutlier = state_NewState/state_NewEnergy/state_NewEnergyWithOutlier;
 * centerProjectedTo R*3; projectedTo R*16. */
int hslam_ba_linearize(hslam_ba* ba, int fixLinearization, double* energy_sum,
                       int8_t* newState, float* newEnergy, float* newEnergyWithOutlier,
                       float* centerProjectedTo, float* projectedTo);
/* What the lineariser keeps of RawResidualJacobian (Include/OptimizationClasses.h:48-77).  0 (default): all 74 floats +
 * projectedTo, as the reference.  1, "GN-loop mode": the per-pixel rows resF[8], JIdx[2][8], JabF[2][8] (160 B) and
 * projectedTo[8] (64 B, written at OptimizationClasses.cpp:150-151 and read nowhere) are NOT stored for the residuals
 * that are re-linearised every iteration — accumulate{A,SC}F, applyRes/takeData and resubstitute only read the 34
 * geometric / summed floats.  Their only readers, fixLinearizationF, addPoint<1>/<2> and calcLEnergyPt, work on
 * linearised residuals: hslam_ba_fix_linearization re-forms the rows for the residuals it fixes, at the window's
 * current state (the reference re-linearises right before fixLinearizationF as well, Src/Mapping.cpp:286-293).  In mode
 * 1 hslam_ba_residual_results returns stale per-pixel rows for non-linearised residuals and hslam_ba_linearize
 * returns stale projectedTo. */
int hslam_ba_set_jacobian_mode(hslam_ba* ba, int mode);
/* System::setNewFrameEnergyTH (Src/FullSystemOptimize.cpp:60-101), which System::linearizeAll calls after EVERY
 * linearisation pass (:124): over the non-linearised residuals whose target is the newest keyframe and whose
 * state_NewEnergyWithOutlier >= 0, the element std::nth_element puts at index (int)(setting_frameEnergyTHN * count)
 * (an exact k-th order statistic: radix select on the device, no sort, no host copy of the energies), then
 * th = sqrtf(nth) * FacMedian; th = 26 * ConstWeight + th * (1 - ConstWeight); th = th * th * overallWeight^2 in
 * exactly-rounded fp32; 12*12*patternNum when there is no such residual (:73-77).  The settings come from the
 * context's hslam_params.  Like the reference, the call also STORES the value as the window's
 * frameEnergyTH[newest_frame_idx] (what the next linearisation reads); th_out (host, nullable) receives it.
 * The per-residual energies never leave HBM. */
int hslam_ba_new_frame_energy_th(hslam_ba* ba, int newest_frame_idx, float* th_out);
/* System::applyRes_Reductor -> PointFrameResidual::applyRes(true) + takeData
 * (FullSystemOptimize.cpp:55-59; OptimizationClasses.cpp:235-256, .h:155-161). */
int hslam_ba_apply_res(hslam_ba* ba);
/* PointFrameResidual::fixLinearizationF (OptimizationClasses.cpp:258-284) on
 * the listed residual ids (host pointer): res_toZeroF = resF - J*delta,
 * isLinearized = true. */
int hslam_ba_fix_linearization(hslam_ba* ba, const int32_t* residual_ids, int count);

#define HSLAM_ACC_A   0   /* EnergyFunctional::accumulateAF_MT  (EnergyFunctional.cpp:155-175) */
#define HSLAM_ACC_L   1   /* EnergyFunctional::accumulateLF_MT  (EnergyFunctional.cpp:178-197) */
#define HSLAM_ACC_SC  2   /* EnergyFunctional::accumulateSCF_MT (EnergyFunctional.cpp:203-220) */
/* addPoint<mode> + stitchDoubleMT for the chosen system.  H is (8n+4)^2
 * col-major, b is (8n+4); host pointers.  nres (nullable) = resInA / resInL.
 * SC requires A and L to have been accumulated for the current state (it reads
 * Hdd/bd/Hcd_acc{A,L}F exactly like AccumulatedSCHessian.cpp:23-31). */
int hslam_ba_accumulate(hslam_ba* ba, int which, double* H, double* b, int* nres);
/* The three systems in one pass (what solveSystemF needs,
 * EnergyFunctional.cpp:717-724): Hb_out = [H_A|b_A|H_L|b_L|H_SC|b_SC], each H
 * (8n+4)^2 col-major followed by its b. nres2 (nullable) = {resInA,resInL}. */
int hslam_ba_accumulate_all(hslam_ba* ba, double* Hb_out_host, int* nres2);
/* Device-resident variants: results stay in HBM (device pointer to
 * 3*((8n+4)^2+(8n+4)) doubles owned by the handle) for an all-reduce. */
int hslam_ba_accumulate_all_device(hslam_ba* ba, double** Hb_dev_out, int* count_out);
/* One fused GN-iteration pass: linearize(fixLinearization=0) + applyRes +
 * accumulate_all, enqueued back-to-back with no host round trip. */
int hslam_ba_iteration_device(hslam_ba* ba, double** Hb_dev_out, int* count_out);
/* Asynchronous use of a handle: between begin_async and wait every call only ENQUEUES its copies
 * and kernels (host output buffers are filled, and must stay valid, until hslam_ba_wait returns).
 * Without begin_async each call that returns data synchronises once before returning. */
int hslam_ba_begin_async(hslam_ba* ba);
int hslam_ba_wait(hslam_ba* ba);
/* A batch of independent windows of one context (BASELINE.json configs[4]: "batched 64 independent
 * synthetic windows"): one GN iteration (linearizeAll + applyRes + accumulateAF/LF/SCF) of EVERY
 * window in FOUR launches; results stay in each window's device buffers (hslam_ba_download_Hb,
 * hslam_ba_point_results ...).  Replaces running System::optimize's inner loop window by window. */
typedef struct hslam_ba_batch hslam_ba_batch;
int hslam_ba_batch_create(hslam_ctx* ctx, hslam_ba* const* windows, int count, hslam_ba_batch** out);
int hslam_ba_batch_iteration(hslam_ba_batch* batch);
int hslam_ba_batch_destroy(hslam_ba_batch* batch);
/* The same GN iteration with HOST buffers, for every window of the batch in one call (the form
 * System::optimize's loop body takes when the windows' state lives on the host: setPrecalcValues /
 * setDeltaF results up, linearizeAll + applyRes + accumulate{A,L,SC}F_MT, energies / residual states /
 * the three stitched systems down; FullSystemOptimize.cpp:380-497, EnergyFunctional.cpp:717-724).
 * All inputs of all windows travel in ONE host->device copy, all outputs in one device->host copy
 * (or straight into caller buffers obtained from hslam_host_alloc); the call returns when the outputs
 * are valid.  io[i] belongs to the i-th window of the batch; null members are skipped. */
typedef struct hslam_ba_step_io {
  /* inputs (host pointers, caller order) */
  const hslam_window* window;      /* tables + scalars, as hslam_ba_set_window (n_frames must not change) */
  const float* idepth;             /* P */
  const float* deltaF;             /* P */
  const float* idepth_zero;        /* P */
  /* outputs (host pointers, caller order) */
  double* energy;                  /* 1: stats[0] of linearizeAll */
  int8_t* newState;                /* R */
  float* newEnergy;                /* R */
  float* newEnergyWithOutlier;     /* R */
  double* Hb;                      /* [H_A|b_A|H_L|b_L|H_SC|b_SC] as hslam_ba_accumulate_all */
  int32_t* nres2;                  /* {resInA, resInL} */
  /* System::setNewFrameEnergyTH on the device after the linearisation (see hslam_ba_new_frame_energy_th): when
   * frameEnergyTH_out is not null, the new threshold of frame `newest_frame` is computed, stored in the window and
   * returned here (1 float) — the caller then needs neither newEnergy nor newEnergyWithOutlier on the host. */
  float* frameEnergyTH_out;
  int32_t newest_frame;
  /* non-zero, with idepth given and deltaF null: the device forms deltaF = idepth - idepth_zero itself (the point part
   * of EnergyFunctional::setDeltaF, EnergyFunctional.cpp:148; one exactly-rounded float subtraction, bit-identical to the
   * host's), so the deltaF array does not cross PCIe. */
  int32_t derive_deltaF;
} hslam_ba_step_io;
int hslam_ba_batch_step_host(hslam_ba_batch* batch, const hslam_ba_step_io* io /* one per window */);
/* Zero-copy buffers: point every OUTPUT member of io[0..count) (and step_out[i], if not null) at the window's
 * slot in the batch's own page-locked result block, and the per-point INPUT members idepth / deltaF at their slots
 * of the page-locked upload block (the caller writes the new point state there in place; idepth_zero is set to
 * null = unchanged).  Results then arrive with two bulk copies per call and
 * no per-array work; the caller reads them in place (e.g. Eigen::Map).  The pointers stay valid until a
 * member's topology changes or the batch is destroyed. */
int hslam_ba_batch_io_bind(hslam_ba_batch* batch, hslam_ba_step_io* io, float** step_out /*nullable*/);
/* Zero-copy window tables: point the table members of `w` (precalc, adHTdeltaF, adHost, adTarget, framePrior,
 * frameDeltaPrior) at window i's slots of the batch's page-locked upload block.  The caller's setPrecalcValues /
 * setAdjointsF / setDeltaF then write the tables THERE (e.g. through Eigen::Map), and hslam_ba_batch_step_host uploads
 * them without a staging copy.  Scalars, frames and frameEnergyTH of `w` stay the caller's.  Valid until a member's
 * topology changes or the batch is destroyed. */
int hslam_ba_batch_window_bind(hslam_ba_batch* batch, int i, hslam_window* w);
/* Device-resident GN loop: with newest_frame >= 0 every hslam_ba_batch_iteration also runs System::setNewFrameEnergyTH
 * (FullSystemOptimize.cpp:60-101; part of linearizeAll, :124) for frame `newest_frame` of every window: the new threshold
 * is written into the window's device-side frameEnergyTH table, where the next linearisation reads it (it is replaced by
 * the caller's table at the next hslam_ba_set_window / step_host with a window).  -1 switches it off (default). */
int hslam_ba_batch_set_newest_frame(hslam_ba_batch* batch, int newest_frame);
/* EnergyFunctional::resubstituteF_MT for every window of the batch: x[i] = (8n+4) host doubles,
 * step_out[i] = P host floats (nullable). */
int hslam_ba_batch_resubstitute_host(hslam_ba_batch* batch, const double* const* x, float* const* step_out);
/* Copy the device Hb block to the host (after hslam_ba_allreduce / an external all-reduce). */
int hslam_ba_download_Hb(hslam_ba* ba, double* Hb_out_host);
/* Energy (stats[0] of the last fused linearise pass) and {resInA, resInL} as the device holds them (summed over the
 * ranks after hslam_ba_allreduce).  Host pointers, each nullable. */
int hslam_ba_download_stats(hslam_ba* ba, double* energy, int* nres2);
/* Sharded window (BASELINE.json configs[4]: one window's points split over the GPUs of a box).  Every
 * rank runs the hot path on its shard (hslam_ba_iteration_device / hslam_ba_accumulate_all_device) and then sums the
 * stitched [H_A|b_A|H_L|b_L|H_SC|b_SC] block — plus energy, resInA, resInL — across ranks with ONE
 * ncclAllReduce(sum, fp64) on the context stream; afterwards every rank's block (hslam_ba_download_Hb) and its energy /
 * counts are the window's.  CONTRACT: the priors (cPrior, frame priors) enter the L system on every rank that has
 * hslam_ba_set_add_priors(ba, 1) — exactly ONE rank of the communicator must (AccumulatedTopHessian.cpp:268 "only do
 * this on one thread").  System::setNewFrameEnergyTH of a sharded window needs the per-residual energies of all
 * shards: request them through hslam_ba_linearize on every rank (hslam_ba_new_frame_energy_th is per shard).
 * nccl_comm is an ncclComm_t.  libnccl is resolved at run time (no link-time dependency): the library named by the
 * environment variable HSLAM_NCCL_LIB, else the libnccl already loaded in the process, else libnccl.so.2.  The
 * hslam_nccl_* helpers create the communicator through the same library (id from rank 0, distributed by the caller). */
int hslam_nccl_get_unique_id(char id_out[128]);
int hslam_nccl_comm_init(hslam_ctx* ctx, const char id[128], int n_ranks, int rank, void** nccl_comm_out);
int hslam_nccl_comm_destroy(void* nccl_comm);
int hslam_ba_allreduce(hslam_ba* ba, void* nccl_comm);
/* hslam_ba_iteration_device of this rank's shard (one CUDA graph launch) followed by hslam_ba_allreduce on the same
 * stream: what a rank runs per GN iteration of a sharded window. */
int hslam_ba_iteration_allreduce(hslam_ba* ba, void* nccl_comm);
/* How many thread blocks per stage this window should fill (0 = default, two waves over the SMs).  The
 * residual / point lists are cut into that many chunks, each reduced into one fp32 partial, the partials
 * are folded in chunk order in fp64: the value changes the grouping of the fp32 sums (results agree to
 * fp32 round-off, each setting is deterministic).  hslam_ba_batch_create sets it for members that still
 * have the default, so that the whole batch launches about two blocks per SM and stage. */
int hslam_ba_set_parallelism(hslam_ba* ba, int target_blocks);
int hslam_ba_get_parallelism(hslam_ba* ba, int* target_blocks_out);
/* In a sharded window only one rank may add the priors (cPrior, frame priors):
 * 1 (default) = add, 0 = skip. */
int hslam_ba_set_add_priors(hslam_ba* ba, int add_priors);

/* Side effects of addPoint on the points (AccumulatedTopHessian.cpp:122-139,
 * AccumulatedSCHessian.cpp:16-33), needed by resubstituteFPt / flagPointsForRemoval
 * / makeCoarseDepthL0.  Host pointers, caller order, each nullable. */
int hslam_ba_point_results(hslam_ba* ba, float* HdiF, float* bdSumF, float* idepth_hessian,
                           float* Hdd_accAF, float* bd_accAF, float* Hcd_accAF /*P*4*/,
                           float* Hdd_accLF, float* bd_accLF, float* Hcd_accLF /*P*4*/);
/* Per-residual state after applyRes: state_state, isActive, JpJdF (R*8),
 * and the raw Jacobian (R*74 in RawResidualJacobian member order: resF[8],
 * Jpdxi[2][6], Jpdc[2][4], Jpdd[2], JIdx[2][8], JabF[2][8], JIdx2[4], JabJIdx[4],
 * Jab2[4]; 2x2 blocks col-major). Each nullable. */
int hslam_ba_residual_results(hslam_ba* ba, int8_t* state_state, uint8_t* is_active,
                              float* JpJdF, float* J74, float* res_toZeroF /*R*8*/);

/* EnergyFunctional::marginalizePointsF (Src/EnergyFunctional.cpp:545-609), its accumulate + stitch part, for the
 * points listed in point_ids (caller order): priorF *= prior_scale (setting_idepthFixPriorMargFac),
 * AccumulatedTopHessianSSE::addPoint<2> and AccumulatedSCHessianSSE::addPoint(p, shiftPriorToZero=false) over
 * them, then stitchDouble(usePrior=false) of both (AccumulatedTopHessian.cpp:148-215, AccumulatedSCHessian.cpp:
 * 135-195).  Outputs (host, col-major, each nullable): M/Mb, Msc/Mbsc; the caller forms H = M - Msc, b = Mb - Mbsc
 * and keeps the HM/bM algebra.  nres = accSSE_top_A->nres[0] (resInM increment).  The residuals of these points
 * must have been fixed with hslam_ba_fix_linearization.   */
int hslam_ba_marginalize_points(hslam_ba* ba, const int32_t* point_ids, int count, float prior_scale,
                                double* M, double* Mb, double* Msc, double* Mbsc, int* nres);
/* EnergyFunctional::calcLEnergyF_MT / calcLEnergyPt (Src/EnergyFunctional.cpp:289-364): energy of the linearised
 * residuals and of the priors at the current deltas.   */
int hslam_ba_calc_l_energy(hslam_ba* ba, double* energy_out);
/* EnergyFunctional::resubstituteF_MT / resubstituteFPt
 * (EnergyFunctional.cpp:222-274): x is the (8n+4) solution; step_out[P] =
 * efPoint->step.   */
int hslam_ba_resubstitute(hslam_ba* ba, const double* x_host, float* step_out_host);

/* ---------------------------------------------------------------------------
 * Immature point tracing
 * ImmaturePoint::traceOn (Src/ImmaturePoint.cpp:40-351) for every immature point of the window against the newest
 * frame, as System::traceNewCoarse drives it (Src/Mapping.cpp:494-538): per host keyframe the caller passes
 * KRKi = K * R(host->frame) * K^-1 (row-major 3x3), Kt = K * t and the affine pair of fromToVecExposure
 * (Mapping.cpp:509-517), per point the members of ImmaturePoint (Include/ImmaturePoint.h:33-66). */
#define HSLAM_IPS_GOOD 0           /* enum ImmaturePointStatus, Include/ImmaturePoint.h:24-30 */
#define HSLAM_IPS_OOB 1
#define HSLAM_IPS_OUTLIER 2
#define HSLAM_IPS_SKIPPED 3
#define HSLAM_IPS_BADCONDITION 4
#define HSLAM_IPS_UNINITIALIZED 5
typedef struct hslam_trace_settings {      /* Src/Settings.cpp:68,85-93 */
  float maxPixSearch, slackInterval, stepsize, minImprovementFactor, GNThreshold, extraSlackOnTH, huberTH;
  int minTraceTestRadius, GNIterations;
} hslam_trace_settings;
void hslam_trace_settings_default(hslam_trace_settings* s);
typedef struct hslam_immature_soa {        /* HOST pointers; the last six arrays are read AND written */
  int n_points;
  const int32_t* host_idx;        /* n  index into the per-host tables */
  const float* u;                 /* n */
  const float* v;                 /* n */
  const float* gradH;             /* n*4 Mat22f col-major */
  const float* color;             /* n*8 */
  const float* weights;           /* n*8 */
  const float* energyTH;          /* n */
  float* idepth_min;              /* n */
  float* idepth_max;              /* n  (may be +inf / NaN as in the reference) */
  float* quality;                 /* n */
  int8_t* lastTraceStatus;        /* n  HSLAM_IPS_* */
  float* lastTraceUV;             /* n*2 */
  float* lastTracePixelInterval;  /* n */
} hslam_immature_soa;
int hslam_trace_immature(hslam_ctx* ctx, const hslam_frame* frame, const hslam_trace_settings* settings /*nullable*/,
                         int n_hosts, const float* KRKi /*n_hosts*9*/, const float* Kt /*n_hosts*3*/,
                         const float* aff /*n_hosts*2*/, const hslam_immature_soa* pts);

/* System::optimizeImmaturePoint (Src/FullSystemOptPoint.cpp:24-175) with ImmaturePoint::linearizeResidual
 * (Src/ImmaturePoint.cpp:389-451) for a list of activation candidates (System::activatePointsMT, Src/Mapping.cpp:330-480):
 * the 1-D Gauss-Newton on the inverse depth against every other keyframe.  RtAff: per ordered pair (host + target*n_frames)
 * 14 floats = FrameFramePrecalc::PRE_RTll (row-major 3x3), PRE_tTll, PRE_aff_mode (Src/OptimizationClasses.cpp:13-39).
 * Outputs per point: ok_out (1 = the reference would return a MapPoint), idepth_out (the optimised inverse depth:
 * setIdepthZero / setIdepth), res_state_out[n_frames] = final state of the temporary residual into each frame
 * (HSLAM_RES_IN => the caller creates the PointFrameResidual; the host's own slot is HSLAM_RES_OOB). */
typedef struct hslam_activate_soa {
  int n_points;
  const int32_t* host_idx;        /* n */
  const float* u;                 /* n */
  const float* v;                 /* n */
  const float* color;             /* n*8 */
  const float* weights;           /* n*8 */
  const float* energyTH;          /* n */
  const float* idepth_min;        /* n */
  const float* idepth_max;        /* n */
  int8_t* ok_out;                 /* n */
  float* idepth_out;              /* n */
  int8_t* res_state_out;          /* n*n_frames */
} hslam_activate_soa;
int hslam_activate_immature(hslam_ctx* ctx, int n_frames, const hslam_frame* const* frames, float fx, float fy, float cx,
                            float cy, float fxi, float fyi, const float* RtAff /*n*n*14*/,
                            float minIdepthH_act /*setting_minIdepthH_act = 100*/, int GNIts /*setting_GNItsOnPointActivation = 3*/,
                            int minObs, const hslam_activate_soa* pts);

#ifdef __cplusplus
}
#endif
#endif /* HSLAM_B200_H */
