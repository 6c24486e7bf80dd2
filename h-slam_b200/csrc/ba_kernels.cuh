// ba_kernels.cuh — device code of the windowed-BA hot path (included by ba.cu only).
//
// One Gauss-Newton iteration of one window — or of a whole BATCH of independent windows — is FOUR
// launches: linearise(+applyRes) -> top Hessian A and L (+ per-pair fold) -> Schur fold (+ per-point
// sums, + per-host fold) -> fp64 stitch.  Every kernel takes an array of per-window descriptors plus
// a block-prefix table, so a batch fills the 148 SMs with one grid per stage instead of W small ones.
//
// Replaces (reference file:line)
//   PointFrameResidual::linearize / applyRes / takeData / fixLinearizationF
//       Src/OptimizationClasses.cpp:43-284, Include/OptimizationClasses.h:155-161
//   AccumulatedTopHessianSSE::addPoint<mode> + AccumulatorApprox      Src/AccumulatedTopHessian.cpp:21-141
//   AccumulatedSCHessianSSE::addPoint + AccumulatorXX/X (Schur fold)  Src/AccumulatedSCHessian.cpp:10-53
//   stitchDoubleInternal / stitchDoubleMT (top and SC)                AccumulatedTopHessian.cpp:218-280, .h:69-117,
//                                                                     AccumulatedSCHessian.cpp:54-133, .h:70-111
//   EnergyFunctional::resubstituteFPt                                  Src/EnergyFunctional.cpp:249-274
//
// Determinism: no floating-point atomics.  Fixed-shape warp trees, fixed-order cross-warp sums,
// per-chunk partials folded BY CHUNK INDEX (the last-arriving block of a pair/host does the fold, the
// order of the sum does not depend on which block that is), gather-style stitch.
#pragma once
#include <cuda.h>   // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint)
#include "common.cuh"
#include <type_traits>

namespace hs {

constexpr int JS = 80;                 // floats per residual row: 74 of RawResidualJacobian + JI_r(2) Jab_r(2) rr(1) + pad
constexpr int J_RESF = 0, J_XI0 = 8, J_XI1 = 14, J_C0 = 20, J_C1 = 24, J_DD = 28, J_IDX0 = 30, J_IDX1 = 38,
              J_AB0 = 46, J_AB1 = 54, J_IDX2 = 62, J_ABIDX = 66, J_AB2 = 70, J_JIR = 74, J_JABR = 76, J_RR = 78;
// Lineariser launch shape, measured on B200 with the pair-image gather (64 cfg2 windows = 5.4 M residuals per launch):
//   64 threads x 8 blocks/SM (128 regs, no spills) 1015 us | 64 x 10 (96 regs) 1028 | 32 x 20 1026 | 128 x 5 1041 | 128 x 4 1047
//   64 x 9 / 96 x 7 (spills) 1180.   With the float4 image (four 16-byte loads per tap): 128 x 6 (80 regs) 1094, 64 x 12 1081.
// Round 2 (tile schedule, GN-loop mode, 3.76 M residuals per launch): 32 x 16 628 us | 64 x 8 635 | 128 x 4 642 — with one
// warp per tile a warp retires as soon as its own residuals are done and the last tile of a pair wastes less.
#ifndef LIN_MINB
#define LIN_MINB 16
#endif
#ifndef LIN_THREADS_
#define LIN_THREADS_ 32
#endif
constexpr int LIN_THREADS = LIN_THREADS_;
// Streaming (evict-first) stores for the big per-residual output streams: they are written once per
// iteration and are larger than what L2 can usefully keep, while the target images that the gathers hit
// (7 x 4.9 MB per window) should stay L2-resident.
#ifndef LIN_STREAM_STORES
#define LIN_STREAM_STORES 1
#endif
#if defined(LIN_EXP_NOSTORE)   /* timing experiment only: results are wrong */
#define STJ(ptr, v) do { if ((v) == 1.2345e-33f) *(ptr) = (v); } while (0)
#elif LIN_STREAM_STORES
#define STJ(ptr, v) __stcs((ptr), (v))
#else
#define STJ(ptr, v) (*(ptr) = (v))
#endif
constexpr int TOP_THREADS = 128;
constexpr int TOP_PART = 96;           // 91 accumulators + nres + pad (fp32 chunk partials)
constexpr int TOP_DENSE = 176;         // folded fp64 accumulator of a pair: dense symmetric 13x13 (169) + nres at [169] + pad
constexpr int SC_THREADS = 256;
#ifndef SC_PTS_
#define SC_PTS_ 64
#endif
constexpr int SC_PTS = SC_PTS_;        // points staged in shared memory per sub-tile (32 / 64 / 128)
// Stitch blocks: the sandwiches of an output block are dealt round-robin to ST_SLOTS = 8 accumulator SLOTS; a block of
// eight warps gives every warp one slot, a block of four warps two — the slots are summed in slot order, so the result
// does not depend on the warp count.  Measured on the 64-window batch (12 288 output blocks, most of them with two
// sandwiches): 16 warps 134 us | 8: 72 | 4: 56 | 2: 50; a window that runs alone (192 blocks, up to 38 sandwiches each)
// is latency-bound and wants the eight warps.  So: four warps for a batch, eight for a single window, same bits.
constexpr int ST_SLOTS = 8;

__constant__ int c_pat[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};

struct BAWin {
  int n, w, h;
  int n_rcp;                 // ceil(2^16 / n): pair / n == (pair * n_rcp) >> 16 for pair < n*n <= 256 (no integer division per residual)
  float fx, fy, cx, cy, fxi, fyi;
  float huber, outlierTH, scale_idepth, scale_f, scale_c, affA, affB;
  float cDeltaF[4];
  float thN, thFac, thConst, thOverall;   // setting_frameEnergyTH{N,FacMedian,ConstWeight}, setting_overallEnergyTHWeight
  const float4* img[HSLAM_MAX_FRAMES];
  const float4* pimg[HSLAM_MAX_FRAMES];   // level-0 pair images (hslam_frame::pair0): what the lineariser gathers from
  float frameEnergyTH[HSLAM_MAX_FRAMES];
  const float* precalc;      // n*n*28
  const float* adHTdeltaF;   // n*n*8
};

struct BARes {               // per-residual SoA, sorted by (host,target) pair, stride = rs
  int R, rs;
  const int* point;          // sorted point index
  const int* pair;           // host + target*n
  int8_t* state_state;
  int8_t* newState;
  uint8_t* isLin;
  uint8_t* isActive;
  float* state_energy;
  float* newEnergy;
  float* newEnergyWO;
  float* J;                  // [JS][rs]
  const int* slotpos;        // float offset of the residual's 32-byte records in JpJdF / prec: point * pstr + 8 * slot
  int pstr;                  // floats per point in the record arrays: 8 * (maxres + 1), see sc_pstr
  float* JpJdF;              // [P][pstr]: one 32-byte record per (point, residual slot), in the point's own order: the Schur
                             // kernel and the back-substitution read a point's residuals as ONE contiguous span
  float* prec;               // [P][pstr]: per (point, slot) {bd, Hdd, Hcd[4], flag (0 inactive / 1 active A / 2 active L), target}:
                             // the residual's contribution to the point sums (AccumulatedTopHessian.cpp:114-117), written by addPoint<mode>
  float* rtz;                // [8][rs] res_toZeroF
  float* cpt;                // [3][rs] centerProjectedTo
  float* pto;                // [16][rs] projectedTo
};

struct BAPts {               // per-point SoA, sorted by host
  int P;
  const float4* uvid;        // {u, v, idepth, idepth_zero}
  const float4* color;       // [P][2] float4
  const float4* weights;     // [P][2] float4
  const float* priorF;
  const float* deltaF;
  const int* host;
  const int* res_begin;      // CSR over sorted residual positions, in the point's own order
  const int* res_index;
  float* HdiF; float* bdSumF; float* idepth_hessian;
  float* Hdd[2]; float* bd[2]; float4* Hcd[2];
  float* step;
};

struct TopChunk { int pair, start, count, mode; };   // mode = schedule index: 0 = A, 1 = L, 2 = marginalisation
struct TopTile { int start, count, chunk, pair; };  // <= 128 residuals of one pair; chunk = partial it is reduced into (global index)
struct TopSched { const TopTile* tiles; const int* block_begin; int blocks, pad; };   // tile schedule of one mode (see ba_top_kernel)
struct LinTile { int start, count, pair, pad; };    // <= LIN_THREADS non-linearised residuals of one (host,target) pair
struct PtChunk { int host, start, count, pad; };

struct StitchArgs {
  int n, LD;
  const double* adHost;      // n*n*64 col-major
  const double* adTarget;
  const double* accTop;      // [3][n*n][TOP_PART] fp64 (A, L, marginalisation)
  int L_set;                 // accumulator set the L system is stitched from: 1 (accumulateLF) or 2 (marginalizePointsF)
  const double* G;           // [host][LD*LD]
  double* Hb;                // 3 x (N*N + N)
  int* nres;                 // [2]
  double cPrior[4];
  float cDeltaF[4];
  const double* framePrior;  // n*8
  const double* frameDeltaPrior;
  int add_priors;
  int has_L;                 // 0: no linearised residual exists, the L accumulators are all zero
};

// Everything the four kernels need about ONE window.
struct BADesc {
  BAWin W;
  BARes r;
  BAPts p;
  // linearise
  double* epart; unsigned int* ticket; double* energy; int lin_warps;   // lin_warps = warps per block * linearise blocks
  const LinTile* lintiles;   // tile schedule of the lineariser: the non-linearised residuals, pair by pair
  int jmode;                 // 0: every Jacobian float is stored; 1: GN-loop mode, the per-pixel rows (resF, JIdx, JabF) and
                             //    projectedTo are not stored (only fixLinearizationF / mode L read them: re-formed on demand)
  // top
  const TopChunk* topchunks; float* toppartial; unsigned int* pair_ticket; const int* pair_chunk_begin; double* accTop;
  TopSched sched[3];         // tile schedules: A over the non-linearised, L over the linearised, marginalisation over all residuals
  // schur
  const PtChunk* ptchunks; float* scpartial; unsigned int* host_ticket; const int* host_chunk_begin; double* G; int LD;
  int maxres;                // longest residual list of a point (<= HSLAM_MAX_FRAMES)
  // stitch
  StitchArgs S;
  // host <-> device layout conversion (caller order <-> sorted order) and back-substitution
  // marginalisation pass (EnergyFunctional::marginalizePointsF): per sorted point 1 = to marginalise; null otherwise
  const unsigned char* marg;
  int shift_prior;           // SC addPoint's shiftPriorToZero (1 in accumulateSCF_MT, 0 in marginalizePointsF)
  const int* r_s2o;          // sorted residual position -> caller residual id
  const int* p_s2o;          // sorted point position -> caller point id
  double* x;                 // (8n+4) solution of the reduced system (uploaded by resubstitute)
  float* xc;                 // [4]      calibration step, float
  float* xAd;                // [n*n][8] per-pair adjoint-mapped frame steps (EnergyFunctional.cpp:236-238)
  // 2-D tensor maps over the SoA Jacobian J[JS][rs] (inner dimension = residual): boxes of TOP_ROW residuals x 22 rows
  // (rows 8..29) and x 17 rows (rows 62..78): the two TMA loads that fill one stage of the top kernel's ring
  alignas(64) CUtensorMap tmapJ[4];   // [2]: J rows 8..78 (mode L box of 71 rows), [3]: res_toZeroF[8][rs] (box of 8 rows)
};

// Per-window byte offsets into the batch staging blocks (one H2D block `up`, one D2H block `dn`);
// a negative offset means "absent / not requested".
struct BAIO {
  long long up_precalc, up_adHTd, up_adH, up_adT, up_fprior, up_fdprior, up_idepth, up_deltaF, up_idz, up_x;
  long long dn_energy, dn_nres, dn_Hb, dn_newEnergy, dn_newEnergyWO, dn_newState, dn_step;
  long long dn_th;           // new frameEnergyTH of frame `newest` (System::setNewFrameEnergyTH), 1 float
  int newest, pad;           // pad: 1 = deltaF is formed on the device (hslam_ba_step_io::derive_deltaF)
};

__device__ __forceinline__ int find_window(const int* __restrict__ prefix, int nwin, int b, int& local) {
  if (nwin == 1) { local = b; return 0; }
  if (nwin < 0) { const int per = -nwin; const int w = b / per; local = b - w * per; return w; }   // every window owns `per` blocks
  int lo = 0, hi = nwin;   // prefix[lo] <= b < prefix[hi]
  while (hi - lo > 1) {
    int m = (lo + hi) >> 1;
    if (prefix[m] <= b) lo = m; else hi = m;
  }
  local = b - prefix[lo];
  return lo;
}

// ------------------------------------------------------------------------------------------------
// K4: PointFrameResidual::linearize (+ applyRes/takeData when APPLY) — one thread per residual.
// Exactly-rounded primitives in the operation order of the source: states, energies and all 74
// Jacobian floats are bit-identical to the float-faithful oracle.  The eight pattern projections are
// computed first, then all 32 texel loads are issued together (memory-level parallelism), then the
// per-pixel arithmetic runs in pattern order.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void take_data(const BARes& r, int k) {  // OptimizationClasses.h:155-161
  const float* J = r.J + k;
  const int s = r.rs;
  float d0 = J[J_DD * s], d1 = J[(J_DD + 1) * s];
  float JJ0 = xadd(xmul(J[(J_IDX2 + 0) * s], d0), xmul(J[(J_IDX2 + 2) * s], d1));
  float JJ1 = xadd(xmul(J[(J_IDX2 + 1) * s], d0), xmul(J[(J_IDX2 + 3) * s], d1));
  float o[8];
#pragma unroll
  for (int i = 0; i < 6; ++i) o[i] = xadd(xmul(J[(J_XI0 + i) * s], JJ0), xmul(J[(J_XI1 + i) * s], JJ1));
  o[6] = xadd(xmul(J[(J_ABIDX + 0) * s], d0), xmul(J[(J_ABIDX + 2) * s], d1));
  o[7] = xadd(xmul(J[(J_ABIDX + 1) * s], d0), xmul(J[(J_ABIDX + 3) * s], d1));
  float4* dst = reinterpret_cast<float4*>(r.JpJdF + r.slotpos[k]);
  dst[0] = make_float4(o[0], o[1], o[2], o[3]); dst[1] = make_float4(o[4], o[5], o[6], o[7]);
}

__device__ __forceinline__ void apply_one(const BARes& r, int k) {  // OptimizationClasses.cpp:235-256
  if (r.state_state[k] == HSLAM_RES_OOB) return;
  int8_t ns = r.newState[k];
  if (ns == HSLAM_RES_IN) { r.isActive[k] = 1; take_data(r, k); } else { r.isActive[k] = 0; }
  r.state_state[k] = ns;
  r.state_energy[k] = r.newEnergy[k];
}

// Indexing: block b of a window owns lintiles[b]: <= LIN_THREADS NON-linearised residuals of ONE (host,target) pair (the
// sorted order keeps the linearised residuals of a pair behind the others), so linearised residuals cost nothing and no
// store of a warp is a partial sector.  With `ids` (sorted positions, one window, J-refresh pass of
// hslam_ba_fix_linearization in GN-loop mode): thread i re-forms every Jacobian float of residual ids[i] and writes nothing else.
template <bool APPLY, bool FULLJ>
__global__ void __launch_bounds__(LIN_THREADS, LIN_MINB) ba_linearize_kernel(const BADesc* __restrict__ descs,
                                                                   const int* __restrict__ prefix, int nwin, int block0,
                                                                   const int* __restrict__ ids, int n_ids) {
#ifndef LIN_PARK
#define LIN_PARK 0   /* measured: parking colours/weights in smem 149 us vs 146 us in registers */
#endif
#if LIN_PARK
  __shared__ float lin_park[16][LIN_THREADS];
#endif
  int lb;
  const int win = find_window(prefix, nwin, blockIdx.x + block0, lb);   // block0: a launch may cover a sub-range of the batch
  const BADesc& D = descs[win];
  const BAWin& W = D.W;
  const BARes& r = D.r;
  const int s = r.rs;
  const bool refresh = ids != nullptr;
  int k, pair;
  bool valid;
  if (refresh) {
    const int i = lb * LIN_THREADS + threadIdx.x;
    valid = i < n_ids;
    k = valid ? ids[i] : 0;
    pair = valid ? r.pair[k] : 0;
  } else {
    const LinTile tl = D.lintiles[lb];
    valid = (int)threadIdx.x < tl.count;
    k = tl.start + threadIdx.x;
    pair = tl.pair;
#ifdef LIN_EXP_NOTILE   /* timing experiment only (wrong results): the residual index does not wait for the tile descriptor */
    k = min(lb * LIN_THREADS + (int)threadIdx.x, r.R - 1); valid = true;
#endif
  }
  double energy_ret = 0.0;
  // every per-residual input is loaded up front (independent loads, one round trip)
  const int8_t st0 = valid ? r.state_state[k] : (int8_t)HSLAM_RES_OOB;
  const float old_energy = valid ? r.state_energy[k] : 0.f;
#ifdef LIN_EXP_NOPOINT  /* timing experiment only (wrong results): the point record does not wait for the point index */
  const int p = k / 6;
#elif defined(LIN_EXP_NOPOINT2)   /* no dependency, but 32 distinct consecutive points per warp (the real traffic) */
  const int p = k % D.p.P;
#elif defined(LIN_EXP_DEPPOINT)   /* dependency kept, 6 residuals share a point (less traffic) */
  const int p = valid ? (r.point[k] >> 30) + k / 6 : 0;
#else
  const int p = valid ? r.point[k] : 0;
#endif
  const int sp = (APPLY && valid && !refresh) ? r.slotpos[k] : 0;
  if (valid) {
    float* J = r.J + k;
    if (!refresh) STJ(r.newEnergyWO + k, -1.f);
    bool oob = (st0 == HSLAM_RES_OOB);
    const int tgtI = (pair * W.n_rcp) >> 16, hostI = pair - tgtI * W.n;   // pair = host + target * n
    float energyLeft = 0.f, wJI2 = 0.f;
    if (!oob) {
      const float4* pc = reinterpret_cast<const float4*>(W.precalc + (size_t)HSLAM_PRECALC_STRIDE * pair);
      const float4 q0 = __ldg(pc), q1 = __ldg(pc + 1), q2 = __ldg(pc + 2), q3 = __ldg(pc + 3), q4 = __ldg(pc + 4),
                   q5 = __ldg(pc + 5), q6 = __ldg(pc + 6);
      // [0..8] KRKi, [9..11] Kt, [12..20] R0, [21..23] t0, [24] a, [25] b, [26] b0
      const float KRKi[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
      const float Kt[3] = {q2.y, q2.z, q2.w};
      const float R0[9] = {q3.x, q3.y, q3.z, q3.w, q4.x, q4.y, q4.z, q4.w, q5.x};
      const float t0[3] = {q5.y, q5.z, q5.w};
      const float affLL0 = q6.x, affLL1 = q6.y, b0 = q6.z;
      const float4 pu = __ldg(D.p.uvid + p);
      // host colours / weights are needed only after the gather, but their addresses are known now:
      // issue them with the point record so their latency overlaps the projections
      const float4 c0 = __ldg(D.p.color + 2 * p), c1 = __ldg(D.p.color + 2 * p + 1);
      const float4 w0 = __ldg(D.p.weights + 2 * p), w1 = __ldg(D.p.weights + 2 * p + 1);
#if LIN_PARK
      // parked in shared memory (component-major: conflict-free) until the pattern loop: 16 registers less
      // through the gather phase, i.e. more texel loads in flight per thread
      {
        float* pk = &lin_park[0][threadIdx.x];
        pk[0 * LIN_THREADS] = c0.x; pk[1 * LIN_THREADS] = c0.y; pk[2 * LIN_THREADS] = c0.z; pk[3 * LIN_THREADS] = c0.w;
        pk[4 * LIN_THREADS] = c1.x; pk[5 * LIN_THREADS] = c1.y; pk[6 * LIN_THREADS] = c1.z; pk[7 * LIN_THREADS] = c1.w;
        pk[8 * LIN_THREADS] = w0.x; pk[9 * LIN_THREADS] = w0.y; pk[10 * LIN_THREADS] = w0.z; pk[11 * LIN_THREADS] = w0.w;
        pk[12 * LIN_THREADS] = w1.x; pk[13 * LIN_THREADS] = w1.y; pk[14 * LIN_THREADS] = w1.z; pk[15 * LIN_THREADS] = w1.w;
      }
#endif
      const float u_pt = pu.x, v_pt = pu.y, idp = pu.z, idz = pu.w;
      const float wm3 = (float)(W.w - 3), hm3 = (float)(W.h - 3);
      // ---- centre projection at the FEJ point, DirectProjection.h:20-38
      const float Kl0 = xmul(xsub(xadd(u_pt, 0.f), W.cx), W.fxi), Kl1 = xmul(xsub(xadd(v_pt, 0.f), W.cy), W.fyi);
      const float p0 = xadd(xdot3(R0[0], R0[1], R0[2], Kl0, Kl1, 1.f), xmul(t0[0], idz));
      const float p1 = xadd(xdot3(R0[3], R0[4], R0[5], Kl0, Kl1, 1.f), xmul(t0[1], idz));
      const float p2 = xadd(xdot3(R0[6], R0[7], R0[8], Kl0, Kl1, 1.f), xmul(t0[2], idz));
      const float drescale = xdiv(1.0f, p2);
      const float new_idepth = xmul(idz, drescale);
      if (!(drescale > 0.f)) oob = true;
      const float u = xmul(p0, drescale), v = xmul(p1, drescale);
      const float Ku = xadd(xmul(u, W.fx), W.cx), Kv = xadd(xmul(v, W.fy), W.cy);
      if (!oob && !(Ku > 1.1f && Kv > 1.1f && Ku < wm3 && Kv < hm3)) oob = true;
      if (!oob) {
        if (!refresh) { STJ(r.cpt + k, Ku); STJ(r.cpt + s + k, Kv); STJ(r.cpt + 2 * s + k, new_idepth); }
        // OptimizationClasses.cpp:85-121
        float ddx = xmul(xmul(xmul(drescale, xsub(t0[0], xmul(t0[2], u))), W.scale_idepth), W.fx);
        float ddy = xmul(xmul(xmul(drescale, xsub(t0[1], xmul(t0[2], v))), W.scale_idepth), W.fy);
        float cx2 = xmul(drescale, xsub(xmul(R0[6], u), R0[0]));
        float cx3 = xmul(xmul(xmul(W.fx, drescale), xsub(xmul(R0[7], u), R0[1])), W.fyi);
        float cx0 = xmul(Kl0, cx2), cx1 = xmul(Kl1, cx3);
        float cy2 = xmul(xmul(xmul(W.fy, drescale), xsub(xmul(R0[6], v), R0[3])), W.fxi);
        float cy3 = xmul(drescale, xsub(xmul(R0[7], v), R0[4]));
        float cy0 = xmul(Kl0, cy2), cy1 = xmul(Kl1, cy3);
        cx0 = xmul(xadd(cx0, u), W.scale_f); cx1 = xmul(cx1, W.scale_f);
        cx2 = xmul(xadd(cx2, 1.f), W.scale_c); cx3 = xmul(cx3, W.scale_c);
        cy0 = xmul(cy0, W.scale_f); cy1 = xmul(xadd(cy1, v), W.scale_f);
        cy2 = xmul(cy2, W.scale_c); cy3 = xmul(xadd(cy3, 1.f), W.scale_c);
        // Jpdxi rows (OptimizationClasses.cpp:109-121); re-formed from (u, v, new_idepth) at takeData time
        // instead of being kept live across the pattern loop (register pressure)
#define HS_XI0(i) ((i) == 0 ? xmul(new_idepth, W.fx) : (i) == 1 ? 0.f : (i) == 2 ? xmul(xmul(-new_idepth, u), W.fx) : \
                   (i) == 3 ? xmul(xmul(-u, v), W.fx) : (i) == 4 ? xmul(xadd(1.f, xmul(u, u)), W.fx) : xmul(-v, W.fx))
#define HS_XI1(i) ((i) == 0 ? 0.f : (i) == 1 ? xmul(new_idepth, W.fy) : (i) == 2 ? xmul(xmul(-new_idepth, v), W.fy) : \
                   (i) == 3 ? xmul(-xadd(1.f, xmul(v, v)), W.fy) : (i) == 4 ? xmul(xmul(u, v), W.fy) : xmul(u, W.fy))
#pragma unroll
        for (int i = 0; i < 6; ++i) { STJ(J + (J_XI0 + i) * s, HS_XI0(i)); STJ(J + (J_XI1 + i) * s, HS_XI1(i)); }
        STJ(J + (J_C0 + 0) * s, cx0); STJ(J + (J_C0 + 1) * s, cx1); STJ(J + (J_C0 + 2) * s, cx2); STJ(J + (J_C0 + 3) * s, cx3);
        STJ(J + (J_C1 + 0) * s, cy0); STJ(J + (J_C1 + 1) * s, cy1); STJ(J + (J_C1 + 2) * s, cy2); STJ(J + (J_C1 + 3) * s, cy3);
        STJ(J + J_DD * s, ddx); STJ(J + (J_DD + 1) * s, ddy);
        // ---- 8-pixel pattern at the current state, :141-204.  Pass 1: projections + bounds.
        float Kup[8], Kvp[8];
        int n_ok = 8;   // number of leading pattern pixels that are in bounds (the source returns at the first OOB)
#pragma unroll
        for (int idx = 0; idx < 8; ++idx) {
          float px = xadd(u_pt, (float)c_pat[idx][0]), py = xadd(v_pt, (float)c_pat[idx][1]);
          float a0 = xadd(xdot3(KRKi[0], KRKi[1], KRKi[2], px, py, 1.f), xmul(Kt[0], idp));
          float a1 = xadd(xdot3(KRKi[3], KRKi[4], KRKi[5], px, py, 1.f), xmul(Kt[1], idp));
          float a2 = xadd(xdot3(KRKi[6], KRKi[7], KRKi[8], px, py, 1.f), xmul(Kt[2], idp));
          Kup[idx] = xdiv(a0, a2); Kvp[idx] = xdiv(a1, a2);
          bool in = (Kup[idx] > 1.1f && Kvp[idx] > 1.1f && Kup[idx] < wm3 && Kvp[idx] < hm3);
          if (!in && n_ok == 8) n_ok = idx;
        }
        // the source stores projectedTo before it tests the next pixel: pixels [0, n_ok) are written
        if (FULLJ) {
#pragma unroll
          for (int idx = 0; idx < 8; ++idx)
            if (idx < n_ok) { STJ(r.pto + (2 * idx) * s + k, Kup[idx]); STJ(r.pto + (2 * idx + 1) * s + k, Kvp[idx]); }
        }
        if (n_ok < 8) {
          oob = true;
          // Jacobian entries of the pixels before the first OOB one are written by the source as well;
          // they are never read for an OOB residual, so they are left untouched here.
        } else {
          // Pass 2: all 32 texel loads in flight at once.
#ifndef LIN_PAIR_IMAGE
#define LIN_PAIR_IMAGE 1
#endif
#if LIN_PAIR_IMAGE
          const float4* img = W.pimg[tgtI];
#else
          const float4* img = W.img[tgtI];
#endif
          Hit3 hit[8];
          const uint64_t pol = l2_keep_policy();
#pragma unroll
#ifdef LIN_EXP_NOGATHER   /* timing experiment only: every gather hits the same few texels */
          for (int idx = 0; idx < 8; ++idx) hit[idx] = interp33<HS_TEXEL_L2_KEEP != 0>(img, 8.25f + idx, 8.5f, W.w, pol);
#elif LIN_PAIR_IMAGE
          for (int idx = 0; idx < 8; ++idx) hit[idx] = interp33_pair(img, Kup[idx], Kvp[idx], W.w, pol);
#else
          for (int idx = 0; idx < 8; ++idx) hit[idx] = interp33<HS_TEXEL_L2_KEEP != 0>(img, Kup[idx], Kvp[idx], W.w, pol);
#endif
#if !LIN_PARK
          const float color[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
          const float wgt[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#define LIN_COLOR(i) color[i]
#define LIN_WGT(i) wgt[i]
#else
#define LIN_COLOR(i) lin_park[i][threadIdx.x]
#define LIN_WGT(i) lin_park[8 + (i)][threadIdx.x]
#endif
          float II00 = 0.f, II11 = 0.f, II10 = 0.f, AI00 = 0.f, AI01 = 0.f, AI10 = 0.f, AI11 = 0.f;
          float AA00 = 0.f, AA01 = 0.f, AA11 = 0.f;
          float JIr0 = 0.f, JIr1 = 0.f, Jabr0 = 0.f, Jabr1 = 0.f, rr = 0.f;
          bool nonfinite = false;
#pragma unroll
          for (int idx = 0; idx < 8; ++idx) {
            if (nonfinite) continue;
            const float hI = hit[idx].I, hdx = hit[idx].dx, hdy = hit[idx].dy;
            if (!isfinite(hI)) { nonfinite = true; continue; }   // :157-161
            const float col = LIN_COLOR(idx);
            float residual = xsub(hI, xadd(xmul(affLL0, col), affLL1));
            float drdA = xsub(col, b0);
            float w = xsqrt(xdiv(W.outlierTH, xadd(W.outlierTH, xadd(xmul(hdx, hdx), xmul(hdy, hdy)))));
            w = xmul(0.5f, xadd(w, LIN_WGT(idx)));
            float ar = fabsf(residual);
            float hw = ar < W.huber ? 1.f : xdiv(W.huber, ar);
            energyLeft = xadd(energyLeft, xmul(xmul(xmul(xmul(xmul(w, w), hw), residual), residual), xsub(2.f, hw)));
            if (hw < 1.f) hw = xsqrt(hw);
            hw = xmul(hw, w);
            float h1 = xmul(hdx, hw), h2 = xmul(hdy, hw);
            float rF = xmul(residual, hw), ja = xmul(drdA, hw);
            float jab0 = W.affA < 0.f ? 0.f : ja, jab1 = W.affB < 0.f ? 0.f : hw;
            if (FULLJ) {   // the per-pixel rows: read only by fixLinearizationF, mode L and calcLEnergyPt
              STJ(J + (J_RESF + idx) * s, rF);
              STJ(J + (J_IDX0 + idx) * s, h1);
              STJ(J + (J_IDX1 + idx) * s, h2);
              STJ(J + (J_AB0 + idx) * s, jab0);
              STJ(J + (J_AB1 + idx) * s, jab1);
            }
            II00 = xadd(II00, xmul(h1, h1)); II11 = xadd(II11, xmul(h2, h2)); II10 = xadd(II10, xmul(h1, h2));
            AI00 = xadd(AI00, xmul(ja, h1)); AI01 = xadd(AI01, xmul(ja, h2));
            AI10 = xadd(AI10, xmul(hw, h1)); AI11 = xadd(AI11, xmul(hw, h2));
            AA00 = xadd(AA00, xmul(xmul(xmul(drdA, drdA), hw), hw));
            AA01 = xadd(AA01, xmul(xmul(drdA, hw), hw));
            AA11 = xadd(AA11, xmul(hw, hw));
            // wJI2_sum uses hitColor AFTER the *= hw (OptimizationClasses.cpp:174-175 precede :197)
            wJI2 = xadd(wJI2, xmul(xmul(hw, hw), xadd(xmul(h1, h1), xmul(h2, h2))));
            // sums addPoint<0> forms from resF (AccumulatedTopHessian.cpp:84-94), same order
            JIr0 = xadd(JIr0, xmul(rF, h1)); JIr1 = xadd(JIr1, xmul(rF, h2));
            Jabr0 = xadd(Jabr0, xmul(rF, jab0)); Jabr1 = xadd(Jabr1, xmul(rF, jab1));
            rr = xadd(rr, xmul(rF, rF));
          }
          if (nonfinite) {
            oob = true;
          } else {
            STJ(J + (J_IDX2 + 0) * s, II00); STJ(J + (J_IDX2 + 1) * s, II10); STJ(J + (J_IDX2 + 2) * s, II10); STJ(J + (J_IDX2 + 3) * s, II11);
            STJ(J + (J_ABIDX + 0) * s, AI00); STJ(J + (J_ABIDX + 1) * s, AI10); STJ(J + (J_ABIDX + 2) * s, AI01); STJ(J + (J_ABIDX + 3) * s, AI11);
            STJ(J + (J_AB2 + 0) * s, AA00); STJ(J + (J_AB2 + 1) * s, AA01); STJ(J + (J_AB2 + 2) * s, AA01); STJ(J + (J_AB2 + 3) * s, AA11);
            STJ(J + (J_JIR + 0) * s, JIr0); STJ(J + (J_JIR + 1) * s, JIr1); STJ(J + (J_JABR + 0) * s, Jabr0); STJ(J + (J_JABR + 1) * s, Jabr1);
            STJ(J + J_RR * s, rr);
            // ---- classification (:221-231) and, when APPLY, applyRes(true)/takeData (:235-256, .h:155-161)
            // straight from registers
            if (refresh) return;   // J-refresh pass: states, energies and JpJdF stay what the last pass left
            STJ(r.newEnergyWO + k, energyLeft);
            const float th = fmaxf(W.frameEnergyTH[hostI], W.frameEnergyTH[tgtI]);
            int8_t ns = HSLAM_RES_IN;
            if (energyLeft > th || wJI2 < 2.f) { energyLeft = th; ns = HSLAM_RES_OUT; }
            STJ(r.newEnergy + k, energyLeft);
            energy_ret = (double)energyLeft;
            r.newState[k] = ns;
            if (APPLY) {   // the old state cannot be OOB here
              if (ns == HSLAM_RES_IN) {
                r.isActive[k] = 1;
                const float JJ0 = xadd(xmul(II00, ddx), xmul(II10, ddy)), JJ1 = xadd(xmul(II10, ddx), xmul(II11, ddy));
                float o[8];
#pragma unroll
                for (int i = 0; i < 6; ++i) o[i] = xadd(xmul(HS_XI0(i), JJ0), xmul(HS_XI1(i), JJ1));
                o[6] = xadd(xmul(AI00, ddx), xmul(AI01, ddy));
                o[7] = xadd(xmul(AI10, ddx), xmul(AI11, ddy));
                stg256(r.JpJdF + sp, make_float4(o[0], o[1], o[2], o[3]), make_float4(o[4], o[5], o[6], o[7]));   // one 32-byte record, in the point's slot
              } else {
                r.isActive[k] = 0;
              }
              r.state_state[k] = ns;
              r.state_energy[k] = energyLeft;
            }
          }
        }
      }
    }
    if (oob && !refresh) {
      r.newState[k] = HSLAM_RES_OOB;
      energy_ret = (double)old_energy;
      if (APPLY && st0 != HSLAM_RES_OOB) {
        // went OOB in THIS pass: applyRes copies the newEnergy of its last non-OOB linearisation
        r.isActive[k] = 0;
        r.state_state[k] = HSLAM_RES_OOB;
        r.state_energy[k] = r.newEnergy[k];
      }
    }
  }
  // ---- energy: warp tree in fp64, one partial per WARP, no fence and no block barrier: a warp retires
  // as soon as its residuals are done (a __threadfence here made every warp wait for its ~100 Jacobian
  // stores to drain: ~20 % of the kernel).  The partials are summed in index order by the next kernel
  // of the stream (ba_energy_kernel, or the stitch kernel in the fused/batched path).
  if (refresh) return;
  const double e = warp_sum(energy_ret);
  if ((threadIdx.x & 31) == 0) D.epart[lb * (LIN_THREADS / 32) + (threadIdx.x >> 5)] = e;
}

// Fixed-order sum of the per-warp energy partials of ONE window by one block (any block size >= 128 that is a
// multiple of 32).  `red` is shared scratch of 16 doubles.
__device__ __forceinline__ void energy_sum_block(const BADesc& D, double* red) {
  const int nparts = D.lin_warps;
  // eight independent loads per round, four accumulators: the sum is one block's work and latency-bound — for the
  // 2 M-residual window (62 500 partials) the dependent-add loop it replaces was ~100 us of the 0.55 ms iteration
  // (the first 128 threads do it, whatever the block size: the order must not depend on who calls)
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  const int st = 128;
  int b = threadIdx.x < 128 ? (int)threadIdx.x : nparts;
  for (; b + 7 * st < nparts; b += 8 * st) {
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldcg(D.epart + b + j * st);
    a0 += v[0]; a1 += v[1]; a2 += v[2]; a3 += v[3];
    a0 += v[4]; a1 += v[5]; a2 += v[6]; a3 += v[7];
  }
  for (; b < nparts; b += st) a0 += __ldcg(D.epart + b);
  double acc = (a0 + a1) + (a2 + a3);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0 && threadIdx.x < 128) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 4; ++w) t += red[w];
    *D.energy = t;
    if (D.S.Hb) { const int N = 8 * D.S.n + 4; D.S.Hb[3 * ((size_t)N * N + N)] = t; }   // rides with the systems through hslam_ba_allreduce
  }
}

__global__ void __launch_bounds__(256) ba_energy_kernel(const BADesc* __restrict__ descs) {
  __shared__ double red[16];
  energy_sum_block(descs[blockIdx.x], red);
}


// ------------------------------------------------------------------------------------------------
// System::setNewFrameEnergyTH (Src/FullSystemOptimize.cpp:60-101) for ONE window by one block: exact k-th order
// statistic of state_NewEnergyWithOutlier over the non-linearised residuals into the newest frame (they are one
// contiguous range of the pair-sorted arrays: pair = host + target*n), k = (int)(THN * count), by a 4-round MSB-first
// radix select on the float bit patterns (all candidates are >= 0, so the unsigned order is the float order) — the
// value std::nth_element leaves at index k, without sorting and without the energies leaving HBM.  Then the
// threshold arithmetic of :86-91 in exactly-rounded fp32.  Integer histogram atomics only: deterministic.
// ------------------------------------------------------------------------------------------------
constexpr int TH_THREADS = 512;
__device__ __forceinline__ float energy_th_block(const BADesc& D, int newest) {
  __shared__ unsigned int hist[256];
  __shared__ int sh[6];   // lo, hi, count, prefix, rank, (unused)
  const BARes& r = D.r;
  const BAWin& W = D.W;
  const int n = W.n, R = r.R;
  if (threadIdx.x < 2) {
    // lower_bound of pair >= (newest + threadIdx.x) * n over the pair-sorted residual array
    const int key = (newest + (int)threadIdx.x) * n;
    int lo = 0, hi = R;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (r.pair[m] < key) lo = m + 1; else hi = m; }
    sh[threadIdx.x] = lo;
  }
  if (threadIdx.x == 2) sh[2] = 0;
  __syncthreads();
  const int lo = sh[0], hi = sh[1];
  int c = 0;
  for (int k = lo + threadIdx.x; k < hi; k += TH_THREADS)
    if (r.isLin[k] == 0 && r.newEnergyWO[k] >= 0.f) c++;
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&sh[2], c);
  __syncthreads();
  const int cnt = sh[2];
  if (cnt == 0) return (float)(12 * 12 * HSLAM_PATTERN_NUM);   // :73-77
  if (threadIdx.x == 0) { sh[3] = 0; sh[4] = (int)xmul(W.thN, (float)cnt); }   // :80 nthIdx = setting_frameEnergyTHN * size
  unsigned int mask = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
    const unsigned int prefix = (unsigned int)sh[3];
    for (int k = lo + threadIdx.x; k < hi; k += TH_THREADS) {
      const float e = r.newEnergyWO[k];
      if (r.isLin[k] == 0 && e >= 0.f) {
        const unsigned int key = __float_as_uint(e);
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      // warp 0: lane l owns bins 8l..8l+7; exclusive scan of the lane sums, the lane whose span holds the rank resolves it
      const int rank = sh[4];
      unsigned int h8[8], tot = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { h8[j] = hist[8 * threadIdx.x + j]; tot += h8[j]; }
      unsigned int inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o); if ((int)threadIdx.x >= o) inc += v; }
      unsigned int cum = inc - tot;
      if ((unsigned int)rank >= cum && (unsigned int)rank < inc) {
        int b = 0;
        bool open = true;   // rank < cum + tot: the walk stops inside this lane's eight bins
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (open && (unsigned int)rank >= cum + h8[j]) { cum += h8[j]; b = j + 1; } else open = false;
        }
        sh[3] = (int)(prefix | ((unsigned int)(8 * threadIdx.x + b) << shift));
        sh[4] = rank - (int)cum;
      }
    }
    mask |= 255u << shift;
    __syncthreads();
  }
  const float nth = xsqrt(__uint_as_float((unsigned int)sh[3]));   // :86
  float th = xmul(nth, W.thFac);                                    // :88
  th = xadd(xmul(26.0f, W.thConst), xmul(th, xsub(1.0f, W.thConst)));   // :89
  th = xmul(th, th);                                                // :90
  th = xmul(th, xmul(W.thOverall, W.thOverall));                    // :91
  return th;
}
// One block per window.  `newest` < 0: taken from io[win].newest.  The result goes to th_dev[win] (nullable), to the
// window's slot of the batch D2H block (io / dn, nullable) and INTO the device descriptor (the next linearisation of
// the window reads it, as the reference's does from Frame::frameEnergyTH).
__global__ void __launch_bounds__(TH_THREADS) ba_energy_th_kernel(BADesc* __restrict__ descs, int newest, float* __restrict__ th_dev,
                                                                  const BAIO* __restrict__ io, char* __restrict__ dn) {
  const int win = blockIdx.x;
  if (io && io[win].dn_th < 0) return;
  const int nf = newest >= 0 ? newest : io[win].newest;
  const float th = energy_th_block(descs[win], nf);
  if (threadIdx.x == 0) {
    descs[win].W.frameEnergyTH[nf] = th;
    if (th_dev) th_dev[win] = th;
    if (io) *reinterpret_cast<float*>(dn + io[win].dn_th) = th;
  }
}

__global__ void ba_apply_kernel(BARes r) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < r.R && !r.isLin[k]) apply_one(r, k);
}

// PointFrameResidual::fixLinearizationF, OptimizationClasses.cpp:258-284 (ids = sorted positions)
__global__ void ba_fixlin_kernel(BAWin W, BARes r, BAPts pts, const int* __restrict__ ids, int count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  int k = ids[i];
  const int s = r.rs;
  const float* J = r.J + k;
  const float* dp = W.adHTdeltaF + 8 * r.pair[k];
  float dx = 0.f, dy = 0.f, cxs = 0.f, cys = 0.f;
#pragma unroll
  for (int j = 0; j < 6; ++j) { dx = xadd(dx, xmul(J[(J_XI0 + j) * s], dp[j])); dy = xadd(dy, xmul(J[(J_XI1 + j) * s], dp[j])); }
#pragma unroll
  for (int j = 0; j < 4; ++j) { cxs = xadd(cxs, xmul(J[(J_C0 + j) * s], W.cDeltaF[j])); cys = xadd(cys, xmul(J[(J_C1 + j) * s], W.cDeltaF[j])); }
  float dd = pts.deltaF[r.point[k]];
  float Jpx = xadd(xadd(dx, cxs), xmul(J[J_DD * s], dd));
  float Jpy = xadd(xadd(dy, cys), xmul(J[(J_DD + 1) * s], dd));
  float da = dp[6], db = dp[7];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float rtz = J[(J_RESF + j) * s];
    rtz = xsub(rtz, xmul(J[(J_IDX0 + j) * s], Jpx));
    rtz = xsub(rtz, xmul(J[(J_IDX1 + j) * s], Jpy));
    rtz = xsub(rtz, xmul(J[(J_AB0 + j) * s], da));
    rtz = xsub(rtz, xmul(J[(J_AB1 + j) * s], db));
    r.rtz[j * s + k] = rtz;
  }
  r.isLin[k] = 1;
}

// ------------------------------------------------------------------------------------------------
// K5: AccumulatedTopHessianSSE::addPoint<mode> — per-pair 13x13 relative accumulators, modes A and L
// in ONE launch.  A block owns a chunk of ONE (mode,pair) residual segment; each thread accumulates
// its residuals into 91 registers ([C(4);xi(6)]^2 upper = 55, [C;xi] x [a,b,r] = 30, [a,b,r]^2 upper
// = 6), then warp shuffles -> shared memory -> fixed-order cross-warp sum -> per-chunk partial.  The
// last block of a (mode,pair) to finish folds that pair's chunks in chunk order into fp64.
// ------------------------------------------------------------------------------------------------
// PART 0 threads own the 55 entries of the 10x10 block, PART 1 threads the other 36 (+ nres and the
// per-residual point terms): half the accumulator registers per thread, twice the warps in flight.
// Loads that only feed the other part's entries are dead code in each instantiation.
template <int MODE, int PART>
__device__ __forceinline__ void top_accumulate(const BADesc& D, const TopChunk& ch, float (&acc)[56], int& nres) {
  const BAWin& W = D.W;
  const BARes& r = D.r;
  const int s = r.rs;
  float dp[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (MODE == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) dp[j] = __ldg(W.adHTdeltaF + 8 * ch.pair + j);
  }
  for (int k = ch.start + (threadIdx.x & (TOP_THREADS - 1)); k < ch.start + ch.count; k += TOP_THREADS) {
    // The residual filter (AccumulatedTopHessian.cpp:37-49) is applied AFTER the loads: every load of an
    // iteration is independent of the flags, so an iteration costs one memory round trip instead of two.  A
    // filtered-out residual contributes exact zeros (its x, y and bottom-right terms are zeroed).
    const bool lin = r.isLin[k] != 0, act = r.isActive[k] != 0;
    // mode 2 = addPoint<2> (marginalisation): active residuals of the points flagged in D.marg, resApprox = res_toZeroF
    const bool use = (MODE == 0) ? (!lin && act) : (MODE == 1) ? (lin && act) : (act && D.marg[r.point[k]] != 0);
    const float* J = r.J + k;
    float x[10], y[10];
#pragma unroll
    for (int j = 0; j < 4; ++j) { x[j] = J[(J_C0 + j) * s]; y[j] = J[(J_C1 + j) * s]; }
#pragma unroll
    for (int j = 0; j < 6; ++j) { x[4 + j] = J[(J_XI0 + j) * s]; y[4 + j] = J[(J_XI1 + j) * s]; }
    const float d0 = J[J_DD * s], d1 = J[(J_DD + 1) * s];
    float a = J[(J_IDX2 + 0) * s], b = J[(J_IDX2 + 2) * s], c = J[(J_IDX2 + 3) * s];
    float AI00 = 0.f, AI10 = 0.f, AI01 = 0.f, AI11 = 0.f, AA00 = 0.f, AA01 = 0.f, AA11 = 0.f;
    float JIr0 = 0.f, JIr1 = 0.f, Jabr0 = 0.f, Jabr1 = 0.f, rr = 0.f;
    if (PART == 1) {
      AI00 = J[(J_ABIDX + 0) * s]; AI10 = J[(J_ABIDX + 1) * s]; AI01 = J[(J_ABIDX + 2) * s]; AI11 = J[(J_ABIDX + 3) * s];
      AA00 = J[(J_AB2 + 0) * s]; AA01 = J[(J_AB2 + 2) * s]; AA11 = J[(J_AB2 + 3) * s];
    }
    if (MODE == 0) {
      if (PART == 1) {
        JIr0 = J[(J_JIR + 0) * s]; JIr1 = J[(J_JIR + 1) * s];
        Jabr0 = J[(J_JABR + 0) * s]; Jabr1 = J[(J_JABR + 1) * s]; rr = J[J_RR * s];
      }
    } else if (PART == 1) {
      // resApprox = res_toZeroF + J*delta, AccumulatedTopHessian.cpp:63-81 (sequential sums)
      float dx = 0.f, dy = 0.f, cxs = 0.f, cys = 0.f;
#pragma unroll
      for (int j = 0; j < 6; ++j) { dx = xadd(dx, xmul(x[4 + j], dp[j])); dy = xadd(dy, xmul(y[4 + j], dp[j])); }
#pragma unroll
      for (int j = 0; j < 4; ++j) { cxs = xadd(cxs, xmul(x[j], W.cDeltaF[j])); cys = xadd(cys, xmul(y[j], W.cDeltaF[j])); }
      float dd = D.p.deltaF[r.point[k]];
      float Jpx = xadd(xadd(dx, cxs), xmul(d0, dd)), Jpy = xadd(xadd(dy, cys), xmul(d1, dd));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float i0 = J[(J_IDX0 + j) * s], i1 = J[(J_IDX1 + j) * s], a0 = J[(J_AB0 + j) * s], a1 = J[(J_AB1 + j) * s];
        float rz = r.rtz[j * s + k];
        if (MODE == 1) {   // mode 2 takes res_toZeroF as it is (AccumulatedTopHessian.cpp:61)
          rz = xadd(rz, xmul(i0, Jpx)); rz = xadd(rz, xmul(i1, Jpy));
          rz = xadd(rz, xmul(a0, dp[6])); rz = xadd(rz, xmul(a1, dp[7]));
        }
        JIr0 = xadd(JIr0, xmul(rz, i0)); JIr1 = xadd(JIr1, xmul(rz, i1));
        Jabr0 = xadd(Jabr0, xmul(rz, a0)); Jabr1 = xadd(Jabr1, xmul(rz, a1));
        rr = xadd(rr, xmul(rz, rz));
      }
    }
    if (!use) {
#pragma unroll
      for (int j = 0; j < 10; ++j) { x[j] = 0.f; y[j] = 0.f; }
      a = b = c = AI00 = AI10 = AI01 = AI11 = JIr0 = JIr1 = 0.f;   // stale Jacobians may hold anything
      AA00 = AA01 = AA11 = Jabr0 = Jabr1 = rr = 0.f;
    }
    if (PART == 0) {
      // 10x10 upper triangle: [x y] [[a b],[b c]] [x y]^T
      int idx = 0;
#pragma unroll
      for (int rI = 0; rI < 10; ++rI) {
        float px = fmaf(a, x[rI], b * y[rI]), py = fmaf(b, x[rI], c * y[rI]);
#pragma unroll
        for (int cI = rI; cI < 10; ++cI, ++idx) acc[idx] += fmaf(x[cI], px, y[cI] * py);
      }
    } else {
      // 10x3 top-right: columns a, b, r  (MatrixAccumulators.h:859-901); entries 55..84 -> acc[0..29]
#pragma unroll
      for (int rI = 0; rI < 10; ++rI) {
        acc[3 * rI + 0] += fmaf(x[rI], AI00, y[rI] * AI01);
        acc[3 * rI + 1] += fmaf(x[rI], AI10, y[rI] * AI11);
        acc[3 * rI + 2] += fmaf(x[rI], JIr0, y[rI] * JIr1);
      }
      acc[30] += AA00; acc[31] += AA01; acc[32] += Jabr0; acc[33] += AA11; acc[34] += Jabr1; acc[35] += rr;
      const float tgt = (float)((ch.pair * W.n_rcp) >> 16);
      float* o = r.prec + r.slotpos[k];   // the residual's 32-byte record in its point's slot
      if (!use) {
        // modes A / L own every residual of their tiles: an inactive one gets a record that says so.  The
        // marginalisation pass (2) leaves the records of the points it does not touch alone.
        if (MODE != 2) stg256(o, make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, tgt));
        continue;
      }
      nres++;
      // per-residual contribution to the point sums (:114-117), exact ops: they are summed per point in
      // the point's own order by the Schur kernel, which keeps Hdd/bd/Hcd bit-identical to the oracle.
      float Ji0 = xadd(xmul(a, d0), xmul(b, d1)), Ji1 = xadd(xmul(b, d0), xmul(c, d1));
      stg256(o, make_float4(xadd(xmul(JIr0, d0), xmul(JIr1, d1)), xadd(xmul(Ji0, d0), xmul(Ji1, d1)),
                            xadd(xmul(x[0], Ji0), xmul(y[0], Ji1)), xadd(xmul(x[1], Ji0), xmul(y[1], Ji1))),
             make_float4(xadd(xmul(x[2], Ji0), xmul(y[2], Ji1)), xadd(xmul(x[3], Ji0), xmul(y[3], Ji1)), lin ? 2.f : 1.f, tgt));
    }
  }
}

// ---- tile-scheduled top-Hessian kernel -------------------------------------------------------------------
// The sorted residual array is cut into TILES of <= 128 residuals that never cross a (host,target) pair.
// Block b of a window owns the contiguous run of tiles [top_block_begin[b], top_block_begin[b+1]): every
// block of a window has the same amount of work (no wave-quantisation tail), a batch launches exactly as
// many blocks as the device holds.  A maximal run of tiles of one pair inside a block is a CHUNK: it is
// reduced into one fp32 partial, and the last block of a pair folds that pair's partials in chunk order.
// Mode-A tiles are staged through shared memory with cp.async, double-buffered: the 39 Jacobian
// components of tile i+1 stream in while tile i is accumulated (one block-wide set of loads instead of two
// per-thread sets, and no load latency on the critical path).  Mode-L tiles (linearised residuals: few,
// 80 components each) keep the direct per-thread path.
constexpr int TOP_NC = 39;   // staged components: J[8..29] (Jpdxi, Jpdc, Jpdd) and J[62..78] (the 2x2 sums, JI_r, Jab_r, rr)
__device__ __forceinline__ int top_comp(int c) { return c < 22 ? J_XI0 + c : J_IDX2 + (c - 22); }
// stage rows: components 0..21 in rows 0..21, components 22..38 in rows 24..40 (row 24 starts on a 128-byte boundary:
// the destination of the second TMA box)
constexpr int TS_XI0 = 0, TS_XI1 = 6, TS_C0 = 12, TS_C1 = 16, TS_DD = 20, TS_IDX2 = 24, TS_ABIDX = 28, TS_AB2 = 32,
              TS_JIR = 36, TS_JABR = 38, TS_RR = 40;
constexpr int TOP_NR = 41;   // stage rows
__device__ __forceinline__ int top_row(int c) { return c < 22 ? c : c + 2; }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- mbarrier + bulk-copy (TMA engine, no tensor map: rows are contiguous) primitives of the top kernel's ring
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, unsigned parity) {
  unsigned ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
  if (mbar_try_wait(b, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(b, parity))
    if (clock64() - t0 > 40000000000LL) __trap();   // ~20 s: a protocol bug must fail the launch, not hang the device
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Staging ring of the top kernel.  A tile starts anywhere in the sorted array (pair segments are not
// aligned), cp.async wants 16-byte-aligned sources: the ring rows hold the 4-aligned span that covers the
// tile, [start & ~3, start + count) rounded up, at most 128 + 3 + 3 -> TOP_ROW columns; residual l of the tile
// sits in column (start & 3) + l.
constexpr int TOP_STAGES = 4;
constexpr int TOP_ROW = 136;
constexpr int TOP_FROW = 160;   // flag rows: the 16-byte-aligned span that covers a tile, <= 128 + 15 + 15 bytes
template <int NR> struct alignas(128) TopStageT {
  float v[NR][TOP_ROW];
  unsigned char lin[TOP_FROW], act[TOP_FROW];
  int sp[TOP_ROW];           // slotpos of the tile's residuals (same 4-aligned span as the Jacobian rows)
};
using TopStage = TopStageT<TOP_NR>;
static_assert(sizeof(TopStage) % 128 == 0 && (24 * TOP_ROW * 4) % 128 == 0, "TMA box destinations are 128-byte aligned");
// Mode L (linearised residuals) stages J rows 8..78 in stage rows 0..70 (one tensor-map box of 71 rows: the geometric
// rows, the per-pixel rows JIdx / JabF and the 2x2 sums) and res_toZeroF in rows 72..79 (a second box): 80 rows.
constexpr int TL_XI0 = 0, TL_XI1 = 6, TL_C0 = 12, TL_C1 = 16, TL_DD = 20, TL_IDX0 = 22, TL_IDX1 = 30, TL_AB0 = 38, TL_AB1 = 46,
              TL_IDX2 = 54, TL_ABIDX = 58, TL_AB2 = 62, TL_RTZ = 72;
constexpr int TOP_NR_L = 80, TOP_NJ_L = 71;
using TopStageL = TopStageT<TOP_NR_L>;
static_assert(sizeof(TopStageL) % 128 == 0 && (TL_RTZ * TOP_ROW * 4) % 128 == 0, "TMA box destinations are 128-byte aligned");
constexpr int TOP_STAGES_L = 2;

// packed index of the 91 accumulators -> (row, col), row <= col, of the 13x13 block [C(4); xi(6); a, b; r]:
// 0..54 the row-major upper triangle of the 10x10, 55..84 the 10x3 top-right, 85..90 the upper triangle of the 3x3
__device__ __forceinline__ void top_unpack_index(int t, int& r, int& c) {
  if (t < 55) { r = 0; while (t >= 10 - r) { t -= 10 - r; ++r; } c = r + t; }
  else if (t < 85) { r = (t - 55) / 3; c = 10 + (t - 55) % 3; }
  else { const int u = t - 85; r = u < 3 ? 10 : (u < 5 ? 11 : 12); c = u < 3 ? 10 + u : (u < 5 ? 8 + u : 12); }
}
// chunk partial -> global.  The per-(mode, pair) fold of the partials into fp64 (chunks in index order) is a separate,
// wide launch (fold_top_slot, run by extra blocks of ba_fold_sc_kernel): folding here by "the last block of a pair" put a
// __threadfence (= wait for the tile's record stores), an atomic round trip and two more barriers into EVERY flush — 6.4 us
// per flush measured, during which half of the SM's warps and the block's TMA ring stood still.
__device__ __forceinline__ void top_flush(const BADesc& D, int chunk, int part, float (&acc)[56], int nres, float (*red)[33]) {
  // two shuffle levels take a warp's 32 lanes to 8, the remaining 8 lanes x 4 warps = 32 values per
  // accumulator are summed in a fixed order by one thread per accumulator (a full shuffle tree per
  // accumulator costs 10 instructions x 91 per warp; this is 4 x 91 + a 32-term sum by 92 threads)
  const int lane = threadIdx.x & 31, wid = (threadIdx.x >> 5) & (TOP_THREADS / 32 - 1);
  const int nacc = part == 0 ? 55 : 36, base = part == 0 ? 0 : 55;
#pragma unroll
  for (int i = 0; i < 55; ++i) {
    if (i < nacc) {
      float v = acc[i];
      v += __shfl_down_sync(0xffffffffu, v, 16);
      v += __shfl_down_sync(0xffffffffu, v, 8);
      if (lane < 8) red[base + i][wid * 8 + lane] = v;
    }
  }
  if (part == 1) {
    int q = warp_sum(nres);
    if (lane == 0) red[91][wid] = __int_as_float(q);
  }
  __syncthreads();
  float* out = D.toppartial + (size_t)chunk * TOP_PART;
  if (threadIdx.x < 92) {
    int t = threadIdx.x;
    if (t == 91) {
      int n = 0;
#pragma unroll
      for (int w = 0; w < TOP_THREADS / 32; ++w) n += __float_as_int(red[91][w]);
      out[91] = __int_as_float(n);
    } else {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 32; ++w) v += red[t][w];
      out[t] = v;
    }
  }
  __syncthreads();   // `red` is reused by the next chunk
}

// Fold of the chunk partials of one (mode, pair) slot into the dense fp64 13x13 accumulator that the stitch reads, chunks
// in index order; thread t < 92 owns accumulator t (91 = the residual count).  A slot without chunks reads as zero.
__device__ __forceinline__ void fold_top_slot(const BADesc& D, int slot, int t) {
  if (t >= 92) return;
  const int cb = D.pair_chunk_begin[slot], ce = D.pair_chunk_begin[slot + 1];
  double* dst = D.accTop + (size_t)slot * TOP_DENSE;
  if (t == 91) {
    int n = 0;
    for (int c = cb; c < ce; ++c) n += __float_as_int(__ldcg(D.toppartial + (size_t)c * TOP_PART + 91));
    dst[169] = (double)n;
  } else {
    double sum = 0.0;
    int c = cb;
    for (; c + 4 <= ce; c += 4) {   // four independent loads in flight, summed in chunk order
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = __ldcg(D.toppartial + (size_t)(c + q) * TOP_PART + t);
#pragma unroll
      for (int q = 0; q < 4; ++q) sum += (double)v[q];
    }
    for (; c < ce; ++c) sum += (double)__ldcg(D.toppartial + (size_t)c * TOP_PART + t);
    int rr, cc;
    top_unpack_index(t, rr, cc);
    dst[rr * 13 + cc] = sum; dst[cc * 13 + rr] = sum;
  }
}

// 10x10 upper triangle of [x y] [[a b],[b c]] [x y]^T into acc[0..54] (PART 0 threads).  Packed fp32 (FMUL2 / FFMA2 /
// FADD2): two neighbouring columns per instruction, every lane rounded exactly like  acc += fmaf(x, px, y * py).
__device__ __forceinline__ void top_acc_part0(float (&acc)[56], const float (&x)[10], const float (&y)[10], float a, float b, float c) {
  float2* acc2 = reinterpret_cast<float2*>(acc);
  int idx = 0;
#pragma unroll
  for (int rI = 0; rI < 10; ++rI) {
    const float px = fmaf(a, x[rI], b * y[rI]), py = fmaf(b, x[rI], c * y[rI]);
    int cI = rI;
    if (idx & 1) { acc[idx] += fmaf(x[cI], px, y[cI] * py); ++cI; ++idx; }
    const float2 px2 = make_float2(px, px), py2 = make_float2(py, py);
#pragma unroll
    for (; cI + 1 < 10; cI += 2, idx += 2) {
      const float2 t = __fmul2_rn(make_float2(y[cI], y[cI + 1]), py2);
      const float2 u = __ffma2_rn(make_float2(x[cI], x[cI + 1]), px2, t);
      acc2[idx >> 1] = __fadd2_rn(acc2[idx >> 1], u);
    }
    if (cI < 10) { acc[idx] += fmaf(x[cI], px, y[cI] * py); ++idx; }
  }
}
// 10x3 top-right (columns a, b, r; MatrixAccumulators.h:859-901) into acc[0..29] and the 3x3 bottom-right into acc[30..35]
// (PART 1 threads); packed: two rows (six accumulators) as three float2 triplets, each lane rounded like the scalar sequence.
__device__ __forceinline__ void top_acc_part1(float (&acc)[56], const float (&x)[10], const float (&y)[10], float AI00, float AI10,
                                              float AI01, float AI11, float JIr0, float JIr1, float AA00, float AA01, float AA11,
                                              float Jabr0, float Jabr1, float rr) {
  float2* acc2 = reinterpret_cast<float2*>(acc);
  const float2 A0 = make_float2(AI00, AI10), B0 = make_float2(AI01, AI11);
  const float2 A1 = make_float2(JIr0, AI00), B1 = make_float2(JIr1, AI01);
  const float2 A2 = make_float2(AI10, JIr0), B2 = make_float2(AI11, JIr1);
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    const float xr = x[2 * q], yr = y[2 * q], xs = x[2 * q + 1], ys = y[2 * q + 1];
    acc2[3 * q + 0] = __fadd2_rn(acc2[3 * q + 0], __ffma2_rn(make_float2(xr, xr), A0, __fmul2_rn(make_float2(yr, yr), B0)));
    acc2[3 * q + 1] = __fadd2_rn(acc2[3 * q + 1], __ffma2_rn(make_float2(xr, xs), A1, __fmul2_rn(make_float2(yr, ys), B1)));
    acc2[3 * q + 2] = __fadd2_rn(acc2[3 * q + 2], __ffma2_rn(make_float2(xs, xs), A2, __fmul2_rn(make_float2(ys, ys), B2)));
  }
  acc[30] += AA00; acc[31] += AA01; acc[32] += Jabr0; acc[33] += AA11; acc[34] += Jabr1; acc[35] += rr;
}

template <int MODE>
__global__ void __launch_bounds__(2 * TOP_THREADS, 2) ba_top_kernel(const BADesc* __restrict__ descs,
                                                                    const int* __restrict__ prefix, int nwin) {
  extern __shared__ __align__(128) unsigned char top_smem[];
  // modes A and L run on the staging ring (L with wider, fewer stages), the marginalisation pass on the direct path
  using Stage = typename std::conditional<MODE == 1, TopStageL, TopStage>::type;
  constexpr int STAGES = MODE == 1 ? TOP_STAGES_L : TOP_STAGES;
  constexpr bool RING = MODE != 2;
  // 128-byte alignment in the shared window (the launch reserves 128 spare bytes)
  Stage* ring = reinterpret_cast<Stage*>(top_smem + ((128u - (smem_u32(top_smem) & 127u)) & 127u));
  __shared__ float red[92][33];   // [accumulator][4 warps x 8 lanes], padded: conflict-free both ways
  int lb;
  const int win = find_window(prefix, nwin, blockIdx.x, lb);
  const BADesc& D = descs[win];
  const BARes& r = D.r;
  constexpr int mode = MODE;   // every mode has its own tile schedule: A over the non-linearised residuals, L over the
                               // linearised ones, marginalisation (2) over all of them
  const TopSched& SC = D.sched[MODE];
  if (lb >= SC.blocks) return;   // a batch launches max-over-windows blocks per window
  const TopTile* __restrict__ toptiles = SC.tiles;
  const int t0 = SC.block_begin[lb], t1 = SC.block_begin[lb + 1];
  const int part = threadIdx.x / TOP_THREADS;   // warp-uniform
  const int l = threadIdx.x & (TOP_THREADS - 1);
  __align__(8) float acc[56];   // 55 used (+1: packed pairs)
#pragma unroll
  for (int i = 0; i < 56; ++i) acc[i] = 0.f;
  int nres = 0, cur = -1;

  // The component rows of a tile are contiguous spans of the SoA Jacobian: TMA tensor-map loads into one stage of a
  // shared-memory ring, completing on the stage's `full` mbarrier.  The warps take turns as producer (warp j & 7
  // issues tile j + DEPTH, lanes in parallel), every warp releases a stage on its `empty` mbarrier when it has
  // read its residuals: no block-wide barrier in the tile loop, a warp only ever waits for data.
#ifndef TOP_DEPTH_
#define TOP_DEPTH_ (TOP_STAGES - 2)
#endif
  constexpr int DEPTH = MODE == 1 ? 1 : TOP_DEPTH_;   // S-2: the refilled stage was consumed two tiles ago, its release is old news
  constexpr int NW = 2 * TOP_THREADS / 32;
  __shared__ __align__(8) uint64_t full_bar[TOP_STAGES], empty_bar[TOP_STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (RING) {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], NW); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
  }
  auto produce = [&](int jp) {   // called by one whole warp
    const int tp = t0 + jp;
    if (tp >= t1) return;
    const int stg = jp % STAGES, use = jp / STAGES;
    if (use > 0) mbar_wait(&empty_bar[stg], (unsigned)(use - 1) & 1u);
    constexpr unsigned jbytes = (MODE == 1 ? (TOP_NJ_L + 8) : TOP_NC) * TOP_ROW * 4;
    if (lane == 0) mbar_expect_tx(&full_bar[stg], jbytes + 2 * TOP_FROW + TOP_ROW * 4);
    __syncwarp();
    const TopTile tlp = toptiles[tp];
    Stage& st = ring[stg];
    // the two flag rows and the record positions ride along (a consumer that loaded them itself had a DRAM round trip
    // on its critical path)
    if (lane == 2) bulk_g2s(st.lin, r.isLin + (tlp.start & ~15), TOP_FROW, &full_bar[stg]);
    else if (lane == 3) bulk_g2s(st.act, r.isActive + (tlp.start & ~15), TOP_FROW, &full_bar[stg]);
    else if (lane == 4) bulk_g2s(st.sp, r.slotpos + (tlp.start & ~3), TOP_ROW * 4, &full_bar[stg]);
    // two tensor-map loads per stage: the TMA unit serves one request in ~46 cycles whatever its size, one bulk copy
    // per component row made the kernel TMA-issue-bound (3.4 TB/s)
    if (MODE == 1) {
      if (lane == 0) tma_load_2d(&st.v[0][0], &D.tmapJ[2], tlp.start & ~3, J_XI0, &full_bar[stg]);
      else if (lane == 1) tma_load_2d(&st.v[TL_RTZ][0], &D.tmapJ[3], tlp.start & ~3, 0, &full_bar[stg]);
    } else {
      if (lane == 0) tma_load_2d(&st.v[0][0], &D.tmapJ[0], tlp.start & ~3, J_XI0, &full_bar[stg]);
      else if (lane == 1) tma_load_2d(&st.v[24][0], &D.tmapJ[1], tlp.start & ~3, J_IDX2, &full_bar[stg]);
    }
  };
  if (RING && warp == 0) {
#pragma unroll
    for (int i = 0; i < DEPTH; ++i) produce(i);
  }
  float dp[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // mode L: adHTdeltaF of the current tile's pair
  int dp_pair = -1;

  TopTile tl_next = t0 < t1 ? toptiles[t0] : TopTile{0, 0, 0, 0};
  for (int t = t0; t < t1; ++t) {
    const TopTile tl = tl_next;
    if (t + 1 < t1) tl_next = toptiles[t + 1];   // the next descriptor is in flight while this tile is consumed
    const int j = t - t0;
    const int chunk = tl.chunk;
    if (chunk != cur) {
      if (cur >= 0) {
        top_flush(D, cur, part, acc, nres, red);
#pragma unroll
        for (int i = 0; i < 55; ++i) acc[i] = 0.f;
        nres = 0;
      }
      cur = chunk;
    }
    if (!RING) {
      const TopChunk ch = {tl.pair, tl.start, tl.count, 1};
      if (part == 0) top_accumulate<(MODE ? MODE : 1), 0>(D, ch, acc, nres); else top_accumulate<(MODE ? MODE : 1), 1>(D, ch, acc, nres);
      continue;
    }
    if (warp == (j & (NW - 1))) produce(j + DEPTH);
    // mode L: what does not come through the ring is requested before the wait (its latency hides behind the stage's)
    float dd = 0.f;
    if (MODE == 1 && part == 1) {
      if (l < tl.count) dd = D.p.deltaF[r.point[tl.start + l]];
      if (tl.pair != dp_pair) {
        dp_pair = tl.pair;
#pragma unroll
        for (int q = 0; q < 8; ++q) dp[q] = __ldg(D.W.adHTdeltaF + 8 * tl.pair + q);
      }
    }
    mbar_wait(&full_bar[j % STAGES], (unsigned)(j / STAGES) & 1u);
    const Stage& S = ring[j % STAGES];
    const int col = (tl.start & 3) + l, cf = (tl.start & 15) + l;
    constexpr int RC0 = MODE == 1 ? TL_C0 : TS_C0, RC1 = MODE == 1 ? TL_C1 : TS_C1, RXI0 = MODE == 1 ? TL_XI0 : TS_XI0,
                  RXI1 = MODE == 1 ? TL_XI1 : TS_XI1, RIDX2 = MODE == 1 ? TL_IDX2 : TS_IDX2, RABIDX = MODE == 1 ? TL_ABIDX : TS_ABIDX,
                  RAB2 = MODE == 1 ? TL_AB2 : TS_AB2, RDD = MODE == 1 ? TL_DD : TS_DD;
    // AccumulatedTopHessian.cpp:37-49: mode 0 = !isLinearized && isActive, mode 1 = isLinearized && isActive
    if (l < tl.count && (S.lin[cf] != 0) == (MODE == 1) && S.act[cf] != 0) {
      float x[10], y[10];
#pragma unroll
      for (int q = 0; q < 4; ++q) { x[q] = S.v[RC0 + q][col]; y[q] = S.v[RC1 + q][col]; }
#pragma unroll
      for (int q = 0; q < 6; ++q) { x[4 + q] = S.v[RXI0 + q][col]; y[4 + q] = S.v[RXI1 + q][col]; }
      const float a = S.v[RIDX2 + 0][col], b = S.v[RIDX2 + 2][col], c = S.v[RIDX2 + 3][col];
      if (part == 0) {
        top_acc_part0(acc, x, y, a, b, c);
      } else {
        const float AI00 = S.v[RABIDX + 0][col], AI10 = S.v[RABIDX + 1][col], AI01 = S.v[RABIDX + 2][col], AI11 = S.v[RABIDX + 3][col];
        const float AA00 = S.v[RAB2 + 0][col], AA01 = S.v[RAB2 + 2][col], AA11 = S.v[RAB2 + 3][col];
        const float d0 = S.v[RDD][col], d1 = S.v[RDD + 1][col];
        float JIr0, JIr1, Jabr0, Jabr1, rr;
        if (MODE == 1) {
          // resApprox = res_toZeroF + J*delta, AccumulatedTopHessian.cpp:63-81 (sequential sums)
          float dx = 0.f, dy = 0.f, cxs = 0.f, cys = 0.f;
#pragma unroll
          for (int q = 0; q < 6; ++q) { dx = xadd(dx, xmul(x[4 + q], dp[q])); dy = xadd(dy, xmul(y[4 + q], dp[q])); }
#pragma unroll
          for (int q = 0; q < 4; ++q) { cxs = xadd(cxs, xmul(x[q], D.W.cDeltaF[q])); cys = xadd(cys, xmul(y[q], D.W.cDeltaF[q])); }
          const float Jpx = xadd(xadd(dx, cxs), xmul(d0, dd)), Jpy = xadd(xadd(dy, cys), xmul(d1, dd));
          JIr0 = JIr1 = Jabr0 = Jabr1 = rr = 0.f;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float i0 = S.v[TL_IDX0 + q][col], i1 = S.v[TL_IDX1 + q][col], a0 = S.v[TL_AB0 + q][col], a1 = S.v[TL_AB1 + q][col];
            float rz = S.v[TL_RTZ + q][col];
            rz = xadd(rz, xmul(i0, Jpx)); rz = xadd(rz, xmul(i1, Jpy));
            rz = xadd(rz, xmul(a0, dp[6])); rz = xadd(rz, xmul(a1, dp[7]));
            JIr0 = xadd(JIr0, xmul(rz, i0)); JIr1 = xadd(JIr1, xmul(rz, i1));
            Jabr0 = xadd(Jabr0, xmul(rz, a0)); Jabr1 = xadd(Jabr1, xmul(rz, a1));
            rr = xadd(rr, xmul(rz, rz));
          }
        } else {
          JIr0 = S.v[TS_JIR][col]; JIr1 = S.v[TS_JIR + 1][col]; Jabr0 = S.v[TS_JABR][col]; Jabr1 = S.v[TS_JABR + 1][col];
          rr = S.v[TS_RR][col];
        }
        top_acc_part1(acc, x, y, AI00, AI10, AI01, AI11, JIr0, JIr1, AA00, AA01, AA11, Jabr0, Jabr1, rr);
        nres++;
        // per-residual contribution to the point sums (:114-117), exact ops: they are summed per point in
        // the point's own order by the Schur kernel, which keeps Hdd/bd/Hcd bit-identical to the oracle.
        const float Ji0 = xadd(xmul(a, d0), xmul(b, d1)), Ji1 = xadd(xmul(b, d0), xmul(c, d1));
        // the residual's 32-byte record in its point's slot: one 256-bit store
#if defined(TOP_EXP_LINREC)      /* timing experiment only: records in residual order (coalesced) */
        stg256(r.prec + 8 * (size_t)(tl.start + l),
#elif defined(TOP_EXP_NOREC)     /* timing experiment only: no record store */
        if (rr == 123456.f) stg256(r.prec + S.sp[col],
#else
        stg256(r.prec + S.sp[col],
#endif
               make_float4(xadd(xmul(JIr0, d0), xmul(JIr1, d1)), xadd(xmul(Ji0, d0), xmul(Ji1, d1)),
                           xadd(xmul(x[0], Ji0), xmul(y[0], Ji1)), xadd(xmul(x[1], Ji0), xmul(y[1], Ji1))),
               make_float4(xadd(xmul(x[2], Ji0), xmul(y[2], Ji1)), xadd(xmul(x[3], Ji0), xmul(y[3], Ji1)), MODE == 1 ? 2.f : 1.f,
                           (float)((tl.pair * D.W.n_rcp) >> 16)));
      }
    } else if (part == 1 && l < tl.count) {
      // inactive residual: its record says so (the Schur kernel and the point sums skip it)
      stg256(r.prec + S.sp[col], make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, (float)((tl.pair * D.W.n_rcp) >> 16)));
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[j % STAGES]);   // this warp is done with the stage
  }
  if (cur >= 0) top_flush(D, cur, part, acc, nres, red);
}

// Per-point sums of Hdd/bd/Hcd over the point's residuals in the point's own order (the tail of
// addPoint<mode>, AccumulatedTopHessian.cpp:122-139).  Only used by the stand-alone accumulate(A|L)
// entry points; the fused path does this inside the Schur kernel.
__global__ void ba_point_sum_kernel(BARes r, BAPts pts, int do_A, int do_L, int maxres) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pts.P) return;
  float bdv[2] = {0.f, 0.f}, Hddv[2] = {0.f, 0.f}, Hcdv[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  const float4* rec = reinterpret_cast<const float4*>(r.prec + (size_t)p * r.pstr);
  for (int j = 0; j < maxres; ++j) {
    const float4 q0 = rec[2 * j], q1 = rec[2 * j + 1];
    if (q1.z == 0.f) continue;            // no residual in this slot, or inactive
    const int m = q1.z == 2.f ? 1 : 0;
    if ((m == 0 && !do_A) || (m == 1 && !do_L)) continue;
    bdv[m] = xadd(bdv[m], q0.x);
    Hddv[m] = xadd(Hddv[m], q0.y);
    Hcdv[m][0] = xadd(Hcdv[m][0], q0.z); Hcdv[m][1] = xadd(Hcdv[m][1], q0.w);
    Hcdv[m][2] = xadd(Hcdv[m][2], q1.x); Hcdv[m][3] = xadd(Hcdv[m][3], q1.y);
  }
  if (do_A) { pts.Hdd[0][p] = Hddv[0]; pts.bd[0][p] = bdv[0]; pts.Hcd[0][p] = make_float4(Hcdv[0][0], Hcdv[0][1], Hcdv[0][2], Hcdv[0][3]); }
  if (do_L) { pts.Hdd[1][p] = Hddv[1]; pts.bd[1][p] = bdv[1]; pts.Hcd[1][p] = make_float4(Hcdv[1][0], Hcdv[1][1], Hcdv[1][2], Hcdv[1][3]); }
}

// ------------------------------------------------------------------------------------------------
// K6: AccumulatedSCHessianSSE::addPoint — the per-point Schur fold (+ the tail of addPoint<mode>:
// per-point sums of Hdd/bd/Hcd over the point's residuals in the point's own order).
// For the points of one host the reference's accD[(h,t1,t2)], accE[(h,t1)], accEB[(h,t1)], accHcc,
// accbc are all blocks of ONE weighted Gram matrix  G_h = sum_p HdiF_p * a_p a_p^T  of the vector
// a_p = [JpJdF(p->t) for t = 0..n-1 (zero where there is no active residual) ; Hcd(4) ; bdSumF].
// A block stages SC_PTS such vectors in shared memory and every thread owns 4x4 register tiles of
// the upper triangle of G; points are visited sequentially, so the sum order is fixed.  fp32 SIMT:
// the matrix is only (8n+5) wide — tensor cores are deliberately not used (north_star).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tri_decode(int t, int NT, int& ti, int& tj) {
  // t-th tile of the upper triangle enumerated row by row: row ti has NT-ti tiles
  int row = 0, rem = t;
  while (rem >= NT - row) { rem -= NT - row; row++; }
  ti = row; tj = row + rem;
}

// Data path.  The residual records of a point — JpJdF (32 B) and {bd, Hdd, Hcd[4], flag, target} (32 B) per residual
// slot — are stored POINT-MAJOR (BARes::JpJdF / prec, pstr floats per point), so the records of SC_PTS consecutive points
// are two contiguous spans: one elected thread fetches them with two bulk copies (cp.async.bulk, TMA engine) that
// complete on an mbarrier, for sub-tile i+1 while sub-tile i is in the Gram phase.  No slot table, no flag loads, no
// gathers, no global load on the critical path.
// Phase 1 (four lanes per point, only __syncwarp inside): the six ordered per-point sums {bd, Hdd, Hcd[0..3]} for modes
// A and L IN THE POINT'S OWN ORDER (AccumulatedTopHessian.cpp:35-139), the Schur scalars (AccumulatedSCHessian.cpp:12-33),
// and the row  s_p = sqrt(HdiF_p) * a_p  into shared memory, so that the Gram phase needs ONE matrix:
// G_h = sum_p s_p s_p^T  (= sum_p HdiF_p a_p a_p^T up to fp32 rounding; HdiF / bdSumF / Hdd / bd / Hcd themselves are
// written exactly as before).
// Phase 2: 4x4 register tiles of the upper triangle of G WITHOUT the two tile rows/columns of the host's own 8-block (a
// point has no residual into its host).  Each tile is owned by TWO threads (tid and tid+128), one for the even and one
// for the odd points of the sub-tile.  MAXT = tiles per thread: (NT-2)(NT-1)/2 <= MAXT*128 (1: n <= 7, 2: n <= 10, 5: n <= 16).
// Shared memory (floats): A[SC_PTS][LD+4] | jp[SC_PTS*pstr] | pt[SC_PTS*pstr] | mbarrier.
// floats per point in the record arrays: 32-byte records (256-bit stores) -> a multiple of 8; maxres + 1 records make the
// stride an odd multiple of 8 for even maxres (2-way instead of 4-way bank conflicts on the scalar reads of phase 1)
__host__ __device__ inline int sc_pstr(int maxres) { return 8 * (maxres + 1); }
__host__ __device__ inline size_t sc_smem_floats(int LD, int maxres) {
  return (size_t)SC_PTS * (LD + 4) + 2 * (size_t)SC_PTS * sc_pstr(maxres) + 4;
}

template <int MAXT>
#ifndef SC_MINB
#define SC_MINB 3
#endif
__global__ void __launch_bounds__(SC_THREADS, MAXT <= 2 ? SC_MINB : 2)
ba_schur_kernel(const BADesc* __restrict__ descs, const int* __restrict__ prefix, int nwin) {
  static_assert(SC_PTS * 4 == SC_THREADS, "phase 1 gives four lanes to every point of a sub-tile");
  extern __shared__ __align__(16) float smem[];
  int lb;
  const int win = find_window(prefix, nwin, blockIdx.x, lb);
  const BADesc& D = descs[win];
  const BARes& r = D.r;
  const BAPts& pts = D.p;
  const int LD = D.LD, LDP = LD + 4;
  const int maxres = D.maxres, PSTR = r.pstr;
  float* A = smem;                               // [SC_PTS][LDP]  sqrt(HdiF_p) * a_p
  float* jp = A + SC_PTS * LDP;                  // [SC_PTS][PSTR] staged JpJdF records
  float* pt = jp + SC_PTS * PSTR;                // [SC_PTS][PSTR] staged point-term records
  uint64_t* full = reinterpret_cast<uint64_t*>(pt + SC_PTS * PSTR);
  const PtChunk ch = D.ptchunks[lb];
  const int n = D.W.n;
  const int NT = LD / 4, NTR = NT - 2, ntiles = NTR * (NTR + 1) / 2;
  const int half = threadIdx.x >> 7;             // 0: even points of the sub-tile, 1: odd points
  int tI[MAXT], tJ[MAXT];
  __align__(8) float acc[MAXT][16];
#pragma unroll
  for (int j = 0; j < MAXT; ++j) {
    int tile = (threadIdx.x & 127) + j * 128;
    tI[j] = tJ[j] = -1;
    if (tile < ntiles) {
      tri_decode(tile, NTR, tI[j], tJ[j]);
      if (tI[j] >= 2 * ch.host) tI[j] += 2;   // skip the host's own two tile rows / columns
      if (tJ[j] >= 2 * ch.host) tJ[j] += 2;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[j][i] = 0.f;
  }
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  auto fetch = [&](int base) {   // thread 0: the records of sub-tile [base, base + SC_PTS) of the chunk
    const int cnt = min(SC_PTS, ch.count - base);
    if (cnt <= 0) return;
    const unsigned bytes = (unsigned)cnt * PSTR * 4;
    mbar_expect_tx(full, 2 * bytes);
    bulk_g2s(jp, r.JpJdF + (size_t)(ch.start + base) * PSTR, bytes, full);
    bulk_g2s(pt, r.prec + (size_t)(ch.start + base) * PSTR, bytes, full);
  };
  if (threadIdx.x == 0) fetch(0);
  const int lp = threadIdx.x >> 2, q = threadIdx.x & 3;   // phase 1: point of the sub-tile, lane within the point
  // priorF / deltaF of this thread's point, loaded one sub-tile ahead (lanes 0 and 1 of a point need them)
  float pr_n = 0.f, de_n = 0.f;
  if (q < 2 && lp < ch.count) { pr_n = pts.priorF[ch.start + lp]; if (q == 0) de_n = pts.deltaF[ch.start + lp]; }
  unsigned parity = 0;
  for (int base = 0; base < ch.count; base += SC_PTS, parity ^= 1u) {
    const int cnt = min(SC_PTS, ch.count - base);
    const float prc = pr_n, dec = de_n;
    {
      const int nb = base + SC_PTS + lp;
      pr_n = de_n = 0.f;
      if (q < 2 && nb < ch.count) { pr_n = pts.priorF[ch.start + nb]; if (q == 0) de_n = pts.deltaF[ch.start + nb]; }
    }
    mbar_wait(full, parity);
    // ---- phase 1
    {
      float* row = A + lp * LDP;
      // zero the row (blocks without an active residual, the host's own block, the padding columns)
      for (int i = q; i < LDP / 4; i += 4) reinterpret_cast<float4*>(row)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int p = ch.start + base + lp;
      const bool live = lp < cnt && (!D.marg || D.marg[p] != 0);   // marginalisation pass: flagged points only
      float hdi = 0.f, mine0 = 0.f, mine1 = 0.f;   // this lane's entries of a_p: component q, and q+4 for lanes 0 / 1
      if (live) {
        const float* pl = pt + lp * PSTR;
        // branch-free ordered sums: a slot that is absent / inactive adds +0 (exact)
        float s0[2] = {0.f, 0.f}, s1[2] = {0.f, 0.f};
        int ngood = 0;
#pragma unroll 2
        for (int slot = 0; slot < maxres; ++slot) {
          const float f = pl[8 * slot + 6], v0 = pl[8 * slot + q], v1 = pl[8 * slot + (q < 2 ? q + 4 : q)];
          ngood += (f != 0.f) ? 1 : 0;
          s0[0] = xadd(s0[0], f == 1.f ? v0 : 0.f); s0[1] = xadd(s0[1], f == 2.f ? v0 : 0.f);
          s1[0] = xadd(s1[0], f == 1.f ? v1 : 0.f); s1[1] = xadd(s1[1], f == 2.f ? v1 : 0.f);
        }
        // components: 0 = bd, 1 = Hdd, 2..5 = Hcd[0..3]; lane q owns q, lanes 0 / 1 also own 4 / 5
        if (q == 0) {
          pts.bd[0][p] = s0[0]; pts.bd[1][p] = s0[1];
          if (ngood > 0) {
            mine0 = xadd(s0[0], s0[1]);
            if (D.shift_prior) mine0 = xadd(mine0, xmul(prc, dec));   // shiftPriorToZero (EnergyFunctional.cpp:209)
          }
          pts.bdSumF[p] = mine0;
        } else if (q == 1) {
          pts.Hdd[0][p] = s0[0]; pts.Hdd[1][p] = s0[1];
          float idh = 0.f;
          if (ngood > 0) {
            float H = xadd(xadd(s0[0], s0[1]), prc);
            if (H < 1e-10f) H = 1e-10f;
            idh = H;
            hdi = xdiv(1.0f, H);
          }
          pts.HdiF[p] = hdi; pts.idepth_hessian[p] = idh;
        } else {
          reinterpret_cast<float*>(pts.Hcd[0])[4 * (size_t)p + (q - 2)] = s0[0];
          reinterpret_cast<float*>(pts.Hcd[1])[4 * (size_t)p + (q - 2)] = s0[1];
          mine0 = xadd(s0[0], s0[1]);
        }
        if (q < 2) {
          reinterpret_cast<float*>(pts.Hcd[0])[4 * (size_t)p + (q + 2)] = s1[0];
          reinterpret_cast<float*>(pts.Hcd[1])[4 * (size_t)p + (q + 2)] = s1[1];
          mine1 = xadd(s1[0], s1[1]);
        }
      }
      // the four lanes of a point share sqrt(HdiF) (lane 1 holds HdiF)
      const float sw = sqrtf(__shfl_sync(0xffffffffu, hdi, (threadIdx.x & 28) + 1));
      __syncwarp();   // the row is zeroed
      if (live) {
        const int colb = 8 * n;   // a_p tail: Hcd[0..3] at 8n..8n+3, bdSum at 8n+4
        if (q == 0) { row[colb + 4] = sw * mine0; row[colb + 2] = sw * mine1; }
        else if (q == 1) { row[colb + 3] = sw * mine1; }
        else { row[colb + (q - 2)] = sw * mine0; }
        const float* pl = pt + lp * PSTR;
        const float* jl = jp + lp * PSTR;
        for (int slot = q; slot < maxres; slot += 4) {
          if (pl[8 * slot + 6] == 0.f) continue;
          const int t = (int)pl[8 * slot + 7];
          const float4 j0 = *reinterpret_cast<const float4*>(jl + 8 * slot), j1 = *reinterpret_cast<const float4*>(jl + 8 * slot + 4);
          *reinterpret_cast<float4*>(row + 8 * t) = make_float4(sw * j0.x, sw * j0.y, sw * j0.z, sw * j0.w);
          *reinterpret_cast<float4*>(row + 8 * t + 4) = make_float4(sw * j1.x, sw * j1.y, sw * j1.z, sw * j1.w);
        }
      }
    }
    __syncthreads();   // rows complete; the staged records are free again
    if (threadIdx.x == 0) fetch(base + SC_PTS);   // lands while the Gram phase runs
    // ---- phase 2: G += sum_p s_p s_p^T, 4x4 register tiles, this thread's half of the points in order
#pragma unroll
    for (int j = 0; j < MAXT; ++j) {
      if (tI[j] >= 0) {
        const float* ap = A + 4 * tI[j];
        const float* bp = A + 4 * tJ[j];
#pragma unroll 4
        for (int l2 = half; l2 < cnt; l2 += 2) {
          const float4 av = *reinterpret_cast<const float4*>(ap + l2 * LDP);
          const float4 bv = *reinterpret_cast<const float4*>(bp + l2 * LDP);
          const float a4[4] = {av.x, av.y, av.z, av.w};
          // packed fp32 FMA (sm_100 FFMA2): two accumulators per instruction, each product rounded exactly like fmaf
          float2* acc2 = reinterpret_cast<float2*>(acc[j]);
          const float2 b01 = make_float2(bv.x, bv.y), b23 = make_float2(bv.z, bv.w);
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) {
            const float2 aa = make_float2(a4[ii], a4[ii]);
            acc2[2 * ii] = __ffma2_rn(aa, b01, acc2[2 * ii]);
            acc2[2 * ii + 1] = __ffma2_rn(aa, b23, acc2[2 * ii + 1]);
          }
        }
      }
    }
    __syncthreads();   // A is free again
  }
  // the odd-point halves go through shared memory (A is free now), the even-point threads add and store
  float* out = D.scpartial + (size_t)lb * LD * LD;
#pragma unroll
  for (int j = 0; j < MAXT; ++j) {
    float* xch = A + (size_t)((threadIdx.x & 127) + j * 128) * 16;   // <= MAXT*128*16 floats (the host sizes the block for it)
    if (tI[j] >= 0 && half == 1) {
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
        *reinterpret_cast<float4*>(xch + 4 * ii) = make_float4(acc[j][4 * ii], acc[j][4 * ii + 1], acc[j][4 * ii + 2], acc[j][4 * ii + 3]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < MAXT; ++j) {
    const float* xch = A + (size_t)((threadIdx.x & 127) + j * 128) * 16;
    if (tI[j] >= 0 && half == 0) {
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) {
        const float4 o = *reinterpret_cast<const float4*>(xch + 4 * ii);
        *reinterpret_cast<float4*>(out + (size_t)(4 * tI[j] + ii) * LD + 4 * tJ[j]) =
            make_float4(acc[j][4 * ii] + o.x, acc[j][4 * ii + 1] + o.y, acc[j][4 * ii + 2] + o.z, acc[j][4 * ii + 3] + o.w);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Schur fold, 8x8 register tiles (n <= 10).  Same data path and the same phase 1 as ba_schur_kernel; the Gram phase is
// re-blocked so that it no longer lives on shared-memory bandwidth (the 4x4 version: L1 wavefronts 86 %, two LDS.128 per
// 8 FFMA2):
//   * a thread owns ONE 8x8 tile of the upper block triangle (8-blocks = the frames + the tail block [Hcd | bdSum | pad],
//     without the host's own block) and 1/SLICES of the sub-tile's points: 4 LDS.128 per 32 FFMA2;
//   * the lanes of a warp are different tiles of the SAME point, and a row of A is stored with the low halves of all
//     8-blocks first and the high halves behind them (sc8_col), so one operand load of a warp touches one contiguous
//     <= 128-byte piece of one row: a single wavefront with broadcast;
//   * the SLICES partial tiles of a thread group are summed through shared memory in a fixed binary tree at the end of
//     the chunk (the staging buffers are free then).
//   * the staged records are double-buffered: the bulk copies of sub-tile i+1 are issued BEFORE phase 1 of sub-tile i
//     (single-buffered, a block's time per sub-tile was phase 1 + the copy latency: 91 us of the launch).
// SLICES = 8: <= 32 tiles (n <= 7); SLICES = 4: <= 64 tiles (n <= 10).  Shared memory (floats): 2 mbarriers (4) |
// A[SC_PTS][LD+4] | {jp, pt} x 2 stages, at least 4 + 8192 for the final reduction.
__host__ __device__ inline int sc8_col(int c, int LD) { return ((c & 4) ? (LD >> 1) : 0) + ((c >> 3) << 2) + (c & 3); }
__host__ __device__ inline size_t sc8_smem_floats(int LD, int maxres) {
  const size_t a = (size_t)SC_PTS * (LD + 4) + 4 * (size_t)SC_PTS * sc_pstr(maxres);   // A + two stages of {jp, pt}
  return 4 + (a > 8192 ? a : 8192);
}
template <int SLICES>
__global__ void __launch_bounds__(SC_THREADS, 2)
ba_schur8_kernel(const BADesc* __restrict__ descs, const int* __restrict__ prefix, int nwin) {
  static_assert(SC_PTS * 4 == SC_THREADS, "phase 1 gives four lanes to every point of a sub-tile");
  constexpr int TL = SC_THREADS / SLICES;   // tile lanes per slice
  extern __shared__ __align__(16) float smem[];
  int lb;
  const int win = find_window(prefix, nwin, blockIdx.x, lb);
  const BADesc& D = descs[win];
  const BARes& r = D.r;
  const BAPts& pts = D.p;
  const int LD = D.LD, LDP = LD + 4, HALF = LD >> 1;
  const int maxres = D.maxres, PSTR = r.pstr;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  float* A = smem + 4;                           // [SC_PTS][LDP]  sqrt(HdiF_p) * a_p, columns permuted by sc8_col
  float* stage0 = A + SC_PTS * LDP;              // two stages of { jp[SC_PTS][PSTR] JpJdF records | pt[SC_PTS][PSTR] point-term records }
  const int STG = 2 * SC_PTS * PSTR;
  const PtChunk ch = D.ptchunks[lb];
  const int n = D.W.n;
  const int NB = LD >> 3, NBR = NB - 1, ntiles = NBR * (NBR + 1) / 2;
  const int slice = threadIdx.x / TL, tile = threadIdx.x % TL;
  int bi = -1, bj = -1;
  if (tile < ntiles) {
    tri_decode(tile, NBR, bi, bj);
    if (bi >= ch.host) bi++;   // skip the host's own block row / column
    if (bj >= ch.host) bj++;
  }
  __align__(8) float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1); mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  auto fetch = [&](int base, int stg) {   // thread 0: the records of sub-tile [base, base + SC_PTS) of the chunk
    const int cnt = min(SC_PTS, ch.count - base);
    if (cnt <= 0) return;
    const unsigned bytes = (unsigned)cnt * PSTR * 4;
    float* jpd = stage0 + stg * STG;
    mbar_expect_tx(&full[stg], 2 * bytes);
    bulk_g2s(jpd, r.JpJdF + (size_t)(ch.start + base) * PSTR, bytes, &full[stg]);
    bulk_g2s(jpd + SC_PTS * PSTR, r.prec + (size_t)(ch.start + base) * PSTR, bytes, &full[stg]);
  };
  if (threadIdx.x == 0) fetch(0, 0);
  const int lp = threadIdx.x >> 2, q = threadIdx.x & 3;   // phase 1: point of the sub-tile, lane within the point
  float pr_n = 0.f, de_n = 0.f;
  if (q < 2 && lp < ch.count) { pr_n = pts.priorF[ch.start + lp]; if (q == 0) de_n = pts.deltaF[ch.start + lp]; }
  for (int base = 0, it = 0; base < ch.count; base += SC_PTS, ++it) {
    const int cnt = min(SC_PTS, ch.count - base);
    const int stg = it & 1;
    // the other stage was consumed by phase 1 of the previous sub-tile (two block barriers ago): refill it now, the
    // copy runs under this sub-tile's phase 1 and Gram phase
    if (threadIdx.x == 0) fetch(base + SC_PTS, stg ^ 1);
    const float* jp = stage0 + stg * STG;
    const float* pt = jp + SC_PTS * PSTR;
    const float prc = pr_n, dec = de_n;
    {
      const int nb = base + SC_PTS + lp;
      pr_n = de_n = 0.f;
      if (q < 2 && nb < ch.count) { pr_n = pts.priorF[ch.start + nb]; if (q == 0) de_n = pts.deltaF[ch.start + nb]; }
    }
    mbar_wait(&full[stg], (unsigned)(it >> 1) & 1u);
    // ---- phase 1 (as in ba_schur_kernel; only the column positions in the row differ)
    {
      float* row = A + lp * LDP;
      for (int i = q; i < LDP / 4; i += 4) reinterpret_cast<float4*>(row)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int p = ch.start + base + lp;
#ifdef SC8_EXP_NOP1
      const bool live = lp < cnt && (!D.marg || D.marg[p] != 0) && cnt == 12345;
#else
      const bool live = lp < cnt && (!D.marg || D.marg[p] != 0);
#endif
      float hdi = 0.f, mine0 = 0.f, mine1 = 0.f;
      if (live) {
        const float* pl = pt + lp * PSTR;
        float s0[2] = {0.f, 0.f}, s1[2] = {0.f, 0.f};
        int ngood = 0;
        // the first six slots fully unrolled (all 18 loads in flight at once); a slot index beyond the point's list reads
        // the always-zero pad record (flag 0: adds +0, exact)
        const int q1 = q < 2 ? q + 4 : q;
#pragma unroll
        for (int slot = 0; slot < 6; ++slot) {
          const int sl = 8 * min(slot, maxres);
          const float f = pl[sl + 6], v0 = pl[sl + q], v1 = pl[sl + q1];
          ngood += (f != 0.f) ? 1 : 0;
          s0[0] = xadd(s0[0], f == 1.f ? v0 : 0.f); s0[1] = xadd(s0[1], f == 2.f ? v0 : 0.f);
          s1[0] = xadd(s1[0], f == 1.f ? v1 : 0.f); s1[1] = xadd(s1[1], f == 2.f ? v1 : 0.f);
        }
        for (int slot = 6; slot < maxres; ++slot) {
          const float f = pl[8 * slot + 6], v0 = pl[8 * slot + q], v1 = pl[8 * slot + q1];
          ngood += (f != 0.f) ? 1 : 0;
          s0[0] = xadd(s0[0], f == 1.f ? v0 : 0.f); s0[1] = xadd(s0[1], f == 2.f ? v0 : 0.f);
          s1[0] = xadd(s1[0], f == 1.f ? v1 : 0.f); s1[1] = xadd(s1[1], f == 2.f ? v1 : 0.f);
        }
        if (q == 0) {
          pts.bd[0][p] = s0[0]; pts.bd[1][p] = s0[1];
          if (ngood > 0) {
            mine0 = xadd(s0[0], s0[1]);
            if (D.shift_prior) mine0 = xadd(mine0, xmul(prc, dec));
          }
          pts.bdSumF[p] = mine0;
        } else if (q == 1) {
          pts.Hdd[0][p] = s0[0]; pts.Hdd[1][p] = s0[1];
          float idh = 0.f;
          if (ngood > 0) {
            float H = xadd(xadd(s0[0], s0[1]), prc);
            if (H < 1e-10f) H = 1e-10f;
            idh = H;
            hdi = xdiv(1.0f, H);
          }
          pts.HdiF[p] = hdi; pts.idepth_hessian[p] = idh;
        } else {
          reinterpret_cast<float*>(pts.Hcd[0])[4 * (size_t)p + (q - 2)] = s0[0];
          reinterpret_cast<float*>(pts.Hcd[1])[4 * (size_t)p + (q - 2)] = s0[1];
          mine0 = xadd(s0[0], s0[1]);
        }
        if (q < 2) {
          reinterpret_cast<float*>(pts.Hcd[0])[4 * (size_t)p + (q + 2)] = s1[0];
          reinterpret_cast<float*>(pts.Hcd[1])[4 * (size_t)p + (q + 2)] = s1[1];
          mine1 = xadd(s1[0], s1[1]);
        }
      }
      const float sw = sqrtf(__shfl_sync(0xffffffffu, hdi, (threadIdx.x & 28) + 1));
      __syncwarp();   // the row is zeroed
      if (live) {
        // a_p tail (logical columns 8n .. 8n+4): Hcd[0..3] = low half of block n, bdSum = first entry of its high half
        float* lo = row + 4 * n;
        float* hi = row + HALF + 4 * n;
        if (q == 0) { hi[0] = sw * mine0; lo[2] = sw * mine1; }
        else if (q == 1) { lo[3] = sw * mine1; }
        else { lo[q - 2] = sw * mine0; }
        const float* pl = pt + lp * PSTR;
        const float* jl = jp + lp * PSTR;
        for (int slot = q; slot < maxres; slot += 4) {
          if (pl[8 * slot + 6] == 0.f) continue;
          const int t = (int)pl[8 * slot + 7];
          const float4 j0 = *reinterpret_cast<const float4*>(jl + 8 * slot), j1 = *reinterpret_cast<const float4*>(jl + 8 * slot + 4);
          *reinterpret_cast<float4*>(row + 4 * t) = make_float4(sw * j0.x, sw * j0.y, sw * j0.z, sw * j0.w);
          *reinterpret_cast<float4*>(row + HALF + 4 * t) = make_float4(sw * j1.x, sw * j1.y, sw * j1.z, sw * j1.w);
        }
      }
    }
    __syncthreads();   // rows complete
    // ---- phase 2: this thread's 8x8 tile over its slice of the points, in point order
#ifdef SC8_EXP_NOGRAM
    if (bi >= 0 && cnt == 12345) {
#else
    if (bi >= 0) {
#endif
      const float* ap = A + 4 * bi;
      const float* bp = A + 4 * bj;
#pragma unroll 2
      for (int l2 = slice; l2 < cnt; l2 += SLICES) {
        const float4 a0 = *reinterpret_cast<const float4*>(ap + l2 * LDP), a1 = *reinterpret_cast<const float4*>(ap + l2 * LDP + HALF);
        const float4 b0 = *reinterpret_cast<const float4*>(bp + l2 * LDP), b1 = *reinterpret_cast<const float4*>(bp + l2 * LDP + HALF);
        const float a8[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 aa = make_float2(a8[i], a8[i]);
          float2* row2 = reinterpret_cast<float2*>(acc[i]);
#pragma unroll
          for (int j = 0; j < 4; ++j) row2[j] = __ffma2_rn(aa, bb[j], row2[j]);   // packed fp32 FMA, each lane rounded like fmaf
        }
      }
    }
    __syncthreads();   // A is free again
  }
  // ---- sum of the SLICES partial tiles: binary tree through shared memory (everything staged is consumed by now)
  float* scr = smem + 4;
#pragma unroll
  for (int half = SLICES / 2; half >= 1; half >>= 1) {
    if (slice >= half && slice < 2 * half) {
      float* dst = scr + (size_t)(slice - half) * 64 * TL + tile;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[(i * 8 + j) * TL] = acc[i][j];
    }
    __syncthreads();
    if (slice < half) {
      const float* src = scr + (size_t)slice * 64 * TL + tile;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] += src[(i * 8 + j) * TL];
    }
    __syncthreads();
  }
  if (slice == 0 && bi >= 0) {
    float* out = D.scpartial + (size_t)lb * LD * LD + (size_t)(8 * bi) * LD + 8 * bj;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      *reinterpret_cast<float4*>(out + (size_t)i * LD) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(out + (size_t)i * LD + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
  }
}

// Fold of the Schur chunk partials into fp64 G[host], chunks in index order.  A separate, wide launch:
// folding inside the Schur kernel by "the last block of a host" put ~0.5 MB of dependent reads per host
// on ONE block at the very end of the kernel (measured: 165 us of a 190 us launch for 8 windows).
// grid: per window n * fold_sc_blocks(LD) blocks (+ the top folds).
// Blocks behind the Schur part fold the top accumulators (fold_top_slot): two (mode, pair) slots per block, the modes
// of bit mask `top_modes` (1 = A, 2 = L, 4 = marginalisation) in bit order.  with_sc = 0: top folds only.
__host__ __device__ inline int fold_sc_blocks(int LD) { return (LD * LD / 4 + 255) / 256; }   // per host: one thread per four elements
__host__ __device__ inline int fold_top_blocks(int n, int top_modes) {
  const int nm = (top_modes & 1) + ((top_modes >> 1) & 1) + ((top_modes >> 2) & 1);
  return (nm * n * n + 1) / 2;
}
__global__ void __launch_bounds__(256) ba_fold_sc_kernel(const BADesc* __restrict__ descs,
                                                         const int* __restrict__ prefix, int nwin, int top_modes, int with_sc) {
  int lb;
  const int win = find_window(prefix, nwin, blockIdx.x, lb);
  const BADesc& D = descs[win];
  const int LD = D.LD, per = fold_sc_blocks(LD);
  const int nsc = with_sc ? D.W.n * per : 0;
  if (lb >= nsc) {
    const int nn = D.W.n * D.W.n;
    int idx = (lb - nsc) * 2 + (threadIdx.x >> 7);
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      if (!((top_modes >> m) & 1)) continue;
      if (idx < nn) { fold_top_slot(D, m * nn + idx, threadIdx.x & 127); return; }
      idx -= nn;
    }
    return;
  }
  // four neighbouring elements of a row per thread (LD is a multiple of 8: a float4 never leaves its row or its 4x4 tile)
  const int host = lb / per;
  const int i = 4 * ((lb % per) * 256 + threadIdx.x);
  if (i >= LD * LD) return;
  const int row = i / LD, col = i % LD;
  if ((col >> 2) < (row >> 2)) return;   // lower tiles are not produced
  const int cb = D.host_chunk_begin[host], ce = D.host_chunk_begin[host + 1];
  const float* src = D.scpartial + (size_t)cb * LD * LD + i;
  const size_t stride = (size_t)LD * LD;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int c = cb;
  for (; c + 4 <= ce; c += 4) {   // four independent 16-byte loads in flight, summed in chunk order
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)j * stride));
#pragma unroll
    for (int j = 0; j < 4; ++j) { s0 += (double)v[j].x; s1 += (double)v[j].y; s2 += (double)v[j].z; s3 += (double)v[j].w; }
    src += 4 * stride;
  }
  for (; c < ce; ++c) {
    const float4 v = __ldcg(reinterpret_cast<const float4*>(src));
    s0 += (double)v.x; s1 += (double)v.y; s2 += (double)v.z; s3 += (double)v.w;
    src += stride;
  }
  // G_h is symmetric; the Schur kernel produces its upper tiles only: store the mirror too, the stitch reads a dense matrix
  double* G = D.G + (size_t)host * LD * LD;
  reinterpret_cast<double2*>(G + i)[0] = make_double2(s0, s1);
  reinterpret_cast<double2*>(G + i)[1] = make_double2(s2, s3);
  if ((col >> 2) > (row >> 2)) {
    G[(size_t)col * LD + row] = s0; G[(size_t)(col + 1) * LD + row] = s1;
    G[(size_t)(col + 2) * LD + row] = s2; G[(size_t)(col + 3) * LD + row] = s3;
  }
}

// ------------------------------------------------------------------------------------------------
// fp64 tensor-core step D(8x8) += A(8x4) * B(4x8) of one warp (mma.sync.m8n8k4.f64, DMMA): lane (g = lane>>2, t = lane&3)
// holds A[g][t], B[t][g] and the accumulator pair D[g][2t], D[g][2t+1].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// (o0, o1) += A1 * M * A2^T for one warp, as four DMMAs: T = A1 * M (two k-steps), T from accumulator layout to A-operand
// layout with two quad shuffles, then o += T * A2^T (two k-steps).  A1, A2: 8x8 column-major (the adjoints); M(row, col) =
// Mb[row * ldm + col] (dense 13x13 top accumulator or dense Gram matrix).  6 loads per lane, one memory round trip.
__device__ __forceinline__ void sandwich_dmma(const double* __restrict__ A1, const double* __restrict__ A2,
                                              const double* __restrict__ Mb, int ldm, int lane, double& o0, double& o1) {
  const int g = lane >> 2, t = lane & 3;
  const double a1lo = A1[g + 8 * t], a1hi = A1[g + 8 * (t + 4)];
  const double mlo = Mb[t * ldm + g], mhi = Mb[(t + 4) * ldm + g];
  const double a2lo = A2[g + 8 * t], a2hi = A2[g + 8 * (t + 4)];
  double T0 = 0.0, T1 = 0.0;
  dmma884(T0, T1, a1lo, mlo);
  dmma884(T0, T1, a1hi, mhi);
  // T[g][t] lives in lane (quad | t>>1), slot t&1; T[g][t+4] two lanes further
  const int src = (lane & ~3) | (t >> 1);
  const double s0 = __shfl_sync(0xffffffffu, T0, src), s1 = __shfl_sync(0xffffffffu, T1, src);
  const double u0 = __shfl_sync(0xffffffffu, T0, src + 2), u1 = __shfl_sync(0xffffffffu, T1, src + 2);
  const double tlo = (t & 1) ? s1 : s0, thi = (t & 1) ? u1 : u0;
  dmma884(o0, o1, tlo, a2lo);
  dmma884(o0, o1, thi, a2hi);
}

// K7: stitch — relative accumulators -> (8n+4)^2 system in fp64, gather form.  One block per output 8x8 block (a,b) of one
// system; index n is the calibration block.  The Ad * M * Ad^T sandwiches that land in a block are dealt round-robin to the
// block's warps (four fp64 tensor-core steps each, accumulated in the warp's DMMA accumulators), then the warps' partial
// blocks are summed in warp order.  The 8x8x8 fp64 products are the one true contraction on the path (and exact fp64 FMAs).
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) ba_stitch_kernel(const BADesc* __restrict__ descs,
                                                               const int* __restrict__ prefix, int nwin,
                                                               int sys_begin, int sys_end, int with_energy) {
  static_assert(WARPS == 4 || WARPS == 8, "a warp owns one or two of the eight accumulator slots");
  int lb;
  const int win = find_window(prefix, nwin, blockIdx.x, lb);
  const StitchArgs& S = descs[win].S;
  const int n = S.n, N = 8 * n + 4, LD = S.LD;
  const int nb1 = n + 1;
  const int sys = lb / (nb1 * nb1), ab = lb % (nb1 * nb1);
  const int a = ab % nb1, b = ab / nb1;
  if (sys < sys_begin || sys >= sys_end) return;   // systems (0=A, 1=L, 2=SC) this launch produces
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ double part[ST_SLOTS][64];
  double* H = S.Hb + (size_t)sys * ((size_t)N * N + N);
  double* bv = H + (size_t)N * N;
  const int nn = n * n;
  auto AdH = [&](int h, int t) { return S.adHost + 64 * (size_t)(h + t * n); };
  auto AdT = [&](int h, int t) { return S.adTarget + 64 * (size_t)(h + t * n); };
  double o0 = 0.0, o1 = 0.0;   // slot `warp`
  double p0 = 0.0, p1 = 0.0;   // slot `warp + 4` (four-warp blocks)
  int ci = 0;   // contribution counter, identical in every warp
#define CONTRIB(A1, A2, MB, LDM)                                                                          \
  do {                                                                                                    \
    const int sl_ = ci & (ST_SLOTS - 1);                                                                  \
    if (sl_ == warp) sandwich_dmma((A1), (A2), (MB), (LDM), lane, o0, o1);                                \
    else if (WARPS == 4 && sl_ == warp + 4) sandwich_dmma((A1), (A2), (MB), (LDM), lane, p0, p1);         \
    ci++;                                                                                                 \
  } while (0)
  // the slots' accumulator pairs (row g, columns 2t, 2t+1) -> part[slot][row + 8 col], summed in slot order by 64 threads
  auto reduce_block = [&](auto&& store) {
    const int g = lane >> 2, t = lane & 3;
    part[warp][g + 8 * (2 * t)] = o0; part[warp][g + 8 * (2 * t + 1)] = o1;
    if (WARPS == 4) { part[warp + 4][g + 8 * (2 * t)] = p0; part[warp + 4][g + 8 * (2 * t + 1)] = p1; }
    __syncthreads();
    if (threadIdx.x < 64) {
      double o = 0.0;
#pragma unroll
      for (int w = 0; w < ST_SLOTS; ++w) o += part[w][threadIdx.x];
      store(threadIdx.x & 7, threadIdx.x >> 3, o);
    }
  };

  if (sys < 2) {
    const double* acc = S.accTop + (size_t)(sys == 0 ? 0 : S.L_set) * nn * TOP_DENSE;
    auto ACC = [&](int h, int t) { return acc + (size_t)(h + t * n) * TOP_DENSE; };
    const bool usePrior = (sys == 1) && S.add_priors;
    const bool live = (sys == 0) || S.has_L;   // without linearised residuals the L accumulators are zero
    constexpr int FB = 4 * 13 + 4;              // the 8x8 frame block of the dense 13x13 starts at (4,4)
    if (a < n && b < n) {
      if (a > b) return;                        // written by the mirror block (AccumulatedTopHessian.h:105-116)
      if (live) {
        if (a == b) {
          // S[a,a] = sum_t AdH[a,t] M[a,t] AdH[a,t]^T + sum_h AdT[h,a] M[h,a] AdT[h,a]^T + AdH[a,a] M[a,a] AdT[a,a]^T
          for (int t = 0; t < n; ++t) {
            const double* m = ACC(a, t);
            if (m[169] == 0.0) continue;   // acc.num == 0 (AccumulatedTopHessian.cpp:243)
            CONTRIB(AdH(a, t), AdH(a, t), m + FB, 13);
          }
          for (int h = 0; h < n; ++h) {
            const double* m = ACC(h, a);
            if (m[169] == 0.0) continue;
            CONTRIB(AdT(h, a), AdT(h, a), m + FB, 13);
          }
          const double* m = ACC(a, a);
          if (m[169] != 0.0) CONTRIB(AdH(a, a), AdT(a, a), m + FB, 13);
        } else {
          // H[a,b] = S[a,b] + S[b,a]^T  (M is symmetric)
          const double* m1 = ACC(a, b);
          if (m1[169] != 0.0) CONTRIB(AdH(a, b), AdT(a, b), m1 + FB, 13);
          const double* m2 = ACC(b, a);
          if (m2[169] != 0.0) CONTRIB(AdT(b, a), AdH(b, a), m2 + FB, 13);
        }
      }
      reduce_block([&](int rr, int cc, double o) {
        if (a == b) {
          if (usePrior && rr == cc) o += S.framePrior[8 * a + rr];
          H[(size_t)(4 + 8 * a + rr) + (size_t)N * (4 + 8 * a + cc)] = o;
        } else {
          H[(size_t)(4 + 8 * a + rr) + (size_t)N * (4 + 8 * b + cc)] = o;
          H[(size_t)(4 + 8 * b + cc) + (size_t)N * (4 + 8 * a + rr)] = o;
        }
      });
    } else if (a < n && b == n) {
      // H[a, 0:4] (8x4), its transpose, and b[a]: mat-vec products, one thread per output element
      if (threadIdx.x < 40) {
        const int r = threadIdx.x & 7, c = threadIdx.x >> 3;
        const int col = c < 4 ? c : 12;
        double o = 0.0;
        if (live) {
          for (int t = 0; t < n; ++t) {
            const double* m = ACC(a, t);
            if (m[169] == 0.0) continue;
            const double* A1 = AdH(a, t);
#pragma unroll
            for (int x = 0; x < 8; ++x) o += A1[r + 8 * x] * m[(4 + x) * 13 + col];
          }
          for (int h = 0; h < n; ++h) {
            const double* m = ACC(h, a);
            if (m[169] == 0.0) continue;
            const double* A1 = AdT(h, a);
#pragma unroll
            for (int x = 0; x < 8; ++x) o += A1[r + 8 * x] * m[(4 + x) * 13 + col];
          }
        }
        if (c < 4) {
          H[(size_t)(4 + 8 * a + r) + (size_t)N * c] = o;
          H[(size_t)c + (size_t)N * (4 + 8 * a + r)] = o;
        } else {
          if (usePrior) o += S.framePrior[8 * a + r] * S.frameDeltaPrior[8 * a + r];
          bv[4 + 8 * a + r] = o;
        }
      }
    } else if (a == n && b == n) {
      if (threadIdx.x < 20) {
        const int r = threadIdx.x & 3, c = threadIdx.x >> 2;
        const int col = c < 4 ? c : 12;
        double o = 0.0;
        if (live)
          for (int k = 0; k < nn; ++k) {
            const double* m = acc + (size_t)k * TOP_DENSE;
            if (m[169] == 0.0) continue;
            o += m[r * 13 + col];
          }
        if (c < 4) {
          if (usePrior && r == c) o += S.cPrior[r];
          H[(size_t)r + (size_t)N * c] = o;
        } else {
          if (usePrior) o += S.cPrior[r] * (double)S.cDeltaF[r];
          bv[r] = o;
        }
      } else if (threadIdx.x == 32) {
        int tot = 0;
        if (live)
          for (int k = 0; k < nn; ++k) tot += (int)acc[(size_t)k * TOP_DENSE + 169];
        S.nres[sys] = tot;   // resInA / resInL (EnergyFunctional.cpp:163,186)
        S.Hb[3 * ((size_t)N * N + N) + 1 + sys] = (double)tot;   // ... and behind the systems, for the sharded window's all-reduce
      }
      if (sys == 0 && with_energy) energy_sum_block(descs[win], &part[0][0]);   // uniform per block
    }
    return;
  }

  // ---- Schur complement system: H_sc = sum_i T_i G_i T_i^T with T_i[a,j] = d(a,i) AdH[i,j] + d(a,j) AdT[i,j]
  // (G_i = dense symmetric Gram matrix of host i, ba_fold_sc_kernel)
  auto Gh = [&](int i) { return S.G + (size_t)i * LD * LD; };
  if (a < n && b < n) {
    // H_sc is symmetric (every G_i is): block (b,a) is the transpose of block (a,b) and is written by the mirror store
    if (a > b) return;
    if (a == b) {   // sum_{j,k} AdH[a,j] D_a[j,k] AdH[a,k]^T   (AccumulatedSCHessian.cpp:108)
      const double* G = Gh(a);
      for (int j = 0; j < n; ++j)
        for (int k = 0; k < n; ++k)
          CONTRIB(AdH(a, j), AdH(a, k), G + (size_t)(8 * j) * LD + 8 * k, LD);
    }
    {   // :111  H[i,k] += AdH[ij] D AdT[ik]^T with i=a, k=b
      const double* G = Gh(a);
      for (int j = 0; j < n; ++j) CONTRIB(AdH(a, j), AdT(a, b), G + (size_t)(8 * j) * LD + 8 * b, LD);
    }
    {   // :110  H[j,i] += AdT[ij] D AdH[ik]^T with j=a, i=b
      const double* G = Gh(b);
      for (int k = 0; k < n; ++k) CONTRIB(AdT(b, a), AdH(b, k), G + (size_t)(8 * a) * LD + 8 * k, LD);
    }
    for (int i = 0; i < n; ++i) {   // :109  H[j,k] += AdT[ij] D AdT[ik]^T with j=a, k=b
      const double* G = Gh(i);
      CONTRIB(AdT(i, a), AdT(i, b), G + (size_t)(8 * a) * LD + 8 * b, LD);
    }
    reduce_block([&](int rr, int cc, double o) {
      H[(size_t)(4 + 8 * a + rr) + (size_t)N * (4 + 8 * b + cc)] = o;
      if (a != b) H[(size_t)(4 + 8 * b + cc) + (size_t)N * (4 + 8 * a + rr)] = o;
    });
  } else if (a < n && b == n) {
    if (threadIdx.x < 40) {   // :86-89: H[i,0:4] += AdH[ij] E[ij], H[j,0:4] += AdT[ij] E[ij]; b likewise with EB
      const int r = threadIdx.x & 7, c = threadIdx.x >> 3;
      double o = 0.0;
      const int col = 8 * n + c;   // Hcd columns 8n..8n+3, bdSum column 8n+4
      {
        const double* G = Gh(a);
        for (int j = 0; j < n; ++j) {
          const double* A1 = AdH(a, j);
#pragma unroll
          for (int x = 0; x < 8; ++x) o += A1[r + 8 * x] * G[(size_t)(8 * j + x) * LD + col];
        }
      }
      for (int i = 0; i < n; ++i) {
        const double* G = Gh(i);
        const double* A1 = AdT(i, a);
#pragma unroll
        for (int x = 0; x < 8; ++x) o += A1[r + 8 * x] * G[(size_t)(8 * a + x) * LD + col];
      }
      if (c < 4) {
        H[(size_t)(4 + 8 * a + r) + (size_t)N * c] = o;
        H[(size_t)c + (size_t)N * (4 + 8 * a + r)] = o;   // .h:106-110
      } else {
        bv[4 + 8 * a + r] = o;
      }
    }
  } else if (a == n && b == n) {
    if (threadIdx.x < 20) {
      const int r = threadIdx.x & 3, c = threadIdx.x >> 2;
      double o = 0.0;
      for (int i = 0; i < n; ++i) o += Gh(i)[(size_t)(8 * n + r) * LD + 8 * n + c];
      if (c < 4) H[(size_t)r + (size_t)N * c] = o; else bv[r] = o;
    }
  }
#undef CONTRIB
}

// Layout conversion for the host API: component-major sorted SoA [comps][stride] -> caller order
// [count][comps] (idx = sorted2orig), so the D2H copy is contiguous and the host only memcpy's.
template <class T>
__global__ void ba_export_kernel(const T* __restrict__ src, int comps, int stride, int count,
                                 const int* __restrict__ sorted2orig, T* __restrict__ dst) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  size_t o = (size_t)sorted2orig[k] * comps;
  for (int c = 0; c < comps; ++c) dst[o + c] = src[(size_t)c * stride + k];
}
template <class T>
__global__ void ba_export_aos_kernel(const T* __restrict__ src, int comps, int count,
                                     const int* __restrict__ sorted2orig, T* __restrict__ dst) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  size_t o = (size_t)sorted2orig[k] * comps;
  for (int c = 0; c < comps; ++c) dst[o + c] = src[(size_t)k * comps + c];
}
// JpJdF records (point-major slots) -> caller order [R][8]
__global__ void ba_export_jpjd_kernel(BARes r, const int* __restrict__ sorted2orig, float* __restrict__ dst) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= r.R) return;
  const float4* src = reinterpret_cast<const float4*>(r.JpJdF + r.slotpos[k]);
  float4* d = reinterpret_cast<float4*>(dst + 8 * (size_t)sorted2orig[k]);
  d[0] = src[0]; d[1] = src[1];
}
// Re-grouping of the sorted residual order (a residual became linearised: it moves behind the non-linearised ones of its
// pair): dst[c][k] = src[c][perm[k]] for component-major arrays (grid.y = component), and the same for records.
template <class T>
__global__ void ba_permute_kernel(const T* __restrict__ src, T* __restrict__ dst, size_t stride, int count, const int* __restrict__ perm) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const size_t o = (size_t)blockIdx.y * stride;
  dst[o + k] = src[o + perm[k]];
}
__global__ void ba_permute_rec32_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int count, const int* __restrict__ perm) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const int q = perm[k];
  dst[2 * (size_t)k] = src[2 * (size_t)q]; dst[2 * (size_t)k + 1] = src[2 * (size_t)q + 1];
}
// and the other way for per-point uploads: caller order [count] -> sorted order
template <class T>
__global__ void ba_import_kernel(const T* __restrict__ src, int count, const int* __restrict__ sorted2orig,
                                 T* __restrict__ dst, int dst_stride_elems, int dst_offset) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  dst[(size_t)k * dst_stride_elems + dst_offset] = src[sorted2orig[k]];
}

// xAd[n*h+t] = x_h^T adHostF[h,t] + x_t^T adTargetF[h,t] in float, index order (EnergyFunctional.cpp:228-238);
// one block per window.  The adjoints are the fp64 tables of the window cast to float element by element,
// exactly like adHostF/adTargetF (EnergyFunctional.cpp:73-74).
__global__ void __launch_bounds__(256) ba_xad_kernel(const BADesc* __restrict__ descs) {
  const BADesc& D = descs[blockIdx.x];
  const int n = D.W.n, N = 8 * n + 4;
  __shared__ float xF[8 * HSLAM_MAX_FRAMES + 4];
  for (int i = threadIdx.x; i < N; i += blockDim.x) xF[i] = (float)D.x[i];
  __syncthreads();
  if (threadIdx.x < 4) D.xc[threadIdx.x] = xF[threadIdx.x];
  for (int idx = threadIdx.x; idx < n * n * 8; idx += blockDim.x) {
    const int c = idx & 7, ht = idx >> 3, h = ht / n, t = ht % n;
    const double* aH = D.S.adHost + 64 * (size_t)(h + n * t);
    const double* aT = D.S.adTarget + 64 * (size_t)(h + n * t);
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      a = xadd(a, xmul(xF[4 + 8 * h + r], (float)aH[r + 8 * c]));
      b = xadd(b, xmul(xF[4 + 8 * t + r], (float)aT[r + 8 * c]));
    }
    D.xAd[8 * (size_t)(n * h + t) + c] = xadd(a, b);
  }
}

// EnergyFunctional::resubstituteFPt (EnergyFunctional.cpp:249-274).  With `io` the step is also written
// in CALLER order into the batch's D2H staging block.
__global__ void __launch_bounds__(128) ba_resub_kernel(const BADesc* __restrict__ descs, const int* __restrict__ prefix,
                                                       int nwin, const BAIO* __restrict__ io, char* __restrict__ dn, int block0) {
  int lb;
  const int win = find_window(prefix, nwin, blockIdx.x + block0, lb);   // block0: a launch may cover a sub-range of the batch
  const BADesc& D = descs[win];
  const BARes& r = D.r;
  const BAPts& pts = D.p;
  const int p = lb * 128 + threadIdx.x;
  if (p >= pts.P) return;
  const int s = r.rs, n = D.W.n;
  const float* xc = D.xc;
  const float* xAd = D.xAd;
  int ngood = 0;
  const int rb = pts.res_begin[p], re = pts.res_begin[p + 1];
  for (int i = rb; i < re; ++i) if (r.isActive[pts.res_index[i]]) ngood++;
  float step = 0.f;
  if (ngood != 0) {
    float b = pts.bdSumF[p];
    float4 ha = pts.Hcd[0][p], hl = pts.Hcd[1][p];
    float dot = 0.f;
    dot = xadd(dot, xmul(xc[0], xadd(ha.x, hl.x))); dot = xadd(dot, xmul(xc[1], xadd(ha.y, hl.y)));
    dot = xadd(dot, xmul(xc[2], xadd(ha.z, hl.z))); dot = xadd(dot, xmul(xc[3], xadd(ha.w, hl.w)));
    b = xsub(b, dot);
    for (int i = rb; i < re; ++i) {
      int k = pts.res_index[i];
      if (!r.isActive[k]) continue;
      int pair = r.pair[k], t = (pair * D.W.n_rcp) >> 16, h = pair - t * n;
      const float* xa = xAd + 8 * (n * h + t);
      float d = 0.f;
      const float4* jp = reinterpret_cast<const float4*>(r.JpJdF + (size_t)p * r.pstr + 8 * (i - rb));   // slot order = the point's own order
      const float4 j0 = jp[0], j1 = jp[1];
      const float jv[8] = {j0.x, j0.y, j0.z, j0.w, j1.x, j1.y, j1.z, j1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) d = xadd(d, xmul(xa[j], jv[j]));
      b = xsub(b, d);
    }
    step = xmul(-b, pts.HdiF[p]);
  }
  pts.step[p] = step;
  if (io && io[win].dn_step >= 0) reinterpret_cast<float*>(dn + io[win].dn_step)[D.p_s2o[p]] = step;
}

// ---- batched host API: ONE staging block up, ONE down (hslam_ba_batch_step_host) -----------------------
// Import: window tables are plain copies into the window's own buffers; per-point arrays arrive in caller
// order and are scattered into the host-sorted point records.  grid = (blocks per window, windows).
template <class T>
__device__ __forceinline__ void staged_copy(T* dst, const char* up, long long off, int count, int tid, int stride) {
  if (off < 0) return;
  const T* src = reinterpret_cast<const T*>(up + off);
  for (int i = tid; i < count; i += stride) dst[i] = src[i];
}
__global__ void __launch_bounds__(256) ba_batch_import_kernel(const BADesc* __restrict__ descs, const BAIO* __restrict__ io,
                                                              const char* __restrict__ up) {
  const BADesc& D = descs[blockIdx.y];
  const BAIO o = io[blockIdx.y];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  const int n = D.W.n, nn = n * n;
  staged_copy(const_cast<float*>(D.W.precalc), up, o.up_precalc, nn * HSLAM_PRECALC_STRIDE, tid, stride);
  staged_copy(const_cast<float*>(D.W.adHTdeltaF), up, o.up_adHTd, nn * 8, tid, stride);
  staged_copy(const_cast<double*>(D.S.adHost), up, o.up_adH, nn * 64, tid, stride);
  staged_copy(const_cast<double*>(D.S.adTarget), up, o.up_adT, nn * 64, tid, stride);
  staged_copy(const_cast<double*>(D.S.framePrior), up, o.up_fprior, n * 8, tid, stride);
  staged_copy(const_cast<double*>(D.S.frameDeltaPrior), up, o.up_fdprior, n * 8, tid, stride);
  staged_copy(D.x, up, o.up_x, 8 * n + 4, tid, stride);
  const int P = D.p.P;
  float* uvid = reinterpret_cast<float*>(const_cast<float4*>(D.p.uvid));
  float* deltaF = const_cast<float*>(D.p.deltaF);
  const float* s_id = o.up_idepth >= 0 ? reinterpret_cast<const float*>(up + o.up_idepth) : nullptr;
  const float* s_iz = o.up_idz >= 0 ? reinterpret_cast<const float*>(up + o.up_idz) : nullptr;
  const float* s_dF = o.up_deltaF >= 0 ? reinterpret_cast<const float*>(up + o.up_deltaF) : nullptr;
  const bool derive = o.pad != 0 && s_id && !s_dF;   // EnergyFunctional::setDeltaF's point part (EnergyFunctional.cpp:148) on the device
  if (s_id || s_iz || s_dF)
    for (int p = tid; p < P; p += stride) {
      const int q = D.p_s2o[p];
      if (s_id) uvid[4 * (size_t)p + 2] = s_id[q];
      if (s_iz) uvid[4 * (size_t)p + 3] = s_iz[q];
      if (s_dF) deltaF[p] = s_dF[q];
      else if (derive) deltaF[p] = xsub(s_id[q], s_iz ? s_iz[q] : uvid[4 * (size_t)p + 3]);   // one exactly-rounded subtraction, as on the host
    }
}
// Export: per-residual results back to caller order, the stitched systems, counts and energy, all into
// the window's slice of the D2H staging block.
// `what`: bit 0 = per-residual arrays (valid as soon as the linearise kernel is done: exported and copied on a
// side stream while the accumulate kernels run), bit 1 = systems / counts / energy (after the stitch).
__global__ void __launch_bounds__(256) ba_batch_export_kernel(const BADesc* __restrict__ descs, const BAIO* __restrict__ io,
                                                              char* __restrict__ dn, int what) {
  const BADesc& D = descs[blockIdx.y];
  const BAIO o = io[blockIdx.y];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  const int R = D.r.R;
  float* oE = o.dn_newEnergy >= 0 ? reinterpret_cast<float*>(dn + o.dn_newEnergy) : nullptr;
  float* oW = o.dn_newEnergyWO >= 0 ? reinterpret_cast<float*>(dn + o.dn_newEnergyWO) : nullptr;
  int8_t* oS = o.dn_newState >= 0 ? reinterpret_cast<int8_t*>(dn + o.dn_newState) : nullptr;
  if ((what & 1) && (oE || oW || oS))
    for (int k = tid; k < R; k += stride) {
      const int q = D.r_s2o[k];
      if (oE) oE[q] = D.r.newEnergy[k];
      if (oW) oW[q] = D.r.newEnergyWO[k];
      if (oS) oS[q] = D.r.newState[k];
    }
  // what & 2: the three systems + energy + counts; & 4: the A and L systems + energy + counts (final before the Schur
  // fold has run); & 8: the Schur-complement system alone
  if (!(what & (2 | 4 | 8))) return;
  if (o.dn_Hb >= 0) {
    const int N = 8 * D.W.n + 4, per = N * N + N;
    const int lo = (what & (2 | 4)) ? 0 : 2 * per, hi = (what & (2 | 8)) ? 3 * per : 2 * per;
    double* dst = reinterpret_cast<double*>(dn + o.dn_Hb);
    for (int i = lo + tid; i < hi; i += stride) dst[i] = D.S.Hb[i];
  }
  if (!(what & (2 | 4))) return;
  if (tid < 2 && o.dn_nres >= 0) reinterpret_cast<int*>(dn + o.dn_nres)[tid] = D.S.nres[tid];
  if (tid == 0 && o.dn_energy >= 0) *reinterpret_cast<double*>(dn + o.dn_energy) = *D.energy;
}

// marginalizePointsF: priorF *= setting_idepthFixPriorMargFac for the flagged points (EnergyFunctional.cpp:564)
__global__ void ba_scale_prior_kernel(BAPts pts, const unsigned char* __restrict__ marg, float scale) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < pts.P && marg[p]) const_cast<float*>(pts.priorF)[p] = xmul(pts.priorF[p], scale);
}

// EnergyFunctional::calcLEnergyPt (EnergyFunctional.cpp:289-347): per point the energy of its linearised, active
// residuals at the current delta, E = sum (2 res_toZeroF + J delta) * J delta, plus deltaF^2 priorF.  One thread per
// point; per-block fp64 partials (summed in block order by the caller).
__global__ void __launch_bounds__(128) ba_lenergy_kernel(const BADesc* __restrict__ descs, double* __restrict__ partial) {
  const BADesc& D = descs[0];
  const BARes& r = D.r;
  const BAPts& pts = D.p;
  const int p = blockIdx.x * 128 + threadIdx.x;
  const int s = r.rs, n = D.W.n;
  float E = 0.f;
  if (p < pts.P) {
    const float dd = pts.deltaF[p];
    for (int i = pts.res_begin[p]; i < pts.res_begin[p + 1]; ++i) {
      const int k = pts.res_index[i];
      if (!r.isLin[k] || !r.isActive[k]) continue;
      const float* J = r.J + k;
      const float* dp = D.W.adHTdeltaF + 8 * r.pair[k];
      float dx = 0.f, dy = 0.f, cxs = 0.f, cys = 0.f;
#pragma unroll
      for (int j = 0; j < 6; ++j) { dx = xadd(dx, xmul(J[(J_XI0 + j) * s], dp[j])); dy = xadd(dy, xmul(J[(J_XI1 + j) * s], dp[j])); }
#pragma unroll
      for (int j = 0; j < 4; ++j) { cxs = xadd(cxs, xmul(J[(J_C0 + j) * s], D.W.cDeltaF[j])); cys = xadd(cys, xmul(J[(J_C1 + j) * s], D.W.cDeltaF[j])); }
      const float jx = xadd(xadd(dx, cxs), xmul(J[J_DD * s], dd)), jy = xadd(xadd(dy, cys), xmul(J[(J_DD + 1) * s], dd));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float Jd = xmul(J[(J_IDX0 + j) * s], jx);
        Jd = xadd(Jd, xmul(J[(J_IDX1 + j) * s], jy));
        Jd = xadd(Jd, xmul(J[(J_AB0 + j) * s], dp[6]));
        Jd = xadd(Jd, xmul(J[(J_AB1 + j) * s], dp[7]));
        float r0 = r.rtz[j * s + k];
        r0 = xadd(xadd(r0, r0), Jd);
        E = xadd(E, xmul(Jd, r0));
      }
    }
    E = xadd(E, xmul(xmul(dd, dd), pts.priorF[p]));
    (void)n;
  }
  __shared__ double red[4];
  double e = warp_sum((double)E);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e;
  __syncthreads();
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0] + red[1] + red[2] + red[3];
}

}  // namespace hs
