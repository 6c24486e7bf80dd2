// tracker.cu — frame-to-keyframe direct image alignment on the device.
//
// Replaces CoarseTracker::calcRes / calcGSSSE / trackNewestCoarse
// (reference Src/CoarseTracker.cpp:329-485, 267-324, 506-683) and the SSE Accumulator9
// (Include/MatrixAccumulators.h:982-1345).
//
// Design (B200): one thread per reference point; the point record {u,v,idepth,color} is ONE float4
// load; the four bilinear texels are four float4 loads of the {I,dx,dy,.} image; residual, Huber
// weight and the 9-column Jacobian stay in registers and the 45 upper-triangular J^T W J products
// are reduced warp-shuffle -> shared-memory -> per-block partial, so the reference's eight
// buf_warped_* arrays are never written.  The per-block partials are folded in a fixed order in
// fp64, which makes H,b bit-identical from run to run.
//
// hslam_tracker_track runs the WHOLE coarse-to-fine Levenberg-Marquardt loop in one cooperative
// launch: after every evaluation a grid barrier is followed by every block redundantly folding the
// partials and doing the 8x8 LDLT solve + SE3 exp + accept/reject on one thread, so there is no
// host round trip between the ~50 evaluations of a frame (SURVEY.md §7 "Latency, not bandwidth").
#include "common.cuh"
#include <cooperative_groups.h>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <algorithm>

namespace cg = cooperative_groups;

namespace hs {

constexpr int TRK_THREADS = 256;
constexpr int TRK_CL_THREADS = 512;   // threads per CTA of the cluster kernel
constexpr int TRK_NACC = 45;      // upper triangle of the 9x9
constexpr int TRK_PART = 56;      // floats per block partial: 45 H, E, flowT, flowRT, flowNum, nE, nSat, nW (+pad)
constexpr int P_E = 45, P_FT = 46, P_FRT = 47, P_FN = 48, P_NE = 49, P_NSAT = 50, P_NW = 51;

struct TrkLevelDev {
  int w, h, n;
  float fx, fy, cx, cy;
  float Ki[9];
  const float4* pc;  // {u, v, idepth, color}
};

// Everything one evaluation needs, prepared once per (pose, aff) by ONE thread.
struct EvalConst {
  float RKi[9], t[3], Ki[9];
  float fx, fy, cx, cy;
  float affLL0, affLL1;     // (float) fromToVecExposure(...)
  float b0;                 // (float) lastRef_aff_g2l.b
  float cutoff, maxEnergy, huber;
  int w, h, n, lvl;
  const float4* pc;
  const float4* img;
};

struct EvalResult {          // folded, fp64
  double acc[TRK_NACC];
  double E, flowT, flowRT, flowNum;
  int nE, nSat, nW, pad;
};

__device__ __forceinline__ void make_eval_const(EvalConst& c, const TrkLevelDev& L, int lvl, const float4* img,
                                                const double* T /*R rm 9 + t 3*/, double aff_a, double aff_b,
                                                float ref_exposure, float new_exposure, double ref_aff_a,
                                                double ref_aff_b, float cutoff, float huber) {
  float R[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = (float)T[i];
  // Mat33f RKi = R.cast<float>() * Ki[lvl]   (CoarseTracker.cpp:345), coefficient order a0+(a1+a2)
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c.RKi[3 * i + j] = xdot3(R[3 * i], R[3 * i + 1], R[3 * i + 2], L.Ki[j], L.Ki[3 + j], L.Ki[6 + j]);
#pragma unroll
  for (int i = 0; i < 3; ++i) c.t[i] = (float)T[9 + i];
#pragma unroll
  for (int i = 0; i < 9; ++i) c.Ki[i] = L.Ki[i];
  c.fx = L.fx; c.fy = L.fy; c.cx = L.cx; c.cy = L.cy;
  // AffLight::fromToVecExposure (GlobalTypes.h:334-346) in fp64, then cast (CoarseTracker.cpp:347)
  float eF = ref_exposure, eT = new_exposure;
  if (eF == 0.f || eT == 0.f) { eF = eT = 1.f; }
  double a = exp(aff_a - ref_aff_a) * (double)eT / (double)eF;
  double b = aff_b - a * ref_aff_b;
  c.affLL0 = (float)a; c.affLL1 = (float)b;
  c.b0 = (float)ref_aff_b;
  c.cutoff = cutoff; c.huber = huber;
  c.maxEnergy = xsub(xmul(xmul(2.0f, huber), cutoff), xmul(huber, huber));  // :354
  c.w = L.w; c.h = L.h; c.n = L.n; c.lvl = lvl; c.pc = L.pc; c.img = img;
}

// One pass over the reference cloud of a level: every thread accumulates its points, the block
// reduces, thread t<TRK_PART writes partial[blockIdx][t].  smem: float[TRK_THREADS/32][TRK_PART].
// `partial` points at THIS block's TRK_PART floats (global memory, or shared memory in the cluster kernel).
template <int THREADS>
__device__ __forceinline__ void eval_block(const EvalConst& c, float* __restrict__ partial, float* red,
                                           int block, int nblocks) {
  float acc[TRK_NACC];
#pragma unroll
  for (int i = 0; i < TRK_NACC; ++i) acc[i] = 0.f;
  float E = 0.f, flowT = 0.f, flowRT = 0.f, flowNum = 0.f;
  int nE = 0, nSat = 0, nW = 0;
  const float wm3 = (float)(c.w - 3), hm3 = (float)(c.h - 3);

#ifndef TRK_UNROLL
#define TRK_UNROLL 1
#endif
  constexpr int kUnroll = TRK_UNROLL;
#pragma unroll kUnroll
  for (int i = block * THREADS + threadIdx.x; i < c.n; i += nblocks * THREADS) {
    float4 p = __ldg(c.pc + i);
    float x = p.x, y = p.y, id = p.z, refColor = p.w;
    // pt = RKi*(x,y,1) + t*id ; CoarseTracker.cpp:377-382
    float pt0 = xadd(xdot3(c.RKi[0], c.RKi[1], c.RKi[2], x, y, 1.0f), xmul(c.t[0], id));
    float pt1 = xadd(xdot3(c.RKi[3], c.RKi[4], c.RKi[5], x, y, 1.0f), xmul(c.t[1], id));
    float pt2 = xadd(xdot3(c.RKi[6], c.RKi[7], c.RKi[8], x, y, 1.0f), xmul(c.t[2], id));
    float u = xdiv(pt0, pt2), v = xdiv(pt1, pt2);
    float Ku = xadd(xmul(c.fx, u), c.cx), Kv = xadd(xmul(c.fy, v), c.cy);
    float new_idepth = xdiv(id, pt2);

    if (c.lvl == 0 && (i & 31) == 0) {  // flow indicators, :384-415
      float k0 = xdot3(c.Ki[0], c.Ki[1], c.Ki[2], x, y, 1.0f);
      float k1 = xdot3(c.Ki[3], c.Ki[4], c.Ki[5], x, y, 1.0f);
      float k2 = xdot3(c.Ki[6], c.Ki[7], c.Ki[8], x, y, 1.0f);
      float t0 = xmul(c.t[0], id), t1 = xmul(c.t[1], id), t2 = xmul(c.t[2], id);
      float q2 = xadd(k2, t2);
      float KuT = xadd(xmul(c.fx, xdiv(xadd(k0, t0), q2)), c.cx), KvT = xadd(xmul(c.fy, xdiv(xadd(k1, t1), q2)), c.cy);
      q2 = xsub(k2, t2);
      float KuT2 = xadd(xmul(c.fx, xdiv(xsub(k0, t0), q2)), c.cx), KvT2 = xadd(xmul(c.fy, xdiv(xsub(k1, t1), q2)), c.cy);
      float r0 = xdot3(c.RKi[0], c.RKi[1], c.RKi[2], x, y, 1.0f);
      float r1 = xdot3(c.RKi[3], c.RKi[4], c.RKi[5], x, y, 1.0f);
      float r2 = xdot3(c.RKi[6], c.RKi[7], c.RKi[8], x, y, 1.0f);
      q2 = xsub(r2, t2);
      float Ku3 = xadd(xmul(c.fx, xdiv(xsub(r0, t0), q2)), c.cx), Kv3 = xadd(xmul(c.fy, xdiv(xsub(r1, t1), q2)), c.cy);
      float a, b;
      a = xsub(KuT, x); b = xsub(KvT, y);   flowT = xadd(flowT, xadd(xmul(a, a), xmul(b, b)));
      a = xsub(KuT2, x); b = xsub(KvT2, y); flowT = xadd(flowT, xadd(xmul(a, a), xmul(b, b)));
      a = xsub(Ku, x); b = xsub(Kv, y);     flowRT = xadd(flowRT, xadd(xmul(a, a), xmul(b, b)));
      a = xsub(Ku3, x); b = xsub(Kv3, y);   flowRT = xadd(flowRT, xadd(xmul(a, a), xmul(b, b)));
      flowNum += 2.f;
    }

    if (!(Ku > 2.f && Kv > 2.f && Ku < wm3 && Kv < hm3 && new_idepth > 0.f)) continue;  // :417
    Hit3 hit = interp33(c.img, Ku, Kv, c.w);
    if (!isfinite(hit.I)) continue;
    float residual = xsub(hit.I, xadd(xmul(c.affLL0, refColor), c.affLL1));
    float ar = fabsf(residual);
    float hw = ar < c.huber ? 1.0f : xdiv(c.huber, ar);
    nE++;
    if (ar > c.cutoff) {
      E = xadd(E, c.maxEnergy);
      nSat++;
      continue;
    }
    E = xadd(E, xmul(xmul(xmul(hw, residual), residual), xsub(2.0f, hw)));
    nW++;
    // Jacobian row, CoarseTracker.cpp:284-305 (contraction allowed from here on: tolerance-checked)
    float dx = hit.dx * c.fx, dy = hit.dy * c.fy;
    float J[9];
    J[0] = new_idepth * dx;
    J[1] = new_idepth * dy;
    J[2] = -new_idepth * (u * dx + v * dy);
    J[3] = -(u * v * dx + dy * (1.f + v * v));
    J[4] = u * v * dy + dx * (1.f + u * u);
    J[5] = u * dy - v * dx;
    J[6] = c.affLL0 * (c.b0 - refColor);
    J[7] = -1.f;
    J[8] = residual;
    int k = 0;
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      float Jw = J[r] * hw;
#pragma unroll
      for (int cc = r; cc < 9; ++cc, ++k) acc[k] = fmaf(Jw, J[cc], acc[k]);
    }
  }

  // Block reduction of the 52 sums.  Two shuffle levels take a warp's 32 lanes to 8; the 8 x (THREADS/32) values of
  // every sum go to shared memory (row stride odd: conflict-free both ways); four threads per sum add a quarter of the
  // row each in index order and are combined by two more shuffle levels — a fixed association, the same on every
  // run.  (Five shuffle levels for every sum in every warp were 31 % of the kernel's instructions.)
  constexpr int RW = THREADS / 4 + 1;          // row stride; THREADS/4 = 8 lanes x THREADS/32 warps entries per sum
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  auto put = [&](int i, float v) {
    v += __shfl_down_sync(0xffffffffu, v, 16);
    v += __shfl_down_sync(0xffffffffu, v, 8);
    if (lane < 8) red[i * RW + wid * 8 + lane] = v;
  };
  auto puti = [&](int i, int v) {
    v += __shfl_down_sync(0xffffffffu, v, 16);
    v += __shfl_down_sync(0xffffffffu, v, 8);
    if (lane < 8) red[i * RW + wid * 8 + lane] = __int_as_float(v);
  };
#pragma unroll
  for (int i = 0; i < TRK_NACC; ++i) put(i, acc[i]);
  put(P_E, E); put(P_FT, flowT); put(P_FRT, flowRT); put(P_FN, flowNum);
  puti(P_NE, nE); puti(P_NSAT, nSat); puti(P_NW, nW);
  __syncthreads();
  if (threadIdx.x < 4 * TRK_PART) {            // 224 threads = 7 whole warps (full-mask shuffles); rows 52..55 are padding
    const int t = threadIdx.x >> 2, part = threadIdx.x & 3;
    constexpr int Q = THREADS / 16;            // entries per quarter row
    const float* row = red + t * RW + part * Q;
    if (t >= P_NE) {
      int q = 0;
#pragma unroll 8
      for (int j = 0; j < Q; ++j) q += __float_as_int(row[j]);
      q += __shfl_down_sync(0xffffffffu, q, 2, 4);
      q += __shfl_down_sync(0xffffffffu, q, 1, 4);
      if (part == 0 && t <= P_NW) partial[t] = __int_as_float(q);
    } else {
      float v = 0.f;
#pragma unroll 8
      for (int j = 0; j < Q; ++j) v += row[j];
      v += __shfl_down_sync(0xffffffffu, v, 2, 4);
      v += __shfl_down_sync(0xffffffffu, v, 1, 4);
      if (part == 0) partial[t] = v;
    }
    __threadfence();  // partials visible device-wide before the block signals (ticket / grid barrier)
  }
  __syncthreads();
}

// Deterministic fold of nblocks partials into fp64; executed by one whole block.
// smem: double fold[4][TRK_PART].
// `get(b, t)` returns element t of block b's partial (global memory or a peer CTA's shared memory).
template <class GET>
__device__ __forceinline__ void fold_partials(GET get, int nblocks, EvalResult& out, double (*fold)[TRK_PART]) {
  int t = threadIdx.x & 63, q = threadIdx.x >> 6;  // 4 interleaved groups of 64 threads
  if (t < TRK_PART && q < 4) {
    if (t >= P_NE && t <= P_NW) {
      long long s = 0;
      for (int b = q; b < nblocks; b += 4) s += __float_as_int(get(b, t));
      fold[q][t] = (double)s;
    } else {
      double s = 0.0;
      for (int b = q; b < nblocks; b += 4) s += (double)get(b, t);
      fold[q][t] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x < TRK_PART) {
    int i = threadIdx.x;
    double s = (fold[0][i] + fold[1][i]) + (fold[2][i] + fold[3][i]);
    if (i < TRK_NACC) out.acc[i] = s;
    else if (i == P_E) out.E = s;
    else if (i == P_FT) out.flowT = s;
    else if (i == P_FRT) out.flowRT = s;
    else if (i == P_FN) out.flowNum = s;
    else if (i == P_NE) out.nE = (int)s;
    else if (i == P_NSAT) out.nSat = (int)s;
    else if (i == P_NW) out.nW = (int)s;
  }
  __syncthreads();
}

// ---------------------------------------------------------------- single evaluation (calcRes / calcGS)
struct EvalArgs {
  TrkLevelDev L;
  int lvl;
  const float4* img;
  double T[12];
  double aff_a, aff_b, ref_aff_a, ref_aff_b;
  float ref_exposure, new_exposure, cutoff, huber;
};

__global__ void __launch_bounds__(TRK_THREADS) trk_eval_kernel(EvalArgs a, float* __restrict__ partial,
                                                               unsigned int* __restrict__ ticket,
                                                               EvalResult* __restrict__ result) {
  __shared__ EvalConst c;
  __shared__ float red[TRK_PART * (TRK_THREADS / 4 + 1)];
  __shared__ double fold[4][TRK_PART];
  __shared__ EvalResult res;
  __shared__ bool last;
  if (threadIdx.x == 0)
    make_eval_const(c, a.L, a.lvl, a.img, a.T, a.aff_a, a.aff_b, a.ref_exposure, a.new_exposure, a.ref_aff_a,
                    a.ref_aff_b, a.cutoff, a.huber);
  __syncthreads();
  eval_block<TRK_THREADS>(c, partial + (size_t)blockIdx.x * TRK_PART, red, blockIdx.x, gridDim.x);
  // last block folds (threadfence + ticket); fold order is by block index, not arrival order
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    fold_partials([&](int b, int t) { return __ldcg(partial + (size_t)b * TRK_PART + t); }, gridDim.x, res, fold);
    if (threadIdx.x == 0) {
      *result = res;
      *ticket = 0;
    }
  }
}

// ---------------------------------------------------------------- fp64 helpers for the LM loop
__device__ void d_ldlt_solve(const double* Ain, const double* rhs, int n, double* x) {
  // symmetric diagonal pivoting, same scheme as the oracle (stands in for Eigen::LDLT)
  double A[64];
  int perm[8];
  for (int i = 0; i < n * n; ++i) A[i] = Ain[i];
  for (int i = 0; i < n; ++i) perm[i] = i;
#define AT(r, c) A[(r) + (c) * n]
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = fabs(AT(k, k));
    for (int i = k + 1; i < n; ++i)
      if (fabs(AT(i, i)) > best) { best = fabs(AT(i, i)); p = i; }
    if (p != k) {
      for (int c = 0; c < n; ++c) { double t = AT(k, c); AT(k, c) = AT(p, c); AT(p, c) = t; }
      for (int r = 0; r < n; ++r) { double t = AT(r, k); AT(r, k) = AT(r, p); AT(r, p) = t; }
      int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
    }
    double d = AT(k, k);
    if (d == 0.0) continue;
    for (int i = k + 1; i < n; ++i) AT(i, k) /= d;
    for (int j = k + 1; j < n; ++j)
      for (int i = j; i < n; ++i) {
        AT(i, j) -= AT(i, k) * d * AT(j, k);
        AT(j, i) = AT(i, j);
      }
  }
  double y[8];
  for (int i = 0; i < n; ++i) y[i] = rhs[perm[i]];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) y[i] -= AT(i, j) * y[j];
  for (int i = 0; i < n; ++i) y[i] = (AT(i, i) != 0.0) ? y[i] / AT(i, i) : 0.0;
  for (int i = n - 1; i >= 0; --i)
    for (int j = i + 1; j < n; ++j) y[i] -= AT(j, i) * y[j];
  for (int i = 0; i < n; ++i) x[perm[i]] = y[i];
#undef AT
}

// Unpivoted LDL^T of the damped (SPD) system in shared memory; lanes 0..n-1 each own a row.
// A is overwritten by L (strict lower) and D (diagonal); y holds rhs on entry and x on exit.
__device__ __forceinline__ void warp_ldlt_solve(double (*A)[9], double* y, int n, int lane) {
  for (int k = 0; k < n; ++k) {
    double d = A[k][k];
    double lik = 0.0;
    const bool mine = lane > k && lane < n;
    if (mine) lik = A[lane][k] / d;
    __syncwarp();
    if (mine) A[lane][k] = lik;
    __syncwarp();
    if (mine)
      for (int j = k + 1; j <= lane; ++j) A[lane][j] -= lik * d * A[j][k];
    __syncwarp();
  }
  for (int i = 0; i < n; ++i) {          // L z = rhs
    double zi = y[i];
    __syncwarp();
    if (lane > i && lane < n) y[lane] -= A[lane][i] * zi;
    __syncwarp();
  }
  if (lane < n) y[lane] = y[lane] / A[lane][lane];   // D
  __syncwarp();
  for (int i = n - 1; i >= 0; --i) {     // L^T x = z
    double xi = y[i];
    __syncwarp();
    if (lane < i) y[lane] -= A[i][lane] * xi;
    __syncwarp();
  }
}

// The same factorisation and solves with each row in REGISTERS and warp shuffles instead of shared memory and
// __syncwarp between dependent steps (the shared-memory version cost ~7000 cycles per solve: eight divisions, each
// followed by a serial loop of dependent shared-memory read-modify-writes).  Same operations on the same values in
// the same order per element.  All 32 lanes call; lanes >= n only take part in the shuffles.
__device__ __forceinline__ void warp_ldlt_solve_reg(double (*A)[9], double* y, int n, int lane) {
  const int row = lane & 7;
  double a[8], t[8], yy = y[row];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = A[row][j]; t[j] = 0.0; }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < n) {   // uniform
      const double d = __shfl_sync(0xffffffffu, a[k], k);
      const bool mine = lane > k && lane < n;
      double lik = 0.0;
      if (mine) { lik = a[k] / d; a[k] = lik; }
#pragma unroll
      for (int j = k + 1; j < 8; ++j) {
        const double ljk = __shfl_sync(0xffffffffu, a[k], j);   // L[j][k]
        if (mine && j <= lane) a[j] -= lik * d * ljk;
        if (lane == k) t[j] = ljk;                              // column k of L, for the transposed solve
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {          // L z = rhs
    if (i < n) {
      const double zi = __shfl_sync(0xffffffffu, yy, i);
      if (lane > i && lane < n) yy -= a[i] * zi;
    }
  }
  double diag = a[0];
#pragma unroll
  for (int j = 1; j < 8; ++j) if (row == j) diag = a[j];
  if (lane < n) yy = yy / diag;          // D
#pragma unroll
  for (int i = 7; i >= 0; --i) {         // L^T x = z
    if (i < n) {
      const double xi = __shfl_sync(0xffffffffu, yy, i);
      if (lane < i) yy -= t[i] * xi;
    }
  }
  if (lane < n) y[lane] = yy;
  __syncwarp();
}

__device__ void d_matmul3(const double* A, const double* B, double* C) {
  double t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  for (int i = 0; i < 9; ++i) C[i] = t[i];
}

// Sophus SE3::exp (se3.hpp:407-427 / so3.hpp:343-369), T = [R row-major | t]
__device__ void d_se3_exp(const double* xi, double* T) {
  const double* ups = xi;
  const double* om = xi + 3;
  double theta_sq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  double theta = sqrt(theta_sq), half = 0.5 * theta, imag, real;
  const double eps = 1e-10;
  if (theta < eps) {
    double t4 = theta_sq * theta_sq;
    imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * t4;
    real = 1.0 - 0.5 * theta_sq + (1.0 / 384.0) * t4;
  } else {
    imag = sin(half) / theta;
    real = cos(half);
  }
  double qw = real, qx = imag * om[0], qy = imag * om[1], qz = imag * om[2];
  double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  double twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx;
  double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  T[0] = 1 - (tyy + tzz); T[1] = txy - twz; T[2] = txz + twy;
  T[3] = txy + twz; T[4] = 1 - (txx + tzz); T[5] = tyz - twx;
  T[6] = txz - twy; T[7] = tyz + twx; T[8] = 1 - (txx + tyy);
  double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0}, O2[9], V[9];
  d_matmul3(O, O, O2);
  if (theta < eps) {
    for (int i = 0; i < 9; ++i) V[i] = T[i];
  } else {
    double a = (1 - cos(theta)) / theta_sq, b = (theta - sin(theta)) / (theta_sq * theta);
    for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0 ? 1.0 : 0.0) + a * O[i] + b * O2[i];
  }
  for (int i = 0; i < 3; ++i) T[9 + i] = V[3 * i] * ups[0] + V[3 * i + 1] * ups[1] + V[3 * i + 2] * ups[2];
}

__device__ void d_se3_mul(const double* A, const double* B, double* C) {
  double R[9], t[3];
  d_matmul3(A, B, R);
  for (int i = 0; i < 3; ++i) t[i] = A[3 * i] * B[9] + A[3 * i + 1] * B[10] + A[3 * i + 2] * B[11] + A[9 + i];
  for (int i = 0; i < 9; ++i) C[i] = R[i];
  for (int i = 0; i < 3; ++i) C[9 + i] = t[i];
}

// H (8x8 col-major), b from the folded accumulators; CoarseTracker.cpp:309-323
__host__ __device__ inline void make_Hb(const double* acc, int nW, const float* scales8, double* H, double* b) {
  int n = (nW + 3) & ~3;
  double inv = (double)(1.0f / (float)n);
  int k = 0;
  double full[9][9];
  for (int r = 0; r < 9; ++r)
    for (int c = r; c < 9; ++c, ++k) full[r][c] = full[c][r] = acc[k];
  for (int r = 0; r < 8; ++r) {
    for (int c = 0; c < 8; ++c) H[r + 8 * c] = full[r][c] * inv;
    b[r] = full[r][8] * inv;
  }
  for (int c = 0; c < 8; ++c)
    for (int r = 0; r < 8; ++r) H[r + 8 * c] *= (double)scales8[c];
  for (int r = 0; r < 8; ++r)
    for (int c = 0; c < 8; ++c) H[r + 8 * c] *= (double)scales8[r];
  for (int r = 0; r < 8; ++r) b[r] *= (double)scales8[r];
}

// ---------------------------------------------------------------- the fused LM kernel
struct TrackArgs {
  TrkLevelDev L[5];
  const float4* img[5];
  int coarsestLvl;
  int maxIterations[5];
  float scales8[8];   // {rot,rot,rot,trans,trans,trans,a,b} as applied in CoarseTracker.cpp:312-323,597-600
  float coarseCutoffTH, huber, affineOptModeA, affineOptModeB;
  float ref_exposure, new_exposure;
  double ref_aff_a, ref_aff_b;
};

struct TrackIO {           // pinned-host visible struct copied in/out
  double T[12];
  double aff[2];
  double minResForAbort[5];
  double lastResiduals[5];
  double lastFlow[3];
  int ok;
  int iterations[5];
  int evals;
  float new_exposure;          // of THIS member's frame
  const float4* img[5];        // pyramid levels of THIS member's frame (a batch tracks several frames / hypotheses at once)
  long long phase_cycles[4];   // diagnostics (block 0): solve+exp+constants | evaluation | barrier+fold | accept+make_Hb
};

struct LMState {
  double cur[12], cand[12];
  double aff_cur[2], aff_cand[2];
  double H[64], b[8];
  double resOldE, resOldN;       // resOld[0], resOld[1]
  double flowOld[3];
  float lambda, levelCutoffRepeat;
  int lvl, iteration, phase, haveRepeated, done, ok, evals;
  double incnorm;
};

__device__ __forceinline__ double sat_ratio(const EvalResult& r) { return (double)((float)r.nSat / (float)r.nE); }

// One evaluation of the LM loop: every participating block evaluates its share, all blocks meet, every block folds all
// partials in block order (identical result in every block).
template <int THREADS, bool CLUSTER>
__device__ __forceinline__ void track_eval_and_fold(const EvalConst& c, float* gpart, float* spart, float* red,
                                                    double (*fold)[TRK_PART], EvalResult& res, int nb, long long* tk, int brank) {
  const long long t0 = clock64();
  if (CLUSTER) {
    cg::cluster_group cluster = cg::this_cluster();
    if (brank < nb) eval_block<THREADS>(c, spart, red, brank, nb);
    if (threadIdx.x == 0) tk[1] += clock64() - t0;
    cluster.sync();   // release/acquire across the cluster: the shared-memory partials are visible
    fold_partials([&](int b, int t) { return cluster.map_shared_rank(spart, b)[t]; }, nb, res, fold);
  } else {
    if (brank < nb) eval_block<THREADS>(c, gpart + (size_t)brank * TRK_PART, red, brank, nb);
    if (threadIdx.x == 0) tk[1] += clock64() - t0;
    __threadfence();
    cg::this_grid().sync();
    fold_partials([&](int b, int t) { return __ldcg(gpart + (size_t)b * TRK_PART + t); }, nb, res, fold);
  }
}

// CLUSTER = false: cooperative launch over up to 2 blocks per SM, grid barrier per evaluation, partials in global memory.
// CLUSTER = true : ONE thread-block cluster (8 or 16 CTAs) runs the whole frame: the barrier is the hardware cluster
//                  barrier, the per-CTA partials stay in shared memory and are folded through distributed shared
//                  memory, nothing of an evaluation touches global memory except the gathers; the rest of the GPU
//                  stays free for other trackers / hypotheses.
template <int THREADS, bool CLUSTER>
__global__ void __launch_bounds__(THREADS) trk_track_kernel(TrackArgs a, float* __restrict__ partial0,
                                                            float* __restrict__ partial1,
                                                            TrackIO* __restrict__ io) {
  __shared__ EvalConst c;
  __shared__ float red[TRK_PART * (THREADS / 4 + 1)];
  __shared__ float mypart[2][TRK_PART];   // CLUSTER: this CTA's partial of the current / previous evaluation
  __shared__ double fold[4][TRK_PART];
  __shared__ EvalResult res;
  __shared__ LMState s;
  __shared__ int iters[5];
  __shared__ double lastRes[5];
  __shared__ double sA[8][9];
  __shared__ double sy[8];
  __shared__ long long tk[4];   // [0] solve, [1] evaluation (inside [2]), [2] evaluation + barrier + fold, [3] accept
  if (threadIdx.x < 4) tk[threadIdx.x] = 0;
  // CLUSTER: the grid is a BATCH of clusters, cluster m tracks member m (its own frame, initial pose and result slot);
  // the members share nothing, so each is bit-identical to a launch of its own.
  int brank = blockIdx.x, nblk = gridDim.x;
  if (CLUSTER) {
    cg::cluster_group cl = cg::this_cluster();
    brank = (int)cl.block_rank(); nblk = (int)cl.num_blocks();
    io += blockIdx.x / nblk;
  }

  // Every block carries an identical copy of the LM state and takes identical decisions.
  if (threadIdx.x == 0) {
    for (int i = 0; i < 12; ++i) s.cur[i] = io->T[i];
    s.aff_cur[0] = io->aff[0]; s.aff_cur[1] = io->aff[1];
    s.haveRepeated = 0; s.done = 0; s.ok = 0; s.evals = 0;
    for (int i = 0; i < 5; ++i) { iters[i] = 0; lastRes[i] = nan(""); }
    s.flowOld[0] = s.flowOld[1] = s.flowOld[2] = 1000.0;
  }
  __syncthreads();
  int buf = 0;

  for (int lvl = a.coarsestLvl; lvl >= 0; lvl--) {
    const TrkLevelDev& L = a.L[lvl];
    int nb = (L.n + THREADS - 1) / THREADS;
    if (nb > nblk) nb = nblk;
    if (nb < 1) nb = 1;
    // --- initial evaluation with cutoff doubling, CoarseTracker.cpp:530-541
    if (threadIdx.x == 0) { s.levelCutoffRepeat = 1.f; }
    __syncthreads();
    while (true) {
      if (threadIdx.x == 0)
        make_eval_const(c, L, lvl, io->img[lvl], s.cur, s.aff_cur[0], s.aff_cur[1], a.ref_exposure, io->new_exposure,
                        a.ref_aff_a, a.ref_aff_b, xmul(a.coarseCutoffTH, s.levelCutoffRepeat), a.huber);
      __syncthreads();
      track_eval_and_fold<THREADS, CLUSTER>(c, buf ? partial1 : partial0, mypart[buf], red, fold, res, nb, tk, brank);
      buf ^= 1;
      bool again = (sat_ratio(res) > 0.6) && (s.levelCutoffRepeat < 50.f);
      __syncthreads();
      if (threadIdx.x == 0) { s.evals++; if (again) s.levelCutoffRepeat *= 2.f; }
      __syncthreads();
      if (!again) break;
    }
    if (threadIdx.x == 0) {
      make_Hb(res.acc, res.nW, a.scales8, s.H, s.b);
      s.resOldE = res.E; s.resOldN = (double)res.nE;
      s.flowOld[0] = (double)((float)res.flowT / ((float)res.flowNum + 0.1));
      s.flowOld[1] = 0.0;
      s.flowOld[2] = (double)((float)res.flowRT / ((float)res.flowNum + 0.1));
      s.lambda = 0.01f;
    }
    __syncthreads();

    // --- LM iterations, CoarseTracker.cpp:559-646
    for (int iteration = 0; iteration < a.maxIterations[lvl]; iteration++) {
      // 8x8 damped solve by the first warp: lanes 0..7 own one row each of a shared-memory LDLT
      // (the serial thread-0 version spent ~10 us per iteration in local-memory traffic).
      long long tk0 = clock64();
      if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const bool fixA = a.affineOptModeA < 0.f, fixB = a.affineOptModeB < 0.f;
        int nsolve = 8;
        if (fixA && fixB) nsolve = 6; else if (fixA != fixB) nsolve = 7;
        if (lane < 8) {
          const double damp = (double)(1.f + s.lambda);
          for (int cc = 0; cc < 8; ++cc) {
            int sr = lane, sc = cc;
            if (fixA && !fixB) { if (sr == 6) sr = 7; if (sc == 6) sc = 7; }   // :575-587: row/col 6 := row/col 7
            double v = s.H[sr + 8 * sc];
            if (sr == sc) v *= damp;
            sA[lane][cc] = v;
          }
          int sb = (fixA && !fixB && lane == 6) ? 7 : lane;
          sy[lane] = -s.b[sb];
        }
        __syncwarp();
#ifndef TRK_LDLT_REG
#define TRK_LDLT_REG 1
#endif
#if TRK_LDLT_REG
        warp_ldlt_solve_reg(sA, sy, nsolve, lane);
#else
        warp_ldlt_solve(sA, sy, nsolve, lane);
#endif
      }
      if (threadIdx.x == 0) {
        iters[lvl]++;
        double inc[8];
        const bool fixA = a.affineOptModeA < 0.f, fixB = a.affineOptModeB < 0.f;
        for (int i = 0; i < 8; ++i) inc[i] = sy[i];
        if (fixA && fixB) { inc[6] = inc[7] = 0.0; }          // :565-569
        else if (!fixA && fixB) { inc[7] = 0.0; }             // :570-574
        else if (fixA && !fixB) { inc[7] = inc[6]; inc[6] = 0.0; }   // :583-586
        float extrapFac = 1.f;
        if (s.lambda < 0.001f) extrapFac = sqrtf(sqrtf(0.001f / s.lambda));
        double nrm = 0.0, sum = 0.0, incS[8];
        for (int i = 0; i < 8; ++i) { inc[i] *= (double)extrapFac; nrm += inc[i] * inc[i]; }
        for (int i = 0; i < 8; ++i) { incS[i] = inc[i] * (double)a.scales8[i]; sum += incS[i]; }
        if (!isfinite(sum)) for (int i = 0; i < 8; ++i) incS[i] = 0.0;
        double Texp[12];
        d_se3_exp(incS, Texp);
        d_se3_mul(Texp, s.cur, s.cand);
        s.aff_cand[0] = s.aff_cur[0] + incS[6];
        s.aff_cand[1] = s.aff_cur[1] + incS[7];
        s.incnorm = sqrt(nrm);
        make_eval_const(c, L, lvl, io->img[lvl], s.cand, s.aff_cand[0], s.aff_cand[1], a.ref_exposure, io->new_exposure,
                        a.ref_aff_a, a.ref_aff_b, xmul(a.coarseCutoffTH, s.levelCutoffRepeat), a.huber);
      }
      __syncthreads();
      long long tk1 = clock64();
      track_eval_and_fold<THREADS, CLUSTER>(c, buf ? partial1 : partial0, mypart[buf], red, fold, res, nb, tk, brank);
      long long tk3 = clock64();
      buf ^= 1;
      if (threadIdx.x == 0) {
        s.evals++;
        bool accept = (res.E / (double)res.nE) < (s.resOldE / s.resOldN);  // :611
        if (accept) {
          make_Hb(res.acc, res.nW, a.scales8, s.H, s.b);
          s.resOldE = res.E; s.resOldN = (double)res.nE;
          s.flowOld[0] = (double)((float)res.flowT / ((float)res.flowNum + 0.1));
          s.flowOld[2] = (double)((float)res.flowRT / ((float)res.flowNum + 0.1));
          for (int i = 0; i < 12; ++i) s.cur[i] = s.cand[i];
          s.aff_cur[0] = s.aff_cand[0]; s.aff_cur[1] = s.aff_cand[1];
          s.lambda *= 0.5f;
        } else {
          s.lambda *= 4.f;
          if (s.lambda < 0.001f) s.lambda = 0.001f;
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) { tk[0] += tk1 - tk0; tk[2] += tk3 - tk1; tk[3] += clock64() - tk3; }
      if (!(s.incnorm > 1e-3)) break;   // uniform: s is identical in every block
    }

    // --- level epilogue, :648-659
    bool abort_now = false;
    if (threadIdx.x == 0) lastRes[lvl] = (double)sqrtf((float)(s.resOldE / s.resOldN));
    __syncthreads();
    abort_now = lastRes[lvl] > 1.5 * io->minResForAbort[lvl];
    if (abort_now) { if (threadIdx.x == 0) { s.done = 1; s.ok = 0; } __syncthreads(); break; }
    if (s.levelCutoffRepeat > 1.f && !s.haveRepeated) {
      __syncthreads();
      if (threadIdx.x == 0) s.haveRepeated = 1;
      lvl++;
      __syncthreads();
    }
  }

  if (brank == 0 && threadIdx.x == 0) {
    int ok = 0;
    if (!s.done) {
      ok = 1;
      double aa = s.aff_cur[0], ab = s.aff_cur[1];
      if ((a.affineOptModeA != 0.f && fabsf((float)aa) > 1.2f) || (a.affineOptModeB != 0.f && fabsf((float)ab) > 200.f))
        ok = 0;
      float eF = a.ref_exposure, eT = io->new_exposure;
      if (eF == 0.f || eT == 0.f) { eF = eT = 1.f; }
      double ra = exp(aa - a.ref_aff_a) * (double)eT / (double)eF;
      double rb = ab - ra * a.ref_aff_b;
      if ((a.affineOptModeA == 0.f && fabsf(logf((float)ra)) > 1.5f) || (a.affineOptModeB == 0.f && fabsf((float)rb) > 200.f))
        ok = 0;
      for (int i = 0; i < 12; ++i) io->T[i] = s.cur[i];   // lastToNew_out is set before the checks (:663-664)
      io->aff[0] = aa; io->aff[1] = ab;
      if (ok) {
        if (a.affineOptModeA < 0.f) io->aff[0] = 0.0;
        if (a.affineOptModeB < 0.f) io->aff[1] = 0.0;
      }
    }
    io->ok = ok;
    for (int i = 0; i < 5; ++i) { io->lastResiduals[i] = lastRes[i]; io->iterations[i] = iters[i]; }
    for (int i = 0; i < 3; ++i) io->lastFlow[i] = s.flowOld[i];
    io->evals = s.evals;
    for (int i = 0; i < 4; ++i) io->phase_cycles[i] = tk[i];
  }
  if (CLUSTER) cg::this_cluster().sync();   // no CTA leaves while a peer may still read its partial
}

}  // namespace hs

using namespace hs;

struct hslam_tracker {
  hslam_ctx* ctx = nullptr;
  int w = 0, h = 0, levels = 0;
  TrkLevelDev L[HSLAM_MAX_PYR_LEVELS];
  float4* pc[HSLAM_MAX_PYR_LEVELS] = {nullptr};
  int pc_cap[HSLAM_MAX_PYR_LEVELS] = {0};
  float ref_exposure = 1.f;
  double ref_aff_a = 0, ref_aff_b = 0;
  float* partial[2] = {nullptr, nullptr};
  int partial_blocks = 0;
  unsigned int* ticket = nullptr;
  EvalResult* d_result = nullptr;
  EvalResult* h_result = nullptr;   // pinned
  TrackIO* d_io = nullptr;          // [io_cap]: one slot per member of a batched launch
  TrackIO* h_io = nullptr;          // pinned
  int io_cap = 0, n_inflight = 0;
  float4* h_pack = nullptr;         // pinned staging for set_ref
  size_t h_pack_cap = 0;
  int coop_blocks = 0;
  long long last_phase_cycles[4] = {0, 0, 0, 0};
  int last_evals = 0;
  int cluster_ctas = 0;             // 0: cooperative grid kernel; 8 / 16: one thread-block cluster runs the frame
  char* cd_buf = nullptr;           // makeCoarseDepthL0 working set (seeds, pixel lists, maps, scan)
  size_t cd_cap = 0;
};

static int ensure_partials(hslam_tracker* t, int blocks) {
  if (blocks <= t->partial_blocks) return 0;
  for (int i = 0; i < 2; ++i) {
    if (t->partial[i]) cudaFree(t->partial[i]);
    HS_CUDA(cudaMalloc(&t->partial[i], (size_t)blocks * TRK_PART * sizeof(float)));
  }
  t->partial_blocks = blocks;
  return 0;
}

extern "C" {

int hslam_tracker_create(hslam_ctx* ctx, int w, int h, hslam_tracker** out) {
  HS_ARG(ctx && out && w > 0 && h > 0, "hslam_tracker_create: bad argument");
  HS_CUDA(cudaSetDevice(ctx->device));
  hslam_tracker* t = new hslam_tracker();
  t->ctx = ctx; t->w = w; t->h = h;
  memset(t->L, 0, sizeof(t->L));
  HS_CUDA(cudaMalloc(&t->ticket, sizeof(unsigned int)));
  HS_CUDA(cudaMemsetAsync(t->ticket, 0, sizeof(unsigned int), ctx->stream));
  HS_CUDA(cudaMalloc(&t->d_result, sizeof(EvalResult)));
  HS_CUDA(cudaMallocHost(&t->h_result, sizeof(EvalResult)));
  HS_CUDA(cudaMalloc(&t->d_io, sizeof(TrackIO)));
  HS_CUDA(cudaMallocHost(&t->h_io, sizeof(TrackIO)));
  t->io_cap = 1;
  int per_sm = 0;
  HS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trk_track_kernel<TRK_THREADS, false>, TRK_THREADS, 0));
  HS_ARG(per_sm >= 1, "hslam_tracker_create: track kernel does not fit on an SM");
  t->coop_blocks = ctx->sm_count * (per_sm > 2 ? 2 : per_sm);
  int rc = ensure_partials(t, t->coop_blocks > 1024 ? t->coop_blocks : 1024);
  if (rc) return rc;
  // cluster mode: the largest cluster (16 CTAs, non-portable size, else 8) that the device can co-schedule
  t->cluster_ctas = 0;
  const char* mode = getenv("HSLAM_TRACKER_MODE");
  if (!(mode && strcmp(mode, "coop") == 0)) {
    cudaFuncSetAttribute(trk_track_kernel<TRK_CL_THREADS, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int c : {16, 8}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(c); cfg.blockDim = dim3(TRK_CL_THREADS);
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, trk_track_kernel<TRK_CL_THREADS, true>, &cfg) == cudaSuccess && nclusters >= 1) {
        t->cluster_ctas = c;
        break;
      }
    }
    cudaGetLastError();   // a failed probe is not an error of the handle
  }
  *out = t;
  return 0;
}

int hslam_tracker_set_mode(hslam_tracker* t, int cluster_ctas) {
  HS_ARG(t && (cluster_ctas == 0 || cluster_ctas == 8 || cluster_ctas == 16), "hslam_tracker_set_mode: cluster_ctas must be 0, 8 or 16");
  t->cluster_ctas = cluster_ctas;
  return 0;
}
int hslam_tracker_last_profile(hslam_tracker* t, long long phase_cycles[4], int* evals) {
  HS_ARG(t && phase_cycles && evals, "hslam_tracker_last_profile: null");
  memcpy(phase_cycles, t->last_phase_cycles, sizeof(t->last_phase_cycles));
  *evals = t->last_evals;
  return 0;
}
int hslam_tracker_get_mode(hslam_tracker* t, int* cluster_ctas) {
  HS_ARG(t && cluster_ctas, "hslam_tracker_get_mode: null");
  *cluster_ctas = t->cluster_ctas;
  return 0;
}

int hslam_tracker_destroy(hslam_tracker* t) {
  if (!t) return 0;
  cudaSetDevice(t->ctx->device);
  cudaStreamSynchronize(t->ctx->stream);
  for (int l = 0; l < HSLAM_MAX_PYR_LEVELS; ++l) cudaFree(t->pc[l]);
  cudaFree(t->partial[0]); cudaFree(t->partial[1]);
  cudaFree(t->ticket); cudaFree(t->d_result); cudaFreeHost(t->h_result);
  cudaFree(t->d_io); cudaFreeHost(t->h_io);
  if (t->h_pack) cudaFreeHost(t->h_pack);
  if (t->cd_buf) cudaFree(t->cd_buf);
  delete t;
  return 0;
}

static inline float cof3(const float* m, int i, int j) {
  int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  volatile float a = m[3 * i1 + j1] * m[3 * i2 + j2];
  volatile float b = m[3 * i1 + j2] * m[3 * i2 + j1];
  return a - b;
}

// CoarseTracker::makeK (Src/CoarseTracker.cpp:72-101); Ki by Eigen's 3x3 cofactor inverse.
int hslam_tracker_make_K(hslam_tracker* t, int levels, int w0, int h0, float fx0, float fy0, float cx0, float cy0) {
  HS_ARG(t && levels >= 1 && levels <= HSLAM_MAX_PYR_LEVELS, "hslam_tracker_make_K: bad levels");
  t->levels = levels;
  t->L[0].w = w0; t->L[0].h = h0; t->L[0].fx = fx0; t->L[0].fy = fy0; t->L[0].cx = cx0; t->L[0].cy = cy0;
  for (int l = 1; l < levels; ++l) {
    t->L[l].w = w0 >> l; t->L[l].h = h0 >> l;
    t->L[l].fx = (float)(t->L[l - 1].fx * 0.5);
    t->L[l].fy = (float)(t->L[l - 1].fy * 0.5);
    t->L[l].cx = (float)((cx0 + 0.5) / ((int)1 << l) - 0.5);
    t->L[l].cy = (float)((cy0 + 0.5) / ((int)1 << l) - 0.5);
  }
  for (int l = 0; l < levels; ++l) {
    float K[9] = {t->L[l].fx, 0.f, t->L[l].cx, 0.f, t->L[l].fy, t->L[l].cy, 0.f, 0.f, 1.f};
    float c0 = cof3(K, 0, 0), c1 = cof3(K, 1, 0), c2 = cof3(K, 2, 0);
    volatile float p0 = c0 * K[0], p1 = c1 * K[3], p2 = c2 * K[6];
    volatile float s12 = p1 + p2;
    float det = p0 + s12;
    float invdet = 1.0f / det;
    float* r = t->L[l].Ki;
    r[0] = c0 * invdet; r[1] = c1 * invdet; r[2] = c2 * invdet;
    r[3] = cof3(K, 0, 1) * invdet; r[4] = cof3(K, 1, 1) * invdet; r[5] = cof3(K, 2, 1) * invdet;
    r[6] = cof3(K, 0, 2) * invdet; r[7] = cof3(K, 1, 2) * invdet; r[8] = cof3(K, 2, 2) * invdet;
  }
  return 0;
}

int hslam_tracker_get_K(hslam_tracker* t, int lvl, float out4[4], int out_wh[2], float out_Ki[9]) {
  HS_ARG(t && lvl >= 0 && lvl < t->levels, "hslam_tracker_get_K: bad level");
  if (out4) { out4[0] = t->L[lvl].fx; out4[1] = t->L[lvl].fy; out4[2] = t->L[lvl].cx; out4[3] = t->L[lvl].cy; }
  if (out_wh) { out_wh[0] = t->L[lvl].w; out_wh[1] = t->L[lvl].h; }
  if (out_Ki) memcpy(out_Ki, t->L[lvl].Ki, 9 * sizeof(float));
  return 0;
}

int hslam_tracker_set_ref(hslam_tracker* t, int lvl, const float* u, const float* v, const float* idepth,
                          const float* color, int n) {
  HS_ARG(t && lvl >= 0 && lvl < t->levels && n >= 0, "hslam_tracker_set_ref: bad level/count");
  HS_ARG(n == 0 || (u && v && idepth && color), "hslam_tracker_set_ref: null array");
  hslam_ctx* ctx = t->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  if (n > t->pc_cap[lvl]) {
    HS_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(t->pc[lvl]);
    int cap = n + n / 4 + 256;
    HS_CUDA(cudaMalloc(&t->pc[lvl], (size_t)cap * sizeof(float4)));
    t->pc_cap[lvl] = cap;
  }
  if ((size_t)n > t->h_pack_cap) {
    HS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (t->h_pack) cudaFreeHost(t->h_pack);
    t->h_pack_cap = (size_t)n + n / 4 + 256;
    HS_CUDA(cudaMallocHost(&t->h_pack, t->h_pack_cap * sizeof(float4)));
  } else {
    HS_CUDA(cudaStreamSynchronize(ctx->stream));  // previous async copy out of h_pack
  }
  for (int i = 0; i < n; ++i) t->h_pack[i] = make_float4(u[i], v[i], idepth[i], color[i]);
  if (n) HS_CUDA(cudaMemcpyAsync(t->pc[lvl], t->h_pack, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  t->L[lvl].n = n;
  t->L[lvl].pc = t->pc[lvl];
  return 0;
}

int hslam_tracker_set_ref_aff(hslam_tracker* t, float ref_exposure, double a, double b) {
  HS_ARG(t, "hslam_tracker_set_ref_aff: null");
  t->ref_exposure = ref_exposure; t->ref_aff_a = a; t->ref_aff_b = b;
  return 0;
}

static int run_eval(hslam_tracker* t, int lvl, const hslam_frame* f, const double T[12], double aff_a, double aff_b,
                    float new_exposure, float cutoffTH) {
  HS_ARG(t && f && T, "tracker eval: null argument");
  HS_ARG(lvl >= 0 && lvl < t->levels && lvl < f->levels, "tracker eval: level out of range");
  HS_ARG(t->L[lvl].w == (f->w >> lvl) && t->L[lvl].h == (f->h >> lvl), "tracker eval: frame size != tracker K size");
  HS_ARG(f->ctx == t->ctx, "tracker eval: frame belongs to another context");
  hslam_ctx* ctx = t->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  EvalArgs a;
  a.L = t->L[lvl]; a.lvl = lvl; a.img = f->dI[lvl];
  memcpy(a.T, T, sizeof(a.T));
  a.aff_a = aff_a; a.aff_b = aff_b; a.ref_aff_a = t->ref_aff_a; a.ref_aff_b = t->ref_aff_b;
  a.ref_exposure = t->ref_exposure; a.new_exposure = new_exposure; a.cutoff = cutoffTH; a.huber = ctx->prm.huberTH;
  int blocks = (t->L[lvl].n + TRK_THREADS - 1) / TRK_THREADS;
  if (blocks < 1) blocks = 1;
  int cap = ctx->sm_count * 4;
  if (blocks > cap) blocks = cap;
  int rc = ensure_partials(t, blocks);
  if (rc) return rc;
  HS_LAUNCH(ctx, HSLAM_K_TRK_EVAL,
            (trk_eval_kernel<<<blocks, TRK_THREADS, 0, ctx->stream>>>(a, t->partial[0], t->ticket, t->d_result)));
  HS_CUDA(cudaGetLastError());
  HS_CUDA(cudaMemcpyAsync(t->h_result, t->d_result, sizeof(EvalResult), cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int hslam_tracker_calc_res(hslam_tracker* t, int lvl, const hslam_frame* f, const double T[12], double aff_a,
                           double aff_b, float new_exposure, float cutoffTH, double out6[6], int* n_warped_out) {
  HS_ARG(out6, "hslam_tracker_calc_res: null output");
  int rc = run_eval(t, lvl, f, T, aff_a, aff_b, new_exposure, cutoffTH);
  if (rc) return rc;
  const EvalResult& r = *t->h_result;
  out6[0] = r.E;
  out6[1] = r.nE;
  out6[2] = (double)((float)r.flowT / ((float)r.flowNum + 0.1));
  out6[3] = 0;
  out6[4] = (double)((float)r.flowRT / ((float)r.flowNum + 0.1));
  out6[5] = (double)((float)r.nSat / (float)r.nE);
  if (n_warped_out) *n_warped_out = r.nW;
  return 0;
}

static void scales_from(const hslam_params& p, float* s) {
  s[0] = s[1] = s[2] = p.scale_xi_rot; s[3] = s[4] = s[5] = p.scale_xi_trans; s[6] = p.scale_a; s[7] = p.scale_b;
}

int hslam_tracker_calc_gs(hslam_tracker* t, int lvl, const hslam_frame* f, const double T[12], double aff_a,
                          double aff_b, float new_exposure, float cutoffTH, double H[64], double b[8]) {
  HS_ARG(H && b, "hslam_tracker_calc_gs: null output");
  int rc = run_eval(t, lvl, f, T, aff_a, aff_b, new_exposure, cutoffTH);
  if (rc) return rc;
  float s8[8];
  scales_from(t->ctx->prm, s8);
  make_Hb(t->h_result->acc, t->h_result->nW, s8, H, b);
  return 0;
}

// Enqueue one launch that tracks `n` members (frames / pose hypotheses) against this tracker's reference: member m gets
// its own thread-block cluster, its own TrackIO slot (initial pose in, results out) and its own frame.
static int track_enqueue_n(hslam_tracker* t, int n, const hslam_frame* const* frames, const float* new_exposure,
                           const double* T /*n*12*/, const double* aff /*n*2*/, int coarsestLvl, const double minResForAbort[5]) {
  HS_ARG(t && frames && T && aff && n >= 1, "hslam_tracker_track: null argument");
  HS_ARG(coarsestLvl >= 0 && coarsestLvl < 5 && coarsestLvl < t->levels,
         "hslam_tracker_track: coarsestLvl out of range (reference asserts < 5, CoarseTracker.cpp:511)");
  HS_ARG(n == 1 || t->cluster_ctas > 0, "hslam_tracker_track_batch: needs the cluster mode (hslam_tracker_set_mode(trk, 8 or 16))");
  hslam_ctx* ctx = t->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  if (n > t->io_cap) {
    HS_CUDA(cudaStreamSynchronize(ctx->stream));   // nothing of the old slots is in flight
    if (t->d_io) cudaFree(t->d_io);
    if (t->h_io) cudaFreeHost(t->h_io);
    t->d_io = nullptr; t->h_io = nullptr;
    t->io_cap = std::max(n, 2 * t->io_cap);
    HS_CUDA(cudaMalloc(&t->d_io, (size_t)t->io_cap * sizeof(TrackIO)));
    HS_CUDA(cudaMallocHost((void**)&t->h_io, (size_t)t->io_cap * sizeof(TrackIO)));
  }
  TrackArgs a;
  memset(&a, 0, sizeof(a));
  int maxn = 0;
  for (int l = 0; l <= coarsestLvl; ++l) {
    a.L[l] = t->L[l];
    if (t->L[l].n > maxn) maxn = t->L[l].n;
  }
  a.coarsestLvl = coarsestLvl;
  for (int i = 0; i < 5; ++i) a.maxIterations[i] = ctx->prm.max_tracker_iterations[i];
  scales_from(ctx->prm, a.scales8);
  a.coarseCutoffTH = ctx->prm.coarseCutoffTH; a.huber = ctx->prm.huberTH;
  a.affineOptModeA = ctx->prm.affineOptModeA; a.affineOptModeB = ctx->prm.affineOptModeB;
  a.ref_exposure = t->ref_exposure;
  a.ref_aff_a = t->ref_aff_a; a.ref_aff_b = t->ref_aff_b;
  for (int m = 0; m < n; ++m) {
    const hslam_frame* f = frames[m];
    HS_ARG(f && f->ctx == t->ctx, "hslam_tracker_track: frame missing or from another context");
    HS_ARG(coarsestLvl < f->levels, "hslam_tracker_track: coarsestLvl beyond the frame's pyramid");
    HS_ARG(t->L[0].w == f->w && t->L[0].h == f->h, "hslam_tracker_track: frame size != tracker K size");
    TrackIO* io = t->h_io + m;
    memset(io, 0, sizeof(TrackIO));
    memcpy(io->T, T + 12 * m, sizeof(io->T));
    io->aff[0] = aff[2 * m]; io->aff[1] = aff[2 * m + 1];
    for (int i = 0; i < 5; ++i) io->minResForAbort[i] = minResForAbort ? minResForAbort[i] : NAN;
    for (int l = 0; l <= coarsestLvl; ++l) io->img[l] = f->dI[l];
    io->new_exposure = new_exposure ? new_exposure[m] : 1.f;
  }
  HS_CUDA(cudaMemcpyAsync(t->d_io, t->h_io, (size_t)n * sizeof(TrackIO), cudaMemcpyHostToDevice, ctx->stream));
  cudaError_t lerr = cudaSuccess;
  if (t->cluster_ctas > 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(t->cluster_ctas * n); cfg.blockDim = dim3(TRK_CL_THREADS); cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = t->cluster_ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    float* nullp = nullptr;
    HS_LAUNCH(ctx, HSLAM_K_TRK_TRACK, (lerr = cudaLaunchKernelEx(&cfg, trk_track_kernel<TRK_CL_THREADS, true>, a, nullp, nullp, t->d_io)));
  } else {
    int blocks = (maxn + TRK_THREADS - 1) / TRK_THREADS;
    if (blocks < 1) blocks = 1;
    if (blocks > t->coop_blocks) blocks = t->coop_blocks;
    int rc = ensure_partials(t, blocks);
    if (rc) return rc;
    void* args[] = {&a, &t->partial[0], &t->partial[1], &t->d_io};
    HS_LAUNCH(ctx, HSLAM_K_TRK_TRACK, (lerr = cudaLaunchCooperativeKernel((void*)trk_track_kernel<TRK_THREADS, false>, dim3(blocks),
                                                                         dim3(TRK_THREADS), args, 0, ctx->stream)));
  }
  HS_CUDA(lerr);
  HS_CUDA(cudaMemcpyAsync(t->h_io, t->d_io, (size_t)n * sizeof(TrackIO), cudaMemcpyDeviceToHost, ctx->stream));
  t->n_inflight = n;
  return 0;
}

static void track_unpack(hslam_tracker* t, int m, double T_out[12], double aff_out[2], double lastResiduals[5], double lastFlow[3],
                         int* ok, int* iterations_out) {
  const TrackIO* io = t->h_io + m;
  // the reference writes lastToNew_out/aff_g2l_out only when all levels ran (:663-664); the kernel mirrors that
  if (T_out) memcpy(T_out, io->T, sizeof(io->T));
  if (aff_out) { aff_out[0] = io->aff[0]; aff_out[1] = io->aff[1]; }
  if (lastResiduals) memcpy(lastResiduals, io->lastResiduals, sizeof(io->lastResiduals));
  if (lastFlow) memcpy(lastFlow, io->lastFlow, sizeof(io->lastFlow));
  if (ok) *ok = io->ok;
  if (iterations_out) memcpy(iterations_out, io->iterations, sizeof(io->iterations));
}

int hslam_tracker_track_enqueue(hslam_tracker* t, const hslam_frame* f, float new_exposure, const double T[12],
                                const double aff[2], int coarsestLvl, const double minResForAbort[5]) {
  HS_ARG(t && f && T && aff, "hslam_tracker_track: null argument");
  return track_enqueue_n(t, 1, &f, &new_exposure, T, aff, coarsestLvl, minResForAbort);
}

int hslam_tracker_track_finish(hslam_tracker* t, double T_out[12], double aff_out[2], double lastResiduals[5],
                               double lastFlow[3], int* ok, int* iterations_out) {
  HS_ARG(t, "hslam_tracker_track_finish: null");
  HS_CUDA(cudaSetDevice(t->ctx->device));
  HS_CUDA(cudaStreamSynchronize(t->ctx->stream));
  track_unpack(t, 0, T_out, aff_out, lastResiduals, lastFlow, ok, iterations_out);
  memcpy(t->last_phase_cycles, t->h_io->phase_cycles, sizeof(t->h_io->phase_cycles));
  t->last_evals = t->h_io->evals;
  return 0;
}

int hslam_tracker_track_batch(hslam_tracker* t, int n, const hslam_frame* const* frames, const float* new_exposure,
                              double* T_io, double* aff_io, int coarsestLvl, const double minResForAbort[5],
                              double* lastResiduals, double* lastFlow, int* ok, int* iterations_out) {
  HS_ARG(t && n >= 1 && frames && T_io && aff_io, "hslam_tracker_track_batch: bad argument");
  int rc = track_enqueue_n(t, n, frames, new_exposure, T_io, aff_io, coarsestLvl, minResForAbort);
  if (rc) return rc;
  HS_CUDA(cudaStreamSynchronize(t->ctx->stream));
  for (int m = 0; m < n; ++m)
    track_unpack(t, m, T_io + 12 * m, aff_io + 2 * m, lastResiduals ? lastResiduals + 5 * m : nullptr,
                 lastFlow ? lastFlow + 3 * m : nullptr, ok ? ok + m : nullptr, iterations_out ? iterations_out + 5 * m : nullptr);
  memcpy(t->last_phase_cycles, t->h_io->phase_cycles, sizeof(t->h_io->phase_cycles));
  t->last_evals = t->h_io->evals;
  return 0;
}

int hslam_tracker_track(hslam_tracker* t, const hslam_frame* f, float new_exposure, double T_io[12], double aff_io[2],
                        int coarsestLvl, const double minResForAbort[5], double lastResiduals[5], double lastFlow[3],
                        int* ok, int* iterations_out) {
  int rc = hslam_tracker_track_enqueue(t, f, new_exposure, T_io, aff_io, coarsestLvl, minResForAbort);
  if (rc) return rc;
  return hslam_tracker_track_finish(t, T_io, aff_io, lastResiduals, lastFlow, ok, iterations_out);
}

}  // extern "C"

// ================================================================================================
// CoarseTracker::makeCoarseDepthL0 on the device (reference Src/CoarseTracker.cpp:105-263):
// splat the active points that are IN the newest keyframe, sum-pyramid, one dilation per level, normalise and
// compact into the per-level reference clouds pc_* — which the tracker reads from HBM anyway, so the per-keyframe
// host round trip (point side effects down, five pc arrays up) disappears.
//   * splat: the source accumulates `idepth += new_idepth*weight` in point order; two points in one pixel are rare
//     but must add in that order.  Points are linked per pixel with atomicExch, the pixel's list head then adds
//     its list in ascending point index: deterministic and bit-identical to the sequential loop, no sort.
//   * compaction keeps the source's raster order with a three-kernel ordered scan (block counts, scan of the
//     counts, ballot-ranked emit).
// ================================================================================================
namespace hs {

struct CdSeed { int u, v; float idepth, weight; };

__global__ void cd_link_kernel(const CdSeed* __restrict__ seeds, int n, int w, int h, int* __restrict__ head, int* __restrict__ next) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const CdSeed s = seeds[i];
  next[i] = -2;                                            // not part of any list
  if (s.u < 0 || s.v < 0 || s.u >= w || s.v >= h) return;
  next[i] = atomicExch(&head[s.u + w * s.v], i);
}
__global__ void cd_splat_kernel(const CdSeed* __restrict__ seeds, int n, int w, const int* __restrict__ head,
                                const int* __restrict__ next, float* __restrict__ idepth, float* __restrict__ wsum) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || next[i] == -2) return;
  const int pix = seeds[i].u + w * seeds[i].v;
  if (head[pix] != i) return;                              // one owner per pixel
  float accI = 0.f, accW = 0.f;
  int last = -1;
  while (true) {                                           // ascending point index = the source's loop order
    int best = 0x7fffffff;
    for (int j = i; j >= 0; j = next[j]) if (j > last && j < best) best = j;
    if (best == 0x7fffffff) break;
    accI = xadd(accI, xmul(seeds[best].idepth, seeds[best].weight));
    accW = xadd(accW, seeds[best].weight);
    last = best;
  }
  idepth[pix] = accI; wsum[pix] = accW;
}
__global__ void cd_down_kernel(const float* __restrict__ idm, const float* __restrict__ wsm, int wlm1, float* __restrict__ idl,
                               float* __restrict__ wsl, int wl, int hl) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= wl * hl) return;
  const int x = i % wl, y = i / wl, b = 2 * x + 2 * y * wlm1;
  idl[i] = xadd(xadd(xadd(idm[b], idm[b + 1]), idm[b + wlm1]), idm[b + wlm1 + 1]);
  wsl[i] = xadd(xadd(xadd(wsm[b], wsm[b + 1]), wsm[b + wlm1]), wsm[b + wlm1 + 1]);
}
// one dilation: pixels without weight take the mean of their weighted neighbours (diagonal on levels 0-1, 4-neighbourhood
// above); reads only pixels WITH weight, writes only pixels WITHOUT: order-free.  `wbak` is the copy of the weights.
__global__ void cd_dilate_kernel(float* __restrict__ idl, float* __restrict__ wsl, const float* __restrict__ wbak, int wl, int hl,
                                 int diagonal) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + wl;
  const int tot = wl * hl, wh = tot - wl;
  if (i >= wh || wbak[i] > 0.f) return;
  const int o0 = diagonal ? 1 + wl : 1, o1 = diagonal ? -1 - wl : -1, o2 = diagonal ? wl - 1 : wl, o3 = diagonal ? -wl + 1 : -wl;
  float sum = 0.f, num = 0.f, numn = 0.f;
  const int offs[4] = {o0, o1, o2, o3};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int j = i + offs[k];
    if (j < tot && wbak[j] > 0.f) { sum = xadd(sum, idl[j]); num = xadd(num, wbak[j]); numn += 1.f; }   // j == tot: the source reads one past the end
  }
  if (numn > 0.f) { idl[i] = xdiv(sum, numn); wsl[i] = xdiv(num, numn); }
}
__device__ __forceinline__ bool cd_valid(const float* idl, const float* wsl, const float4* dI, int wl, int hl, int i, float4& out) {
  if (i >= wl * hl) return false;
  const int x = i % wl, y = i / wl;
  if (x < 2 || y < 2 || x >= wl - 2 || y >= hl - 2) return false;
  const float ws = wsl[i];
  if (!(ws > 0.f)) return false;
  const float idp = xdiv(idl[i], ws), col = dI[i].x;
  if (!isfinite(col) || !(idp > 0.f)) return false;
  out = make_float4((float)x, (float)y, idp, col);
  return true;
}
constexpr int CD_BLOCK = 1024;
__global__ void __launch_bounds__(CD_BLOCK) cd_count_kernel(const float* idl, const float* wsl, const float4* dI, int wl, int hl, int* blockcnt) {
  float4 o;
  const bool v = cd_valid(idl, wsl, dI, wl, hl, blockIdx.x * CD_BLOCK + threadIdx.x, o);
  const int c = __syncthreads_count(v);
  if (threadIdx.x == 0) blockcnt[blockIdx.x] = c;
}
__global__ void __launch_bounds__(1024) cd_scan_kernel(int* blockcnt, int nblocks, int* total_out) {   // one block, exclusive scan in place
  __shared__ int carry, wsum[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nblocks ? blockcnt[i] : 0;
    int incl = v;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int s = wsum[lane], is = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, is, o); if (lane >= o) is += t; }
      wsum[lane] = is - s;                                  // exclusive per warp
    }
    __syncthreads();
    const int excl = carry + wsum[wid] + incl - v;
    if (i < nblocks) blockcnt[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}
__global__ void __launch_bounds__(CD_BLOCK) cd_emit_kernel(const float* idl, const float* wsl, const float4* dI, int wl, int hl,
                                                            const int* __restrict__ blockoff, float4* __restrict__ pc) {
  __shared__ int woff[CD_BLOCK / 32];
  float4 o;
  const bool v = cd_valid(idl, wsl, dI, wl, hl, blockIdx.x * CD_BLOCK + threadIdx.x, o);
  const unsigned m = __ballot_sync(0xffffffffu, v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) woff[wid] = __popc(m);
  __syncthreads();
  if (wid == 0) {
    int s = woff[lane], is = s;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) { int t = __shfl_up_sync(0xffffffffu, is, k); if (lane >= k) is += t; }
    woff[lane] = is - s;
  }
  __syncthreads();
  if (v) pc[blockoff[blockIdx.x] + woff[wid] + __popc(m & ((1u << lane) - 1u))] = o;
}
// seeds from host arrays: int u = cpt[0] + 0.5f, weight = sqrtf(1e-3 / (HdiF + 1e-12))   (CoarseTracker.cpp:120-123)
__global__ void cd_seed_kernel(const float* __restrict__ cpt3, const float* __restrict__ hdif, int n, CdSeed* __restrict__ seeds) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  CdSeed s;
  s.u = (int)xadd(cpt3[3 * i], 0.5f); s.v = (int)xadd(cpt3[3 * i + 1], 0.5f);
  s.idepth = cpt3[3 * i + 2];
  s.weight = xsqrt((float)(1e-3 / ((double)hdif[i] + 1e-12)));
  seeds[i] = s;
}

}  // namespace hs

int hs_ba_gather_coarse_inputs(hslam_ba* ba, hslam_ctx* ctx, const int32_t* residual_ids, int count, float* d_cpt3, float* d_hdif);   // ba.cu

static int cd_run(hslam_tracker* t, const hslam_frame* ref, const hs::CdSeed* d_seeds, int n, int* d_head, int* d_next, float* d_maps,
                  int* d_blockcnt, int* d_total) {
  using namespace hs;
  hslam_ctx* ctx = t->ctx;
  cudaStream_t st = ctx->stream;
  const int levels = t->levels;
  // maps: per level idepth | wsum, then one wbak of level-0 size
  std::vector<float*> idp(levels), wsm(levels);
  size_t off = 0;
  for (int l = 0; l < levels; ++l) { size_t px = (size_t)t->L[l].w * t->L[l].h; idp[l] = d_maps + off; wsm[l] = d_maps + off + px; off += 2 * px; }
  float* wbak = d_maps + off;
  const int w0 = t->L[0].w, h0 = t->L[0].h;
  HS_CUDA(cudaMemsetAsync(d_maps, 0, off * sizeof(float), st));
  HS_CUDA(cudaMemsetAsync(d_head, 0xff, (size_t)w0 * h0 * sizeof(int), st));
  if (n > 0) {
    HS_LAUNCH(ctx, HSLAM_K_COARSE_DEPTH, (cd_link_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_seeds, n, w0, h0, d_head, d_next)));
    HS_LAUNCH(ctx, HSLAM_K_COARSE_DEPTH, (cd_splat_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_seeds, n, w0, d_head, d_next, idp[0], wsm[0])));
  }
  for (int l = 1; l < levels; ++l) {
    const int wl = t->L[l].w, hl = t->L[l].h;
    HS_LAUNCH(ctx, HSLAM_K_COARSE_DEPTH, (cd_down_kernel<<<(wl * hl + 255) / 256, 256, 0, st>>>(idp[l - 1], wsm[l - 1], t->L[l - 1].w, idp[l], wsm[l], wl, hl)));
  }
  for (int l = 0; l < levels; ++l) {
    const int wl = t->L[l].w, hl = t->L[l].h;
    HS_CUDA(cudaMemcpyAsync(wbak, wsm[l], (size_t)wl * hl * sizeof(float), cudaMemcpyDeviceToDevice, st));
    HS_LAUNCH(ctx, HSLAM_K_COARSE_DEPTH, (cd_dilate_kernel<<<(wl * hl + 255) / 256, 256, 0, st>>>(idp[l], wsm[l], wbak, wl, hl, l < 2 ? 1 : 0)));
  }
  for (int l = 0; l < levels; ++l) {
    const int wl = t->L[l].w, hl = t->L[l].h, nb = (wl * hl + CD_BLOCK - 1) / CD_BLOCK;
    if (wl * hl > t->pc_cap[l]) {
      HS_CUDA(cudaStreamSynchronize(st));
      cudaFree(t->pc[l]);
      HS_CUDA(cudaMalloc(&t->pc[l], (size_t)wl * hl * sizeof(float4)));
      t->pc_cap[l] = wl * hl;
    }
    HS_LAUNCH(ctx, HSLAM_K_COARSE_DEPTH, (cd_count_kernel<<<nb, CD_BLOCK, 0, st>>>(idp[l], wsm[l], ref->dI[l], wl, hl, d_blockcnt)));
    HS_LAUNCH(ctx, HSLAM_K_COARSE_DEPTH, (cd_scan_kernel<<<1, 1024, 0, st>>>(d_blockcnt, nb, d_total + l)));
    HS_LAUNCH(ctx, HSLAM_K_COARSE_DEPTH, (cd_emit_kernel<<<nb, CD_BLOCK, 0, st>>>(idp[l], wsm[l], ref->dI[l], wl, hl, d_blockcnt, t->pc[l])));
  }
  HS_CUDA(cudaGetLastError());
  int counts[HSLAM_MAX_PYR_LEVELS];
  HS_CUDA(cudaMemcpyAsync(counts, d_total, levels * sizeof(int), cudaMemcpyDeviceToHost, st));
  HS_CUDA(cudaStreamSynchronize(st));
  for (int l = 0; l < levels; ++l) { t->L[l].n = counts[l]; t->L[l].pc = t->pc[l]; }
  return 0;
}

static int cd_buffers(hslam_tracker* t, int n, hs::CdSeed** seeds, int** head, int** next, float** maps, int** blockcnt, int** total,
                      float** in4) {
  // one allocation, grown on demand: seeds | next | in (4 floats per point) | head | maps | blockcnt | totals
  size_t px = 0, px0 = (size_t)t->L[0].w * t->L[0].h;
  for (int l = 0; l < t->levels; ++l) px += (size_t)t->L[l].w * t->L[l].h;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t cap_n = (size_t)std::max(n, 1);
  const size_t b_seed = al(cap_n * sizeof(hs::CdSeed)), b_next = al(cap_n * 4), b_in = al(cap_n * 16), b_head = al(px0 * 4),
               b_maps = al((2 * px + px0) * 4), b_cnt = al((px0 / hs::CD_BLOCK + 2) * 4), b_tot = 256;
  const size_t total_b = b_seed + b_next + b_in + b_head + b_maps + b_cnt + b_tot;
  if (total_b > t->cd_cap) {
    HS_CUDA(cudaStreamSynchronize(t->ctx->stream));
    if (t->cd_buf) cudaFree(t->cd_buf);
    t->cd_buf = nullptr; t->cd_cap = 0;
    HS_CUDA(cudaMalloc((void**)&t->cd_buf, total_b + total_b / 4));
    t->cd_cap = total_b + total_b / 4;
  }
  char* p = t->cd_buf;
  *seeds = (hs::CdSeed*)p; p += b_seed; *next = (int*)p; p += b_next; *in4 = (float*)p; p += b_in; *head = (int*)p; p += b_head;
  *maps = (float*)p; p += b_maps; *blockcnt = (int*)p; p += b_cnt; *total = (int*)p;
  return 0;
}

extern "C" {

int hslam_tracker_make_coarse_depth(hslam_tracker* t, const hslam_frame* lastRef, int n, const float* centerProjectedTo,
                                    const float* HdiF) {
  HS_ARG(t && lastRef && t->levels > 0 && n >= 0 && (n == 0 || (centerProjectedTo && HdiF)), "hslam_tracker_make_coarse_depth: bad argument");
  HS_ARG(lastRef->ctx == t->ctx && lastRef->levels >= t->levels && lastRef->w == t->L[0].w && lastRef->h == t->L[0].h,
         "hslam_tracker_make_coarse_depth: reference frame does not match the tracker (context, size or levels)");
  hslam_ctx* ctx = t->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  hs::CdSeed* seeds; int *head, *next, *blockcnt, *total; float *maps, *in4;
  int rc = cd_buffers(t, n, &seeds, &head, &next, &maps, &blockcnt, &total, &in4);
  if (rc) return rc;
  if (n > 0) {
    if ((size_t)n > t->h_pack_cap) {
      HS_CUDA(cudaStreamSynchronize(ctx->stream));
      if (t->h_pack) cudaFreeHost(t->h_pack);
      t->h_pack_cap = (size_t)n + n / 4 + 256;
      HS_CUDA(cudaMallocHost(&t->h_pack, t->h_pack_cap * sizeof(float4)));
    } else {
      HS_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    float* hp = (float*)t->h_pack;
    memcpy(hp, centerProjectedTo, (size_t)n * 12);
    memcpy(hp + 3 * (size_t)n, HdiF, (size_t)n * 4);
    HS_CUDA(cudaMemcpyAsync(in4, hp, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
    HS_LAUNCH(ctx, HSLAM_K_COARSE_DEPTH, (hs::cd_seed_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(in4, in4 + 3 * (size_t)n, n, seeds)));
  }
  return cd_run(t, lastRef, seeds, n, head, next, maps, blockcnt, total);
}

int hslam_tracker_make_coarse_depth_from_ba(hslam_tracker* t, const hslam_frame* lastRef, hslam_ba* ba, const int32_t* residual_ids,
                                            int count) {
  HS_ARG(t && lastRef && ba && t->levels > 0 && count >= 0 && (count == 0 || residual_ids), "hslam_tracker_make_coarse_depth_from_ba: bad argument");
  HS_ARG(lastRef->ctx == t->ctx && lastRef->levels >= t->levels && lastRef->w == t->L[0].w && lastRef->h == t->L[0].h,
         "hslam_tracker_make_coarse_depth_from_ba: reference frame does not match the tracker (context, size or levels)");
  hslam_ctx* ctx = t->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  hs::CdSeed* seeds; int *head, *next, *blockcnt, *total; float *maps, *in4;
  int rc = cd_buffers(t, count, &seeds, &head, &next, &maps, &blockcnt, &total, &in4);
  if (rc) return rc;
  if (count > 0) {
    rc = hs_ba_gather_coarse_inputs(ba, ctx, residual_ids, count, in4, in4 + 3 * (size_t)count);
    if (rc) return rc;
    HS_LAUNCH(ctx, HSLAM_K_COARSE_DEPTH, (hs::cd_seed_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>(in4, in4 + 3 * (size_t)count, count, seeds)));
  }
  return cd_run(t, lastRef, seeds, count, head, next, maps, blockcnt, total);
}

int hslam_tracker_get_ref(hslam_tracker* t, int lvl, int capacity, float* u, float* v, float* idepth, float* color, int* n_out) {
  HS_ARG(t && lvl >= 0 && lvl < t->levels && n_out, "hslam_tracker_get_ref: bad argument");
  HS_CUDA(cudaSetDevice(t->ctx->device));
  const int n = t->L[lvl].n;
  *n_out = n;
  if (n == 0 || !u) return 0;
  HS_ARG(capacity >= n && v && idepth && color, "hslam_tracker_get_ref: output arrays too small");
  std::vector<float4> tmp(n);
  HS_CUDA(cudaMemcpyAsync(tmp.data(), t->pc[lvl], (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, t->ctx->stream));
  HS_CUDA(cudaStreamSynchronize(t->ctx->stream));
  for (int i = 0; i < n; ++i) { u[i] = tmp[i].x; v[i] = tmp[i].y; idepth[i] = tmp[i].z; color[i] = tmp[i].w; }
  return 0;
}

}  // extern "C"
