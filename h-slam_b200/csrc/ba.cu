// ba.cu — windowed photometric bundle-adjustment linearisation on the device.
//
// Replaces (reference file:line)
//   PointFrameResidual::linearize / applyRes / takeData / fixLinearizationF
//       Src/OptimizationClasses.cpp:43-284, Include/OptimizationClasses.h:155-161
//   AccumulatedTopHessianSSE::addPoint<mode> + AccumulatorApprox      Src/AccumulatedTopHessian.cpp:21-141
//   AccumulatedSCHessianSSE::addPoint + AccumulatorXX/X (Schur fold)  Src/AccumulatedSCHessian.cpp:10-53
//   stitchDoubleInternal / stitchDoubleMT (top and SC)                AccumulatedTopHessian.cpp:218-280, .h:69-117,
//                                                                     AccumulatedSCHessian.cpp:54-133, .h:70-111
//   EnergyFunctional::accumulateAF_MT/LF_MT/SCF_MT, resubstituteF_MT  Src/EnergyFunctional.cpp:155-274
//   IndexThreadReduce (replaced by the CUDA grid)                     Include/IndexThreadReduce.h
//
// Layout in HBM.  Residuals are re-ordered once per topology change by (host,target) pair, so a
// block of the linearise kernel shares one target image and one precalc record, and a block of the
// top-Hessian kernel reduces into ONE relative 13x13 accumulator.  Everything per residual is SoA with
// a padded stride (component-major), so every load/store of a warp is one coalesced 128-byte line.
// Points are ordered by host so a Schur block owns the n^2 8x8 blocks of one host.
//
// Determinism.  No floating-point atomics anywhere: fixed-shape warp trees, fixed-order cross-warp
// sums, per-block partials folded by block index in fp64, and a gather-style stitch (one thread per
// output element, contributions visited in a fixed order).
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <dlfcn.h>

namespace hs {

constexpr int JS = 80;                 // floats per residual row: 74 of RawResidualJacobian + JI_r(2) Jab_r(2) rr(1) + pad
constexpr int J_RESF = 0, J_XI0 = 8, J_XI1 = 14, J_C0 = 20, J_C1 = 24, J_DD = 28, J_IDX0 = 30, J_IDX1 = 38,
              J_AB0 = 46, J_AB1 = 54, J_IDX2 = 62, J_ABIDX = 66, J_AB2 = 70, J_JIR = 74, J_JABR = 76, J_RR = 78;
constexpr int TOP_THREADS = 128;
constexpr int TOP_PART = 96;           // 91 accumulators + nres + pad
constexpr int SC_THREADS = 256;
constexpr int SC_PTS = 64;             // points staged in shared memory per sub-tile

__constant__ int c_pat[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};

struct BAWin {
  int n, w, h;
  float fx, fy, cx, cy, fxi, fyi;
  float huber, outlierTH, scale_idepth, scale_f, scale_c, affA, affB;
  float cDeltaF[4];
  const float4* img[HSLAM_MAX_FRAMES];
  float frameEnergyTH[HSLAM_MAX_FRAMES];
  const float* precalc;      // n*n*28
  const float* adHTdeltaF;   // n*n*8
};

struct BARes {               // per-residual SoA, sorted order, stride = rs
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
  float* JpJdF;              // [8][rs]
  float* rtz;                // [8][rs] res_toZeroF
  float* cpt;                // [3][rs] centerProjectedTo
  float* pto;                // [16][rs] projectedTo
  float* ptc[2];             // [6][rs] per-residual contribution to (bd,Hdd,Hcd[4]) for mode A / L
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

// ------------------------------------------------------------------------------------------------
// K4: PointFrameResidual::linearize (+ applyRes/takeData when APPLY) — one thread per residual.
// The whole routine uses exactly-rounded primitives in the operation order of the source, so every
// output (states, energies, all 74 Jacobian floats) is bit-identical to the float-faithful oracle.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void take_data(const BARes& r, int k) {  // OptimizationClasses.h:155-161
  const float* J = r.J + k;
  const int s = r.rs;
  float d0 = J[J_DD * s], d1 = J[(J_DD + 1) * s];
  float JJ0 = xadd(xmul(J[(J_IDX2 + 0) * s], d0), xmul(J[(J_IDX2 + 2) * s], d1));
  float JJ1 = xadd(xmul(J[(J_IDX2 + 1) * s], d0), xmul(J[(J_IDX2 + 3) * s], d1));
#pragma unroll
  for (int i = 0; i < 6; ++i)
    r.JpJdF[i * s + k] = xadd(xmul(J[(J_XI0 + i) * s], JJ0), xmul(J[(J_XI1 + i) * s], JJ1));
  r.JpJdF[6 * s + k] = xadd(xmul(J[(J_ABIDX + 0) * s], d0), xmul(J[(J_ABIDX + 2) * s], d1));
  r.JpJdF[7 * s + k] = xadd(xmul(J[(J_ABIDX + 1) * s], d0), xmul(J[(J_ABIDX + 3) * s], d1));
}

__device__ __forceinline__ void apply_one(const BARes& r, int k) {  // OptimizationClasses.cpp:235-256
  if (r.state_state[k] == HSLAM_RES_OOB) return;
  int8_t ns = r.newState[k];
  if (ns == HSLAM_RES_IN) { r.isActive[k] = 1; take_data(r, k); } else { r.isActive[k] = 0; }
  r.state_state[k] = ns;
  r.state_energy[k] = r.newEnergy[k];
}

template <bool APPLY>
__global__ void __launch_bounds__(128) ba_linearize_kernel(BAWin W, BARes r, BAPts pts, double* __restrict__ epart,
                                                           unsigned int* __restrict__ ticket,
                                                           double* __restrict__ energy_out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = r.rs;
  double energy_ret = 0.0;
  if (k < r.R && !r.isLin[k]) {
    float* J = r.J + k;
    r.newEnergyWO[k] = -1.f;
    bool oob = (r.state_state[k] == HSLAM_RES_OOB);
    const float old_energy = r.state_energy[k];
    int p = r.point[k], pair = r.pair[k];
    int hostI = pair % W.n, tgtI = pair / W.n;
    float energyLeft = 0.f, wJI2 = 0.f;
    if (!oob) {
      const float4* pc = reinterpret_cast<const float4*>(W.precalc + (size_t)HSLAM_PRECALC_STRIDE * pair);
      float4 q0 = __ldg(pc), q1 = __ldg(pc + 1), q2 = __ldg(pc + 2), q3 = __ldg(pc + 3), q4 = __ldg(pc + 4),
             q5 = __ldg(pc + 5), q6 = __ldg(pc + 6);
      // [0..8] KRKi, [9..11] Kt, [12..20] R0, [21..23] t0, [24] a, [25] b, [26] b0
      const float KRKi[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
      const float Kt[3] = {q2.y, q2.z, q2.w};
      const float R0[9] = {q3.x, q3.y, q3.z, q3.w, q4.x, q4.y, q4.z, q4.w, q5.x};
      const float t0[3] = {q5.y, q5.z, q5.w};
      const float affLL0 = q6.x, affLL1 = q6.y, b0 = q6.z;
      float4 pu = __ldg(pts.uvid + p);
      const float u_pt = pu.x, v_pt = pu.y, idp = pu.z, idz = pu.w;
      const float wm3 = (float)(W.w - 3), hm3 = (float)(W.h - 3);
      // ---- centre projection at the FEJ point, DirectProjection.h:20-38
      float Kl0 = xmul(xsub(xadd(u_pt, 0.f), W.cx), W.fxi), Kl1 = xmul(xsub(xadd(v_pt, 0.f), W.cy), W.fyi);
      float p0 = xadd(xdot3(R0[0], R0[1], R0[2], Kl0, Kl1, 1.f), xmul(t0[0], idz));
      float p1 = xadd(xdot3(R0[3], R0[4], R0[5], Kl0, Kl1, 1.f), xmul(t0[1], idz));
      float p2 = xadd(xdot3(R0[6], R0[7], R0[8], Kl0, Kl1, 1.f), xmul(t0[2], idz));
      float drescale = xdiv(1.0f, p2);
      float new_idepth = xmul(idz, drescale);
      if (!(drescale > 0.f)) oob = true;
      float u = xmul(p0, drescale), v = xmul(p1, drescale);
      float Ku = xadd(xmul(u, W.fx), W.cx), Kv = xadd(xmul(v, W.fy), W.cy);
      if (!oob && !(Ku > 1.1f && Kv > 1.1f && Ku < wm3 && Kv < hm3)) oob = true;
      if (!oob) {
        r.cpt[k] = Ku; r.cpt[s + k] = Kv; r.cpt[2 * s + k] = new_idepth;
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
        J[(J_XI0 + 0) * s] = xmul(new_idepth, W.fx);
        J[(J_XI0 + 1) * s] = 0.f;
        J[(J_XI0 + 2) * s] = xmul(xmul(-new_idepth, u), W.fx);
        J[(J_XI0 + 3) * s] = xmul(xmul(-u, v), W.fx);
        J[(J_XI0 + 4) * s] = xmul(xadd(1.f, xmul(u, u)), W.fx);
        J[(J_XI0 + 5) * s] = xmul(-v, W.fx);
        J[(J_XI1 + 0) * s] = 0.f;
        J[(J_XI1 + 1) * s] = xmul(new_idepth, W.fy);
        J[(J_XI1 + 2) * s] = xmul(xmul(-new_idepth, v), W.fy);
        J[(J_XI1 + 3) * s] = xmul(-xadd(1.f, xmul(v, v)), W.fy);
        J[(J_XI1 + 4) * s] = xmul(xmul(u, v), W.fy);
        J[(J_XI1 + 5) * s] = xmul(u, W.fy);
        J[(J_C0 + 0) * s] = cx0; J[(J_C0 + 1) * s] = cx1; J[(J_C0 + 2) * s] = cx2; J[(J_C0 + 3) * s] = cx3;
        J[(J_C1 + 0) * s] = cy0; J[(J_C1 + 1) * s] = cy1; J[(J_C1 + 2) * s] = cy2; J[(J_C1 + 3) * s] = cy3;
        J[J_DD * s] = ddx; J[(J_DD + 1) * s] = ddy;
        // ---- 8-pixel pattern at the current state, :141-204
        const float4* img = W.img[tgtI];
        float4 c0 = __ldg(pts.color + 2 * p), c1 = __ldg(pts.color + 2 * p + 1);
        float4 w0 = __ldg(pts.weights + 2 * p), w1 = __ldg(pts.weights + 2 * p + 1);
        const float color[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        const float wgt[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        float II00 = 0.f, II11 = 0.f, II10 = 0.f, AI00 = 0.f, AI01 = 0.f, AI10 = 0.f, AI11 = 0.f;
        float AA00 = 0.f, AA01 = 0.f, AA11 = 0.f;
        float JIr0 = 0.f, JIr1 = 0.f, Jabr0 = 0.f, Jabr1 = 0.f, rr = 0.f;
#pragma unroll
        for (int idx = 0; idx < 8; ++idx) {
          if (oob) break;
          float px = xadd(u_pt, (float)c_pat[idx][0]), py = xadd(v_pt, (float)c_pat[idx][1]);
          float a0 = xadd(xdot3(KRKi[0], KRKi[1], KRKi[2], px, py, 1.f), xmul(Kt[0], idp));
          float a1 = xadd(xdot3(KRKi[3], KRKi[4], KRKi[5], px, py, 1.f), xmul(Kt[1], idp));
          float a2 = xadd(xdot3(KRKi[6], KRKi[7], KRKi[8], px, py, 1.f), xmul(Kt[2], idp));
          float Kup = xdiv(a0, a2), Kvp = xdiv(a1, a2);
          if (!(Kup > 1.1f && Kvp > 1.1f && Kup < wm3 && Kvp < hm3)) { oob = true; break; }
          r.pto[(2 * idx) * s + k] = Kup;
          r.pto[(2 * idx + 1) * s + k] = Kvp;
          Hit3 hit = interp33(img, Kup, Kvp, W.w);
          float residual = xsub(hit.I, xadd(xmul(affLL0, color[idx]), affLL1));
          float drdA = xsub(color[idx], b0);
          if (!isfinite(hit.I)) { oob = true; break; }
          float w = xsqrt(xdiv(W.outlierTH, xadd(W.outlierTH, xadd(xmul(hit.dx, hit.dx), xmul(hit.dy, hit.dy)))));
          w = xmul(0.5f, xadd(w, wgt[idx]));
          float ar = fabsf(residual);
          float hw = ar < W.huber ? 1.f : xdiv(W.huber, ar);
          energyLeft = xadd(energyLeft, xmul(xmul(xmul(xmul(xmul(w, w), hw), residual), residual), xsub(2.f, hw)));
          if (hw < 1.f) hw = xsqrt(hw);
          hw = xmul(hw, w);
          float h1 = xmul(hit.dx, hw), h2 = xmul(hit.dy, hw);
          float rF = xmul(residual, hw), ja = xmul(drdA, hw);
          J[(J_RESF + idx) * s] = rF;
          J[(J_IDX0 + idx) * s] = h1;
          J[(J_IDX1 + idx) * s] = h2;
          float jab0 = W.affA < 0.f ? 0.f : ja, jab1 = W.affB < 0.f ? 0.f : hw;
          J[(J_AB0 + idx) * s] = jab0;
          J[(J_AB1 + idx) * s] = jab1;
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
        if (!oob) {
          J[(J_IDX2 + 0) * s] = II00; J[(J_IDX2 + 1) * s] = II10; J[(J_IDX2 + 2) * s] = II10; J[(J_IDX2 + 3) * s] = II11;
          J[(J_ABIDX + 0) * s] = AI00; J[(J_ABIDX + 1) * s] = AI10; J[(J_ABIDX + 2) * s] = AI01; J[(J_ABIDX + 3) * s] = AI11;
          J[(J_AB2 + 0) * s] = AA00; J[(J_AB2 + 1) * s] = AA01; J[(J_AB2 + 2) * s] = AA01; J[(J_AB2 + 3) * s] = AA11;
          J[(J_JIR + 0) * s] = JIr0; J[(J_JIR + 1) * s] = JIr1; J[(J_JABR + 0) * s] = Jabr0; J[(J_JABR + 1) * s] = Jabr1;
          J[J_RR * s] = rr;
        }
      }
    }
    if (oob) {
      r.newState[k] = HSLAM_RES_OOB;
      energy_ret = (double)old_energy;
    } else {
      r.newEnergyWO[k] = energyLeft;
      float th = fmaxf(W.frameEnergyTH[hostI], W.frameEnergyTH[tgtI]);
      if (energyLeft > th || wJI2 < 2.f) { energyLeft = th; r.newState[k] = HSLAM_RES_OUT; }
      else r.newState[k] = HSLAM_RES_IN;
      r.newEnergy[k] = energyLeft;
      energy_ret = (double)energyLeft;
    }
    if (APPLY) apply_one(r, k);
  }
  // ---- energy: block tree in fp64, per-block partial, last block folds in block order
  __shared__ double wsumd[4];
  __shared__ bool last;
  double e = warp_sum(energy_ret);
  if ((threadIdx.x & 31) == 0) wsumd[threadIdx.x >> 5] = e;
  __syncthreads();
  if (threadIdx.x == 0) {
    epart[blockIdx.x] = (wsumd[0] + wsumd[1]) + (wsumd[2] + wsumd[3]);
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence();
    double acc = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) acc += __ldcg(epart + b);
    acc = warp_sum(acc);
    if (threadIdx.x == 0) { *energy_out = acc; *ticket = 0; }
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
// K5: AccumulatedTopHessianSSE::addPoint<mode> — per-pair 13x13 relative accumulators.
// A block owns a chunk of ONE pair's residual segment; each thread accumulates its residuals into 91
// registers ([C(4);xi(6)]^2 upper = 55, [C;xi] x [a,b,r] = 30, [a,b,r]^2 upper = 6), then warp
// shuffles -> shared memory -> fixed-order cross-warp sum -> per-chunk partial.
// ------------------------------------------------------------------------------------------------
struct TopChunk { int pair, start, count, pad; };

template <int MODE>
__global__ void __launch_bounds__(TOP_THREADS) ba_top_kernel(BAWin W, BARes r, BAPts pts,
                                                             const TopChunk* __restrict__ chunks,
                                                             float* __restrict__ partial /*[nchunks][TOP_PART]*/) {
  const TopChunk ch = chunks[blockIdx.x];
  const int s = r.rs;
  float acc[91];
#pragma unroll
  for (int i = 0; i < 91; ++i) acc[i] = 0.f;
  int nres = 0;
  float dp[8];
  if (MODE == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) dp[j] = __ldg(W.adHTdeltaF + 8 * ch.pair + j);
  }
  for (int k = ch.start + threadIdx.x; k < ch.start + ch.count; k += TOP_THREADS) {
    bool lin = r.isLin[k] != 0, act = r.isActive[k] != 0;
    if (MODE == 0 ? (lin || !act) : (!lin || !act)) continue;
    const float* J = r.J + k;
    float x[10], y[10];
#pragma unroll
    for (int j = 0; j < 4; ++j) { x[j] = J[(J_C0 + j) * s]; y[j] = J[(J_C1 + j) * s]; }
#pragma unroll
    for (int j = 0; j < 6; ++j) { x[4 + j] = J[(J_XI0 + j) * s]; y[4 + j] = J[(J_XI1 + j) * s]; }
    const float d0 = J[J_DD * s], d1 = J[(J_DD + 1) * s];
    const float a = J[(J_IDX2 + 0) * s], b = J[(J_IDX2 + 2) * s], c = J[(J_IDX2 + 3) * s];
    const float AI00 = J[(J_ABIDX + 0) * s], AI10 = J[(J_ABIDX + 1) * s], AI01 = J[(J_ABIDX + 2) * s],
                AI11 = J[(J_ABIDX + 3) * s];
    const float AA00 = J[(J_AB2 + 0) * s], AA01 = J[(J_AB2 + 2) * s], AA11 = J[(J_AB2 + 3) * s];
    float JIr0, JIr1, Jabr0, Jabr1, rr;
    if (MODE == 0) {
      JIr0 = J[(J_JIR + 0) * s]; JIr1 = J[(J_JIR + 1) * s];
      Jabr0 = J[(J_JABR + 0) * s]; Jabr1 = J[(J_JABR + 1) * s]; rr = J[J_RR * s];
    } else {
      // resApprox = res_toZeroF + J*delta, AccumulatedTopHessian.cpp:63-81 (sequential sums)
      float dx = 0.f, dy = 0.f, cxs = 0.f, cys = 0.f;
#pragma unroll
      for (int j = 0; j < 6; ++j) { dx = xadd(dx, xmul(x[4 + j], dp[j])); dy = xadd(dy, xmul(y[4 + j], dp[j])); }
#pragma unroll
      for (int j = 0; j < 4; ++j) { cxs = xadd(cxs, xmul(x[j], W.cDeltaF[j])); cys = xadd(cys, xmul(y[j], W.cDeltaF[j])); }
      float dd = pts.deltaF[r.point[k]];
      float Jpx = xadd(xadd(dx, cxs), xmul(d0, dd)), Jpy = xadd(xadd(dy, cys), xmul(d1, dd));
      JIr0 = JIr1 = Jabr0 = Jabr1 = rr = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float i0 = J[(J_IDX0 + j) * s], i1 = J[(J_IDX1 + j) * s], a0 = J[(J_AB0 + j) * s], a1 = J[(J_AB1 + j) * s];
        float rz = r.rtz[j * s + k];
        rz = xadd(rz, xmul(i0, Jpx)); rz = xadd(rz, xmul(i1, Jpy));
        rz = xadd(rz, xmul(a0, dp[6])); rz = xadd(rz, xmul(a1, dp[7]));
        JIr0 = xadd(JIr0, xmul(rz, i0)); JIr1 = xadd(JIr1, xmul(rz, i1));
        Jabr0 = xadd(Jabr0, xmul(rz, a0)); Jabr1 = xadd(Jabr1, xmul(rz, a1));
        rr = xadd(rr, xmul(rz, rz));
      }
    }
    // 10x10 upper triangle: [x y] [[a b],[b c]] [x y]^T
    int idx = 0;
#pragma unroll
    for (int rI = 0; rI < 10; ++rI) {
      float px = fmaf(a, x[rI], b * y[rI]), py = fmaf(b, x[rI], c * y[rI]);
#pragma unroll
      for (int cI = rI; cI < 10; ++cI, ++idx) acc[idx] += fmaf(x[cI], px, y[cI] * py);
    }
    // 10x3 top-right: columns a, b, r  (MatrixAccumulators.h:859-901)
#pragma unroll
    for (int rI = 0; rI < 10; ++rI) {
      acc[55 + 3 * rI + 0] += fmaf(x[rI], AI00, y[rI] * AI01);
      acc[55 + 3 * rI + 1] += fmaf(x[rI], AI10, y[rI] * AI11);
      acc[55 + 3 * rI + 2] += fmaf(x[rI], JIr0, y[rI] * JIr1);
    }
    acc[85] += AA00; acc[86] += AA01; acc[87] += Jabr0; acc[88] += AA11; acc[89] += Jabr1; acc[90] += rr;
    nres++;
    // per-residual contribution to the point sums (:114-117), exact ops: they are summed per point in
    // the point's own order by ba_point_sum_kernel, which keeps Hdd/bd/Hcd bit-identical to the oracle.
    float Ji0 = xadd(xmul(a, d0), xmul(b, d1)), Ji1 = xadd(xmul(b, d0), xmul(c, d1));
    float* o = r.ptc[MODE] + k;
    o[0] = xadd(xmul(JIr0, d0), xmul(JIr1, d1));
    o[s] = xadd(xmul(Ji0, d0), xmul(Ji1, d1));
#pragma unroll
    for (int j = 0; j < 4; ++j) o[(2 + j) * s] = xadd(xmul(x[j], Ji0), xmul(y[j], Ji1));
  }
  __shared__ float red[TOP_THREADS / 32][TOP_PART];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 91; ++i) {
    float v = warp_sum(acc[i]);
    if (lane == 0) red[wid][i] = v;
  }
  int q = warp_sum(nres);
  if (lane == 0) red[wid][91] = __int_as_float(q);
  __syncthreads();
  if (threadIdx.x < 92) {
    int t = threadIdx.x;
    float* out = partial + (size_t)blockIdx.x * TOP_PART;
    if (t == 91) {
      int n = 0;
#pragma unroll
      for (int w = 0; w < TOP_THREADS / 32; ++w) n += __float_as_int(red[w][91]);
      out[91] = __int_as_float(n);
    } else {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < TOP_THREADS / 32; ++w) v += red[w][t];
      out[t] = v;
    }
  }
}

// Per-point sums of Hdd/bd/Hcd over the point's residuals in the point's own order
// (the tail of addPoint<mode>, AccumulatedTopHessian.cpp:122-139).  One thread per point.
__global__ void ba_point_sum_kernel(BARes r, BAPts pts, int do_A, int do_L) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pts.P) return;
  const int s = r.rs;
  float bdv[2] = {0.f, 0.f}, Hddv[2] = {0.f, 0.f}, Hcdv[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (int i = pts.res_begin[p]; i < pts.res_begin[p + 1]; ++i) {
    int k = pts.res_index[i];
    if (!r.isActive[k]) continue;
    int m = r.isLin[k] ? 1 : 0;
    if ((m == 0 && !do_A) || (m == 1 && !do_L)) continue;
    const float* o = r.ptc[m] + k;
    bdv[m] = xadd(bdv[m], o[0]);
    Hddv[m] = xadd(Hddv[m], o[s]);
#pragma unroll
    for (int j = 0; j < 4; ++j) Hcdv[m][j] = xadd(Hcdv[m][j], o[(2 + j) * s]);
  }
  if (do_A) { pts.Hdd[0][p] = Hddv[0]; pts.bd[0][p] = bdv[0]; pts.Hcd[0][p] = make_float4(Hcdv[0][0], Hcdv[0][1], Hcdv[0][2], Hcdv[0][3]); }
  if (do_L) { pts.Hdd[1][p] = Hddv[1]; pts.bd[1][p] = bdv[1]; pts.Hcd[1][p] = make_float4(Hcdv[1][0], Hcdv[1][1], Hcdv[1][2], Hcdv[1][3]); }
}

// ------------------------------------------------------------------------------------------------
// K6: AccumulatedSCHessianSSE::addPoint — the per-point Schur fold.
// For the points of one host the reference's accD[(h,t1,t2)], accE[(h,t1)], accEB[(h,t1)], accHcc,
// accbc are all blocks of ONE weighted Gram matrix  G_h = sum_p HdiF_p * a_p a_p^T  of the vector
// a_p = [JpJdF(p->t) for t = 0..n-1 (zero where there is no active residual) ; Hcd(4) ; bdSumF].
// A block stages SC_PTS such vectors in shared memory and every thread owns 4x4 register tiles of
// G; points are visited sequentially, so the sum order is fixed.  fp32 SIMT: the matrix is only
// (8n+5) wide — tensor cores are deliberately not used (north_star).
// ------------------------------------------------------------------------------------------------
struct PtChunk { int host, start, count, pad; };

__global__ void __launch_bounds__(SC_THREADS) ba_schur_kernel(BAWin W, BARes r, BAPts pts,
                                                              const PtChunk* __restrict__ chunks, int LD,
                                                              float* __restrict__ partial /*[nchunks][LD*LD]*/) {
  extern __shared__ __align__(16) float smem[];
  float* A = smem;                       // [SC_PTS][LD]
  float* wv = smem + SC_PTS * LD;        // [SC_PTS] HdiF
  const PtChunk ch = chunks[blockIdx.x];
  const int s = r.rs, n = W.n;
  const int NT = LD / 4, ntiles = NT * NT;
  // register tiles: tile id = tid + j*SC_THREADS
  constexpr int MAXT = 5;                // supports NT*NT <= 5*256 (n <= 16 -> LD=136 -> 1156 tiles)
  float acc[MAXT][16];
#pragma unroll
  for (int j = 0; j < MAXT; ++j)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[j][i] = 0.f;

  for (int base = 0; base < ch.count; base += SC_PTS) {
    int cnt = min(SC_PTS, ch.count - base);
    // zero the staging tile
    for (int i = threadIdx.x; i < SC_PTS * LD; i += SC_THREADS) A[i] = 0.f;
    __syncthreads();
    // phase 1: 4 threads per point: lane q handles residuals q, q+4, ... of the point's list
    {
      int lp = threadIdx.x >> 2, q = threadIdx.x & 3;
      int ngood = 0;
      const int p = ch.start + base + lp;
      if (lp < cnt) {
        int rb = pts.res_begin[p], re = pts.res_begin[p + 1];
        for (int i = rb + q; i < re; i += 4) {
          int k = pts.res_index[i];
          if (!r.isActive[k]) continue;
          ngood++;
          int t = r.pair[k] / n;
          float* dst = A + lp * LD + 8 * t;
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = r.JpJdF[j * s + k];
        }
      }
      ngood += __shfl_xor_sync(0xffffffffu, ngood, 1);   // all 32 lanes take part
      ngood += __shfl_xor_sync(0xffffffffu, ngood, 2);
      if (lp < cnt) {
        if (q == 0) {
          // AccumulatedSCHessian.cpp:12-33
          float hdi = 0.f, bds = 0.f, idh = 0.f;
          float4 hcd = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ngood > 0) {
            float H = xadd(xadd(pts.Hdd[0][p], pts.Hdd[1][p]), pts.priorF[p]);
            if (H < 1e-10f) H = 1e-10f;
            idh = H;
            hdi = xdiv(1.0f, H);
            bds = xadd(pts.bd[0][p], pts.bd[1][p]);
            bds = xadd(bds, xmul(pts.priorF[p], pts.deltaF[p]));  // shiftPriorToZero = true (EnergyFunctional.cpp:209)
            float4 ha = pts.Hcd[0][p], hl = pts.Hcd[1][p];
            hcd = make_float4(xadd(ha.x, hl.x), xadd(ha.y, hl.y), xadd(ha.z, hl.z), xadd(ha.w, hl.w));
          }
          pts.HdiF[p] = hdi; pts.bdSumF[p] = bds; pts.idepth_hessian[p] = idh;
          float* dst = A + lp * LD + 8 * n;
          dst[0] = hcd.x; dst[1] = hcd.y; dst[2] = hcd.z; dst[3] = hcd.w; dst[4] = bds;
          wv[lp] = hdi;
        }
      } else if (lp < SC_PTS && q == 0) {
        wv[lp] = 0.f;
      }
    }
    __syncthreads();
    // phase 2: G += sum_p (w_p a_p) a_p^T, 4x4 register tiles, points in order
#pragma unroll
    for (int j = 0; j < MAXT; ++j) {
      int tile = threadIdx.x + j * SC_THREADS;
      if (tile < ntiles) {
        int ti = tile / NT, tj = tile % NT;
        if (tj >= ti) {   // symmetric: only the upper-triangular tiles are computed
          for (int lp = 0; lp < cnt; ++lp) {
            float w = wv[lp];
            float4 av = *reinterpret_cast<const float4*>(A + lp * LD + 4 * ti);
            float4 bv = *reinterpret_cast<const float4*>(A + lp * LD + 4 * tj);
            float a4[4] = {w * av.x, w * av.y, w * av.z, w * av.w};
            float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int ii = 0; ii < 4; ++ii)
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) acc[j][4 * ii + jj] = fmaf(a4[ii], b4[jj], acc[j][4 * ii + jj]);
          }
        }
      }
    }
    __syncthreads();
  }
  float* out = partial + (size_t)blockIdx.x * LD * LD;
#pragma unroll
  for (int j = 0; j < MAXT; ++j) {
    int tile = threadIdx.x + j * SC_THREADS;
    if (tile < ntiles) {
      int ti = tile / NT, tj = tile % NT;
      if (tj >= ti) {
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
          *reinterpret_cast<float4*>(out + (size_t)(4 * ti + ii) * LD + 4 * tj) =
              make_float4(acc[j][4 * ii], acc[j][4 * ii + 1], acc[j][4 * ii + 2], acc[j][4 * ii + 3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fold of the per-chunk partials into fp64 relative accumulators, by chunk index (fixed order).
//   top : accTop[mode][pair][96]  (91 values + nres)
//   sc  : G[host][LD*LD] (upper-triangular tiles valid)
// ------------------------------------------------------------------------------------------------
__global__ void ba_fold_top_kernel(const float* __restrict__ partial, const int* __restrict__ pair_chunk_begin,
                                   int npairs, double* __restrict__ accTop, int* __restrict__ nres_out) {
  int pair = blockIdx.x, t = threadIdx.x;
  int b = pair_chunk_begin[pair], e = pair_chunk_begin[pair + 1];
  if (t < 91) {
    double sum = 0.0;
    for (int c = b; c < e; ++c) sum += (double)partial[(size_t)c * TOP_PART + t];
    accTop[(size_t)pair * TOP_PART + t] = sum;
  } else if (t == 91) {
    int n = 0;
    for (int c = b; c < e; ++c) n += __float_as_int(partial[(size_t)c * TOP_PART + 91]);
    accTop[(size_t)pair * TOP_PART + 91] = (double)n;
    atomicAdd(nres_out, n);   // integer: order-independent
  }
}

__global__ void ba_fold_sc_kernel(const float* __restrict__ partial, const int* __restrict__ host_chunk_begin, int LD,
                                  double* __restrict__ G) {
  int host = blockIdx.y;
  int b = host_chunk_begin[host], e = host_chunk_begin[host + 1];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= LD * LD) return;
  int row = i / LD, col = i % LD;
  if ((col >> 2) < (row >> 2)) return;   // lower tiles are not produced
  double sum = 0.0;
  for (int c = b; c < e; ++c) sum += (double)partial[(size_t)c * LD * LD + i];
  G[(size_t)host * LD * LD + i] = sum;
}

// ------------------------------------------------------------------------------------------------
// K7: stitch — relative accumulators -> (8n+4)^2 system in fp64, gather form.
// grid = (n+1, n+1, 3 systems); block (a,b): a,b < n are frame blocks, index n is the calibration
// block.  64 threads: thread = (row r, col c) of an 8x8 output block.
// ------------------------------------------------------------------------------------------------
struct StitchArgs {
  int n, LD;
  const double* adHost;      // n*n*64 col-major
  const double* adTarget;
  const double* accTop[2];   // [pair][96] fp64 (A, L)
  const double* G;           // [host][LD*LD]
  double* Hb;                // 3 x (N*N + N)
  double cPrior[4];
  float cDeltaF[4];
  const double* framePrior;  // n*8
  const double* frameDeltaPrior;
  int add_priors;
  int sys_begin, sys_end;    // systems (0=A, 1=L, 2=SC) this launch produces
};

// element (r,c) of the 13x13 relative accumulator from its packed 91-value form
__device__ __forceinline__ double acc13(const double* a, int r, int c) {
  if (r > c) { int t = r; r = c; c = t; }
  if (c < 10) return a[r * 10 - (r * (r - 1)) / 2 + (c - r)];   // row-major upper triangle of 10x10
  if (r < 10) return a[55 + 3 * r + (c - 10)];
  const int br[3][3] = {{85, 86, 87}, {86, 88, 89}, {87, 89, 90}};
  return a[br[r - 10][c - 10]];
}
__device__ __forceinline__ double Gsym(const double* G, int LD, int i, int j) {
  return ((j >> 2) >= (i >> 2)) ? G[(size_t)i * LD + j] : G[(size_t)j * LD + i];
}

// out[r][c] = sum_{x,y} A1[r,x] * M(x,y) * A2[c,y]  for one thread (r,c), two-stage through smem
template <class MF>
__device__ __forceinline__ double sandwich(const double* A1, const double* A2, MF M, double (*tmp)[9], int r, int c) {
  // stage 1: tmp[x][c] = sum_y M(x,y) * A2[c,y]   (thread (r,c) computes x=r)
  double t = 0.0;
#pragma unroll
  for (int y = 0; y < 8; ++y) t += M(r, y) * A2[c + 8 * y];
  __syncthreads();
  tmp[r][c] = t;
  __syncthreads();
  double o = 0.0;
#pragma unroll
  for (int x = 0; x < 8; ++x) o += A1[r + 8 * x] * tmp[x][c];
  return o;
}

__global__ void __launch_bounds__(64) ba_stitch_kernel(StitchArgs S) {
  const int n = S.n, N = 8 * n + 4, LD = S.LD;
  const int a = blockIdx.x, b = blockIdx.y, sys = blockIdx.z;
  if (sys < S.sys_begin || sys >= S.sys_end) return;
  const int r = threadIdx.x & 7, c = threadIdx.x >> 3;
  __shared__ double tmp[8][9];
  double* H = S.Hb + (size_t)sys * ((size_t)N * N + N);
  double* bv = H + (size_t)N * N;
  auto AdH = [&](int h, int t) { return S.adHost + 64 * (size_t)(h + t * n); };
  auto AdT = [&](int h, int t) { return S.adTarget + 64 * (size_t)(h + t * n); };

  if (sys < 2) {
    const double* acc = S.accTop[sys];
    auto ACC = [&](int h, int t) { return acc + (size_t)(h + t * n) * TOP_PART; };
    const bool usePrior = (sys == 1) && S.add_priors;
    if (a < n && b < n) {
      double o = 0.0;
      if (a == b) {
        // S[a,a] = sum_t AdH[a,t] M[a,t] AdH[a,t]^T + sum_h AdT[h,a] M[h,a] AdT[h,a]^T + AdH[a,a] M[a,a] AdT[a,a]^T
        for (int t = 0; t < n; ++t) {
          const double* m = ACC(a, t);
          if (m[91] == 0.0) continue;   // acc.num == 0 (AccumulatedTopHessian.cpp:243)
          o += sandwich(AdH(a, t), AdH(a, t), [&](int x, int y) { return acc13(m, 4 + x, 4 + y); }, tmp, r, c);
        }
        for (int h = 0; h < n; ++h) {
          const double* m = ACC(h, a);
          if (m[91] == 0.0) continue;
          o += sandwich(AdT(h, a), AdT(h, a), [&](int x, int y) { return acc13(m, 4 + x, 4 + y); }, tmp, r, c);
        }
        {
          const double* m = ACC(a, a);
          if (m[91] != 0.0)
            o += sandwich(AdH(a, a), AdT(a, a), [&](int x, int y) { return acc13(m, 4 + x, 4 + y); }, tmp, r, c);
        }
        if (usePrior && r == c) o += S.framePrior[8 * a + r];
        H[(size_t)(4 + 8 * a + r) + (size_t)N * (4 + 8 * a + c)] = o;
      } else if (a < b) {
        // H[a,b] = S[a,b] + S[b,a]^T ; H[b,a] = H[a,b]^T   (AccumulatedTopHessian.h:105-116)
        const double* m1 = ACC(a, b);
        if (m1[91] != 0.0)
          o += sandwich(AdH(a, b), AdT(a, b), [&](int x, int y) { return acc13(m1, 4 + x, 4 + y); }, tmp, r, c);
        const double* m2 = ACC(b, a);
        if (m2[91] != 0.0)   // (AdH[b,a] M AdT[b,a]^T)^T = AdT[b,a] M^T AdH[b,a]^T, M symmetric
          o += sandwich(AdT(b, a), AdH(b, a), [&](int x, int y) { return acc13(m2, 4 + y, 4 + x); }, tmp, r, c);
        H[(size_t)(4 + 8 * a + r) + (size_t)N * (4 + 8 * b + c)] = o;
        H[(size_t)(4 + 8 * b + c) + (size_t)N * (4 + 8 * a + r)] = o;
      }
    } else if (a < n && b == n) {
      // H[a, 0:4] (8x4), its transpose, and b[a]
      if (c < 5) {
        double o = 0.0;
        for (int t = 0; t < n; ++t) {
          const double* m = ACC(a, t);
          if (m[91] == 0.0) continue;
          const double* A1 = AdH(a, t);
          for (int x = 0; x < 8; ++x) o += A1[r + 8 * x] * (c < 4 ? acc13(m, 4 + x, c) : acc13(m, 4 + x, 12));
        }
        for (int h = 0; h < n; ++h) {
          const double* m = ACC(h, a);
          if (m[91] == 0.0) continue;
          const double* A1 = AdT(h, a);
          for (int x = 0; x < 8; ++x) o += A1[r + 8 * x] * (c < 4 ? acc13(m, 4 + x, c) : acc13(m, 4 + x, 12));
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
      if (r < 4 && c < 5) {
        double o = 0.0;
        for (int k = 0; k < n * n; ++k) {
          const double* m = acc + (size_t)k * TOP_PART;
          if (m[91] == 0.0) continue;
          o += (c < 4) ? acc13(m, r, c) : acc13(m, r, 12);
        }
        if (c < 4) {
          if (usePrior && r == c) o += S.cPrior[r];
          H[(size_t)r + (size_t)N * c] = o;
        } else {
          if (usePrior) o += S.cPrior[r] * (double)S.cDeltaF[r];
          bv[r] = o;
        }
      }
    }
    return;
  }

  // ---- Schur complement system: H_sc = sum_i T_i G_i T_i^T with T_i[a,j] = d(a,i) AdH[i,j] + d(a,j) AdT[i,j]
  auto Gh = [&](int i) { return S.G + (size_t)i * LD * LD; };
  if (a < n && b < n) {
    double o = 0.0;
    if (a == b) {   // term 1: sum_{j,k} AdH[a,j] D_a[j,k] AdH[a,k]^T   (AccumulatedSCHessian.cpp:108)
      const double* G = Gh(a);
      for (int j = 0; j < n; ++j)
        for (int k = 0; k < n; ++k)
          o += sandwich(AdH(a, j), AdH(a, k), [&](int x, int y) { return Gsym(G, LD, 8 * j + x, 8 * k + y); }, tmp, r, c);
    }
    {   // term 4 in my ordering = :111  H[i,k] += AdH[ij] D AdT[ik]^T with i=a, k=b
      const double* G = Gh(a);
      for (int j = 0; j < n; ++j)
        o += sandwich(AdH(a, j), AdT(a, b), [&](int x, int y) { return Gsym(G, LD, 8 * j + x, 8 * b + y); }, tmp, r, c);
    }
    {   // :110  H[j,i] += AdT[ij] D AdH[ik]^T with j=a, i=b
      const double* G = Gh(b);
      for (int k = 0; k < n; ++k)
        o += sandwich(AdT(b, a), AdH(b, k), [&](int x, int y) { return Gsym(G, LD, 8 * a + x, 8 * k + y); }, tmp, r, c);
    }
    for (int i = 0; i < n; ++i) {   // :109  H[j,k] += AdT[ij] D AdT[ik]^T with j=a, k=b
      const double* G = Gh(i);
      o += sandwich(AdT(i, a), AdT(i, b), [&](int x, int y) { return Gsym(G, LD, 8 * a + x, 8 * b + y); }, tmp, r, c);
    }
    H[(size_t)(4 + 8 * a + r) + (size_t)N * (4 + 8 * b + c)] = o;
  } else if (a < n && b == n) {
    if (c < 5) {   // :86-89: H[i,0:4] += AdH[ij] E[ij], H[j,0:4] += AdT[ij] E[ij]; b likewise with EB
      double o = 0.0;
      const int col = 8 * n + c;   // Hcd columns 8n..8n+3, bdSum column 8n+4
      {
        const double* G = Gh(a);
        for (int j = 0; j < n; ++j) {
          const double* A1 = AdH(a, j);
          for (int x = 0; x < 8; ++x) o += A1[r + 8 * x] * Gsym(G, LD, 8 * j + x, col);
        }
      }
      for (int i = 0; i < n; ++i) {
        const double* G = Gh(i);
        const double* A1 = AdT(i, a);
        for (int x = 0; x < 8; ++x) o += A1[r + 8 * x] * Gsym(G, LD, 8 * a + x, col);
      }
      if (c < 4) {
        H[(size_t)(4 + 8 * a + r) + (size_t)N * c] = o;
        H[(size_t)c + (size_t)N * (4 + 8 * a + r)] = o;   // .h:106-110
      } else {
        bv[4 + 8 * a + r] = o;
      }
    }
  } else if (a == n && b == n) {
    if (r < 4 && c < 5) {
      double o = 0.0;
      for (int i = 0; i < n; ++i) o += Gsym(Gh(i), LD, 8 * n + r, 8 * n + c);
      if (c < 4) H[(size_t)r + (size_t)N * c] = o; else bv[r] = o;
    }
  }
}

// EnergyFunctional::resubstituteFPt (EnergyFunctional.cpp:249-274); xAd[n*h+t] uploaded by the host
__global__ void ba_resub_kernel(BAWin W, BARes r, BAPts pts, const float* __restrict__ xc, const float* __restrict__ xAd) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pts.P) return;
  const int s = r.rs, n = W.n;
  int ngood = 0;
  for (int i = pts.res_begin[p]; i < pts.res_begin[p + 1]; ++i) if (r.isActive[pts.res_index[i]]) ngood++;
  if (ngood == 0) { pts.step[p] = 0.f; return; }
  float b = pts.bdSumF[p];
  float4 ha = pts.Hcd[0][p], hl = pts.Hcd[1][p];
  float dot = 0.f;
  dot = xadd(dot, xmul(xc[0], xadd(ha.x, hl.x))); dot = xadd(dot, xmul(xc[1], xadd(ha.y, hl.y)));
  dot = xadd(dot, xmul(xc[2], xadd(ha.z, hl.z))); dot = xadd(dot, xmul(xc[3], xadd(ha.w, hl.w)));
  b = xsub(b, dot);
  for (int i = pts.res_begin[p]; i < pts.res_begin[p + 1]; ++i) {
    int k = pts.res_index[i];
    if (!r.isActive[k]) continue;
    int pair = r.pair[k], h = pair % n, t = pair / n;
    const float* xa = xAd + 8 * (n * h + t);
    float d = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) d = xadd(d, xmul(xa[j], r.JpJdF[j * s + k]));
    b = xsub(b, d);
  }
  pts.step[p] = xmul(-b, pts.HdiF[p]);
}

}  // namespace hs

using namespace hs;

// ====================================================================================== host side
template <class T> struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = n + n / 8 + 64;
    cudaError_t e = cudaMalloc(&p, cap * sizeof(T));
    if (e != cudaSuccess) { hs::set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e)); cap = 0; return HSLAM_ERR_NOMEM; }
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct hslam_ba {
  hslam_ctx* ctx = nullptr;
  int n = 0, P = 0, R = 0, rs = 0, LD = 0, N = 0;
  bool have_window = false, have_topology = false;
  int add_priors = 1;
  BAWin win;
  // window tables
  DevBuf<float> d_precalc, d_adHTd;
  DevBuf<double> d_adH, d_adT, d_fprior, d_fdprior;
  double cPrior[4];
  std::vector<double> h_adH, h_adT;
  // points
  DevBuf<float4> d_uvid, d_color, d_weights, d_Hcd[2];
  DevBuf<float> d_priorF, d_deltaF, d_HdiF, d_bdSumF, d_idh, d_Hdd[2], d_bd[2], d_step;
  DevBuf<int> d_phost, d_res_begin, d_res_index;
  std::vector<int> p_sorted2orig, p_orig2sorted;
  std::vector<float4> h_uvid;
  // residuals
  DevBuf<int> d_rpoint, d_rpair;
  DevBuf<int8_t> d_state, d_newState;
  DevBuf<uint8_t> d_isLin, d_isActive;
  DevBuf<float> d_senergy, d_newEnergy, d_newEnergyWO, d_J, d_JpJdF, d_rtz, d_cpt, d_pto, d_ptc[2];
  std::vector<int> r_sorted2orig, r_orig2sorted;
  int n_linearized = 0;
  // chunks / partials
  DevBuf<TopChunk> d_topchunks;
  DevBuf<int> d_pair_chunk_begin;
  DevBuf<float> d_toppartial;
  int n_topchunks = 0;
  DevBuf<PtChunk> d_ptchunks;
  DevBuf<int> d_host_chunk_begin;
  DevBuf<float> d_scpartial;
  int n_ptchunks = 0;
  DevBuf<double> d_accTop[2], d_G, d_Hb, d_epart, d_energy;
  DevBuf<int> d_nres;          // [2]
  DevBuf<unsigned int> d_ticket;
  DevBuf<float> d_xc, d_xAd;
  DevBuf<int> d_ids;
  // pinned staging
  void* h_pin = nullptr;
  size_t h_pin_cap = 0;
  BARes res;
  BAPts pts;
};

static int pin_ensure(hslam_ba* ba, size_t bytes) {
  if (bytes <= ba->h_pin_cap) return 0;
  if (ba->h_pin) cudaFreeHost(ba->h_pin);
  ba->h_pin_cap = bytes + bytes / 4 + 4096;
  HS_CUDA(cudaMallocHost(&ba->h_pin, ba->h_pin_cap));
  return 0;
}

// Upload a host array through pinned staging (synchronous w.r.t. the staging buffer).
template <class T> static int upload(hslam_ba* ba, T* dst, const T* src, size_t n) {
  if (!n) return 0;
  int rc = pin_ensure(ba, n * sizeof(T));
  if (rc) return rc;
  memcpy(ba->h_pin, src, n * sizeof(T));
  HS_CUDA(cudaMemcpyAsync(dst, ba->h_pin, n * sizeof(T), cudaMemcpyHostToDevice, ba->ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ba->ctx->stream));
  return 0;
}
template <class T> static int download(hslam_ba* ba, T* dst_host, const T* src_dev, size_t n) {
  if (!n) return 0;
  int rc = pin_ensure(ba, n * sizeof(T));
  if (rc) return rc;
  HS_CUDA(cudaMemcpyAsync(ba->h_pin, src_dev, n * sizeof(T), cudaMemcpyDeviceToHost, ba->ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ba->ctx->stream));
  memcpy(dst_host, ba->h_pin, n * sizeof(T));
  return 0;
}

#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

static void fill_views(hslam_ba* ba) {
  BARes& r = ba->res;
  r.R = ba->R; r.rs = ba->rs;
  r.point = ba->d_rpoint.p; r.pair = ba->d_rpair.p;
  r.state_state = ba->d_state.p; r.newState = ba->d_newState.p; r.isLin = ba->d_isLin.p; r.isActive = ba->d_isActive.p;
  r.state_energy = ba->d_senergy.p; r.newEnergy = ba->d_newEnergy.p; r.newEnergyWO = ba->d_newEnergyWO.p;
  r.J = ba->d_J.p; r.JpJdF = ba->d_JpJdF.p; r.rtz = ba->d_rtz.p; r.cpt = ba->d_cpt.p; r.pto = ba->d_pto.p;
  r.ptc[0] = ba->d_ptc[0].p; r.ptc[1] = ba->d_ptc[1].p;
  BAPts& p = ba->pts;
  p.P = ba->P; p.uvid = ba->d_uvid.p; p.color = ba->d_color.p; p.weights = ba->d_weights.p;
  p.priorF = ba->d_priorF.p; p.deltaF = ba->d_deltaF.p; p.host = ba->d_phost.p;
  p.res_begin = ba->d_res_begin.p; p.res_index = ba->d_res_index.p;
  p.HdiF = ba->d_HdiF.p; p.bdSumF = ba->d_bdSumF.p; p.idepth_hessian = ba->d_idh.p;
  for (int m = 0; m < 2; ++m) { p.Hdd[m] = ba->d_Hdd[m].p; p.bd[m] = ba->d_bd[m].p; p.Hcd[m] = ba->d_Hcd[m].p; }
  p.step = ba->d_step.p;
}

extern "C" {

int hslam_ba_create(hslam_ctx* ctx, hslam_ba** out) {
  HS_ARG(ctx && out, "hslam_ba_create: null argument");
  HS_CUDA(cudaSetDevice(ctx->device));
  hslam_ba* ba = new hslam_ba();
  ba->ctx = ctx;
  memset(&ba->win, 0, sizeof(ba->win));
  memset(&ba->res, 0, sizeof(ba->res));
  memset(&ba->pts, 0, sizeof(ba->pts));
  TRY(ba->d_ticket.ensure(4));
  TRY(ba->d_nres.ensure(4));
  TRY(ba->d_energy.ensure(2));
  HS_CUDA(cudaMemsetAsync(ba->d_ticket.p, 0, 4 * sizeof(unsigned int), ctx->stream));
  *out = ba;
  return 0;
}

int hslam_ba_destroy(hslam_ba* ba) {
  if (!ba) return 0;
  cudaSetDevice(ba->ctx->device);
  cudaStreamSynchronize(ba->ctx->stream);
  ba->d_precalc.release(); ba->d_adHTd.release(); ba->d_adH.release(); ba->d_adT.release();
  ba->d_fprior.release(); ba->d_fdprior.release();
  ba->d_uvid.release(); ba->d_color.release(); ba->d_weights.release();
  ba->d_priorF.release(); ba->d_deltaF.release(); ba->d_HdiF.release(); ba->d_bdSumF.release(); ba->d_idh.release();
  for (int m = 0; m < 2; ++m) { ba->d_Hdd[m].release(); ba->d_bd[m].release(); ba->d_Hcd[m].release(); ba->d_ptc[m].release(); ba->d_accTop[m].release(); }
  ba->d_step.release(); ba->d_phost.release(); ba->d_res_begin.release(); ba->d_res_index.release();
  ba->d_rpoint.release(); ba->d_rpair.release(); ba->d_state.release(); ba->d_newState.release();
  ba->d_isLin.release(); ba->d_isActive.release(); ba->d_senergy.release(); ba->d_newEnergy.release();
  ba->d_newEnergyWO.release(); ba->d_J.release(); ba->d_JpJdF.release(); ba->d_rtz.release(); ba->d_cpt.release();
  ba->d_pto.release(); ba->d_topchunks.release(); ba->d_pair_chunk_begin.release(); ba->d_toppartial.release();
  ba->d_ptchunks.release(); ba->d_host_chunk_begin.release(); ba->d_scpartial.release(); ba->d_G.release();
  ba->d_Hb.release(); ba->d_epart.release(); ba->d_energy.release(); ba->d_nres.release(); ba->d_ticket.release();
  ba->d_xc.release(); ba->d_xAd.release(); ba->d_ids.release();
  if (ba->h_pin) cudaFreeHost(ba->h_pin);
  delete ba;
  return 0;
}

int hslam_ba_set_add_priors(hslam_ba* ba, int add) {
  HS_ARG(ba, "hslam_ba_set_add_priors: null");
  ba->add_priors = add ? 1 : 0;
  return 0;
}

int hslam_ba_set_window(hslam_ba* ba, const hslam_window* w) {
  HS_ARG(ba && w, "hslam_ba_set_window: null argument");
  HS_ARG(w->n_frames >= 1 && w->n_frames <= HSLAM_MAX_FRAMES, "hslam_ba_set_window: n_frames out of range");
  HS_ARG(w->frames && w->frameEnergyTH && w->precalc && w->adHost && w->adTarget && w->adHTdeltaF && w->framePrior &&
             w->frameDeltaPrior, "hslam_ba_set_window: null table");
  HS_ARG(!ba->have_topology || w->n_frames == ba->n, "hslam_ba_set_window: n_frames changed, call set_topology again");
  hslam_ctx* ctx = ba->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  const int n = w->n_frames;
  ba->n = n; ba->N = 8 * n + 4; ba->LD = ((8 * n + 5) + 3) & ~3;
  BAWin& W = ba->win;
  W.n = n; W.w = w->w; W.h = w->h;
  W.fx = w->fx; W.fy = w->fy; W.cx = w->cx; W.cy = w->cy; W.fxi = w->fxi; W.fyi = w->fyi;
  W.huber = ctx->prm.huberTH; W.outlierTH = ctx->prm.outlierTHSumComponent; W.scale_idepth = ctx->prm.scale_idepth;
  W.scale_f = ctx->prm.scale_f; W.scale_c = ctx->prm.scale_c; W.affA = ctx->prm.affineOptModeA; W.affB = ctx->prm.affineOptModeB;
  for (int i = 0; i < 4; ++i) { W.cDeltaF[i] = w->cDeltaF[i]; ba->cPrior[i] = w->cPrior[i]; }
  for (int f = 0; f < n; ++f) {
    HS_ARG(w->frames[f] && w->frames[f]->ctx == ctx, "hslam_ba_set_window: frame missing or from another context");
    HS_ARG(w->frames[f]->w == w->w && w->frames[f]->h == w->h, "hslam_ba_set_window: frame size mismatch");
    W.img[f] = w->frames[f]->dI[0];
    W.frameEnergyTH[f] = w->frameEnergyTH[f];
  }
  // one pinned staging block, one stream-ordered set of copies
  size_t nn = (size_t)n * n;
  TRY(ba->d_precalc.ensure(nn * HSLAM_PRECALC_STRIDE)); TRY(ba->d_adHTd.ensure(nn * 8));
  TRY(ba->d_adH.ensure(nn * 64)); TRY(ba->d_adT.ensure(nn * 64));
  TRY(ba->d_fprior.ensure(n * 8)); TRY(ba->d_fdprior.ensure(n * 8));
  size_t bytes = nn * HSLAM_PRECALC_STRIDE * 4 + nn * 8 * 4 + nn * 64 * 8 * 2 + (size_t)n * 8 * 8 * 2;
  TRY(pin_ensure(ba, bytes));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  char* s = (char*)ba->h_pin;
  auto put = [&](void* dst, const void* src, size_t b) -> int {
    memcpy(s, src, b);
    cudaError_t e = cudaMemcpyAsync(dst, s, b, cudaMemcpyHostToDevice, ctx->stream);
    s += (b + 15) & ~(size_t)15;
    return e == cudaSuccess ? 0 : -1;
  };
  int bad = 0;
  bad |= put(ba->d_adH.p, w->adHost, nn * 64 * 8);
  bad |= put(ba->d_adT.p, w->adTarget, nn * 64 * 8);
  bad |= put(ba->d_fprior.p, w->framePrior, (size_t)n * 64);
  bad |= put(ba->d_fdprior.p, w->frameDeltaPrior, (size_t)n * 64);
  bad |= put(ba->d_precalc.p, w->precalc, nn * HSLAM_PRECALC_STRIDE * 4);
  bad |= put(ba->d_adHTd.p, w->adHTdeltaF, nn * 8 * 4);
  if (bad) { hs::set_error("hslam_ba_set_window: H2D copy failed"); return HSLAM_ERR_CUDA; }
  HS_CUDA(cudaStreamSynchronize(ctx->stream));   // staging buffer is reused by later uploads
  ba->h_adH.assign(w->adHost, w->adHost + nn * 64);
  ba->h_adT.assign(w->adTarget, w->adTarget + nn * 64);
  W.precalc = ba->d_precalc.p; W.adHTdeltaF = ba->d_adHTd.p;
  TRY(ba->d_Hb.ensure(3 * ((size_t)ba->N * ba->N + ba->N)));
  TRY(ba->d_accTop[0].ensure(nn * TOP_PART)); TRY(ba->d_accTop[1].ensure(nn * TOP_PART));
  TRY(ba->d_G.ensure((size_t)n * ba->LD * ba->LD));
  ba->have_window = true;
  return 0;
}

int hslam_ba_set_topology(hslam_ba* ba, const hslam_point_soa* pt, const hslam_residual_soa* rs) {
  HS_ARG(ba && pt && rs, "hslam_ba_set_topology: null argument");
  HS_ARG(ba->have_window, "hslam_ba_set_topology: call hslam_ba_set_window first");
  const int P = pt->n_points, R = rs->n_residuals, n = ba->n;
  HS_ARG(P >= 0 && R >= 0, "hslam_ba_set_topology: negative count");
  HS_ARG(P == 0 || (pt->u && pt->v && pt->idepth && pt->idepth_zero && pt->color && pt->weights && pt->priorF &&
                    pt->deltaF && pt->host_idx && pt->res_begin), "hslam_ba_set_topology: null point array");
  HS_ARG(R == 0 || (pt->res_index && rs->point_idx && rs->target_idx && rs->state_state && rs->is_linearized &&
                    rs->is_active), "hslam_ba_set_topology: null residual array");
  hslam_ctx* ctx = ba->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  HS_ARG(P == 0 || pt->res_begin[P] == R, "hslam_ba_set_topology: CSR does not cover every residual exactly once");
  // ---- points sorted by host (stable counting sort)
  std::vector<int> hostCount(n + 1, 0);
  for (int p = 0; p < P; ++p) {
    HS_ARG(pt->host_idx[p] >= 0 && pt->host_idx[p] < n, "hslam_ba_set_topology: host_idx out of range");
    hostCount[pt->host_idx[p] + 1]++;
  }
  for (int h = 0; h < n; ++h) hostCount[h + 1] += hostCount[h];
  std::vector<int> hostBegin(hostCount);
  ba->p_sorted2orig.assign(P, 0); ba->p_orig2sorted.assign(P, 0);
  {
    std::vector<int> cur(hostCount.begin(), hostCount.end() - 1);
    for (int p = 0; p < P; ++p) { int sidx = cur[pt->host_idx[p]]++; ba->p_sorted2orig[sidx] = p; ba->p_orig2sorted[p] = sidx; }
  }
  // ---- residuals: walk points in sorted order, then stable counting sort by pair
  std::vector<int> walk; walk.reserve(R);
  std::vector<char> seen(R, 0);
  for (int sp = 0; sp < P; ++sp) {
    int p = ba->p_sorted2orig[sp];
    for (int i = pt->res_begin[p]; i < pt->res_begin[p + 1]; ++i) {
      int r = pt->res_index[i];
      HS_ARG(r >= 0 && r < R && !seen[r], "hslam_ba_set_topology: residual index invalid or listed twice");
      HS_ARG(rs->point_idx[r] == p, "hslam_ba_set_topology: residual.point_idx disagrees with the point's list");
      HS_ARG(rs->target_idx[r] >= 0 && rs->target_idx[r] < n, "hslam_ba_set_topology: target_idx out of range");
      seen[r] = 1;
      walk.push_back(r);
    }
  }
  HS_ARG((int)walk.size() == R, "hslam_ba_set_topology: some residuals belong to no point");
  std::vector<int> pairBegin(n * n + 1, 0);
  auto pair_of = [&](int r) { return pt->host_idx[rs->point_idx[r]] + rs->target_idx[r] * n; };
  for (int r : walk) pairBegin[pair_of(r) + 1]++;
  for (int k = 0; k < n * n; ++k) pairBegin[k + 1] += pairBegin[k];
  ba->r_sorted2orig.assign(R, 0); ba->r_orig2sorted.assign(R, 0);
  {
    std::vector<int> cur(pairBegin.begin(), pairBegin.end() - 1);
    for (int r : walk) { int k = cur[pair_of(r)]++; ba->r_sorted2orig[k] = r; ba->r_orig2sorted[r] = k; }
  }
  ba->P = P; ba->R = R; ba->rs = (R + 31) & ~31;
  if (ba->rs == 0) ba->rs = 32;
  const size_t rsz = ba->rs;
  // ---- allocate
  TRY(ba->d_uvid.ensure(P + 1)); TRY(ba->d_color.ensure(2 * (size_t)P + 2)); TRY(ba->d_weights.ensure(2 * (size_t)P + 2));
  TRY(ba->d_priorF.ensure(P + 1)); TRY(ba->d_deltaF.ensure(P + 1)); TRY(ba->d_phost.ensure(P + 1));
  TRY(ba->d_res_begin.ensure(P + 2)); TRY(ba->d_res_index.ensure(R + 1));
  TRY(ba->d_HdiF.ensure(P + 1)); TRY(ba->d_bdSumF.ensure(P + 1)); TRY(ba->d_idh.ensure(P + 1)); TRY(ba->d_step.ensure(P + 1));
  for (int m = 0; m < 2; ++m) { TRY(ba->d_Hdd[m].ensure(P + 1)); TRY(ba->d_bd[m].ensure(P + 1)); TRY(ba->d_Hcd[m].ensure(P + 1)); TRY(ba->d_ptc[m].ensure(6 * rsz)); }
  TRY(ba->d_rpoint.ensure(rsz)); TRY(ba->d_rpair.ensure(rsz)); TRY(ba->d_state.ensure(rsz)); TRY(ba->d_newState.ensure(rsz));
  TRY(ba->d_isLin.ensure(rsz)); TRY(ba->d_isActive.ensure(rsz)); TRY(ba->d_senergy.ensure(rsz)); TRY(ba->d_newEnergy.ensure(rsz));
  TRY(ba->d_newEnergyWO.ensure(rsz)); TRY(ba->d_J.ensure((size_t)JS * rsz)); TRY(ba->d_JpJdF.ensure(8 * rsz));
  TRY(ba->d_rtz.ensure(8 * rsz)); TRY(ba->d_cpt.ensure(3 * rsz)); TRY(ba->d_pto.ensure(16 * rsz));
  // ---- pack + upload points
  {
    std::vector<float4> uvid(P), col(2 * (size_t)P), wgt(2 * (size_t)P);
    std::vector<float> prior(P), delta(P);
    std::vector<int> host(P), rb(P + 1), ri(R);
    int pos = 0;
    for (int sp = 0; sp < P; ++sp) {
      int p = ba->p_sorted2orig[sp];
      uvid[sp] = make_float4(pt->u[p], pt->v[p], pt->idepth[p], pt->idepth_zero[p]);
      const float* c = pt->color + 8 * (size_t)p; const float* w = pt->weights + 8 * (size_t)p;
      col[2 * sp] = make_float4(c[0], c[1], c[2], c[3]); col[2 * sp + 1] = make_float4(c[4], c[5], c[6], c[7]);
      wgt[2 * sp] = make_float4(w[0], w[1], w[2], w[3]); wgt[2 * sp + 1] = make_float4(w[4], w[5], w[6], w[7]);
      prior[sp] = pt->priorF[p]; delta[sp] = pt->deltaF[p]; host[sp] = pt->host_idx[p];
      rb[sp] = pos;
      for (int i = pt->res_begin[p]; i < pt->res_begin[p + 1]; ++i) ri[pos++] = ba->r_orig2sorted[pt->res_index[i]];
    }
    rb[P] = pos;
    ba->h_uvid = uvid;
    TRY(upload(ba, ba->d_uvid.p, uvid.data(), P)); TRY(upload(ba, ba->d_color.p, col.data(), 2 * (size_t)P));
    TRY(upload(ba, ba->d_weights.p, wgt.data(), 2 * (size_t)P)); TRY(upload(ba, ba->d_priorF.p, prior.data(), P));
    TRY(upload(ba, ba->d_deltaF.p, delta.data(), P)); TRY(upload(ba, ba->d_phost.p, host.data(), P));
    TRY(upload(ba, ba->d_res_begin.p, rb.data(), P + 1)); TRY(upload(ba, ba->d_res_index.p, ri.data(), R));
  }
  // ---- pack + upload residuals
  {
    std::vector<int> rp(R), rpair(R);
    std::vector<int8_t> st(R);
    std::vector<uint8_t> lin(R), act(R);
    std::vector<float> se(R);
    ba->n_linearized = 0;
    for (int k = 0; k < R; ++k) {
      int r = ba->r_sorted2orig[k];
      rp[k] = ba->p_orig2sorted[rs->point_idx[r]];
      rpair[k] = pair_of(r);
      st[k] = rs->state_state[r]; lin[k] = rs->is_linearized[r] ? 1 : 0; act[k] = rs->is_active[r] ? 1 : 0;
      se[k] = rs->state_energy ? (float)rs->state_energy[r] : 0.f;
      ba->n_linearized += lin[k];
    }
    TRY(upload(ba, ba->d_rpoint.p, rp.data(), R)); TRY(upload(ba, ba->d_rpair.p, rpair.data(), R));
    TRY(upload(ba, ba->d_state.p, st.data(), R)); TRY(upload(ba, ba->d_isLin.p, lin.data(), R));
    TRY(upload(ba, ba->d_isActive.p, act.data(), R)); TRY(upload(ba, ba->d_senergy.p, se.data(), R));
  }
  HS_CUDA(cudaMemsetAsync(ba->d_J.p, 0, (size_t)JS * rsz * sizeof(float), ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_JpJdF.p, 0, 8 * rsz * sizeof(float), ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_rtz.p, 0, 8 * rsz * sizeof(float), ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_cpt.p, 0, 3 * rsz * sizeof(float), ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_pto.p, 0, 16 * rsz * sizeof(float), ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_newState.p, HSLAM_RES_OUT, rsz, ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_newEnergy.p, 0, rsz * sizeof(float), ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_newEnergyWO.p, 0, rsz * sizeof(float), ctx->stream));
  for (int m = 0; m < 2; ++m) {
    HS_CUDA(cudaMemsetAsync(ba->d_ptc[m].p, 0, 6 * rsz * sizeof(float), ctx->stream));
    HS_CUDA(cudaMemsetAsync(ba->d_Hdd[m].p, 0, (P + 1) * sizeof(float), ctx->stream));
    HS_CUDA(cudaMemsetAsync(ba->d_bd[m].p, 0, (P + 1) * sizeof(float), ctx->stream));
    HS_CUDA(cudaMemsetAsync(ba->d_Hcd[m].p, 0, (P + 1) * sizeof(float4), ctx->stream));
  }
  HS_CUDA(cudaMemsetAsync(ba->d_HdiF.p, 0, (P + 1) * sizeof(float), ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_bdSumF.p, 0, (P + 1) * sizeof(float), ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_idh.p, 0, (P + 1) * sizeof(float), ctx->stream));
  // ---- chunk tables.  Target ~2 waves of blocks over the SMs, chunk a multiple of the block size.
  {
    int target_blocks = 2 * ctx->sm_count;
    int chunk = ((R + target_blocks - 1) / target_blocks + TOP_THREADS - 1) / TOP_THREADS * TOP_THREADS;
    if (chunk < TOP_THREADS) chunk = TOP_THREADS;
    std::vector<TopChunk> ch;
    std::vector<int> pcb(n * n + 1, 0);
    for (int k = 0; k < n * n; ++k) {
      pcb[k] = (int)ch.size();
      for (int s0 = pairBegin[k]; s0 < pairBegin[k + 1]; s0 += chunk)
        ch.push_back(TopChunk{k, s0, std::min(chunk, pairBegin[k + 1] - s0), 0});
    }
    pcb[n * n] = (int)ch.size();
    ba->n_topchunks = (int)ch.size();
    TRY(ba->d_topchunks.ensure(ch.size() + 1)); TRY(ba->d_pair_chunk_begin.ensure(pcb.size()));
    TRY(ba->d_toppartial.ensure((ch.size() + 1) * TOP_PART));
    TRY(upload(ba, ba->d_topchunks.p, ch.data(), ch.size())); TRY(upload(ba, ba->d_pair_chunk_begin.p, pcb.data(), pcb.size()));
  }
  {
    int target_blocks = 2 * ctx->sm_count;
    int chunk = ((P + target_blocks - 1) / target_blocks + SC_PTS - 1) / SC_PTS * SC_PTS;
    if (chunk < SC_PTS) chunk = SC_PTS;
    std::vector<PtChunk> ch;
    std::vector<int> hcb(n + 1, 0);
    for (int h = 0; h < n; ++h) {
      hcb[h] = (int)ch.size();
      for (int s0 = hostBegin[h]; s0 < hostBegin[h + 1]; s0 += chunk)
        ch.push_back(PtChunk{h, s0, std::min(chunk, hostBegin[h + 1] - s0), 0});
    }
    hcb[n] = (int)ch.size();
    ba->n_ptchunks = (int)ch.size();
    TRY(ba->d_ptchunks.ensure(ch.size() + 1)); TRY(ba->d_host_chunk_begin.ensure(hcb.size()));
    TRY(ba->d_scpartial.ensure((ch.size() + 1) * (size_t)ba->LD * ba->LD));
    TRY(upload(ba, ba->d_ptchunks.p, ch.data(), ch.size())); TRY(upload(ba, ba->d_host_chunk_begin.p, hcb.data(), hcb.size()));
  }
  TRY(ba->d_epart.ensure((size_t)(ba->rs / 128 + 2)));
  fill_views(ba);
  ba->have_topology = true;
  return 0;
}

int hslam_ba_update_points(hslam_ba* ba, const float* idepth, const float* deltaF, const float* idepth_zero) {
  HS_ARG(ba && ba->have_topology, "hslam_ba_update_points: no topology");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  const int P = ba->P;
  if (idepth || idepth_zero) {
    for (int sp = 0; sp < P; ++sp) {
      int p = ba->p_sorted2orig[sp];
      if (idepth) ba->h_uvid[sp].z = idepth[p];
      if (idepth_zero) ba->h_uvid[sp].w = idepth_zero[p];
    }
    TRY(upload(ba, ba->d_uvid.p, ba->h_uvid.data(), P));
  }
  if (deltaF) {
    std::vector<float> d(P);
    for (int sp = 0; sp < P; ++sp) d[sp] = deltaF[ba->p_sorted2orig[sp]];
    TRY(upload(ba, ba->d_deltaF.p, d.data(), P));
  }
  return 0;
}

static int launch_linearize(hslam_ba* ba, bool apply) {
  hslam_ctx* ctx = ba->ctx;
  int blocks = (ba->R + 127) / 128;
  if (blocks < 1) blocks = 1;
  if (apply)
    HS_LAUNCH(ctx, HSLAM_K_BA_LINEARIZE, (ba_linearize_kernel<true><<<blocks, 128, 0, ctx->stream>>>(
                                             ba->win, ba->res, ba->pts, ba->d_epart.p, ba->d_ticket.p, ba->d_energy.p)));
  else
    HS_LAUNCH(ctx, HSLAM_K_BA_LINEARIZE, (ba_linearize_kernel<false><<<blocks, 128, 0, ctx->stream>>>(
                                             ba->win, ba->res, ba->pts, ba->d_epart.p, ba->d_ticket.p, ba->d_energy.p)));
  HS_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"

// download a per-residual SoA component block [comps][rs] into caller order [R][comps]
template <class T> static int download_res(hslam_ba* ba, const T* dev, int comps, T* out) {
  if (!out) return 0;
  const size_t rs = ba->rs;
  TRY(pin_ensure(ba, comps * rs * sizeof(T)));
  HS_CUDA(cudaMemcpyAsync(ba->h_pin, dev, comps * rs * sizeof(T), cudaMemcpyDeviceToHost, ba->ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ba->ctx->stream));
  const T* src = (const T*)ba->h_pin;
  for (int k = 0; k < ba->R; ++k) {
    size_t r = ba->r_sorted2orig[k];
    for (int c = 0; c < comps; ++c) out[r * comps + c] = src[c * rs + k];
  }
  return 0;
}

extern "C" {

int hslam_ba_linearize(hslam_ba* ba, int fixLinearization, double* energy_sum, int8_t* newState, float* newEnergy,
                       float* newEnergyWithOutlier, float* centerProjectedTo, float* projectedTo) {
  HS_ARG(ba && ba->have_topology && ba->have_window, "hslam_ba_linearize: window/topology not set");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  TRY(launch_linearize(ba, fixLinearization != 0));
  if (energy_sum) TRY(download(ba, energy_sum, ba->d_energy.p, 1));
  TRY(download_res(ba, ba->d_newState.p, 1, newState));
  TRY(download_res(ba, ba->d_newEnergy.p, 1, newEnergy));
  TRY(download_res(ba, ba->d_newEnergyWO.p, 1, newEnergyWithOutlier));
  TRY(download_res(ba, ba->d_cpt.p, 3, centerProjectedTo));
  TRY(download_res(ba, ba->d_pto.p, 16, projectedTo));
  HS_CUDA(cudaStreamSynchronize(ba->ctx->stream));
  return 0;
}

int hslam_ba_apply_res(hslam_ba* ba) {
  HS_ARG(ba && ba->have_topology, "hslam_ba_apply_res: no topology");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  int blocks = (ba->R + 255) / 256;
  if (blocks < 1) blocks = 1;
  HS_LAUNCH(ba->ctx, HSLAM_K_BA_APPLY, (ba_apply_kernel<<<blocks, 256, 0, ba->ctx->stream>>>(ba->res)));
  HS_CUDA(cudaGetLastError());
  return 0;
}

int hslam_ba_fix_linearization(hslam_ba* ba, const int32_t* ids, int count) {
  HS_ARG(ba && ba->have_topology && (count == 0 || ids), "hslam_ba_fix_linearization: bad argument");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  if (count == 0) return 0;
  std::vector<int> sorted(count);
  for (int i = 0; i < count; ++i) {
    HS_ARG(ids[i] >= 0 && ids[i] < ba->R, "hslam_ba_fix_linearization: id out of range");
    sorted[i] = ba->r_orig2sorted[ids[i]];
  }
  TRY(ba->d_ids.ensure(count));
  TRY(upload(ba, ba->d_ids.p, sorted.data(), count));
  HS_LAUNCH(ba->ctx, HSLAM_K_BA_FIXLIN, (ba_fixlin_kernel<<<(count + 127) / 128, 128, 0, ba->ctx->stream>>>(
                                            ba->win, ba->res, ba->pts, ba->d_ids.p, count)));
  HS_CUDA(cudaGetLastError());
  ba->n_linearized = -1;  // unknown without a count: run both modes
  return 0;
}

static int launch_top(hslam_ba* ba, int mode) {
  hslam_ctx* ctx = ba->ctx;
  HS_CUDA(cudaMemsetAsync(ba->d_nres.p + mode, 0, sizeof(int), ctx->stream));
  if (ba->n_topchunks > 0) {
    float* part = ba->d_toppartial.p;
    if (mode == 0)
      HS_LAUNCH(ctx, HSLAM_K_BA_TOP, (ba_top_kernel<0><<<ba->n_topchunks, TOP_THREADS, 0, ctx->stream>>>(
                                         ba->win, ba->res, ba->pts, ba->d_topchunks.p, part)));
    else
      HS_LAUNCH(ctx, HSLAM_K_BA_TOP, (ba_top_kernel<1><<<ba->n_topchunks, TOP_THREADS, 0, ctx->stream>>>(
                                         ba->win, ba->res, ba->pts, ba->d_topchunks.p, part)));
  }
  HS_LAUNCH(ctx, HSLAM_K_BA_FOLD_TOP, (ba_fold_top_kernel<<<ba->n * ba->n, 96, 0, ctx->stream>>>(
                                          ba->d_toppartial.p, ba->d_pair_chunk_begin.p, ba->n * ba->n,
                                          ba->d_accTop[mode].p, ba->d_nres.p + mode)));
  int pblocks = (ba->P + 127) / 128;
  if (pblocks > 0)
    HS_LAUNCH(ctx, HSLAM_K_BA_POINT_SUM,
              (ba_point_sum_kernel<<<pblocks, 128, 0, ctx->stream>>>(ba->res, ba->pts, mode == 0, mode == 1)));
  HS_CUDA(cudaGetLastError());
  return 0;
}

static int launch_schur(hslam_ba* ba) {
  hslam_ctx* ctx = ba->ctx;
  const int LD = ba->LD;
  size_t smem = (size_t)(SC_PTS * LD + SC_PTS) * sizeof(float);
  if (smem > 48 * 1024)
    HS_CUDA(cudaFuncSetAttribute(ba_schur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (ba->n_ptchunks > 0) {
    HS_LAUNCH(ctx, HSLAM_K_BA_SCHUR, (ba_schur_kernel<<<ba->n_ptchunks, SC_THREADS, smem, ctx->stream>>>(
                                         ba->win, ba->res, ba->pts, ba->d_ptchunks.p, LD, ba->d_scpartial.p)));
  }
  dim3 g((LD * LD + 255) / 256, ba->n);
  HS_LAUNCH(ctx, HSLAM_K_BA_FOLD_SC, (ba_fold_sc_kernel<<<g, 256, 0, ctx->stream>>>(
                                         ba->d_scpartial.p, ba->d_host_chunk_begin.p, LD, ba->d_G.p)));
  HS_CUDA(cudaGetLastError());
  return 0;
}

static int launch_stitch(hslam_ba* ba, int sys_begin, int sys_end) {
  hslam_ctx* ctx = ba->ctx;
  StitchArgs S;
  S.n = ba->n; S.LD = ba->LD;
  S.adHost = ba->d_adH.p; S.adTarget = ba->d_adT.p;
  S.accTop[0] = ba->d_accTop[0].p; S.accTop[1] = ba->d_accTop[1].p; S.G = ba->d_G.p;
  S.framePrior = ba->d_fprior.p; S.frameDeltaPrior = ba->d_fdprior.p;
  for (int i = 0; i < 4; ++i) { S.cPrior[i] = ba->cPrior[i]; S.cDeltaF[i] = ba->win.cDeltaF[i]; }
  S.add_priors = ba->add_priors;
  S.sys_begin = sys_begin; S.sys_end = sys_end;
  const size_t per = (size_t)ba->N * ba->N + ba->N;
  S.Hb = ba->d_Hb.p;
  // zero first: the lower blocks (a>b) of the top systems are written by their mirror block
  HS_CUDA(cudaMemsetAsync(ba->d_Hb.p + per * sys_begin, 0, per * (sys_end - sys_begin) * sizeof(double), ctx->stream));
  dim3 grid(ba->n + 1, ba->n + 1, 3);
  HS_LAUNCH(ctx, HSLAM_K_BA_STITCH, (ba_stitch_kernel<<<grid, 64, 0, ctx->stream>>>(S)));
  HS_CUDA(cudaGetLastError());
  return 0;
}

static int unpack_Hb(hslam_ba* ba, int sys, double* H, double* b) {
  const size_t NN = (size_t)ba->N * ba->N, per = NN + ba->N;
  std::vector<double> tmp(per);
  TRY(download(ba, tmp.data(), ba->d_Hb.p + per * sys, per));
  if (H) memcpy(H, tmp.data(), NN * sizeof(double));
  if (b) memcpy(b, tmp.data() + NN, ba->N * sizeof(double));
  return 0;
}

int hslam_ba_accumulate(hslam_ba* ba, int which, double* H, double* b, int* nres) {
  HS_ARG(ba && ba->have_topology && ba->have_window, "hslam_ba_accumulate: window/topology not set");
  HS_ARG(which >= 0 && which <= 2, "hslam_ba_accumulate: which must be HSLAM_ACC_A/L/SC");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  if (which == HSLAM_ACC_A || which == HSLAM_ACC_L) {
    TRY(launch_top(ba, which));
    TRY(launch_stitch(ba, which, which + 1));
    TRY(unpack_Hb(ba, which, H, b));
    if (nres) TRY(download(ba, nres, ba->d_nres.p + which, 1));
  } else {
    TRY(launch_schur(ba));
    TRY(launch_stitch(ba, 2, 3));
    TRY(unpack_Hb(ba, 2, H, b));
  }
  return 0;
}

static int enqueue_accumulate_all(hslam_ba* ba) {
  TRY(launch_top(ba, 0));
  TRY(launch_top(ba, 1));
  TRY(launch_schur(ba));
  TRY(launch_stitch(ba, 0, 3));
  return 0;
}

int hslam_ba_accumulate_all_device(hslam_ba* ba, double** Hb_dev_out, int* count_out) {
  HS_ARG(ba && ba->have_topology && ba->have_window, "hslam_ba_accumulate_all: window/topology not set");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  TRY(enqueue_accumulate_all(ba));
  if (Hb_dev_out) *Hb_dev_out = ba->d_Hb.p;
  if (count_out) *count_out = 3 * (ba->N * ba->N + ba->N);
  return 0;
}

int hslam_ba_iteration_device(hslam_ba* ba, double** Hb_dev_out, int* count_out) {
  HS_ARG(ba && ba->have_topology && ba->have_window, "hslam_ba_iteration: window/topology not set");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  TRY(launch_linearize(ba, true));
  return hslam_ba_accumulate_all_device(ba, Hb_dev_out, count_out);
}

int hslam_ba_download_Hb(hslam_ba* ba, double* out) {
  HS_ARG(ba && out && ba->have_window, "hslam_ba_download_Hb: bad argument");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  return download(ba, out, ba->d_Hb.p, 3 * ((size_t)ba->N * ba->N + ba->N));
}

int hslam_ba_accumulate_all(hslam_ba* ba, double* Hb_out_host, int* nres2) {
  HS_ARG(Hb_out_host, "hslam_ba_accumulate_all: null output");
  TRY(hslam_ba_accumulate_all_device(ba, nullptr, nullptr));
  TRY(hslam_ba_download_Hb(ba, Hb_out_host));
  if (nres2) TRY(download(ba, nres2, ba->d_nres.p, 2));
  return 0;
}

typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
int hslam_ba_allreduce(hslam_ba* ba, void* nccl_comm) {
  HS_ARG(ba && nccl_comm && ba->have_window, "hslam_ba_allreduce: bad argument");
  static nccl_allreduce_fn fn = nullptr;
  if (!fn) {
    fn = (nccl_allreduce_fn)dlsym(RTLD_DEFAULT, "ncclAllReduce");
    if (!fn) {
      void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (h) fn = (nccl_allreduce_fn)dlsym(h, "ncclAllReduce");
    }
    if (!fn) { hs::set_error("hslam_ba_allreduce: ncclAllReduce not found in the process or libnccl.so.2"); return HSLAM_ERR_NCCL; }
  }
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  size_t count = 3 * ((size_t)ba->N * ba->N + ba->N);
  // ncclDouble = 8, ncclSum = 0 (nccl.h)
  int rc = fn(ba->d_Hb.p, ba->d_Hb.p, count, 8, 0, nccl_comm, ba->ctx->stream);
  if (rc != 0) { hs::set_error("hslam_ba_allreduce: ncclAllReduce failed with code " + std::to_string(rc)); return HSLAM_ERR_NCCL; }
  return 0;
}

}  // extern "C"

template <class T> static int download_pts(hslam_ba* ba, const T* dev, int comps, T* out) {
  if (!out) return 0;
  std::vector<T> tmp((size_t)ba->P * comps);
  TRY(download(ba, tmp.data(), dev, (size_t)ba->P * comps));
  for (int sp = 0; sp < ba->P; ++sp)
    for (int c = 0; c < comps; ++c) out[(size_t)ba->p_sorted2orig[sp] * comps + c] = tmp[(size_t)sp * comps + c];
  return 0;
}

extern "C" {

int hslam_ba_point_results(hslam_ba* ba, float* HdiF, float* bdSumF, float* idepth_hessian, float* Hdd_accAF,
                           float* bd_accAF, float* Hcd_accAF, float* Hdd_accLF, float* bd_accLF, float* Hcd_accLF) {
  HS_ARG(ba && ba->have_topology, "hslam_ba_point_results: no topology");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  TRY(download_pts(ba, ba->d_HdiF.p, 1, HdiF)); TRY(download_pts(ba, ba->d_bdSumF.p, 1, bdSumF));
  TRY(download_pts(ba, ba->d_idh.p, 1, idepth_hessian));
  TRY(download_pts(ba, ba->d_Hdd[0].p, 1, Hdd_accAF)); TRY(download_pts(ba, ba->d_bd[0].p, 1, bd_accAF));
  TRY(download_pts(ba, (const float*)ba->d_Hcd[0].p, 4, Hcd_accAF));
  TRY(download_pts(ba, ba->d_Hdd[1].p, 1, Hdd_accLF)); TRY(download_pts(ba, ba->d_bd[1].p, 1, bd_accLF));
  TRY(download_pts(ba, (const float*)ba->d_Hcd[1].p, 4, Hcd_accLF));
  return 0;
}

int hslam_ba_residual_results(hslam_ba* ba, int8_t* state_state, uint8_t* is_active, float* JpJdF, float* J74,
                              float* res_toZeroF) {
  HS_ARG(ba && ba->have_topology, "hslam_ba_residual_results: no topology");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  TRY(download_res(ba, ba->d_state.p, 1, state_state));
  TRY(download_res(ba, ba->d_isActive.p, 1, is_active));
  TRY(download_res(ba, ba->d_JpJdF.p, 8, JpJdF));
  TRY(download_res(ba, ba->d_rtz.p, 8, res_toZeroF));
  if (J74) {
    std::vector<float> tmp((size_t)ba->R * JS);
    TRY(download_res(ba, ba->d_J.p, JS, tmp.data()));
    for (int r = 0; r < ba->R; ++r) memcpy(J74 + (size_t)r * 74, tmp.data() + (size_t)r * JS, 74 * sizeof(float));
  }
  return 0;
}

int hslam_ba_resubstitute(hslam_ba* ba, const double* x, float* step_out) {
  HS_ARG(ba && x && ba->have_topology && ba->have_window, "hslam_ba_resubstitute: bad argument");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  const int n = ba->n, N = ba->N;
  std::vector<float> xF(N), xAd((size_t)n * n * 8);
  for (int i = 0; i < N; ++i) xF[i] = (float)x[i];
  // EnergyFunctional.cpp:236-238 (float row-vector times float 8x8, index order)
  for (int h = 0; h < n; ++h)
    for (int t = 0; t < n; ++t)
      for (int c = 0; c < 8; ++c) {
        volatile float a = 0.f, b = 0.f;
        for (int r = 0; r < 8; ++r) {
          volatile float pa = xF[4 + 8 * h + r] * (float)ba->h_adH[64 * (size_t)(h + n * t) + r + 8 * c];
          volatile float pb = xF[4 + 8 * t + r] * (float)ba->h_adT[64 * (size_t)(h + n * t) + r + 8 * c];
          a = a + pa; b = b + pb;
        }
        xAd[8 * (size_t)(n * h + t) + c] = a + b;
      }
  TRY(ba->d_xc.ensure(4)); TRY(ba->d_xAd.ensure(xAd.size()));
  TRY(upload(ba, ba->d_xc.p, xF.data(), 4)); TRY(upload(ba, ba->d_xAd.p, xAd.data(), xAd.size()));
  int blocks = (ba->P + 127) / 128;
  if (blocks > 0) {
    HS_LAUNCH(ba->ctx, HSLAM_K_BA_RESUB, (ba_resub_kernel<<<blocks, 128, 0, ba->ctx->stream>>>(
                                             ba->win, ba->res, ba->pts, ba->d_xc.p, ba->d_xAd.p)));
    HS_CUDA(cudaGetLastError());
  }
  TRY(download_pts(ba, ba->d_step.p, 1, step_out));
  return 0;
}

}  // extern "C"
