// trace.cu — epipolar-line search of the immature points on the device.
//
// Replaces ImmaturePoint::traceOn (reference Src/ImmaturePoint.cpp:40-351) as it is driven by
// System::traceNewCoarse (Src/Mapping.cpp:494-538): every immature point of every keyframe of the window is
// traced into the newest frame, on every frame.  
//
// Design (B200): ONE WARP PER POINT.  The projection of the inverse-depth interval, the error-bound tests and the
// step direction are a few dozen flops that every lane computes redundantly (uniform control flow, no shuffles).
// The discrete search — up to 99 steps along the epipolar line, 8 pattern pixels with a bilinear tap each — is
// where the work is: lane i evaluates steps i, i+32, i+64, i+96, so the 32 texel gathers of a step are in flight
// for 32 steps at once; the step positions are formed by the same chain of float additions as the source
// (ptx += dx), so every energy is bit-identical.  Best / second-best are warp reductions with the source's
// tie rule (first minimum wins).  The 3 Gauss-Newton refinements along the line are short and sequential: the
// lanes run them redundantly on broadcast loads.  All arithmetic uses the exactly-rounded primitives of
// common.cuh in the operation order of the source: states, intervals and qualities match the float-faithful
// oracle bit for bit.
#include "common.cuh"
#include <cmath>

namespace hs {

__constant__ int c_tpat[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};

struct TraceArgs {
  const float4* img; int w, h;
  hslam_trace_settings S;
  int n_hosts;
  const float* KRKi; const float* Kt; const float* aff;
  int n;
  const int* host_idx;
  const float* u; const float* v; const float* gradH; const float* color; const float* weights; const float* energyTH;
  float* idmin; float* idmax; float* quality; int8_t* status; float* uv; float* interval;
};

struct Tap4 { float4 tl, tr, bl, br; float wBR, wBL, wTR, wTL; };
__device__ __forceinline__ Tap4 tap4(const float4* __restrict__ img, float x, float y, int w, int h) {
  int ix = (int)x, iy = (int)y;
  const float dx = xsub(x, (float)ix), dy = xsub(y, (float)iy), dxdy = xmul(dx, dy);
  ix = max(0, min(ix, w - 2)); iy = max(0, min(iy, h - 2));   // the source reads beyond the image here; see oracle_trace.cpp
  const float4* bp = img + ix + (size_t)iy * w;
  Tap4 t;
  t.tl = __ldg(bp); t.tr = __ldg(bp + 1); t.bl = __ldg(bp + w); t.br = __ldg(bp + w + 1);
  t.wBR = dxdy; t.wBL = xsub(dy, dxdy); t.wTR = xsub(dx, dxdy); t.wTL = xadd(xsub(xsub(1.0f, dx), dy), dxdy);
  return t;
}
#define TAP_COMP(t, m) xadd(xadd(xadd(xmul((t).wBR, (t).br.m), xmul((t).wBL, (t).bl.m)), xmul((t).wTR, (t).tr.m)), xmul((t).wTL, (t).tl.m))

__global__ void __launch_bounds__(128) trace_kernel(TraceArgs a) {
  const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= a.n) return;
  const int st_in = a.status[p];
  if (st_in == HSLAM_IPS_OOB) return;   // :43
  const hslam_trace_settings& S = a.S;
  const int hi = a.host_idx[p];
  const float* K = a.KRKi + 9 * hi;
  const float* T = a.Kt + 3 * hi;
  const float a0 = a.aff[2 * hi], a1 = a.aff[2 * hi + 1];
  const float u = a.u[p], v = a.v[p];
  float idmin = a.idmin[p], idmax = a.idmax[p];
  const int w = a.w, h = a.h;
  auto fail = [&](int st) {
    if (lane == 0) { a.uv[2 * p] = -1.f; a.uv[2 * p + 1] = -1.f; a.interval[p] = 0.f; a.status[p] = (int8_t)st; }
  };
  const float maxPixSearch = xmul((float)(w + h), S.maxPixSearch);
  float pr[3], ptpMin[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    pr[i] = xdot3(K[3 * i], K[3 * i + 1], K[3 * i + 2], u, v, 1.0f);
    ptpMin[i] = xadd(pr[i], xmul(T[i], idmin));
  }
  const float uMin = xdiv(ptpMin[0], ptpMin[2]), vMin = xdiv(ptpMin[1], ptpMin[2]);
  const float wm5 = (float)(w - 5), hm5 = (float)(h - 5);
  if (!(uMin > 4.f && vMin > 4.f && uMin < wm5 && vMin < hm5)) { fail(HSLAM_IPS_OOB); return; }
  float dist, uMax, vMax;
  const bool finiteMax = isfinite(idmax);
  if (finiteMax) {
    const float q0 = xadd(pr[0], xmul(T[0], idmax)), q1 = xadd(pr[1], xmul(T[1], idmax)), q2 = xadd(pr[2], xmul(T[2], idmax));
    uMax = xdiv(q0, q2); vMax = xdiv(q1, q2);
    if (!(uMax > 4.f && vMax > 4.f && uMax < wm5 && vMax < hm5)) { fail(HSLAM_IPS_OOB); return; }
    const float du = xsub(uMin, uMax), dv = xsub(vMin, vMax);
    dist = xsqrt(xadd(xmul(du, du), xmul(dv, dv)));
    if (dist < S.slackInterval) {   // :86-95
      if (lane == 0) {
        a.uv[2 * p] = xmul(xadd(uMax, uMin), 0.5f); a.uv[2 * p + 1] = xmul(xadd(vMax, vMin), 0.5f);
        a.interval[p] = dist; a.status[p] = HSLAM_IPS_SKIPPED;
      }
      return;
    }
  } else {
    dist = maxPixSearch;
    const float q0 = xadd(pr[0], xmul(T[0], 0.01f)), q1 = xadd(pr[1], xmul(T[1], 0.01f)), q2 = xadd(pr[2], xmul(T[2], 0.01f));
    uMax = xdiv(q0, q2); vMax = xdiv(q1, q2);
    const float ddx = xsub(uMax, uMin), ddy = xsub(vMax, vMin);
    const float d = xdiv(1.0f, xsqrt(xadd(xmul(ddx, ddx), xmul(ddy, ddy))));
    uMax = xadd(uMin, xmul(xmul(dist, ddx), d));
    vMax = xadd(vMin, xmul(xmul(dist, ddy), d));
    if (!(uMax > 4.f && vMax > 4.f && uMax < wm5 && vMax < hm5)) { fail(HSLAM_IPS_OOB); return; }
  }
  if (!(idmin < 0.f || (ptpMin[2] > 0.75f && ptpMin[2] < 1.5f))) { fail(HSLAM_IPS_OOB); return; }   // :134
  float dx = xmul(S.stepsize, xsub(uMax, uMin)), dy = xmul(S.stepsize, xsub(vMax, vMin));
  const float* G = a.gradH + 4 * p;
  const float G00 = G[0], G10 = G[1], G01 = G[2], G11 = G[3];
  const float e0 = xadd(xmul(dx, G00), xmul(dy, G10)), e1 = xadd(xmul(dx, G01), xmul(dy, G11));
  const float ea = xadd(xmul(e0, dx), xmul(e1, dy));
  const float f0 = xadd(xmul(dy, G00), xmul(-dx, G10)), f1 = xadd(xmul(dy, G01), xmul(-dx, G11));
  const float eb = xadd(xmul(f0, dy), xmul(f1, -dx));
  float errorInPixel = xadd(0.2f, xdiv(xmul(0.2f, xadd(ea, eb)), ea));
  if (xmul(errorInPixel, S.minImprovementFactor) > dist && finiteMax) {   // :151-159
    if (lane == 0) {
      a.uv[2 * p] = xmul(xadd(uMax, uMin), 0.5f); a.uv[2 * p + 1] = xmul(xadd(vMax, vMin), 0.5f);
      a.interval[p] = dist; a.status[p] = HSLAM_IPS_BADCONDITION;
    }
    return;
  }
  if (errorInPixel > 10.f) errorInPixel = 10.f;
  dx = xdiv(dx, dist); dy = xdiv(dy, dist);
  if (dist > maxPixSearch) { dist = maxPixSearch; }   // uMax/vMax are not used any more
  int numSteps = (int)xadd(1.9999f, xdiv(dist, S.stepsize));
  const float us = xmul(uMin, 1000.f);
  const float randShift = xsub(us, floorf(us));
  float ptx = xsub(uMin, xmul(randShift, dx)), pty = xsub(vMin, xmul(randShift, dy));
  float rot[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    rot[i][0] = xadd(xmul(K[0], (float)c_tpat[i][0]), xmul(K[1], (float)c_tpat[i][1]));
    rot[i][1] = xadd(xmul(K[3], (float)c_tpat[i][0]), xmul(K[4], (float)c_tpat[i][1]));
  }
  if (!isfinite(dx) || !isfinite(dy)) { fail(HSLAM_IPS_OOB); return; }
  if (numSteps >= 100) numSteps = 99;
  float col[8], wgt[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { col[i] = a.color[8 * p + i]; wgt[i] = a.weights[8 * p + i]; }

  // ---- discrete search: lane L owns steps L, L+32, L+64, L+96
  float myE[4], myX[4], myY[4];
  {
    float x = ptx, y = pty;
    for (int i = 0; i < lane; ++i) { x = xadd(x, dx); y = xadd(y, dy); }   // the source's chain of additions
#pragma unroll
    for (int rd = 0; rd < 4; ++rd) {
      const int i = lane + 32 * rd;
      myE[rd] = 3e38f; myX[rd] = x; myY[rd] = y;
      if (i < numSteps) {
        Tap4 tp[8];
#pragma unroll
        for (int idx = 0; idx < 8; ++idx) tp[idx] = tap4(a.img, xadd(x, rot[idx][0]), xadd(y, rot[idx][1]), w, h);
        float energy = 0.f;
#pragma unroll
        for (int idx = 0; idx < 8; ++idx) {
          const float hit = TAP_COMP(tp[idx], x);
          if (!isfinite(hit)) { energy = xadd(energy, 1e5f); continue; }
          const float residual = xsub(hit, xadd(xmul(a0, col[idx]), a1));
          const float ar = fabsf(residual);
          const float hw = ar < S.huberTH ? 1.f : xdiv(S.huberTH, ar);
          energy = xadd(energy, xmul(xmul(xmul(hw, residual), residual), xsub(2.f, hw)));
        }
        myE[rd] = energy;
      }
      if (rd < 3) for (int k = 0; k < 32; ++k) { x = xadd(x, dx); y = xadd(y, dy); }
    }
  }
  // best = first step with the smallest energy (the source scans upwards with a strict <), start value 1e10
  float bestEnergy = 1e10f;
  int bestIdx = -1;
#pragma unroll
  for (int rd = 0; rd < 4; ++rd) {
    const int i = lane + 32 * rd;
    if (i < numSteps && myE[rd] < bestEnergy) { bestEnergy = myE[rd]; bestIdx = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float oe = __shfl_xor_sync(0xffffffffu, bestEnergy, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bestIdx, o);
    if (oi >= 0 && (bestIdx < 0 || oe < bestEnergy || (oe == bestEnergy && oi < bestIdx))) { bestEnergy = oe; bestIdx = oi; }
  }
  float bestU = 0.f, bestV = 0.f;
  if (bestIdx >= 0) {
    const int owner = bestIdx & 31, rd = bestIdx >> 5;
    const float sx = rd == 0 ? myX[0] : rd == 1 ? myX[1] : rd == 2 ? myX[2] : myX[3];
    const float sy = rd == 0 ? myY[0] : rd == 1 ? myY[1] : rd == 2 ? myY[2] : myY[3];
    bestU = __shfl_sync(0xffffffffu, sx, owner);
    bestV = __shfl_sync(0xffffffffu, sy, owner);
  }
  float secondBest = 1e10f;
#pragma unroll
  for (int rd = 0; rd < 4; ++rd) {
    const int i = lane + 32 * rd;
    if (i < numSteps && (i < bestIdx - S.minTraceTestRadius || i > bestIdx + S.minTraceTestRadius) && myE[rd] < secondBest) secondBest = myE[rd];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) secondBest = fminf(secondBest, __shfl_xor_sync(0xffffffffu, secondBest, o));
  const float newQuality = xdiv(secondBest, bestEnergy);
  if (lane == 0 && (newQuality < a.quality[p] || numSteps > 10)) a.quality[p] = newQuality;

  // ---- Gauss-Newton refinement along the line, :262-316 (all lanes, broadcast loads)
  float uBak = bestU, vBak = bestV, stepBack = 0.f;
  if (S.GNIterations > 0) bestEnergy = 1e5f;
  for (int it = 0; it < S.GNIterations; ++it) {
    float H = 1.f, bb = 0.f, energy = 0.f;
    Tap4 tp[8];
#pragma unroll
    for (int idx = 0; idx < 8; ++idx) tp[idx] = tap4(a.img, xadd(bestU, rot[idx][0]), xadd(bestV, rot[idx][1]), w, h);
#pragma unroll
    for (int idx = 0; idx < 8; ++idx) {
      const float h0 = TAP_COMP(tp[idx], x), h1 = TAP_COMP(tp[idx], y), h2 = TAP_COMP(tp[idx], z);
      if (!isfinite(h0)) { energy = xadd(energy, 1e5f); continue; }
      const float residual = xsub(h0, xadd(xmul(a0, col[idx]), a1));
      const float dRes = xadd(xmul(dx, h1), xmul(dy, h2));
      const float ar = fabsf(residual);
      const float hw = ar < S.huberTH ? 1.f : xdiv(S.huberTH, ar);
      H = xadd(H, xmul(xmul(hw, dRes), dRes));
      bb = xadd(bb, xmul(xmul(hw, residual), dRes));
      energy = xadd(energy, xmul(xmul(xmul(xmul(xmul(wgt[idx], wgt[idx]), hw), residual), residual), xsub(2.f, hw)));
    }
    if (energy > bestEnergy) {
      stepBack = xmul(stepBack, 0.5f);
      bestU = xadd(uBak, xmul(stepBack, dx)); bestV = xadd(vBak, xmul(stepBack, dy));
    } else {
      float step = xdiv(xmul(-1.f, bb), H);
      if (step < -0.5f) step = -0.5f; else if (step > 0.5f) step = 0.5f;
      if (!isfinite(step)) step = 0.f;
      uBak = bestU; vBak = bestV; stepBack = step;
      bestU = xadd(bestU, xmul(step, dx)); bestV = xadd(bestV, xmul(step, dy));
      bestEnergy = energy;
    }
    if (fabsf(stepBack) < S.GNThreshold) break;
  }
  if (!(bestEnergy < xmul(a.energyTH[p], S.extraSlackOnTH))) {   // :320-331
    fail(st_in == HSLAM_IPS_OUTLIER ? HSLAM_IPS_OOB : HSLAM_IPS_OUTLIER);
    return;
  }
  const bool alongU = xmul(dx, dx) > xmul(dy, dy);
  const float best = alongU ? bestU : bestV, dd = alongU ? dx : dy, prc = alongU ? pr[0] : pr[1], Tc = alongU ? T[0] : T[1];
  const float lo = xsub(best, xmul(errorInPixel, dd)), hi2 = xadd(best, xmul(errorInPixel, dd));
  idmin = xdiv(xsub(xmul(pr[2], lo), prc), xsub(Tc, xmul(T[2], lo)));
  idmax = xdiv(xsub(xmul(pr[2], hi2), prc), xsub(Tc, xmul(T[2], hi2)));
  if (idmin > idmax) { const float t = idmin; idmin = idmax; idmax = t; }
  if (lane == 0) { a.idmin[p] = idmin; a.idmax[p] = idmax; }   // assigned before the validity test, like the members
  if (!isfinite(idmin) || !isfinite(idmax) || idmax < 0.f) { fail(HSLAM_IPS_OUTLIER); return; }
  if (lane == 0) {
    a.interval[p] = xmul(2.f, errorInPixel);
    a.uv[2 * p] = bestU; a.uv[2 * p + 1] = bestV;
    a.status[p] = HSLAM_IPS_GOOD;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// System::optimizeImmaturePoint (Src/FullSystemOptPoint.cpp:24-175) + ImmaturePoint::linearizeResidual
// (Src/ImmaturePoint.cpp:389-451): the 1-D Gauss-Newton on the inverse depth that decides the activation of an
// immature point.  One thread per candidate: Hdd / bd are ONE sequential chain of float additions over (residual,
// pattern pixel) in the source (they are passed by reference through all residuals, and an out-of-bounds pixel
// leaves the partial sums of the pixels before it), so the chain is kept in one thread; the candidates of a
// keyframe (~2000 per host) fill the device.
struct ActArgs {
  const float4* img[HSLAM_MAX_FRAMES];
  int n_frames, w, h;
  float fx, fy, cx, cy, fxi, fyi, huber, minIdepthH_act;
  int GNIts, minObs;
  const float* RtAff;
  int n;
  const int* host_idx;
  const float* u; const float* v; const float* color; const float* weights; const float* energyTH;
  const float* idmin; const float* idmax;
  int8_t* ok; float* idepth; int8_t* res_state;
};

struct ActRes { int state, newState; double energy, newEnergy; };

__device__ double act_lin_res(const ActArgs& a, const float4* __restrict__ img, const float* __restrict__ T, float pu, float pv,
                              const float* color, const float* wgt, float energyTH, float slack, ActRes& r, float& Hdd, float& bd,
                              float idepth) {
  if (r.state == HSLAM_RES_OOB) { r.newState = HSLAM_RES_OOB; return r.energy; }
  const float wm3 = (float)(a.w - 3), hm3 = (float)(a.h - 3);
  const float t0 = T[9], t1 = T[10], t2 = T[11], aff0 = T[12], aff1 = T[13];
  float energyLeft = 0.f;
#pragma unroll 1
  for (int idx = 0; idx < 8; ++idx) {
    const float K0 = xmul(xsub(xadd(pu, (float)c_tpat[idx][0]), a.cx), a.fxi), K1 = xmul(xsub(xadd(pv, (float)c_tpat[idx][1]), a.cy), a.fyi);
    const float p0 = xadd(xdot3(T[0], T[1], T[2], K0, K1, 1.0f), xmul(t0, idepth));
    const float p1 = xadd(xdot3(T[3], T[4], T[5], K0, K1, 1.0f), xmul(t1, idepth));
    const float p2 = xadd(xdot3(T[6], T[7], T[8], K0, K1, 1.0f), xmul(t2, idepth));
    const float drescale = xdiv(1.0f, p2);
    if (!(drescale > 0.f)) { r.newState = HSLAM_RES_OOB; return r.energy; }
    const float u = xmul(p0, drescale), v = xmul(p1, drescale);
    const float Ku = xadd(xmul(u, a.fx), a.cx), Kv = xadd(xmul(v, a.fy), a.cy);
    if (!(Ku > 1.1f && Kv > 1.1f && Ku < wm3 && Kv < hm3)) { r.newState = HSLAM_RES_OOB; return r.energy; }
    const Tap4 tp = tap4(img, Ku, Kv, a.w, a.h);
    const float h0 = TAP_COMP(tp, x), h1 = TAP_COMP(tp, y), h2 = TAP_COMP(tp, z);
    if (!isfinite(h0)) { r.newState = HSLAM_RES_OOB; return r.energy; }
    const float residual = xsub(h0, xadd(xmul(aff0, color[idx]), aff1));
    const float ar = fabsf(residual);
    float hw = ar < a.huber ? 1.f : xdiv(a.huber, ar);
    const float ww = xmul(wgt[idx], wgt[idx]);
    energyLeft = xadd(energyLeft, xmul(xmul(xmul(xmul(ww, hw), residual), residual), xsub(2.f, hw)));
    const float dxI = xmul(h1, a.fx), dyI = xmul(h2, a.fy);
    const float d_id = xmul(xadd(xmul(xmul(dxI, drescale), xsub(t0, xmul(t2, u))), xmul(xmul(dyI, drescale), xsub(t1, xmul(t2, v)))), 1.0f);
    hw = xmul(hw, ww);
    Hdd = xadd(Hdd, xmul(xmul(hw, d_id), d_id));
    bd = xadd(bd, xmul(xmul(hw, residual), d_id));
  }
  const float lim = xmul(energyTH, slack);
  if (energyLeft > lim) { energyLeft = lim; r.newState = HSLAM_RES_OUT; } else r.newState = HSLAM_RES_IN;
  r.newEnergy = (double)energyLeft;
  return (double)energyLeft;
}

__global__ void __launch_bounds__(128) activate_kernel(ActArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.n) return;
  const int host = a.host_idx[p], nf = a.n_frames;
  int8_t* out_state = a.res_state + (size_t)p * nf;
  for (int f = 0; f < nf; ++f) out_state[f] = HSLAM_RES_OOB;
  a.ok[p] = 0; a.idepth[p] = 0.f;
  ActRes rs[HSLAM_MAX_FRAMES];
  int tgt[HSLAM_MAX_FRAMES], nres = 0;
  for (int f = 0; f < nf; ++f)
    if (f != host) { rs[nres].energy = rs[nres].newEnergy = 0.0; rs[nres].newState = HSLAM_RES_OUT; rs[nres].state = HSLAM_RES_IN; tgt[nres] = f; nres++; }
  float col[8], wgt[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { col[i] = a.color[8 * (size_t)p + i]; wgt[i] = a.weights[8 * (size_t)p + i]; }
  const float pu = a.u[p], pv = a.v[p], eTH = a.energyTH[p];
  float lastHdd = 0.f, lastbd = 0.f, lastEnergy = 0.f;
  float cur = xmul(xadd(a.idmax[p], a.idmin[p]), 0.5f);
  for (int i = 0; i < nres; ++i) {   // :51-56
    const double e = act_lin_res(a, a.img[tgt[i]], a.RtAff + 14 * (size_t)(host + tgt[i] * nf), pu, pv, col, wgt, eTH, 1000.f, rs[i], lastHdd, lastbd, cur);
    lastEnergy = (float)((double)lastEnergy + e);
    rs[i].state = rs[i].newState; rs[i].energy = rs[i].newEnergy;
  }
  if (!isfinite(lastEnergy) || lastHdd < a.minIdepthH_act) return;
  float lambda = 0.1f;
  for (int it = 0; it < a.GNIts; ++it) {
    float H = lastHdd;
    H = xmul(H, xadd(1.f, lambda));
    const float step = (float)((1.0 / (double)H) * (double)lastbd);
    const float newId = xsub(cur, step);
    float newHdd = 0.f, newbd = 0.f, newEnergy = 0.f;
    for (int i = 0; i < nres; ++i) {
      const double e = act_lin_res(a, a.img[tgt[i]], a.RtAff + 14 * (size_t)(host + tgt[i] * nf), pu, pv, col, wgt, eTH, 1.f, rs[i], newHdd, newbd, newId);
      newEnergy = (float)((double)newEnergy + e);
    }
    if (!isfinite(lastEnergy) || newHdd < a.minIdepthH_act) return;
    if (newEnergy < lastEnergy) {
      cur = newId; lastHdd = newHdd; lastbd = newbd; lastEnergy = newEnergy;
      for (int i = 0; i < nres; ++i) { rs[i].state = rs[i].newState; rs[i].energy = rs[i].newEnergy; }
      lambda = xmul(lambda, 0.5f);
    } else {
      lambda = xmul(lambda, 5.f);
    }
    if ((double)fabsf(step) < 0.0001 * (double)cur) break;
  }
  if (!isfinite(cur)) return;
  int good = 0;
  for (int i = 0; i < nres; ++i) if (rs[i].state == HSLAM_RES_IN) good++;
  if (good < a.minObs || !isfinite(eTH)) return;
  a.ok[p] = 1; a.idepth[p] = cur;
  for (int i = 0; i < nres; ++i) out_state[tgt[i]] = (int8_t)rs[i].state;
}

}  // namespace hs

using namespace hs;

extern "C" {

void hslam_trace_settings_default(hslam_trace_settings* s) {
  if (!s) return;
  s->maxPixSearch = 0.027f;        // Src/Settings.cpp:85
  s->slackInterval = 1.5f;         // :87
  s->stepsize = 1.0f;              // :88
  s->minImprovementFactor = 2.f;   // :89
  s->minTraceTestRadius = 2;       // :90
  s->GNIterations = 3;             // :91
  s->GNThreshold = 0.1f;           // :92
  s->extraSlackOnTH = 1.2f;        // :93
  s->huberTH = 9.f;                // :68
}

int hslam_trace_immature(hslam_ctx* ctx, const hslam_frame* frame, const hslam_trace_settings* settings, int n_hosts,
                         const float* KRKi, const float* Kt, const float* aff, const hslam_immature_soa* pts) {
  HS_ARG(ctx && frame && pts && KRKi && Kt && aff && n_hosts > 0, "hslam_trace_immature: null argument");
  HS_ARG(frame->ctx == ctx, "hslam_trace_immature: frame belongs to another context");
  const int n = pts->n_points;
  HS_ARG(n >= 0, "hslam_trace_immature: negative point count");
  if (n == 0) return 0;
  HS_ARG(pts->u && pts->v && pts->host_idx && pts->gradH && pts->color && pts->weights && pts->energyTH && pts->idepth_min &&
             pts->idepth_max && pts->quality && pts->lastTraceStatus && pts->lastTraceUV && pts->lastTracePixelInterval,
         "hslam_trace_immature: null point array");
  for (int i = 0; i < n; ++i) HS_ARG(pts->host_idx[i] >= 0 && pts->host_idx[i] < n_hosts, "hslam_trace_immature: host_idx out of range");
  HS_CUDA(cudaSetDevice(ctx->device));
  // one device block + one pinned block for everything that crosses (inputs up, the six in/out arrays back)
  const size_t nf = (size_t)n;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
  const size_t o_K = take(n_hosts * 9 * 4), o_T = take(n_hosts * 3 * 4), o_A = take(n_hosts * 2 * 4), o_h = take(nf * 4), o_u = take(nf * 4),
               o_v = take(nf * 4), o_g = take(nf * 16), o_c = take(nf * 32), o_w = take(nf * 32), o_e = take(nf * 4);
  const size_t o_io = off;
  const size_t o_min = take(nf * 4), o_max = take(nf * 4), o_q = take(nf * 4), o_uv = take(nf * 8), o_iv = take(nf * 4), o_st = take(nf);
  const size_t total = off;
  if (total > ctx->trace_cap) {
    if (ctx->trace_dev) cudaFree(ctx->trace_dev);
    if (ctx->trace_host) cudaFreeHost(ctx->trace_host);
    ctx->trace_dev = nullptr; ctx->trace_host = nullptr; ctx->trace_cap = 0;
    HS_CUDA(cudaMalloc((void**)&ctx->trace_dev, total + total / 4));
    HS_CUDA(cudaMallocHost((void**)&ctx->trace_host, total + total / 4));
    ctx->trace_cap = total + total / 4;
  }
  char* hp = ctx->trace_host;
  char* dp = ctx->trace_dev;
  memcpy(hp + o_K, KRKi, n_hosts * 9 * 4); memcpy(hp + o_T, Kt, n_hosts * 3 * 4); memcpy(hp + o_A, aff, n_hosts * 2 * 4);
  memcpy(hp + o_h, pts->host_idx, nf * 4); memcpy(hp + o_u, pts->u, nf * 4); memcpy(hp + o_v, pts->v, nf * 4);
  memcpy(hp + o_g, pts->gradH, nf * 16); memcpy(hp + o_c, pts->color, nf * 32); memcpy(hp + o_w, pts->weights, nf * 32);
  memcpy(hp + o_e, pts->energyTH, nf * 4);
  memcpy(hp + o_min, pts->idepth_min, nf * 4); memcpy(hp + o_max, pts->idepth_max, nf * 4); memcpy(hp + o_q, pts->quality, nf * 4);
  memcpy(hp + o_uv, pts->lastTraceUV, nf * 8); memcpy(hp + o_iv, pts->lastTracePixelInterval, nf * 4);
  memcpy(hp + o_st, pts->lastTraceStatus, nf);
  HS_CUDA(cudaMemcpyAsync(dp, hp, total, cudaMemcpyHostToDevice, ctx->stream));
  TraceArgs a;
  a.img = frame->dI[0]; a.w = frame->w; a.h = frame->h;
  if (settings) a.S = *settings; else hslam_trace_settings_default(&a.S);
  a.n_hosts = n_hosts;
  a.KRKi = (const float*)(dp + o_K); a.Kt = (const float*)(dp + o_T); a.aff = (const float*)(dp + o_A);
  a.n = n; a.host_idx = (const int*)(dp + o_h);
  a.u = (const float*)(dp + o_u); a.v = (const float*)(dp + o_v); a.gradH = (const float*)(dp + o_g);
  a.color = (const float*)(dp + o_c); a.weights = (const float*)(dp + o_w); a.energyTH = (const float*)(dp + o_e);
  a.idmin = (float*)(dp + o_min); a.idmax = (float*)(dp + o_max); a.quality = (float*)(dp + o_q);
  a.uv = (float*)(dp + o_uv); a.interval = (float*)(dp + o_iv); a.status = (int8_t*)(dp + o_st);
  const int blocks = (n * 32 + 127) / 128;
  HS_LAUNCH(ctx, HSLAM_K_TRACE, (trace_kernel<<<blocks, 128, 0, ctx->stream>>>(a)));
  HS_CUDA(cudaGetLastError());
  HS_CUDA(cudaMemcpyAsync(hp + o_io, dp + o_io, total - o_io, cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(pts->idepth_min, hp + o_min, nf * 4); memcpy(pts->idepth_max, hp + o_max, nf * 4); memcpy(pts->quality, hp + o_q, nf * 4);
  memcpy(pts->lastTraceUV, hp + o_uv, nf * 8); memcpy(pts->lastTracePixelInterval, hp + o_iv, nf * 4);
  memcpy(pts->lastTraceStatus, hp + o_st, nf);
  return 0;
}

int hslam_activate_immature(hslam_ctx* ctx, int n_frames, const hslam_frame* const* frames, float fx, float fy, float cx, float cy,
                            float fxi, float fyi, const float* RtAff, float minIdepthH_act, int GNIts, int minObs,
                            const hslam_activate_soa* pts) {
  HS_ARG(ctx && frames && RtAff && pts && n_frames >= 2 && n_frames <= HSLAM_MAX_FRAMES, "hslam_activate_immature: bad argument");
  const int n = pts->n_points;
  HS_ARG(n >= 0, "hslam_activate_immature: negative point count");
  if (n == 0) return 0;
  HS_ARG(pts->host_idx && pts->u && pts->v && pts->color && pts->weights && pts->energyTH && pts->idepth_min && pts->idepth_max &&
             pts->ok_out && pts->idepth_out && pts->res_state_out, "hslam_activate_immature: null point array");
  for (int i = 0; i < n; ++i) HS_ARG(pts->host_idx[i] >= 0 && pts->host_idx[i] < n_frames, "hslam_activate_immature: host_idx out of range");
  HS_CUDA(cudaSetDevice(ctx->device));
  ActArgs a;
  for (int f = 0; f < n_frames; ++f) {
    HS_ARG(frames[f] && frames[f]->ctx == ctx && frames[f]->w == frames[0]->w && frames[f]->h == frames[0]->h,
           "hslam_activate_immature: frame missing, from another context or of another size");
    a.img[f] = frames[f]->dI[0];
  }
  const size_t nf = (size_t)n, nn = (size_t)n_frames * n_frames;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
  const size_t o_T = take(nn * 14 * 4), o_h = take(nf * 4), o_u = take(nf * 4), o_v = take(nf * 4), o_c = take(nf * 32), o_w = take(nf * 32),
               o_e = take(nf * 4), o_min = take(nf * 4), o_max = take(nf * 4);
  const size_t o_io = off;
  const size_t o_ok = take(nf), o_id = take(nf * 4), o_st = take(nf * n_frames);
  const size_t total = off;
  if (total > ctx->trace_cap) {
    HS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->trace_dev) cudaFree(ctx->trace_dev);
    if (ctx->trace_host) cudaFreeHost(ctx->trace_host);
    ctx->trace_dev = nullptr; ctx->trace_host = nullptr; ctx->trace_cap = 0;
    HS_CUDA(cudaMalloc((void**)&ctx->trace_dev, total + total / 4));
    HS_CUDA(cudaMallocHost((void**)&ctx->trace_host, total + total / 4));
    ctx->trace_cap = total + total / 4;
  }
  char* hp = ctx->trace_host;
  char* dp = ctx->trace_dev;
  memcpy(hp + o_T, RtAff, nn * 14 * 4); memcpy(hp + o_h, pts->host_idx, nf * 4); memcpy(hp + o_u, pts->u, nf * 4); memcpy(hp + o_v, pts->v, nf * 4);
  memcpy(hp + o_c, pts->color, nf * 32); memcpy(hp + o_w, pts->weights, nf * 32); memcpy(hp + o_e, pts->energyTH, nf * 4);
  memcpy(hp + o_min, pts->idepth_min, nf * 4); memcpy(hp + o_max, pts->idepth_max, nf * 4);
  HS_CUDA(cudaMemcpyAsync(dp, hp, o_io, cudaMemcpyHostToDevice, ctx->stream));
  a.n_frames = n_frames; a.w = frames[0]->w; a.h = frames[0]->h;
  a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy; a.fxi = fxi; a.fyi = fyi; a.huber = ctx->prm.huberTH; a.minIdepthH_act = minIdepthH_act;
  a.GNIts = GNIts; a.minObs = minObs;
  a.RtAff = (const float*)(dp + o_T); a.n = n; a.host_idx = (const int*)(dp + o_h);
  a.u = (const float*)(dp + o_u); a.v = (const float*)(dp + o_v); a.color = (const float*)(dp + o_c); a.weights = (const float*)(dp + o_w);
  a.energyTH = (const float*)(dp + o_e); a.idmin = (const float*)(dp + o_min); a.idmax = (const float*)(dp + o_max);
  a.ok = (int8_t*)(dp + o_ok); a.idepth = (float*)(dp + o_id); a.res_state = (int8_t*)(dp + o_st);
  HS_LAUNCH(ctx, HSLAM_K_TRACE, (activate_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(a)));
  HS_CUDA(cudaGetLastError());
  HS_CUDA(cudaMemcpyAsync(hp + o_io, dp + o_io, total - o_io, cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(pts->ok_out, hp + o_ok, nf); memcpy(pts->idepth_out, hp + o_id, nf * 4); memcpy(pts->res_state_out, hp + o_st, nf * n_frames);
  return 0;
}

}  // extern "C"
