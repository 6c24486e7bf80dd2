// oracle_trace.cpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY; see hslam_oracle.h).
// Restates ImmaturePoint::traceOn (reference Src/ImmaturePoint.cpp:40-351), the epipolar-line search that
// System::traceNewCoarse (Src/Mapping.cpp:494-538) runs for every immature point of every keyframe on every
// frame.  fp32 scalar code in the operation order of the source, Eigen's a0+(a1+a2) association for the 3-term
// products; compiled with -ffp-contract=off.  Only deviation: the bilinear taps are clamped into the image
// (the source reads whatever lies beyond; such a trace is OOB or an outlier either way).
#include "hslam_oracle.h"
#include "oracle_math.hpp"
#include <cmath>

namespace {

const int PATT[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};  // GlobalTypes.h:181

struct Tap { const float* bp; float wBR, wBL, wTR, wTL; int w; };
inline Tap tap(const float* img /*Vec3f AoS*/, float x, float y, int w, int h) {
  int ix = (int)x, iy = (int)y;
  float dx = x - ix, dy = y - iy, dxdy = dx * dy;
  if (ix < 0) ix = 0; if (iy < 0) iy = 0; if (ix > w - 2) ix = w - 2; if (iy > h - 2) iy = h - 2;
  Tap t;
  t.bp = img + 3 * ((size_t)ix + (size_t)iy * w);
  t.wBR = dxdy; t.wBL = dy - dxdy; t.wTR = dx - dxdy; t.wTL = 1 - dx - dy + dxdy; t.w = w;
  return t;
}
inline float comp(const Tap& t, int c) {   // getInterpolatedElement31/33, GlobalTypes.h:363-388
  return t.wBR * t.bp[3 * (1 + t.w) + c] + t.wBL * t.bp[3 * t.w + c] + t.wTR * t.bp[3 + c] + t.wTL * t.bp[c];
}

}  // namespace

extern "C" int orc_trace_immature(const float* dI0 /*w*h*3*/, int w, int h, const orc_trace_settings* S, int n_hosts,
                                  const float* KRKi /*n_hosts*9 row-major*/, const float* Kt /*n_hosts*3*/,
                                  const float* aff /*n_hosts*2*/, int n, const int32_t* host_idx, const float* u_,
                                  const float* v_, const float* gradH /*n*4: (0,0),(1,0),(0,1),(1,1)*/, const float* color /*n*8*/,
                                  const float* weights /*n*8*/, const float* energyTH, float* idepth_min_io,
                                  float* idepth_max_io, float* quality_io, int8_t* status_io, float* lastTraceUV /*n*2*/,
                                  float* lastTracePixelInterval) {
  for (int p = 0; p < n; ++p) {
    if (status_io[p] == ORC_IPS_OOB) continue;   // :43
    if (host_idx[p] < 0 || host_idx[p] >= n_hosts) return -1;
    const float* K = KRKi + 9 * host_idx[p];
    const float* T = Kt + 3 * host_idx[p];
    const float a0 = aff[2 * host_idx[p]], a1 = aff[2 * host_idx[p] + 1];
    const float u = u_[p], v = v_[p];
    float idmin = idepth_min_io[p], idmax = idepth_max_io[p];
    const float* G = gradH + 4 * p;
    auto fail = [&](int st) { lastTraceUV[2 * p] = -1; lastTraceUV[2 * p + 1] = -1; lastTracePixelInterval[p] = 0; status_io[p] = (int8_t)st; };
    const float maxPixSearch = (w + h) * S->maxPixSearch;
    float pr[3];
    for (int i = 0; i < 3; ++i) pr[i] = orc::sum3(K[3 * i] * u, K[3 * i + 1] * v, K[3 * i + 2] * 1.0f);
    float ptpMin[3];
    for (int i = 0; i < 3; ++i) ptpMin[i] = pr[i] + T[i] * idmin;
    float uMin = ptpMin[0] / ptpMin[2], vMin = ptpMin[1] / ptpMin[2];
    if (!(uMin > 4 && vMin > 4 && uMin < w - 5 && vMin < h - 5)) { fail(ORC_IPS_OOB); continue; }
    float dist, uMax, vMax;
    if (std::isfinite(idmax)) {
      float ptpMax[3];
      for (int i = 0; i < 3; ++i) ptpMax[i] = pr[i] + T[i] * idmax;
      uMax = ptpMax[0] / ptpMax[2]; vMax = ptpMax[1] / ptpMax[2];
      if (!(uMax > 4 && vMax > 4 && uMax < w - 5 && vMax < h - 5)) { fail(ORC_IPS_OOB); continue; }
      dist = (uMin - uMax) * (uMin - uMax) + (vMin - vMax) * (vMin - vMax);
      dist = sqrtf(dist);
      if (dist < S->slackInterval) {   // :86-95
        lastTraceUV[2 * p] = (uMax + uMin) * 0.5f; lastTraceUV[2 * p + 1] = (vMax + vMin) * 0.5f;
        lastTracePixelInterval[p] = dist;
        status_io[p] = ORC_IPS_SKIPPED;
        continue;
      }
    } else {
      dist = maxPixSearch;
      float ptpMax[3];
      for (int i = 0; i < 3; ++i) ptpMax[i] = pr[i] + T[i] * 0.01f;
      uMax = ptpMax[0] / ptpMax[2]; vMax = ptpMax[1] / ptpMax[2];
      float dx = uMax - uMin, dy = vMax - vMin;
      float d = 1.0f / sqrtf(dx * dx + dy * dy);
      uMax = uMin + dist * dx * d;
      vMax = vMin + dist * dy * d;
      if (!(uMax > 4 && vMax > 4 && uMax < w - 5 && vMax < h - 5)) { fail(ORC_IPS_OOB); continue; }
    }
    if (!(idmin < 0 || (ptpMin[2] > 0.75f && ptpMin[2] < 1.5f))) { fail(ORC_IPS_OOB); continue; }   // :134
    float dx = S->stepsize * (uMax - uMin), dy = S->stepsize * (vMax - vMin);
    float e0 = dx * G[0] + dy * G[1], e1 = dx * G[2] + dy * G[3];
    float a = e0 * dx + e1 * dy;
    float f0 = dy * G[0] + (-dx) * G[1], f1 = dy * G[2] + (-dx) * G[3];
    float b = f0 * dy + f1 * (-dx);
    float errorInPixel = 0.2f + 0.2f * (a + b) / a;
    if (errorInPixel * S->minImprovementFactor > dist && std::isfinite(idmax)) {   // :151-159
      lastTraceUV[2 * p] = (uMax + uMin) * 0.5f; lastTraceUV[2 * p + 1] = (vMax + vMin) * 0.5f;
      lastTracePixelInterval[p] = dist;
      status_io[p] = ORC_IPS_BADCONDITION;
      continue;
    }
    if (errorInPixel > 10) errorInPixel = 10;
    dx /= dist; dy /= dist;
    if (dist > maxPixSearch) { uMax = uMin + maxPixSearch * dx; vMax = vMin + maxPixSearch * dy; dist = maxPixSearch; }
    int numSteps = (int)(1.9999f + dist / S->stepsize);
    float randShift = uMin * 1000 - floorf(uMin * 1000);
    float ptx = uMin - randShift * dx, pty = vMin - randShift * dy;
    float rot[8][2];
    for (int i = 0; i < 8; ++i) {
      rot[i][0] = K[0] * (float)PATT[i][0] + K[1] * (float)PATT[i][1];
      rot[i][1] = K[3] * (float)PATT[i][0] + K[4] * (float)PATT[i][1];
    }
    if (!std::isfinite(dx) || !std::isfinite(dy)) { fail(ORC_IPS_OOB); continue; }
    float errors[100];
    float bestU = 0, bestV = 0, bestEnergy = 1e10f;
    int bestIdx = -1;
    if (numSteps >= 100) numSteps = 99;
    for (int i = 0; i < numSteps; ++i) {
      float energy = 0;
      for (int idx = 0; idx < 8; ++idx) {
        float hit = comp(tap(dI0, ptx + rot[idx][0], pty + rot[idx][1], w, h), 0);
        if (!std::isfinite(hit)) { energy += 1e5f; continue; }
        float residual = hit - (a0 * color[8 * p + idx] + a1);
        float hw = fabsf(residual) < S->huberTH ? 1 : S->huberTH / fabsf(residual);
        energy += hw * residual * residual * (2 - hw);
      }
      errors[i] = energy;
      if (energy < bestEnergy) { bestU = ptx; bestV = pty; bestEnergy = energy; bestIdx = i; }
      ptx += dx; pty += dy;
    }
    float secondBest = 1e10f;
    for (int i = 0; i < numSteps; ++i)
      if ((i < bestIdx - S->minTraceTestRadius || i > bestIdx + S->minTraceTestRadius) && errors[i] < secondBest) secondBest = errors[i];
    float newQuality = secondBest / bestEnergy;
    if (newQuality < quality_io[p] || numSteps > 10) quality_io[p] = newQuality;
    // GN refinement along the line, :262-316
    float uBak = bestU, vBak = bestV, gnstepsize = 1, stepBack = 0;
    if (S->GNIterations > 0) bestEnergy = 1e5f;
    for (int it = 0; it < S->GNIterations; ++it) {
      float H = 1, bb = 0, energy = 0;
      for (int idx = 0; idx < 8; ++idx) {
        Tap t = tap(dI0, bestU + rot[idx][0], bestV + rot[idx][1], w, h);
        float h0 = comp(t, 0), h1 = comp(t, 1), h2 = comp(t, 2);
        if (!std::isfinite(h0)) { energy += 1e5f; continue; }
        float residual = h0 - (a0 * color[8 * p + idx] + a1);
        float dResdDist = dx * h1 + dy * h2;
        float hw = fabsf(residual) < S->huberTH ? 1 : S->huberTH / fabsf(residual);
        H += hw * dResdDist * dResdDist;
        bb += hw * residual * dResdDist;
        energy += weights[8 * p + idx] * weights[8 * p + idx] * hw * residual * residual * (2 - hw);
      }
      if (energy > bestEnergy) {
        stepBack *= 0.5f;
        bestU = uBak + stepBack * dx; bestV = vBak + stepBack * dy;
      } else {
        float step = -gnstepsize * bb / H;
        if (step < -0.5f) step = -0.5f; else if (step > 0.5f) step = 0.5f;
        if (!std::isfinite(step)) step = 0;
        uBak = bestU; vBak = bestV; stepBack = step;
        bestU += step * dx; bestV += step * dy;
        bestEnergy = energy;
      }
      if (fabsf(stepBack) < S->GNThreshold) break;
    }
    if (!(bestEnergy < energyTH[p] * S->extraSlackOnTH)) {   // :320-331
      const int prev = status_io[p];
      fail(prev == ORC_IPS_OUTLIER ? ORC_IPS_OOB : ORC_IPS_OUTLIER);
      continue;
    }
    if (dx * dx > dy * dy) {
      idmin = (pr[2] * (bestU - errorInPixel * dx) - pr[0]) / (T[0] - T[2] * (bestU - errorInPixel * dx));
      idmax = (pr[2] * (bestU + errorInPixel * dx) - pr[0]) / (T[0] - T[2] * (bestU + errorInPixel * dx));
    } else {
      idmin = (pr[2] * (bestV - errorInPixel * dy) - pr[1]) / (T[1] - T[2] * (bestV - errorInPixel * dy));
      idmax = (pr[2] * (bestV + errorInPixel * dy) - pr[1]) / (T[1] - T[2] * (bestV + errorInPixel * dy));
    }
    if (idmin > idmax) { float t = idmin; idmin = idmax; idmax = t; }
    idepth_min_io[p] = idmin; idepth_max_io[p] = idmax;   // the source assigns the members before the validity test
    if (!std::isfinite(idmin) || !std::isfinite(idmax) || (idmax < 0)) { fail(ORC_IPS_OUTLIER); continue; }
    lastTracePixelInterval[p] = 2 * errorInPixel;
    lastTraceUV[2 * p] = bestU; lastTraceUV[2 * p + 1] = bestV;
    status_io[p] = ORC_IPS_GOOD;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// System::optimizeImmaturePoint (Src/FullSystemOptPoint.cpp:24-175) + ImmaturePoint::linearizeResidual
// (Src/ImmaturePoint.cpp:389-451) + projectPoint / derive_idepth (Include/DirectProjection.h:7-38): the 1-D
// Gauss-Newton over the inverse depth that decides whether an immature point is activated.  Outputs what the caller
// needs to create the MapPoint and its PointFrameResiduals: activated?, idepth, the final state of the residual into
// every other frame.
namespace {
struct ActRes { int state, newState; double energy, newEnergy; };

double lin_res(const float* dI, int w, int h, float fx, float fy, float cx, float cy, float fxi, float fyi, const float* R /*9 rm*/,
               const float* t, float aff0, float aff1, float huber, float pu, float pv, const float* color, const float* weights,
               float energyTH, float outlierTHSlack, ActRes& r, float& Hdd, float& bd, float idepth) {
  if (r.state == 1) { r.newState = 1; return r.energy; }   // OOB stays OOB (:392-393)
  float energyLeft = 0;
  for (int idx = 0; idx < 8; ++idx) {
    const int dx = PATT[idx][0], dy = PATT[idx][1];
    // projectPoint, DirectProjection.h:20-38
    float K0 = (pu + dx - cx) * fxi, K1 = (pv + dy - cy) * fyi;
    float p0 = orc::sum3(R[0] * K0, R[1] * K1, R[2] * 1.0f) + t[0] * idepth;
    float p1 = orc::sum3(R[3] * K0, R[4] * K1, R[5] * 1.0f) + t[1] * idepth;
    float p2 = orc::sum3(R[6] * K0, R[7] * K1, R[8] * 1.0f) + t[2] * idepth;
    float drescale = 1.0f / p2;
    if (!(drescale > 0)) { r.newState = 1; return r.energy; }
    float u = p0 * drescale, v = p1 * drescale;
    float Ku = u * fx + cx, Kv = v * fy + cy;
    if (!(Ku > 1.1f && Kv > 1.1f && Ku < (w - 3) && Kv < (h - 3))) { r.newState = 1; return r.energy; }
    Tap tp = tap(dI, Ku, Kv, w, h);
    float h0 = comp(tp, 0), h1 = comp(tp, 1), h2 = comp(tp, 2);
    if (!std::isfinite(h0)) { r.newState = 1; return r.energy; }
    float residual = h0 - (aff0 * color[idx] + aff1);
    float hw = fabsf(residual) < huber ? 1 : huber / fabsf(residual);
    energyLeft += weights[idx] * weights[idx] * hw * residual * residual * (2 - hw);
    float dxInterp = h1 * fx, dyInterp = h2 * fy;
    float d_idepth = (dxInterp * drescale * (t[0] - t[2] * u) + dyInterp * drescale * (t[1] - t[2] * v)) * 1.0f;   // SCALE_IDEPTH = 1
    hw *= weights[idx] * weights[idx];
    Hdd += (hw * d_idepth) * d_idepth;
    bd += (hw * residual) * d_idepth;
  }
  if (energyLeft > energyTH * outlierTHSlack) { energyLeft = energyTH * outlierTHSlack; r.newState = 2; }
  else r.newState = 0;
  r.newEnergy = energyLeft;
  return energyLeft;
}
}  // namespace

extern "C" int orc_activate_immature(int n_frames, const float* const* dI0 /*n_frames level-0 Vec3f images*/, int w, int h, float fx,
                                     float fy, float cx, float cy, float fxi, float fyi, const float* RtAff /*n*n*14, host+target*n*/,
                                     float huber, float minIdepthH_act, int GNIts, int minObs, int n, const int32_t* host_idx,
                                     const float* u, const float* v, const float* color, const float* weights, const float* energyTH,
                                     const float* idepth_min, const float* idepth_max, int8_t* ok_out, float* idepth_out,
                                     int8_t* res_state_out /*n*n_frames*/) {
  for (int p = 0; p < n; ++p) {
    const int host = host_idx[p];
    if (host < 0 || host >= n_frames) return -1;
    ok_out[p] = 0; idepth_out[p] = 0;
    for (int f = 0; f < n_frames; ++f) res_state_out[(size_t)p * n_frames + f] = 1;
    ActRes rs[16];
    int tgt[16], nres = 0;
    for (int f = 0; f < n_frames; ++f)
      if (f != host) { rs[nres].energy = rs[nres].newEnergy = 0; rs[nres].newState = 2; rs[nres].state = 0; tgt[nres] = f; nres++; }
    auto eval = [&](float slack, float id, float& Hdd, float& bd) {
      float E = 0;
      for (int i = 0; i < nres; ++i) {
        const float* T = RtAff + 14 * (size_t)(host + tgt[i] * n_frames);
        double e = lin_res(dI0[tgt[i]], w, h, fx, fy, cx, cy, fxi, fyi, T, T + 9, T[12], T[13], huber, u[p], v[p], color + 8 * (size_t)p,
                           weights + 8 * (size_t)p, energyTH[p], slack, rs[i], Hdd, bd, id);
        E = (float)((double)E + e);
      }
      return E;
    };
    float lastHdd = 0, lastbd = 0;
    float currentIdepth = (idepth_max[p] + idepth_min[p]) * 0.5f;
    float lastEnergy = 0;
    for (int i = 0; i < nres; ++i) {   // :51-56 (state committed after every residual)
      const float* T = RtAff + 14 * (size_t)(host + tgt[i] * n_frames);
      double e = lin_res(dI0[tgt[i]], w, h, fx, fy, cx, cy, fxi, fyi, T, T + 9, T[12], T[13], huber, u[p], v[p], color + 8 * (size_t)p,
                         weights + 8 * (size_t)p, energyTH[p], 1000.f, rs[i], lastHdd, lastbd, currentIdepth);
      lastEnergy = (float)((double)lastEnergy + e);
      rs[i].state = rs[i].newState; rs[i].energy = rs[i].newEnergy;
    }
    if (!std::isfinite(lastEnergy) || lastHdd < minIdepthH_act) continue;
    float lambda = 0.1f;
    bool dead = false;
    for (int it = 0; it < GNIts; ++it) {
      float H = lastHdd;
      H *= 1 + lambda;
      float step = (float)((1.0 / H) * lastbd);   // (1.0/H) is a double expression in the source
      float newIdepth = currentIdepth - step;
      float newHdd = 0, newbd = 0;
      float newEnergy = eval(1.f, newIdepth, newHdd, newbd);
      if (!std::isfinite(lastEnergy) || newHdd < minIdepthH_act) { dead = true; break; }
      if (newEnergy < lastEnergy) {
        currentIdepth = newIdepth; lastHdd = newHdd; lastbd = newbd; lastEnergy = newEnergy;
        for (int i = 0; i < nres; ++i) { rs[i].state = rs[i].newState; rs[i].energy = rs[i].newEnergy; }
        lambda *= 0.5f;
      } else {
        lambda *= 5;
      }
      if (fabsf(step) < 0.0001 * currentIdepth) break;
    }
    if (dead || !std::isfinite(currentIdepth)) continue;
    int good = 0;
    for (int i = 0; i < nres; ++i) if (rs[i].state == 0) good++;
    if (good < minObs) continue;
    if (!std::isfinite(energyTH[p])) continue;
    ok_out[p] = 1; idepth_out[p] = currentIdepth;
    for (int i = 0; i < nres; ++i) res_state_out[(size_t)p * n_frames + tgt[i]] = (int8_t)rs[i].state;
  }
  return 0;
}
