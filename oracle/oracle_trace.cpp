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
