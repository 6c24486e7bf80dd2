// Exercises the C++ host mirror (h-slam_b200/host/hslam_host.hpp) the way the reference's System
// would: Frame::CreateDirPyrs, EnergyFunctional::{setAdjointsF/setDeltaF/precalc, makeIDX,
// linearizeAll, applyRes, accumulate{AF,LF,SCF}_MT, resubstituteF_MT} and CoarseTracker::
// {makeK, setCoarseTrackingRef, trackNewestCoarse}.  Input/expected data come from a binary file written
// by tests/test_gpu_host_shim.py; results go to another binary file the Python side checks.
#include <cstdio>
#include <cstdlib>
#include "../../h-slam_b200/host/hslam_host.hpp"

using namespace hslam_b200;

template <class T> static void rd(FILE* f, T* p, size_t n) { if (fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }
template <class T> static void wr(FILE* f, const T* p, size_t n) { fwrite(p, sizeof(T), n, f); }

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  FILE* fi = fopen(argv[1], "rb");
  FILE* fo = fopen(argv[2], "wb");
  if (!fi || !fo) return 2;
  try {
    int hdr[8];   // w, h, n, P, R, levels, pc_total, reserved
    rd(fi, hdr, 8);
    const int w = hdr[0], h = hdr[1], n = hdr[2], P = hdr[3], R = hdr[4], levels = hdr[5];
    CalibData calib;
    calib.Width = w; calib.Height = h;
    float K4[4];
    rd(fi, K4, 4);
    calib.fx = K4[0]; calib.fy = K4[1]; calib.cx = K4[2]; calib.cy = K4[3];
    Context ctx(0);
    // ---- frames
    std::vector<std::unique_ptr<Frame>> frames;
    std::vector<float> img((size_t)w * h);
    for (int f = 0; f < n; ++f) {
      rd(fi, img.data(), img.size());
      frames.emplace_back(new Frame(ctx, w, h, levels));
      frames.back()->CreateDirPyrs(img);
    }
    // ---- window state
    EnergyFunctional ef(ctx);
    ef.frames.resize(n);
    for (int f = 0; f < n; ++f) {
      rd(fi, ef.frames[f].worldToCam_evalPT.data(), 12);
      rd(fi, ef.frames[f].state, 10);
      rd(fi, ef.frames[f].state_zero, 10);
      rd(fi, ef.frames[f].prior, 8);
      rd(fi, ef.frames[f].delta_prior, 8);
      ef.frames[f].frame = frames[f].get();
    }
    rd(fi, ef.cDeltaF, 4);
    rd(fi, ef.cPrior, 4);
    ef.setAdjointsF_setDeltaF_setPrecalc(calib);
    // ---- points / residuals
    ef.allPoints.resize(P);
    for (int p = 0; p < P; ++p) {
      PointRec& q = ef.allPoints[p];
      float v5[5];
      rd(fi, v5, 5);
      q.u = v5[0]; q.v = v5[1]; q.idepth = v5[2]; q.idepth_zero = v5[3]; q.priorF = v5[4];
      rd(fi, q.color, 8); rd(fi, q.weights, 8);
      int hn[2];
      rd(fi, hn, 2);
      q.hostIdx = hn[0];
      q.residuals.resize(hn[1]);
      rd(fi, q.residuals.data(), hn[1]);
    }
    ef.residuals.resize(R);
    for (int r = 0; r < R; ++r) { int pt[2]; rd(fi, pt, 2); ef.residuals[r].point = pt[0]; ef.residuals[r].targetIdx = pt[1]; }
    ef.makeIDX();
    std::vector<int8_t> newState;
    double energy = ef.linearizeAll(false, &newState);
    const float newFrameTH = ef.setNewFrameEnergyTH();   // System::linearizeAll calls it after the pass (FullSystemOptimize.cpp:124)
    ef.applyRes_Reductor();
    MatXX HA, HL, HS; VecX bA, bL, bS;
    ef.accumulateAF_MT(HA, bA, true);
    ef.accumulateLF_MT(HL, bL, true);
    ef.accumulateSCF_MT(HS, bS, true);
    MatXX HA2, HL2, HS2; VecX bA2, bL2, bS2;
    ef.accumulateALL(HA2, bA2, HL2, bL2, HS2, bS2);
    const int N = 8 * n + 4;
    VecX x(N);
    rd(fi, x.data(), N);
    std::vector<float> steps;
    ef.resubstituteF_MT(x, steps);
    // ---- tracker on frame 1 against a reference cloud given in the file
    CoarseTracker trk(ctx, w, h);
    trk.makeK(calib, levels);
    for (int l = 0; l < levels; ++l) {
      int cnt;
      rd(fi, &cnt, 1);
      std::vector<float> a(4 * (size_t)cnt);
      rd(fi, a.data(), a.size());
      trk.setCoarseTrackingRef(l, a.data(), a.data() + cnt, a.data() + 2 * cnt, a.data() + 3 * cnt, cnt);
    }
    trk.setRefFrame(1.0f, AffLight(0, 0), 0);
    SE3 T;
    AffLight aff(0, 0);
    const double nanv[5] = {NAN, NAN, NAN, NAN, NAN};
    bool ok = trk.trackNewestCoarse(*frames[1], T, aff, levels - 1, nanv);
    // ---- write results
    wr(fo, &energy, 1);
    wr(fo, newState.data(), newState.size());
    wr(fo, HA.v.data(), HA.v.size()); wr(fo, bA.data(), bA.size());
    wr(fo, HL.v.data(), HL.v.size()); wr(fo, bL.data(), bL.size());
    wr(fo, HS.v.data(), HS.v.size()); wr(fo, bS.data(), bS.size());
    int same = (HA.v == HA2.v && HL.v == HL2.v && HS.v == HS2.v && bA == bA2 && bL == bL2 && bS == bS2) ? 1 : 0;
    int meta[4] = {same, ef.resInA, ef.resInL, ok ? 1 : 0};
    wr(fo, meta, 4);
    wr(fo, steps.data(), steps.size());
    wr(fo, T.data(), 12);
    double affo[2] = {aff.a, aff.b};
    wr(fo, affo, 2);
    wr(fo, trk.lastResiduals, 5);
    wr(fo, ef.precalc.data(), ef.precalc.size());
    wr(fo, ef.adHost.data(), ef.adHost.size());
    wr(fo, ef.adTarget.data(), ef.adTarget.size());
    wr(fo, ef.adHTdeltaF.data(), ef.adHTdeltaF.size());
    wr(fo, &newFrameTH, 1);
  } catch (const std::exception& e) {
    fprintf(stderr, "FAILED: %s\n", e.what());
    return 1;
  }
  fclose(fi); fclose(fo);
  return 0;
}
