// oracle_ba.cpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY; see hslam_oracle.h).
// Restates the windowed-BA linearisation of the reference:
//   FrameFramePrecalc::set                Src/OptimizationClasses.cpp:13-39
//   PointFrameResidual::linearize         Src/OptimizationClasses.cpp:43-233
//   PointFrameResidual::applyRes/takeData Src/OptimizationClasses.cpp:235-256, Include/OptimizationClasses.h:155-161
//   PointFrameResidual::fixLinearizationF Src/OptimizationClasses.cpp:258-284
//   EnergyFunctional::setAdjointsF/setDeltaF/accumulate*_MT/resubstituteF_MT  Src/EnergyFunctional.cpp:22-274
//   AccumulatedTopHessianSSE::addPoint/stitchDouble*   Src/AccumulatedTopHessian.cpp:21-280, Include/AccumulatedTopHessian.h:69-117
//   AccumulatedSCHessianSSE::addPoint/stitchDouble*    Src/AccumulatedSCHessian.cpp:10-133, Include/AccumulatedSCHessian.h:70-111
//   AccumulatorApprox / AccumulatorXX / AccumulatorX   Include/MatrixAccumulators.h:36-89,177-237,595-972
//   IndexThreadReduce                                  Include/IndexThreadReduce.h:70-207
#include "hslam_oracle.h"
#include "oracle_math.hpp"
#include <vector>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <memory>

using namespace orc;

static const int PAT[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};  // GlobalTypes.h:181

// ---------------------------------------------------------------- window preparation
extern "C" int orc_ba_prepare_window(int n, const double* w2c_evalPT, const double* state, const double* state_zero,
                                     const float* ab_exposure, float fx, float fy, float cx, float cy,
                                     const hslam_params* prm, float* precalc, double* adHost, double* adTarget,
                                     float* adHTdeltaF) {
  if (n < 1 || n > HSLAM_MAX_FRAMES) return -1;
  std::vector<double> PRE_w2c(12 * n), PRE_c2w(12 * n);
  std::vector<double> aff(2 * n), aff0(2 * n);
  for (int f = 0; f < n; ++f) {
    const double* s = state + 10 * f;
    double sc[8];
    for (int i = 0; i < 3; ++i) sc[i] = prm->scale_xi_trans * s[i];     // Frame.h:184-197
    for (int i = 3; i < 6; ++i) sc[i] = prm->scale_xi_rot * s[i];
    sc[6] = prm->scale_a * s[6];
    sc[7] = prm->scale_b * s[7];
    double E[12];
    se3_exp(sc, E);
    se3_mul(E, w2c_evalPT + 12 * f, &PRE_w2c[12 * f]);
    se3_inv(&PRE_w2c[12 * f], &PRE_c2w[12 * f]);
    aff[2 * f] = sc[6]; aff[2 * f + 1] = sc[7];
    aff0[2 * f] = state_zero[10 * f + 6] * prm->scale_a;
    aff0[2 * f + 1] = state_zero[10 * f + 7] * prm->scale_b;
  }
  float K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1}, Ki[9];
  inv3_eigen<float>(K, Ki);
  for (int h = 0; h < n; ++h)
    for (int t = 0; t < n; ++t) {
      int idx = h + t * n;
      float* pc = precalc + HSLAM_PRECALC_STRIDE * idx;
      double inv0[12], ll0[12], ll[12];
      se3_inv(w2c_evalPT + 12 * h, inv0);
      se3_mul(w2c_evalPT + 12 * t, inv0, ll0);                       // OptimizationClasses.cpp:18
      se3_mul(&PRE_w2c[12 * t], &PRE_c2w[12 * h], ll);               // :22
      float R[9], tt[3], KR[9], KRKi[9], Kt[3];
      for (int i = 0; i < 9; ++i) R[i] = (float)ll[i];
      for (int i = 0; i < 3; ++i) tt[i] = (float)ll[9 + i];
      mat3_mul<float>(K, R, KR);
      mat3_mul<float>(KR, Ki, KRKi);                                  // :33
      mat3_vec<float>(K, tt[0], tt[1], tt[2], Kt);                    // :35
      for (int i = 0; i < 9; ++i) pc[i] = KRKi[i];
      for (int i = 0; i < 3; ++i) pc[9 + i] = Kt[i];
      for (int i = 0; i < 9; ++i) pc[12 + i] = (float)ll0[i];
      for (int i = 0; i < 3; ++i) pc[21 + i] = (float)ll0[9 + i];
      double a2[2];
      aff_from_to(ab_exposure[h], ab_exposure[t], aff[2 * h], aff[2 * h + 1], aff[2 * t], aff[2 * t + 1], a2);  // :37
      pc[24] = (float)a2[0]; pc[25] = (float)a2[1];
      pc[26] = (float)aff0[2 * h + 1];                                // :38
      pc[27] = 0;
      // adjoints, EnergyFunctional.cpp:36-61
      double Ad[36];
      se3_adj(ll0, Ad);
      double AH[64], AT[64];  // row-major here, transposed on store
      for (int i = 0; i < 64; ++i) AH[i] = AT[i] = 0;
      for (int i = 0; i < 8; ++i) AH[9 * i] = AT[9 * i] = 1;
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) AH[8 * r + c] = -Ad[6 * c + r];
      double a02[2];
      aff_from_to(ab_exposure[h], ab_exposure[t], aff0[2 * h], aff0[2 * h + 1], aff0[2 * t], aff0[2 * t + 1], a02);
      float affLL0 = (float)a02[0];
      AT[8 * 6 + 6] = -affLL0; AH[8 * 6 + 6] = affLL0;
      AT[8 * 7 + 7] = -1;      AH[8 * 7 + 7] = affLL0;
      double rs[8] = {prm->scale_xi_trans, prm->scale_xi_trans, prm->scale_xi_trans, prm->scale_xi_rot,
                      prm->scale_xi_rot, prm->scale_xi_rot, prm->scale_a, prm->scale_b};
      for (int r = 0; r < 8; ++r)
        for (int c = 0; c < 8; ++c) { AH[8 * r + c] *= rs[r]; AT[8 * r + c] *= rs[r]; }
      for (int r = 0; r < 8; ++r)
        for (int c = 0; c < 8; ++c) {
          adHost[64 * idx + r + 8 * c] = AH[8 * r + c];
          adTarget[64 * idx + r + 8 * c] = AT[8 * r + c];
        }
      // setDeltaF, EnergyFunctional.cpp:136-137 (float row-vector * float 8x8, summed in index order)
      for (int c = 0; c < 8; ++c) {
        float a = 0, b = 0;
        for (int r = 0; r < 8; ++r) {
          float dh = (float)(state[10 * h + r] - state_zero[10 * h + r]);
          float dt = (float)(state[10 * t + r] - state_zero[10 * t + r]);
          a += dh * (float)AH[8 * r + c];
          b += dt * (float)AT[8 * r + c];
        }
        adHTdeltaF[8 * idx + c] = a + b;
      }
    }
  return 0;
}

// ---------------------------------------------------------------- thread pool (IndexThreadReduce)
namespace {
class Pool {
 public:
  explicit Pool(int nt) : nt_(nt), isDone_(nt, true), gotOne_(nt, true) {
    for (int i = 0; i < nt; ++i) th_.emplace_back([this, i] { worker(i); });
  }
  ~Pool() {
    { std::unique_lock<std::mutex> l(m_); running_ = false; todo_.notify_all(); }
    for (auto& t : th_) t.join();
  }
  int size() const { return nt_; }
  // reduce(), IndexThreadReduce.h:70-131
  double reduce(const std::function<void(int, int, double*, int)>& fn, int first, int end, int step) {
    if (step == 0) step = ((end - first) + nt_ - 1) / nt_;
    std::unique_lock<std::mutex> l(m_);
    fn_ = fn; next_ = first; end_ = end; step_ = step; stat_ = 0;
    for (int i = 0; i < nt_; ++i) { isDone_[i] = false; gotOne_[i] = false; }
    todo_.notify_all();
    while (true) {
      done_.wait(l);
      bool all = true;
      for (int i = 0; i < nt_; ++i) all = all && isDone_[i];
      if (all) break;
    }
    next_ = 0; end_ = 0;
    return stat_;
  }

 private:
  void worker(int idx) {  // workerLoop, IndexThreadReduce.h:158-207
    std::unique_lock<std::mutex> l(m_);
    while (running_) {
      int todo = 0;
      bool got = false;
      if (next_ < end_) { todo = next_; next_ += step_; got = true; }
      if (got) {
        l.unlock();
        double s = 0;
        fn_(todo, std::min(todo + step_, end_), &s, idx);
        gotOne_[idx] = true;
        l.lock();
        stat_ += s;
      } else {
        if (!gotOne_[idx]) {
          l.unlock();
          double s = 0;
          fn_(0, 0, &s, idx);
          gotOne_[idx] = true;
          l.lock();
          stat_ += s;
        }
        isDone_[idx] = true;
        done_.notify_all();
        todo_.wait(l);
      }
    }
  }
  int nt_;
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable todo_, done_;
  std::function<void(int, int, double*, int)> fn_;
  int next_ = 0, end_ = 0, step_ = 1;
  bool running_ = true;
  std::vector<char> isDone_, gotOne_;
  double stat_ = 0;
};

// 3-tier accumulator of N scalars (AccumulatorApprox / XX / X semantics). For
// T=double the tiers are kept (harmless) so both precisions share the code.
template <class T, int N> struct Tier {
  T a[N], a1k[N], a1m[N];
  float numIn1, numIn1k, numIn1m;
  size_t num;
  void init() { std::memset(this, 0, sizeof(*this)); }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) {
      for (int i = 0; i < N; ++i) { a1k[i] = a[i] + a1k[i]; a[i] = 0; }
      numIn1k += numIn1; numIn1 = 0;
    }
    if (numIn1k > 1000 || force) {
      for (int i = 0; i < N; ++i) { a1m[i] = a1k[i] + a1m[i]; a1k[i] = 0; }
      numIn1m += numIn1k; numIn1k = 0;
    }
  }
  void counted() { numIn1++; shiftUp(false); }
  void finish() { shiftUp(true); num = (size_t)(numIn1 + numIn1k + numIn1m); }
};

// AccumulatorApprox: 55 (10x10 upper) + 30 (10x3) + 6 (3x3 upper) scalars that
// share ONE counter (the shiftUp of update() moves all three groups).
template <class T> struct AccApprox {
  Tier<T, 91> t;
  void init() { t.init(); }
  void update(const T* x4, const T* x6, const T* y4, const T* y6, T a, T b, T c) {  // MatrixAccumulators.h:754-856
    T x[10], y[10];
    for (int i = 0; i < 4; ++i) { x[i] = x4[i]; y[i] = y4[i]; }
    for (int i = 0; i < 6; ++i) { x[4 + i] = x6[i]; y[4 + i] = y6[i]; }
    int idx = 0;
    for (int r = 0; r < 10; ++r)
      for (int cc = r; cc < 10; ++cc, ++idx)
        t.a[idx] += a * x[cc] * x[r] + c * y[cc] * y[r] + b * (x[cc] * y[r] + y[cc] * x[r]);
    t.num++;
    t.counted();
  }
  void updateTopRight(const T* x4, const T* x6, const T* y4, const T* y6, T TR00, T TR10, T TR01, T TR11, T TR02,
                      T TR12) {  // :859-901
    T x[10], y[10];
    for (int i = 0; i < 4; ++i) { x[i] = x4[i]; y[i] = y4[i]; }
    for (int i = 0; i < 6; ++i) { x[4 + i] = x6[i]; y[4 + i] = y6[i]; }
    for (int r = 0; r < 10; ++r) {
      t.a[55 + 3 * r + 0] += x[r] * TR00 + y[r] * TR10;
      t.a[55 + 3 * r + 1] += x[r] * TR01 + y[r] * TR11;
      t.a[55 + 3 * r + 2] += x[r] * TR02 + y[r] * TR12;
    }
  }
  void updateBotRight(T a00, T a01, T a02, T a11, T a12, T a22) {  // :903-917
    t.a[85] += a00; t.a[86] += a01; t.a[87] += a02; t.a[88] += a11; t.a[89] += a12; t.a[90] += a22;
  }
  // finish(): H 13x13 (col-major == row-major, symmetric) as double, added into out
  void finish_add(double* H) {  // :619-651
    t.finish();
    int idx = 0;
    for (int r = 0; r < 10; ++r)
      for (int c = r; c < 10; ++c, ++idx) {
        H[r + 13 * c] += (double)t.a1m[idx];
        if (c != r) H[c + 13 * r] += (double)t.a1m[idx];
      }
    for (int r = 0; r < 10; ++r)
      for (int c = 0; c < 3; ++c) {
        H[r + 13 * (c + 10)] += (double)t.a1m[55 + 3 * r + c];
        H[(c + 10) + 13 * r] += (double)t.a1m[55 + 3 * r + c];
      }
    const int br[6][2] = {{10, 10}, {10, 11}, {10, 12}, {11, 11}, {11, 12}, {12, 12}};
    for (int k = 0; k < 6; ++k) {
      H[br[k][0] + 13 * br[k][1]] += (double)t.a1m[85 + k];
      if (br[k][0] != br[k][1]) H[br[k][1] + 13 * br[k][0]] += (double)t.a1m[85 + k];
    }
  }
};

struct BABase {
  virtual ~BABase() {}
  virtual int set_window(const orc_window*) = 0;
  virtual int set_topology(const hslam_point_soa*, const hslam_residual_soa*) = 0;
  virtual int update_points(const float*, const float*, const float*) = 0;
  virtual int set_threads(int) = 0;
  virtual int linearize(int, double*, int8_t*, float*, float*, float*, float*) = 0;
  virtual int apply_res() = 0;
  virtual int fix_linearization(const int32_t*, int) = 0;
  virtual int accumulate(int, double*, double*, int*) = 0;
  virtual int point_results(double*, double*, double*, double*, double*, double*, double*, double*, double*) = 0;
  virtual int residual_results(int8_t*, uint8_t*, double*, double*, double*) = 0;
  virtual int raw_top(int, double*) = 0;
  virtual int raw_sc(double*, double*, double*, double*, double*) = 0;
  virtual int resubstitute(const double*, float*) = 0;
  virtual int marginalize_points(const int32_t*, int, float, double*, double*, double*, double*, int*) = 0;
  virtual int calc_l_energy(double*) = 0;
  virtual int new_frame_energy_th(int, float*) = 0;
};

template <class T> struct BA : BABase {
  hslam_params prm;
  // window
  int n = 0, W = 0, Hh = 0;
  T fx, fy, cx, cy, fxi, fyi;
  std::vector<const float*> dI;
  std::vector<float> frameEnergyTH, precalc, adHTdeltaF;
  std::vector<double> adHost, adTarget, framePrior, frameDeltaPrior;
  float cDeltaF[4];
  double cPrior[4];
  // points
  int P = 0;
  std::vector<float> pu, pv, pidepth, pidepth_zero, pcolor, pweights, ppriorF, pdeltaF;
  std::vector<int> phost, res_begin, res_index;
  std::vector<T> HdiF, bdSumF, idepth_hessian, Hdd_accAF, bd_accAF, Hcd_accAF, Hdd_accLF, bd_accLF, Hcd_accLF;
  // residuals
  int R = 0;
  std::vector<int> rpoint, rtarget;
  std::vector<int8_t> state_state, state_NewState;
  std::vector<uint8_t> isLinearized, isActive;
  std::vector<double> state_energy, state_NewEnergy, state_NewEnergyWithOutlier;
  std::vector<T> J, JpJdF, res_toZeroF;
  std::vector<float> centerProjectedTo, projectedTo;
  // accumulators (per thread)
  int nt = 1;
  std::unique_ptr<Pool> pool;
  std::vector<std::vector<AccApprox<T>>> accTop[2];   // [mode][tid][n*n]
  std::vector<int> nresTop[2];
  std::vector<std::vector<Tier<T, 64>>> accD;
  std::vector<std::vector<Tier<T, 32>>> accE;
  std::vector<std::vector<Tier<T, 8>>> accEB;
  std::vector<Tier<T, 16>> accHcc;
  std::vector<Tier<T, 4>> accbc;

  explicit BA(const hslam_params* p) : prm(*p) {}

  int set_threads(int t) override {
    nt = t < 1 ? 1 : t;
    pool.reset(nt > 1 ? new Pool(nt) : nullptr);
    return 0;
  }

  int set_window(const orc_window* w) override {
    if (!w || w->n_frames < 1 || w->n_frames > HSLAM_MAX_FRAMES) return -1;
    n = w->n_frames; W = w->w; Hh = w->h;
    fx = w->fx; fy = w->fy; cx = w->cx; cy = w->cy; fxi = w->fxi; fyi = w->fyi;
    dI.assign(w->dI, w->dI + n);
    frameEnergyTH.assign(w->frameEnergyTH, w->frameEnergyTH + n);
    precalc.assign(w->precalc, w->precalc + (size_t)n * n * HSLAM_PRECALC_STRIDE);
    adHost.assign(w->adHost, w->adHost + (size_t)n * n * 64);
    adTarget.assign(w->adTarget, w->adTarget + (size_t)n * n * 64);
    adHTdeltaF.assign(w->adHTdeltaF, w->adHTdeltaF + (size_t)n * n * 8);
    std::memcpy(cDeltaF, w->cDeltaF, sizeof(cDeltaF));
    std::memcpy(cPrior, w->cPrior, sizeof(cPrior));
    framePrior.assign(w->framePrior, w->framePrior + (size_t)n * 8);
    frameDeltaPrior.assign(w->frameDeltaPrior, w->frameDeltaPrior + (size_t)n * 8);
    return 0;
  }

  int set_topology(const hslam_point_soa* p, const hslam_residual_soa* r) override {
    P = p->n_points; R = r->n_residuals;
    pu.assign(p->u, p->u + P); pv.assign(p->v, p->v + P);
    pidepth.assign(p->idepth, p->idepth + P); pidepth_zero.assign(p->idepth_zero, p->idepth_zero + P);
    pcolor.assign(p->color, p->color + 8 * (size_t)P); pweights.assign(p->weights, p->weights + 8 * (size_t)P);
    ppriorF.assign(p->priorF, p->priorF + P); pdeltaF.assign(p->deltaF, p->deltaF + P);
    phost.assign(p->host_idx, p->host_idx + P);
    res_begin.assign(p->res_begin, p->res_begin + P + 1);
    res_index.assign(p->res_index, p->res_index + res_begin[P]);
    rpoint.assign(r->point_idx, r->point_idx + R); rtarget.assign(r->target_idx, r->target_idx + R);
    state_state.assign(r->state_state, r->state_state + R);
    isLinearized.assign(r->is_linearized, r->is_linearized + R);
    isActive.assign(r->is_active, r->is_active + R);
    state_energy.assign(R, 0.0);
    if (r->state_energy) state_energy.assign(r->state_energy, r->state_energy + R);
    state_NewState.assign(R, HSLAM_RES_OUT);
    state_NewEnergy.assign(R, 0.0); state_NewEnergyWithOutlier.assign(R, -1.0);
    J.assign((size_t)R * 74, T(0)); JpJdF.assign((size_t)R * 8, T(0)); res_toZeroF.assign((size_t)R * 8, T(0));
    centerProjectedTo.assign((size_t)R * 3, 0.f); projectedTo.assign((size_t)R * 16, 0.f);
    for (auto* v : {&HdiF, &bdSumF, &idepth_hessian, &Hdd_accAF, &bd_accAF, &Hdd_accLF, &bd_accLF}) v->assign(P, T(0));
    Hcd_accAF.assign(4 * (size_t)P, T(0)); Hcd_accLF.assign(4 * (size_t)P, T(0));
    return 0;
  }

  int update_points(const float* idepth, const float* deltaF, const float* idepth_zero) override {
    if (idepth) pidepth.assign(idepth, idepth + P);
    if (deltaF) pdeltaF.assign(deltaF, deltaF + P);
    if (idepth_zero) pidepth_zero.assign(idepth_zero, idepth_zero + P);
    return 0;
  }

  // ---------------- PointFrameResidual::linearize, OptimizationClasses.cpp:43-233
  double linearize_one(int r) {
    state_NewEnergyWithOutlier[r] = -1;
    if (state_state[r] == HSLAM_RES_OOB) { state_NewState[r] = HSLAM_RES_OOB; return state_energy[r]; }
    int p = rpoint[r], hostI = phost[p], tgtI = rtarget[r];
    const float* pc = &precalc[(size_t)HSLAM_PRECALC_STRIDE * (hostI + tgtI * n)];
    T KRKi[9], Kt[3], R0[9], t0[3];
    for (int i = 0; i < 9; ++i) { KRKi[i] = pc[i]; R0[i] = pc[12 + i]; }
    for (int i = 0; i < 3; ++i) { Kt[i] = pc[9 + i]; t0[i] = pc[21 + i]; }
    T affLL0 = pc[24], affLL1 = pc[25], b0 = pc[26];
    const float* dIl = dI[tgtI];
    const float* color = &pcolor[8 * (size_t)p];
    const float* weights = &pweights[8 * (size_t)p];
    T* Jr = &J[(size_t)r * 74];
    T energyLeft = 0;
    T SCALE_IDEPTH = prm.scale_idepth, SCALE_F = prm.scale_f, SCALE_C = prm.scale_c;
    T d_xi_x[6], d_xi_y[6], d_C_x[4], d_C_y[4], d_d_x, d_d_y;
    {
      // projectPoint (full), DirectProjection.h:20-38
      T KliP[3] = {((T)pu[p] + 0 - cx) * fxi, ((T)pv[p] + 0 - cy) * fyi, (T)1};
      T ptp[3];
      mat3_vec<T>(R0, KliP[0], KliP[1], KliP[2], ptp);
      T idz = pidepth_zero[p];
      for (int k = 0; k < 3; ++k) ptp[k] = ptp[k] + t0[k] * idz;
      T drescale = (T)1.0 / ptp[2];
      T new_idepth = idz * drescale;
      if (!(drescale > 0)) { state_NewState[r] = HSLAM_RES_OOB; return state_energy[r]; }
      T u = ptp[0] * drescale, v = ptp[1] * drescale;
      T Ku = u * fx + cx, Kv = v * fy + cy;
      if (!(Ku > (T)1.1f && Kv > (T)1.1f && Ku < (T)(W - 3) && Kv < (T)(Hh - 3))) {
        state_NewState[r] = HSLAM_RES_OOB;
        return state_energy[r];
      }
      centerProjectedTo[3 * (size_t)r] = (float)Ku; centerProjectedTo[3 * (size_t)r + 1] = (float)Kv;
      centerProjectedTo[3 * (size_t)r + 2] = (float)new_idepth;
      d_d_x = drescale * (t0[0] - t0[2] * u) * SCALE_IDEPTH * fx;   // :85-86
      d_d_y = drescale * (t0[1] - t0[2] * v) * SCALE_IDEPTH * fy;
      d_C_x[2] = drescale * (R0[6] * u - R0[0]);                     // :89-97
      d_C_x[3] = fx * drescale * (R0[7] * u - R0[1]) * fyi;
      d_C_x[0] = KliP[0] * d_C_x[2];
      d_C_x[1] = KliP[1] * d_C_x[3];
      d_C_y[2] = fy * drescale * (R0[6] * v - R0[3]) * fxi;
      d_C_y[3] = drescale * (R0[7] * v - R0[4]);
      d_C_y[0] = KliP[0] * d_C_y[2];
      d_C_y[1] = KliP[1] * d_C_y[3];
      d_C_x[0] = (d_C_x[0] + u) * SCALE_F;                           // :99-107
      d_C_x[1] *= SCALE_F;
      d_C_x[2] = (d_C_x[2] + 1) * SCALE_C;
      d_C_x[3] *= SCALE_C;
      d_C_y[0] *= SCALE_F;
      d_C_y[1] = (d_C_y[1] + v) * SCALE_F;
      d_C_y[2] *= SCALE_C;
      d_C_y[3] = (d_C_y[3] + 1) * SCALE_C;
      d_xi_x[0] = new_idepth * fx;                                   // :109-121
      d_xi_x[1] = 0;
      d_xi_x[2] = -new_idepth * u * fx;
      d_xi_x[3] = -u * v * fx;
      d_xi_x[4] = (1 + u * u) * fx;
      d_xi_x[5] = -v * fx;
      d_xi_y[0] = 0;
      d_xi_y[1] = new_idepth * fy;
      d_xi_y[2] = -new_idepth * v * fy;
      d_xi_y[3] = -(1 + v * v) * fy;
      d_xi_y[4] = u * v * fy;
      d_xi_y[5] = u * fy;
    }
    for (int i = 0; i < 6; ++i) { Jr[8 + i] = d_xi_x[i]; Jr[14 + i] = d_xi_y[i]; }
    for (int i = 0; i < 4; ++i) { Jr[20 + i] = d_C_x[i]; Jr[24 + i] = d_C_y[i]; }
    Jr[28] = d_d_x; Jr[29] = d_d_y;
    T JIdxJIdx_00 = 0, JIdxJIdx_11 = 0, JIdxJIdx_10 = 0;
    T JabJIdx_00 = 0, JabJIdx_01 = 0, JabJIdx_10 = 0, JabJIdx_11 = 0;
    T JabJab_00 = 0, JabJab_01 = 0, JabJab_11 = 0;
    T wJI2_sum = 0;
    T huber = prm.huberTH, oth = prm.outlierTHSumComponent;
    T idp = pidepth[p];
    for (int idx = 0; idx < 8; ++idx) {
      // projectPoint (short), DirectProjection.h:12-18
      T ptp[3];
      mat3_vec<T>(KRKi, (T)pu[p] + PAT[idx][0], (T)pv[p] + PAT[idx][1], (T)1, ptp);
      for (int k = 0; k < 3; ++k) ptp[k] = ptp[k] + Kt[k] * idp;
      T Ku = ptp[0] / ptp[2], Kv = ptp[1] / ptp[2];
      if (!(Ku > (T)1.1f && Kv > (T)1.1f && Ku < (T)(W - 3) && Kv < (T)(Hh - 3))) {
        state_NewState[r] = HSLAM_RES_OOB;
        return state_energy[r];
      }
      projectedTo[16 * (size_t)r + 2 * idx] = (float)Ku;
      projectedTo[16 * (size_t)r + 2 * idx + 1] = (float)Kv;
      T hit[3];
      {  // getInterpolatedElement33, GlobalTypes.h:377-388
        int ix = (int)Ku, iy = (int)Kv;
        T dx = Ku - ix, dy = Kv - iy, dxdy = dx * dy;
        const float* bp = dIl + 3 * (ix + iy * W);
        T wBR = dxdy, wBL = dy - dxdy, wTR = dx - dxdy, wTL = 1 - dx - dy + dxdy;
        for (int c = 0; c < 3; ++c)
          hit[c] = wBR * (T)bp[3 * (1 + W) + c] + wBL * (T)bp[3 * W + c] + wTR * (T)bp[3 + c] + wTL * (T)bp[c];
      }
      T residual = hit[0] - (T)(affLL0 * (T)color[idx] + affLL1);
      T drdA = ((T)color[idx] - b0);
      if (!std::isfinite((float)hit[0])) { state_NewState[r] = HSLAM_RES_OOB; return state_energy[r]; }
      T w = std::sqrt(oth / (oth + (hit[1] * hit[1] + hit[2] * hit[2])));
      w = (T)0.5f * (w + (T)weights[idx]);
      T ar = std::fabs(residual);
      T hw = ar < huber ? (T)1 : huber / ar;
      energyLeft += w * w * hw * residual * residual * (2 - hw);
      if (hw < 1) hw = std::sqrt(hw);
      hw = hw * w;
      hit[1] *= hw; hit[2] *= hw;
      Jr[idx] = residual * hw;
      Jr[30 + idx] = hit[1]; Jr[38 + idx] = hit[2];
      Jr[46 + idx] = drdA * hw; Jr[54 + idx] = hw;
      JIdxJIdx_00 += hit[1] * hit[1]; JIdxJIdx_11 += hit[2] * hit[2]; JIdxJIdx_10 += hit[1] * hit[2];
      JabJIdx_00 += drdA * hw * hit[1]; JabJIdx_01 += drdA * hw * hit[2];
      JabJIdx_10 += hw * hit[1]; JabJIdx_11 += hw * hit[2];
      JabJab_00 += drdA * drdA * hw * hw; JabJab_01 += drdA * hw * hw; JabJab_11 += hw * hw;
      wJI2_sum += hw * hw * (hit[1] * hit[1] + hit[2] * hit[2]);
      if (prm.affineOptModeA < 0) Jr[46 + idx] = 0;
      if (prm.affineOptModeB < 0) Jr[54 + idx] = 0;
    }
    Jr[62] = JIdxJIdx_00; Jr[63] = JIdxJIdx_10; Jr[64] = JIdxJIdx_10; Jr[65] = JIdxJIdx_11;  // col-major 2x2
    Jr[66] = JabJIdx_00; Jr[67] = JabJIdx_10; Jr[68] = JabJIdx_01; Jr[69] = JabJIdx_11;
    Jr[70] = JabJab_00; Jr[71] = JabJab_01; Jr[72] = JabJab_01; Jr[73] = JabJab_11;
    state_NewEnergyWithOutlier[r] = (double)(float)energyLeft;
    float th = std::max<float>(frameEnergyTH[hostI], frameEnergyTH[tgtI]);
    if (energyLeft > (T)th || wJI2_sum < 2) {
      energyLeft = th;
      state_NewState[r] = HSLAM_RES_OUT;
    } else {
      state_NewState[r] = HSLAM_RES_IN;
    }
    state_NewEnergy[r] = (double)(float)energyLeft;
    return (double)(float)energyLeft;
  }

  void take_data(int r) {  // OptimizationClasses.h:155-161
    const T* Jr = &J[(size_t)r * 74];
    T JJ0 = Jr[62] * Jr[28] + Jr[64] * Jr[29];
    T JJ1 = Jr[63] * Jr[28] + Jr[65] * Jr[29];
    T* o = &JpJdF[(size_t)r * 8];
    for (int i = 0; i < 6; ++i) o[i] = Jr[8 + i] * JJ0 + Jr[14 + i] * JJ1;
    o[6] = Jr[66] * Jr[28] + Jr[68] * Jr[29];
    o[7] = Jr[67] * Jr[28] + Jr[69] * Jr[29];
  }
  void apply_one(int r) {  // OptimizationClasses.cpp:235-256 with copyJacobians=true
    if (state_state[r] == HSLAM_RES_OOB) return;
    if (state_NewState[r] == HSLAM_RES_IN) { isActive[r] = 1; take_data(r); } else { isActive[r] = 0; }
    state_state[r] = state_NewState[r];
    state_energy[r] = state_NewEnergy[r];
  }

  std::vector<int> active_list() {  // FullSystemOptimize.cpp:371-391: residuals with !isLinearized
    std::vector<int> a;
    for (int r = 0; r < R; ++r) if (!isLinearized[r]) a.push_back(r);
    return a;
  }

  int linearize(int fixLin, double* energy_sum, int8_t* newState, float* newEnergy, float* newEnergyWO,
                float* cpt, float* pto) override {
    std::vector<int> act = active_list();
    auto fn = [&](int mn, int mx, double* st, int) {  // linearizeAll_Reductor, FullSystemOptimize.cpp:19-52
      for (int k = mn; k < mx; ++k) {
        *st += linearize_one(act[k]);
        if (fixLin) apply_one(act[k]);
      }
    };
    double e = 0;
    if (pool) e = pool->reduce(fn, 0, (int)act.size(), 0);
    else fn(0, (int)act.size(), &e, 0);
    if (energy_sum) *energy_sum = e;
    for (int r = 0; r < R; ++r) {
      if (newState) newState[r] = state_NewState[r];
      if (newEnergy) newEnergy[r] = (float)state_NewEnergy[r];
      if (newEnergyWO) newEnergyWO[r] = (float)state_NewEnergyWithOutlier[r];
    }
    if (cpt) std::memcpy(cpt, centerProjectedTo.data(), sizeof(float) * 3 * (size_t)R);
    if (pto) std::memcpy(pto, projectedTo.data(), sizeof(float) * 16 * (size_t)R);
    return 0;
  }

  int apply_res() override {
    std::vector<int> act = active_list();
    auto fn = [&](int mn, int mx, double*, int) { for (int k = mn; k < mx; ++k) apply_one(act[k]); };
    double e = 0;
    if (pool) pool->reduce(fn, 0, (int)act.size(), 50);  // FullSystemOptimize.cpp:403
    else fn(0, (int)act.size(), &e, 0);
    return 0;
  }

  int fix_linearization(const int32_t* ids, int count) override {  // OptimizationClasses.cpp:258-284
    for (int k = 0; k < count; ++k) {
      int r = ids[k];
      if (r < 0 || r >= R) return -1;
      int p = rpoint[r];
      const float* dp = &adHTdeltaF[8 * (size_t)(phost[p] + n * rtarget[r])];
      const T* Jr = &J[(size_t)r * 74];
      T dx = 0, dy = 0;
      for (int i = 0; i < 6; ++i) { dx += Jr[8 + i] * (T)dp[i]; dy += Jr[14 + i] * (T)dp[i]; }
      T cx_ = 0, cy_ = 0;
      for (int i = 0; i < 4; ++i) { cx_ += Jr[20 + i] * (T)cDeltaF[i]; cy_ += Jr[24 + i] * (T)cDeltaF[i]; }
      T Jp_delta_x = dx + cx_ + Jr[28] * (T)pdeltaF[p];
      T Jp_delta_y = dy + cy_ + Jr[29] * (T)pdeltaF[p];
      T da = dp[6], db = dp[7];
      for (int i = 0; i < 8; ++i) {
        T rtz = Jr[i];
        rtz = rtz - Jr[30 + i] * Jp_delta_x;
        rtz = rtz - Jr[38 + i] * Jp_delta_y;
        rtz = rtz - Jr[46 + i] * da;
        rtz = rtz - Jr[54 + i] * db;
        res_toZeroF[(size_t)r * 8 + i] = rtz;
      }
      isLinearized[r] = 1;
    }
    return 0;
  }

  // ---------------- AccumulatedTopHessianSSE::addPoint<mode>, AccumulatedTopHessian.cpp:21-141
  void add_point_top(int mode, int p, int tid) {
    T dc[4] = {(T)cDeltaF[0], (T)cDeltaF[1], (T)cDeltaF[2], (T)cDeltaF[3]};
    T dd = pdeltaF[p];
    T bd_acc = 0, Hdd_acc = 0, Hcd_acc[4] = {0, 0, 0, 0};
    for (int k = res_begin[p]; k < res_begin[p + 1]; ++k) {
      int r = res_index[k];
      if (mode == 0) { if (isLinearized[r] || !isActive[r]) continue; }
      if (mode == 1) { if (!isLinearized[r] || !isActive[r]) continue; }
      if (mode == 2) { if (!isActive[r]) continue; }
      const T* Jr = &J[(size_t)r * 74];
      int htIDX = phost[p] + rtarget[r] * n;
      const float* dp = &adHTdeltaF[8 * (size_t)htIDX];
      T resApprox[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (mode == 0) for (int i = 0; i < 8; ++i) resApprox[i] = Jr[i];
      if (mode == 2) for (int i = 0; i < 8; ++i) resApprox[i] = res_toZeroF[(size_t)r * 8 + i];
      if (mode == 1) {
        T dx = 0, dy = 0, cxx = 0, cyy = 0;
        for (int i = 0; i < 6; ++i) { dx += Jr[8 + i] * (T)dp[i]; dy += Jr[14 + i] * (T)dp[i]; }
        for (int i = 0; i < 4; ++i) { cxx += Jr[20 + i] * dc[i]; cyy += Jr[24 + i] * dc[i]; }
        T Jp_delta_x = dx + cxx + Jr[28] * dd;
        T Jp_delta_y = dy + cyy + Jr[29] * dd;
        T da = dp[6], db = dp[7];
        for (int i = 0; i < 8; ++i) {
          T rtz = res_toZeroF[(size_t)r * 8 + i];
          rtz = rtz + Jr[30 + i] * Jp_delta_x;
          rtz = rtz + Jr[38 + i] * Jp_delta_y;
          rtz = rtz + Jr[46 + i] * da;
          rtz = rtz + Jr[54 + i] * db;
          resApprox[i] = rtz;
        }
      }
      T JI_r0 = 0, JI_r1 = 0, Jab_r0 = 0, Jab_r1 = 0, rr = 0;
      for (int i = 0; i < 8; ++i) {
        JI_r0 += resApprox[i] * Jr[30 + i];
        JI_r1 += resApprox[i] * Jr[38 + i];
        Jab_r0 += resApprox[i] * Jr[46 + i];
        Jab_r1 += resApprox[i] * Jr[54 + i];
        rr += resApprox[i] * resApprox[i];
      }
      int am = mode == 0 ? 0 : 1;
      AccApprox<T>& A = accTop[am][tid][htIDX];
      A.update(Jr + 20, Jr + 8, Jr + 24, Jr + 14, Jr[62], Jr[64], Jr[65]);
      A.updateBotRight(Jr[70], Jr[72], Jab_r0, Jr[73], Jab_r1, rr);
      A.updateTopRight(Jr + 20, Jr + 8, Jr + 24, Jr + 14, Jr[66], Jr[68], Jr[67], Jr[69], JI_r0, JI_r1);
      T Ji2_0 = Jr[62] * Jr[28] + Jr[64] * Jr[29];
      T Ji2_1 = Jr[63] * Jr[28] + Jr[65] * Jr[29];
      bd_acc += JI_r0 * Jr[28] + JI_r1 * Jr[29];
      Hdd_acc += Ji2_0 * Jr[28] + Ji2_1 * Jr[29];
      for (int i = 0; i < 4; ++i) Hcd_acc[i] += Jr[20 + i] * Ji2_0 + Jr[24 + i] * Ji2_1;
      nresTop[am][tid]++;
    }
    if (mode == 0) {
      Hdd_accAF[p] = Hdd_acc; bd_accAF[p] = bd_acc;
      for (int i = 0; i < 4; ++i) Hcd_accAF[4 * (size_t)p + i] = Hcd_acc[i];
    } else {
      Hdd_accLF[p] = Hdd_acc; bd_accLF[p] = bd_acc;
      for (int i = 0; i < 4; ++i) Hcd_accLF[4 * (size_t)p + i] = Hcd_acc[i];
    }
    if (mode == 2) {
      Hdd_accAF[p] = 0; bd_accAF[p] = 0;
      for (int i = 0; i < 4; ++i) Hcd_accAF[4 * (size_t)p + i] = 0;
    }
  }

  // ---------------- AccumulatedSCHessianSSE::addPoint, AccumulatedSCHessian.cpp:10-53
  void add_point_sc(int p, bool shiftPriorToZero, int tid) {
    int ngood = 0;
    for (int k = res_begin[p]; k < res_begin[p + 1]; ++k) if (isActive[res_index[k]]) ngood++;
    if (ngood == 0) { HdiF[p] = 0; bdSumF[p] = 0; idepth_hessian[p] = 0; return; }
    T H = Hdd_accAF[p] + Hdd_accLF[p] + (T)ppriorF[p];
    if (H < (T)1e-10) H = (T)1e-10;
    idepth_hessian[p] = H;
    HdiF[p] = (T)(1.0 / H);
    bdSumF[p] = bd_accAF[p] + bd_accLF[p];
    if (shiftPriorToZero) bdSumF[p] += (T)ppriorF[p] * (T)pdeltaF[p];
    T Hcd[4];
    for (int i = 0; i < 4; ++i) Hcd[i] = Hcd_accAF[4 * (size_t)p + i] + Hcd_accLF[4 * (size_t)p + i];
    T hdi = HdiF[p], bds = bdSumF[p];
    {  // accHcc.update(Hcd,Hcd,HdiF): A += (w*L)*R^T, col-major
      Tier<T, 16>& a = accHcc[tid];
      for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) a.a[r + 4 * c] += (hdi * Hcd[r]) * Hcd[c];
      a.counted();
      Tier<T, 4>& b = accbc[tid];
      T w = bds * hdi;
      for (int r = 0; r < 4; ++r) b.a[r] += w * Hcd[r];
      b.counted();
    }
    int n2 = n * n;
    for (int k1 = res_begin[p]; k1 < res_begin[p + 1]; ++k1) {
      int r1 = res_index[k1];
      if (!isActive[r1]) continue;
      int r1ht = phost[p] + rtarget[r1] * n;
      const T* j1 = &JpJdF[(size_t)r1 * 8];
      for (int k2 = res_begin[p]; k2 < res_begin[p + 1]; ++k2) {
        int r2 = res_index[k2];
        if (!isActive[r2]) continue;
        const T* j2 = &JpJdF[(size_t)r2 * 8];
        Tier<T, 64>& d = accD[tid][r1ht + rtarget[r2] * n2];
        for (int c = 0; c < 8; ++c) for (int r = 0; r < 8; ++r) d.a[r + 8 * c] += (hdi * j1[r]) * j2[c];
        d.counted();
      }
      Tier<T, 32>& e = accE[tid][r1ht];
      for (int c = 0; c < 4; ++c) for (int r = 0; r < 8; ++r) e.a[r + 8 * c] += (hdi * j1[r]) * Hcd[c];
      e.counted();
      Tier<T, 8>& eb = accEB[tid][r1ht];
      T w = hdi * bds;
      for (int r = 0; r < 8; ++r) eb.a[r] += w * j1[r];
      eb.counted();
    }
  }

  void for_points(const std::function<void(int, int)>& f) {  // red->reduce(..., 0, allPoints.size(), 50)
    auto fn = [&](int mn, int mx, double*, int tid) { for (int p = mn; p < mx; ++p) f(p, tid); };
    double e = 0;
    if (pool) pool->reduce(fn, 0, P, 50);
    else fn(0, P, &e, 0);
  }

  static void ad_mul(const double* A /*8x8 cm*/, const double* M /*8xk cm*/, int k, double* out /*8xk cm*/) {
    for (int c = 0; c < k; ++c)
      for (int r = 0; r < 8; ++r) {
        double s = 0;
        for (int i = 0; i < 8; ++i) s += A[r + 8 * i] * M[i + 8 * c];
        out[r + 8 * c] = s;
      }
  }
  static void mul_adT(const double* M /*8x8 cm*/, const double* A /*8x8 cm*/, double* out) {  // M * A^T
    for (int c = 0; c < 8; ++c)
      for (int r = 0; r < 8; ++r) {
        double s = 0;
        for (int i = 0; i < 8; ++i) s += M[r + 8 * i] * A[c + 8 * i];
        out[r + 8 * c] = s;
      }
  }

  int raw_top(int which, double* out) override {
    int am = which == 0 ? 0 : 1;
    if (accTop[am].empty()) return -1;
    std::memset(out, 0, sizeof(double) * 169 * (size_t)n * n);
    for (int k = 0; k < n * n; ++k)
      for (int t = 0; t < (int)accTop[am].size(); ++t) {
        AccApprox<T> c = accTop[am][t][k];
        c.finish_add(out + 169 * (size_t)k);
      }
    return 0;
  }
  template <int N> static void sum_tiers(std::vector<std::vector<Tier<T, N>>>& acc, int k, double* out) {
    for (int i = 0; i < N; ++i) out[i] = 0;
    for (auto& th : acc) {
      Tier<T, N> c = th[k];
      c.finish();
      for (int i = 0; i < N; ++i) out[i] += (double)c.a1m[i];
    }
  }
  int raw_sc(double* oD, double* oE, double* oEB, double* oHcc, double* obc) override {
    if (accD.empty()) return -1;
    for (int k = 0; k < n * n * n; ++k) sum_tiers<64>(accD, k, oD + 64 * (size_t)k);
    for (int k = 0; k < n * n; ++k) { sum_tiers<32>(accE, k, oE + 32 * (size_t)k); sum_tiers<8>(accEB, k, oEB + 8 * (size_t)k); }
    for (int i = 0; i < 16; ++i) oHcc[i] = 0;
    for (int i = 0; i < 4; ++i) obc[i] = 0;
    for (size_t t = 0; t < accHcc.size(); ++t) {
      Tier<T, 16> a = accHcc[t]; a.finish();
      Tier<T, 4> b = accbc[t]; b.finish();
      for (int i = 0; i < 16; ++i) oHcc[i] += (double)a.a1m[i];
      for (int i = 0; i < 4; ++i) obc[i] += (double)b.a1m[i];
    }
    return 0;
  }

  // stitchDoubleInternal + stitchDoubleMT, AccumulatedTopHessian.cpp:218-280, .h:69-117
  void stitch_top(int am, bool usePrior, double* H, double* b) {
    int N = 8 * n + 4;
    std::memset(H, 0, sizeof(double) * (size_t)N * N);
    std::memset(b, 0, sizeof(double) * N);
    std::vector<double> raw(169 * (size_t)n * n);
    raw_top(am, raw.data());
    auto Hat = [&](int r, int c) -> double& { return H[r + (size_t)N * c]; };
    for (int k = 0; k < n * n; ++k) {
      int h = k % n, t = k / n, hIdx = 4 + 8 * h, tIdx = 4 + 8 * t;
      const double* acc = &raw[169 * (size_t)k];
      auto A = [&](int r, int c) { return acc[r + 13 * c]; };
      const double* AdH = &adHost[64 * (size_t)k];
      const double* AdT = &adTarget[64 * (size_t)k];
      double M[64], T1[64], T2[64];
      for (int c = 0; c < 8; ++c) for (int r = 0; r < 8; ++r) M[r + 8 * c] = A(4 + r, 4 + c);
      ad_mul(AdH, M, 8, T1); mul_adT(T1, AdH, T2);
      for (int c = 0; c < 8; ++c) for (int r = 0; r < 8; ++r) Hat(hIdx + r, hIdx + c) += T2[r + 8 * c];
      mul_adT(T1, AdT, T2);
      for (int c = 0; c < 8; ++c) for (int r = 0; r < 8; ++r) Hat(hIdx + r, tIdx + c) += T2[r + 8 * c];
      ad_mul(AdT, M, 8, T1); mul_adT(T1, AdT, T2);
      for (int c = 0; c < 8; ++c) for (int r = 0; r < 8; ++r) Hat(tIdx + r, tIdx + c) += T2[r + 8 * c];
      double C[32], o[32];
      for (int c = 0; c < 4; ++c) for (int r = 0; r < 8; ++r) C[r + 8 * c] = A(4 + r, c);
      ad_mul(AdH, C, 4, o);
      for (int c = 0; c < 4; ++c) for (int r = 0; r < 8; ++r) Hat(hIdx + r, c) += o[r + 8 * c];
      ad_mul(AdT, C, 4, o);
      for (int c = 0; c < 4; ++c) for (int r = 0; r < 8; ++r) Hat(tIdx + r, c) += o[r + 8 * c];
      for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) Hat(r, c) += A(r, c);
      double v[8], ov[8];
      for (int r = 0; r < 8; ++r) v[r] = A(4 + r, 12);
      ad_mul(AdH, v, 1, ov);
      for (int r = 0; r < 8; ++r) b[hIdx + r] += ov[r];
      ad_mul(AdT, v, 1, ov);
      for (int r = 0; r < 8; ++r) b[tIdx + r] += ov[r];
      for (int r = 0; r < 4; ++r) b[r] += A(r, 12);
    }
    if (usePrior) {  // :269-279
      for (int i = 0; i < 4; ++i) { Hat(i, i) += cPrior[i]; b[i] += cPrior[i] * (double)cDeltaF[i]; }
      for (int h = 0; h < n; ++h)
        for (int i = 0; i < 8; ++i) {
          Hat(4 + 8 * h + i, 4 + 8 * h + i) += framePrior[8 * h + i];
          b[4 + 8 * h + i] += framePrior[8 * h + i] * frameDeltaPrior[8 * h + i];
        }
    }
    for (int h = 0; h < n; ++h) {  // .h:105-116
      int hIdx = 4 + 8 * h;
      for (int c = 0; c < 8; ++c) for (int r = 0; r < 4; ++r) Hat(r, hIdx + c) = Hat(hIdx + c, r);
      for (int t = h + 1; t < n; ++t) {
        int tIdx = 4 + 8 * t;
        for (int c = 0; c < 8; ++c) for (int r = 0; r < 8; ++r) Hat(hIdx + r, tIdx + c) += Hat(tIdx + c, hIdx + r);
        for (int c = 0; c < 8; ++c) for (int r = 0; r < 8; ++r) Hat(tIdx + r, hIdx + c) = Hat(hIdx + c, tIdx + r);
      }
    }
  }

  // AccumulatedSCHessian.cpp:54-133 + .h:70-111
  void stitch_sc(double* H, double* b) {
    int N = 8 * n + 4, n2 = n * n;
    std::memset(H, 0, sizeof(double) * (size_t)N * N);
    std::memset(b, 0, sizeof(double) * N);
    std::vector<double> D(64 * (size_t)n * n2), E(32 * (size_t)n2), EB(8 * (size_t)n2);
    double Hcc[16], bc[4];
    raw_sc(D.data(), E.data(), EB.data(), Hcc, bc);
    auto Hat = [&](int r, int c) -> double& { return H[r + (size_t)N * c]; };
    for (int k = 0; k < n2; ++k) {
      int i = k % n, j = k / n, iIdx = 4 + 8 * i, jIdx = 4 + 8 * j, ij = i + n * j;
      const double* AdHij = &adHost[64 * (size_t)ij];
      const double* AdTij = &adTarget[64 * (size_t)ij];
      double o[32], ov[8];
      ad_mul(AdHij, &E[32 * (size_t)ij], 4, o);
      for (int c = 0; c < 4; ++c) for (int r = 0; r < 8; ++r) Hat(iIdx + r, c) += o[r + 8 * c];
      ad_mul(AdTij, &E[32 * (size_t)ij], 4, o);
      for (int c = 0; c < 4; ++c) for (int r = 0; r < 8; ++r) Hat(jIdx + r, c) += o[r + 8 * c];
      ad_mul(AdHij, &EB[8 * (size_t)ij], 1, ov);
      for (int r = 0; r < 8; ++r) b[iIdx + r] += ov[r];
      ad_mul(AdTij, &EB[8 * (size_t)ij], 1, ov);
      for (int r = 0; r < 8; ++r) b[jIdx + r] += ov[r];
      for (int kk = 0; kk < n; ++kk) {
        int kIdx = 4 + 8 * kk, ijk = ij + kk * n2, ik = i + n * kk;
        const double* Dm = &D[64 * (size_t)ijk];
        const double* AdHik = &adHost[64 * (size_t)ik];
        const double* AdTik = &adTarget[64 * (size_t)ik];
        double T1[64], T2[64];
        ad_mul(AdHij, Dm, 8, T1);
        mul_adT(T1, AdHik, T2);
        for (int c = 0; c < 8; ++c) for (int r = 0; r < 8; ++r) Hat(iIdx + r, iIdx + c) += T2[r + 8 * c];
        mul_adT(T1, AdTik, T2);
        for (int c = 0; c < 8; ++c) for (int r = 0; r < 8; ++r) Hat(iIdx + r, kIdx + c) += T2[r + 8 * c];
        ad_mul(AdTij, Dm, 8, T1);
        mul_adT(T1, AdTik, T2);
        for (int c = 0; c < 8; ++c) for (int r = 0; r < 8; ++r) Hat(jIdx + r, kIdx + c) += T2[r + 8 * c];
        mul_adT(T1, AdHik, T2);
        for (int c = 0; c < 8; ++c) for (int r = 0; r < 8; ++r) Hat(jIdx + r, iIdx + c) += T2[r + 8 * c];
      }
    }
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) Hat(r, c) += Hcc[r + 4 * c];
    for (int r = 0; r < 4; ++r) b[r] += bc[r];
    for (int h = 0; h < n; ++h)
      for (int c = 0; c < 8; ++c) for (int r = 0; r < 4; ++r) Hat(r, 4 + 8 * h + c) = Hat(4 + 8 * h + c, r);
  }

  int accumulate(int which, double* H, double* b, int* nres) override {
    if (which == HSLAM_ACC_A || which == HSLAM_ACC_L) {
      int am = which;
      accTop[am].assign(nt, std::vector<AccApprox<T>>(n * n));
      nresTop[am].assign(nt, 0);
      for (auto& v : accTop[am]) for (auto& a : v) a.init();
      for_points([&](int p, int tid) { add_point_top(am, p, tid); });
      stitch_top(am, which == HSLAM_ACC_L, H, b);
      int s = 0;
      for (int v : nresTop[am]) s += v;
      if (nres) *nres = s;
      return 0;
    }
    if (which == HSLAM_ACC_SC) {
      accD.assign(nt, std::vector<Tier<T, 64>>(n * n * n));
      accE.assign(nt, std::vector<Tier<T, 32>>(n * n));
      accEB.assign(nt, std::vector<Tier<T, 8>>(n * n));
      accHcc.assign(nt, Tier<T, 16>()); accbc.assign(nt, Tier<T, 4>());
      for (int t = 0; t < nt; ++t) {
        for (auto& a : accD[t]) a.init();
        for (auto& a : accE[t]) a.init();
        for (auto& a : accEB[t]) a.init();
        accHcc[t].init(); accbc[t].init();
      }
      for_points([&](int p, int tid) { add_point_sc(p, true, tid); });
      stitch_sc(H, b);
      return 0;
    }
    return -1;
  }

  int point_results(double* o0, double* o1, double* o2, double* o3, double* o4, double* o5, double* o6, double* o7,
                    double* o8) override {
    for (int p = 0; p < P; ++p) {
      if (o0) o0[p] = HdiF[p]; if (o1) o1[p] = bdSumF[p]; if (o2) o2[p] = idepth_hessian[p];
      if (o3) o3[p] = Hdd_accAF[p]; if (o4) o4[p] = bd_accAF[p];
      if (o6) o6[p] = Hdd_accLF[p]; if (o7) o7[p] = bd_accLF[p];
      for (int i = 0; i < 4; ++i) {
        if (o5) o5[4 * (size_t)p + i] = Hcd_accAF[4 * (size_t)p + i];
        if (o8) o8[4 * (size_t)p + i] = Hcd_accLF[4 * (size_t)p + i];
      }
    }
    return 0;
  }
  int residual_results(int8_t* st, uint8_t* act, double* jp, double* j74, double* rtz) override {
    for (int r = 0; r < R; ++r) {
      if (st) st[r] = state_state[r];
      if (act) act[r] = isActive[r];
      if (jp) for (int i = 0; i < 8; ++i) jp[8 * (size_t)r + i] = JpJdF[8 * (size_t)r + i];
      if (j74) for (int i = 0; i < 74; ++i) j74[74 * (size_t)r + i] = J[74 * (size_t)r + i];
      if (rtz) for (int i = 0; i < 8; ++i) rtz[8 * (size_t)r + i] = res_toZeroF[8 * (size_t)r + i];
    }
    return 0;
  }

  // EnergyFunctional::resubstituteF_MT / resubstituteFPt, EnergyFunctional.cpp:222-274
  int resubstitute(const double* x, float* step_out) override {
    int N = 8 * n + 4;
    std::vector<T> xF(N);
    for (int i = 0; i < N; ++i) xF[i] = (T)(float)x[i];
    std::vector<T> xAd(8 * (size_t)n * n);
    for (int h = 0; h < n; ++h)
      for (int t = 0; t < n; ++t)
        for (int c = 0; c < 8; ++c) {
          T a = 0, b = 0;
          for (int r = 0; r < 8; ++r) {
            a += xF[4 + 8 * h + r] * (T)(float)adHost[64 * (size_t)(h + n * t) + r + 8 * c];
            b += xF[4 + 8 * t + r] * (T)(float)adTarget[64 * (size_t)(h + n * t) + r + 8 * c];
          }
          xAd[8 * (size_t)(n * h + t) + c] = a + b;
        }
    for_points([&](int p, int) {
      int ngood = 0;
      for (int k = res_begin[p]; k < res_begin[p + 1]; ++k) if (isActive[res_index[k]]) ngood++;
      if (ngood == 0) { step_out[p] = 0; return; }
      T b = bdSumF[p];
      T dot = 0;
      for (int i = 0; i < 4; ++i) dot += xF[i] * (Hcd_accAF[4 * (size_t)p + i] + Hcd_accLF[4 * (size_t)p + i]);
      b -= dot;
      for (int k = res_begin[p]; k < res_begin[p + 1]; ++k) {
        int r = res_index[k];
        if (!isActive[r]) continue;
        T d = 0;
        for (int i = 0; i < 8; ++i) d += xAd[8 * (size_t)(phost[p] * n + rtarget[r]) + i] * JpJdF[8 * (size_t)r + i];
        b -= d;
      }
      step_out[p] = (float)(-b * HdiF[p]);
    });
    return 0;
  }

  // EnergyFunctional::marginalizePointsF, Src/EnergyFunctional.cpp:545-609 (the accumulate + stitch part):
  // priorF *= setting_idepthFixPriorMargFac, addPoint<2> + SC addPoint(p, shiftPriorToZero=false) over the points
  // to marginalise on tid 0, stitchDouble(usePrior=false) of both.  The caller forms H = M - Msc, b = Mb - Mbsc.
  int marginalize_points(const int32_t* pids, int count, float prior_scale, double* M, double* Mb, double* Msc, double* Mbsc,
                         int* nres) override {
    accTop[1].assign(nt, std::vector<AccApprox<T>>(n * n));
    nresTop[1].assign(nt, 0);
    for (auto& v : accTop[1]) for (auto& a : v) a.init();
    accD.assign(nt, std::vector<Tier<T, 64>>(n * n * n));
    accE.assign(nt, std::vector<Tier<T, 32>>(n * n));
    accEB.assign(nt, std::vector<Tier<T, 8>>(n * n));
    accHcc.assign(nt, Tier<T, 16>()); accbc.assign(nt, Tier<T, 4>());
    for (int t = 0; t < nt; ++t) {
      for (auto& a : accD[t]) a.init();
      for (auto& a : accE[t]) a.init();
      for (auto& a : accEB[t]) a.init();
      accHcc[t].init(); accbc[t].init();
    }
    for (int i = 0; i < count; ++i) {
      const int p = pids[i];
      if (p < 0 || p >= P) return -1;
      ppriorF[p] *= prior_scale;
      add_point_top(2, p, 0);
      add_point_sc(p, false, 0);
    }
    stitch_top(1, false, M, Mb);
    stitch_sc(Msc, Mbsc);
    if (nres) *nres = nresTop[1][0];
    return 0;
  }

  // EnergyFunctional::calcLEnergyF_MT / calcLEnergyPt, Src/EnergyFunctional.cpp:289-364 (Accumulator11 is a
  // sequential float sum with 1k / 1M shift-up tiers; restated here as a plain sum in T)
  // System::setNewFrameEnergyTH, Src/FullSystemOptimize.cpp:60-101 (called by linearizeAll after every pass, :124)
  int new_frame_energy_th(int newest, float* th_out) override {
    if (newest < 0 || newest >= n) return -1;
    std::vector<float> allResVec;                       // System.h:64 vector<float>
    allResVec.reserve((size_t)R * 2);
    for (int r : active_list())                         // :67 activeResiduals = the non-linearised residuals
      if (state_NewEnergyWithOutlier[r] >= 0 && rtarget[r] == newest) allResVec.push_back((float)state_NewEnergyWithOutlier[r]);
    float th;
    if (allResVec.size() == 0) {
      th = 12 * 12 * HSLAM_PATTERN_NUM;                 // :73-77
    } else {
      int nthIdx = prm.frameEnergyTHN * allResVec.size();   // :80 float * size_t -> float -> int
      std::nth_element(allResVec.begin(), allResVec.begin() + nthIdx, allResVec.end());
      float nthElement = sqrtf(allResVec[nthIdx]);      // :86
      th = nthElement * prm.frameEnergyTHFacMedian;     // :88
      th = 26.0f * prm.frameEnergyTHConstWeight + th * (1 - prm.frameEnergyTHConstWeight);   // :89
      th = th * th;                                     // :90
      th *= prm.overallEnergyTHWeight * prm.overallEnergyTHWeight;   // :91
    }
    frameEnergyTH[newest] = th;
    if (th_out) *th_out = th;
    return 0;
  }

  int calc_l_energy(double* out) override {
    double E = 0;
    for (int f = 0; f < n; ++f)
      for (int i = 0; i < 8; ++i) E += frameDeltaPrior[8 * f + i] * framePrior[8 * f + i] * frameDeltaPrior[8 * f + i];
    for (int i = 0; i < 4; ++i) E += (double)((float)cDeltaF[i] * (float)cPrior[i] * (float)cDeltaF[i]);
    T acc = 0;
    T dc[4] = {(T)cDeltaF[0], (T)cDeltaF[1], (T)cDeltaF[2], (T)cDeltaF[3]};
    for (int p = 0; p < P; ++p) {
      T dd = pdeltaF[p];
      for (int k = res_begin[p]; k < res_begin[p + 1]; ++k) {
        int r = res_index[k];
        if (!isLinearized[r] || !isActive[r]) continue;
        const T* Jr = &J[(size_t)r * 74];
        const float* dp = &adHTdeltaF[8 * (size_t)(phost[p] + rtarget[r] * n)];
        T dx = 0, dy = 0, cxx = 0, cyy = 0;
        for (int i = 0; i < 6; ++i) { dx += Jr[8 + i] * (T)dp[i]; dy += Jr[14 + i] * (T)dp[i]; }
        for (int i = 0; i < 4; ++i) { cxx += Jr[20 + i] * dc[i]; cyy += Jr[24 + i] * dc[i]; }
        T jx = dx + cxx + Jr[28] * dd, jy = dy + cyy + Jr[29] * dd;
        for (int i = 0; i < 8; ++i) {
          T Jdelta = Jr[30 + i] * jx;
          Jdelta = Jdelta + Jr[38 + i] * jy;
          Jdelta = Jdelta + Jr[46 + i] * (T)dp[6];
          Jdelta = Jdelta + Jr[54 + i] * (T)dp[7];
          T r0 = res_toZeroF[(size_t)r * 8 + i];
          r0 = r0 + r0;
          r0 = r0 + Jdelta;
          acc += Jdelta * r0;
        }
      }
      acc += (T)pdeltaF[p] * (T)pdeltaF[p] * (T)ppriorF[p];
    }
    *out = E + (double)acc;
    return 0;
  }
};
}  // namespace

struct orc_ba { std::unique_ptr<BABase> impl; };

extern "C" {
orc_ba* orc_ba_create(int use_double, const hslam_params* prm) {
  orc_ba* b = new orc_ba;
  if (use_double) b->impl.reset(new BA<double>(prm));
  else b->impl.reset(new BA<float>(prm));
  return b;
}
void orc_ba_destroy(orc_ba* ba) { delete ba; }
int orc_ba_set_window(orc_ba* ba, const orc_window* w) { return ba->impl->set_window(w); }
int orc_ba_set_topology(orc_ba* ba, const hslam_point_soa* p, const hslam_residual_soa* r) { return ba->impl->set_topology(p, r); }
int orc_ba_update_points(orc_ba* ba, const float* a, const float* b, const float* c) { return ba->impl->update_points(a, b, c); }
int orc_ba_set_threads(orc_ba* ba, int t) { return ba->impl->set_threads(t); }
int orc_ba_linearize(orc_ba* ba, int fix, double* e, int8_t* ns, float* ne, float* nwo, float* c, float* p) {
  return ba->impl->linearize(fix, e, ns, ne, nwo, c, p);
}
int orc_ba_apply_res(orc_ba* ba) { return ba->impl->apply_res(); }
int orc_ba_fix_linearization(orc_ba* ba, const int32_t* ids, int count) { return ba->impl->fix_linearization(ids, count); }
int orc_ba_accumulate(orc_ba* ba, int which, double* H, double* b, int* nres) { return ba->impl->accumulate(which, H, b, nres); }
int orc_ba_point_results(orc_ba* ba, double* a, double* b, double* c, double* d, double* e, double* f, double* g,
                         double* h, double* i) {
  return ba->impl->point_results(a, b, c, d, e, f, g, h, i);
}
int orc_ba_residual_results(orc_ba* ba, int8_t* s, uint8_t* a, double* j, double* j74, double* rtz) {
  return ba->impl->residual_results(s, a, j, j74, rtz);
}
int orc_ba_raw_top(orc_ba* ba, int which, double* out) { return ba->impl->raw_top(which, out); }
int orc_ba_raw_sc(orc_ba* ba, double* D, double* E, double* EB, double* Hcc, double* bc) {
  return ba->impl->raw_sc(D, E, EB, Hcc, bc);
}
int orc_ba_resubstitute(orc_ba* ba, const double* x, float* step) { return ba->impl->resubstitute(x, step); }
int orc_ba_marginalize_points(orc_ba* ba, const int32_t* pids, int count, float prior_scale, double* M, double* Mb, double* Msc,
                              double* Mbsc, int* nres) {
  return ba->impl->marginalize_points(pids, count, prior_scale, M, Mb, Msc, Mbsc, nres);
}
int orc_ba_calc_l_energy(orc_ba* ba, double* out) { return ba->impl->calc_l_energy(out); }
int orc_ba_new_frame_energy_th(orc_ba* ba, int newest, float* th_out) { return ba->impl->new_frame_energy_th(newest, th_out); }
}
