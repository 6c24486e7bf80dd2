// oracle_tracker.cpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY; see hslam_oracle.h).
// Restates Frame::CreateDirPyrs, CoarseTracker::{makeK, makeCoarseDepthL0,
// calcRes, calcGSSSE, trackNewestCoarse} and Accumulator9 of the reference.
#include "hslam_oracle.h"
#include "oracle_math.hpp"
#include <vector>
#include <cstdio>
#include <cstdlib>

using namespace orc;

extern "C" {

void orc_se3_exp(const double xi[6], double T[12]) { se3_exp(xi, T); }
void orc_se3_log(const double T[12], double xi[6]) { se3_log(T, xi); }
void orc_se3_mul(const double A[12], const double B[12], double C[12]) { se3_mul(A, B, C); }
void orc_se3_inv(const double A[12], double Ai[12]) { se3_inv(A, Ai); }
void orc_se3_adj(const double T[12], double Adj[36]) { se3_adj(T, Adj); }
void orc_aff_from_to(float eF, float eT, double aF, double bF, double aT, double bT, double out2[2]) {
  aff_from_to(eF, eT, aF, bF, aT, bT, out2);
}
void orc_ldlt_solve(const double* A, const double* rhs, int n, double* x) { ldlt_solve(A, rhs, n, x); }

int64_t orc_pyramid_offset(int w, int h, int lvl) {
  int64_t off = 0;
  for (int l = 0; l < lvl; ++l) off += (int64_t)(w >> l) * (h >> l);
  return off;
}

// Frame::CreateDirPyrs, Src/Frame.cpp:104-181
int orc_pyramid(const float* img, int w, int h, int levels, const float* gammaBGrad256, float* dI_out,
                float* absgrad_out) {
  if (!img || !dI_out || levels < 1 || levels > HSLAM_MAX_PYR_LEVELS) return -1;
  for (int lvl = 0; lvl < levels; ++lvl) {
    int wl = w >> lvl, hl = h >> lvl;
    float* dI = dI_out + 3 * orc_pyramid_offset(w, h, lvl);
    float* ag = absgrad_out ? absgrad_out + orc_pyramid_offset(w, h, lvl) : nullptr;
    std::memset(dI, 0, sizeof(float) * 3 * (size_t)wl * hl);  // rows 0/h-1: reference leaves uninitialised
    if (ag) std::memset(ag, 0, sizeof(float) * (size_t)wl * hl);
    if (lvl == 0) {
      for (int i = 0; i < wl * hl; ++i) dI[3 * i] = img[i];  // Frame.cpp:118-119
    } else {
      int wlm1 = w >> (lvl - 1);
      const float* dIm = dI_out + 3 * orc_pyramid_offset(w, h, lvl - 1);
      for (int y = 0; y < hl; ++y)
        for (int x = 0; x < wl; ++x)  // Frame.cpp:137-140, left-to-right float sum
          dI[3 * (x + y * wl)] = 0.25f * (dIm[3 * (2 * x + 2 * y * wlm1)] + dIm[3 * (2 * x + 1 + 2 * y * wlm1)] +
                                          dIm[3 * (2 * x + 2 * y * wlm1 + wlm1)] +
                                          dIm[3 * (2 * x + 1 + 2 * y * wlm1 + wlm1)]);
    }
    int it = wl * (hl - 1);
    for (int idx = wl; idx < it; ++idx) {  // Frame.cpp:143-165
      float dx = 0.5f * (dI[3 * (idx + 1)] - dI[3 * (idx - 1)]);
      float dy = 0.5f * (dI[3 * (idx + wl)] - dI[3 * (idx - wl)]);
      if (!std::isfinite(dx)) dx = 0;
      if (!std::isfinite(dy)) dy = 0;
      dI[3 * idx + 1] = dx;
      dI[3 * idx + 2] = dy;
      if (ag) {
        float a = dx * dx + dy * dy;
        if (gammaBGrad256) {
          // PhotometricUndistorter::getBGradOnly: table lookup at (int)(color+0.5f) clamped to [5,250]
          int c = (int)(dI[3 * idx] + 0.5f);
          if (c < 5) c = 5;
          if (c > 250) c = 250;
          float gw = gammaBGrad256[c];
          a *= gw * gw;
        }
        ag[idx] = a;
      }
    }
  }
  return 0;
}

// CoarseTracker::makeK, Src/CoarseTracker.cpp:72-101
void orc_make_K(int levels, int w0, int h0, float fx0, float fy0, float cx0, float cy0, orc_trk_level* L) {
  L[0].w = w0; L[0].h = h0; L[0].fx = fx0; L[0].fy = fy0; L[0].cx = cx0; L[0].cy = cy0;
  for (int l = 1; l < levels; ++l) {
    L[l].w = w0 >> l; L[l].h = h0 >> l;
    L[l].fx = (float)(L[l - 1].fx * 0.5);
    L[l].fy = (float)(L[l - 1].fy * 0.5);
    L[l].cx = (float)((cx0 + 0.5) / ((int)1 << l) - 0.5);
    L[l].cy = (float)((cy0 + 0.5) / ((int)1 << l) - 0.5);
  }
  for (int l = 0; l < levels; ++l) {
    float K[9] = {L[l].fx, 0.0f, L[l].cx, 0.0f, L[l].fy, L[l].cy, 0.0f, 0.0f, 1.0f};
    std::memcpy(L[l].K, K, sizeof(K));
    inv3_eigen<float>(K, L[l].Ki);
  }
}

// CoarseTracker::makeCoarseDepthL0, Src/CoarseTracker.cpp:105-263
int orc_make_coarse_depth(int levels, int w0, int h0, int n_splat, const int* su, const int* sv,
                          const float* sidepth, const float* sweight, const float* dI_ref, float* const* pc_u,
                          float* const* pc_v, float* const* pc_idepth, float* const* pc_color, int* pc_n) {
  std::vector<std::vector<float>> idepth(levels), wsum(levels), wbak(levels);
  std::vector<int> w(levels), h(levels);
  for (int l = 0; l < levels; ++l) {
    w[l] = w0 >> l; h[l] = h0 >> l;
    idepth[l].assign((size_t)w[l] * h[l] + 1, 0.f);
    wsum[l].assign((size_t)w[l] * h[l] + 1, 0.f);
    wbak[l].assign((size_t)w[l] * h[l] + 1, 0.f);
  }
  for (int i = 0; i < n_splat; ++i) {  // :110-130
    int u = su[i], v = sv[i];
    if (u < 0 || v < 0 || u >= w[0] || v >= h[0]) return -1;
    idepth[0][u + w[0] * v] += sidepth[i] * sweight[i];
    wsum[0][u + w[0] * v] += sweight[i];
  }
  for (int lvl = 1; lvl < levels; ++lvl) {  // :132-157
    int wl = w[lvl], hl = h[lvl], wlm1 = w[lvl - 1];
    for (int y = 0; y < hl; ++y)
      for (int x = 0; x < wl; ++x) {
        int b = 2 * x + 2 * y * wlm1;
        idepth[lvl][x + y * wl] = idepth[lvl - 1][b] + idepth[lvl - 1][b + 1] + idepth[lvl - 1][b + wlm1] +
                                  idepth[lvl - 1][b + wlm1 + 1];
        wsum[lvl][x + y * wl] =
            wsum[lvl - 1][b] + wsum[lvl - 1][b + 1] + wsum[lvl - 1][b + wlm1] + wsum[lvl - 1][b + wlm1 + 1];
      }
  }
  for (int lvl = 0; lvl < 2 && lvl < levels; ++lvl) {  // :161-188 diagonal dilation
    int wh = w[lvl] * h[lvl] - w[lvl], wl = w[lvl];
    float* ws = wsum[lvl].data();
    float* wb = wbak[lvl].data();
    float* id = idepth[lvl].data();
    std::memcpy(wb, ws, sizeof(float) * (size_t)w[lvl] * h[lvl]);
    for (int i = w[lvl]; i < wh; ++i)
      if (wb[i] <= 0) {
        float sum = 0, num = 0, numn = 0;
        if (wb[i + 1 + wl] > 0) { sum += id[i + 1 + wl]; num += wb[i + 1 + wl]; numn++; }
        if (wb[i - 1 - wl] > 0) { sum += id[i - 1 - wl]; num += wb[i - 1 - wl]; numn++; }
        if (wb[i + wl - 1] > 0) { sum += id[i + wl - 1]; num += wb[i + wl - 1]; numn++; }
        if (wb[i - wl + 1] > 0) { sum += id[i - wl + 1]; num += wb[i - wl + 1]; numn++; }
        if (numn > 0) { id[i] = sum / numn; ws[i] = num / numn; }
      }
  }
  for (int lvl = 2; lvl < levels; ++lvl) {  // :192-213 4-neighbour dilation
    int wh = w[lvl] * h[lvl] - w[lvl], wl = w[lvl];
    float* ws = wsum[lvl].data();
    float* wb = wbak[lvl].data();
    float* id = idepth[lvl].data();
    std::memcpy(wb, ws, sizeof(float) * (size_t)w[lvl] * h[lvl]);
    for (int i = w[lvl]; i < wh; ++i)
      if (wb[i] <= 0) {
        float sum = 0, num = 0, numn = 0;
        if (wb[i + 1] > 0) { sum += id[i + 1]; num += wb[i + 1]; numn++; }
        if (wb[i - 1] > 0) { sum += id[i - 1]; num += wb[i - 1]; numn++; }
        if (wb[i + wl] > 0) { sum += id[i + wl]; num += wb[i + wl]; numn++; }
        if (wb[i - wl] > 0) { sum += id[i - wl]; num += wb[i - wl]; numn++; }
        if (numn > 0) { id[i] = sum / numn; ws[i] = num / numn; }
      }
  }
  for (int lvl = 0; lvl < levels; ++lvl) {  // :217-261
    const float* dIr = dI_ref + 3 * orc_pyramid_offset(w0, h0, lvl);
    int wl = w[lvl], hl = h[lvl], n = 0;
    for (int y = 2; y < hl - 2; ++y)
      for (int x = 2; x < wl - 2; ++x) {
        int i = x + y * wl;
        if (wsum[lvl][i] > 0) {
          float idv = idepth[lvl][i] / wsum[lvl][i];
          float col = dIr[3 * i];
          if (!std::isfinite(col) || !(idv > 0)) continue;
          pc_u[lvl][n] = (float)x; pc_v[lvl][n] = (float)y; pc_idepth[lvl][n] = idv; pc_color[lvl][n] = col;
          n++;
        }
      }
    pc_n[lvl] = n;
  }
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------ calcRes
// getInterpolatedElement33, Include/GlobalTypes.h:377-388 (Vec3f AoS image)
template <class T> static inline void interp33(const float* mat, T x, T y, int width, T* out) {
  int ix = (int)x, iy = (int)y;
  T dx = x - ix, dy = y - iy, dxdy = dx * dy;
  const float* bp = mat + 3 * (ix + iy * width);
  T wBR = dxdy, wBL = dy - dxdy, wTR = dx - dxdy, wTL = 1 - dx - dy + dxdy;
  for (int c = 0; c < 3; ++c)
    out[c] = wBR * (T)bp[3 * (1 + width) + c] + wBL * (T)bp[3 * width + c] + wTR * (T)bp[3 + c] + wTL * (T)bp[c];
}

template <class T>
static int calc_res_impl(const orc_trk_level* Ls, int lvl, const float* dINewl, const orc_trk_frame_pair* fp,
                         const hslam_params* prm, const double* refToNew, double aff_a, double aff_b,
                         float cutoffTH_in, double* out6, T* const* warped, int* n_warped, uint8_t* mask) {
  const orc_trk_level& L = Ls[lvl];
  T E = 0;
  int numTermsInE = 0, numTermsInWarped = 0, numSaturated = 0;
  int wl = L.w, hl = L.h;
  T fxl = L.fx, fyl = L.fy, cxl = L.cx, cyl = L.cy;
  T R[9], Ki[9], RKi[9], t[3];
  for (int i = 0; i < 9; ++i) { R[i] = (T)refToNew[i]; Ki[i] = (T)L.Ki[i]; }
  mat3_mul<T>(R, Ki, RKi);  // CoarseTracker.cpp:345
  for (int i = 0; i < 3; ++i) t[i] = (T)refToNew[9 + i];
  double aff2[2];
  aff_from_to(fp->ref_exposure, fp->new_exposure, fp->ref_aff_a, fp->ref_aff_b, aff_a, aff_b, aff2);
  T affLL0 = (T)aff2[0], affLL1 = (T)aff2[1];
  T sumSquaredShiftT = 0, sumSquaredShiftRT = 0, sumSquaredShiftNum = 0;
  T huber = (T)prm->huberTH, cutoffTH = (T)cutoffTH_in;
  T maxEnergy = 2 * huber * cutoffTH - huber * huber;  // :354
  int nl = L.n;
  for (int i = 0; i < nl; ++i) {
    T id = L.pc_idepth[i], x = L.pc_u[i], y = L.pc_v[i];
    if (mask) mask[i] = 0;
    T pt[3];
    mat3_vec<T>(RKi, x, y, (T)1, pt);
    for (int k = 0; k < 3; ++k) pt[k] = pt[k] + t[k] * id;
    T u = pt[0] / pt[2], v = pt[1] / pt[2];
    T Ku = fxl * u + cxl, Kv = fyl * v + cyl;
    T new_idepth = id / pt[2];
    if (lvl == 0 && i % 32 == 0) {  // :384-415
      T p2[3], q[3];
      mat3_vec<T>(Ki, x, y, (T)1, p2);
      for (int k = 0; k < 3; ++k) q[k] = p2[k] + t[k] * id;
      T KuT = fxl * (q[0] / q[2]) + cxl, KvT = fyl * (q[1] / q[2]) + cyl;
      for (int k = 0; k < 3; ++k) q[k] = p2[k] - t[k] * id;
      T KuT2 = fxl * (q[0] / q[2]) + cxl, KvT2 = fyl * (q[1] / q[2]) + cyl;
      mat3_vec<T>(RKi, x, y, (T)1, p2);
      for (int k = 0; k < 3; ++k) q[k] = p2[k] - t[k] * id;
      T Ku3 = fxl * (q[0] / q[2]) + cxl, Kv3 = fyl * (q[1] / q[2]) + cyl;
      sumSquaredShiftT += (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y);
      sumSquaredShiftT += (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
      sumSquaredShiftRT += (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y);
      sumSquaredShiftRT += (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
      sumSquaredShiftNum += 2;
    }
    if (!(Ku > 2 && Kv > 2 && Ku < wl - 3 && Kv < hl - 3 && new_idepth > 0)) continue;  // :417
    T refColor = L.pc_color[i];
    T hit[3];
    interp33<T>(dINewl, Ku, Kv, wl, hit);
    if (!std::isfinite((float)hit[0])) continue;
    T residual = hit[0] - (T)(affLL0 * refColor + affLL1);
    T ar = std::fabs(residual);
    T hw = ar < huber ? (T)1 : huber / ar;
    if (ar > cutoffTH) {
      E += maxEnergy; numTermsInE++; numSaturated++;
      if (mask) mask[i] = 2;
    } else {
      E += hw * residual * residual * (2 - hw);
      numTermsInE++;
      if (warped) {
        warped[0][numTermsInWarped] = new_idepth; warped[1][numTermsInWarped] = u; warped[2][numTermsInWarped] = v;
        warped[3][numTermsInWarped] = hit[1]; warped[4][numTermsInWarped] = hit[2];
        warped[5][numTermsInWarped] = residual; warped[6][numTermsInWarped] = hw;
        warped[7][numTermsInWarped] = refColor;
      }
      numTermsInWarped++;
      if (mask) mask[i] = 1;
    }
  }
  if (n_warped) *n_warped = numTermsInWarped;
  if (warped) {
    int k = numTermsInWarped;
    while (k % 4 != 0) {  // :454-465
      for (int a = 0; a < 8; ++a) warped[a][k] = 0;
      k++;
    }
  }
  out6[0] = (double)E;
  out6[1] = numTermsInE;
  out6[2] = (double)(sumSquaredShiftT / (sumSquaredShiftNum + 0.1));
  out6[3] = 0;
  out6[4] = (double)(sumSquaredShiftRT / (sumSquaredShiftNum + 0.1));
  out6[5] = (double)(numSaturated / (float)numTermsInE);
  return 0;
}

// ------------------------------------------------------------------ calcGSSSE
// Accumulator9 (Include/MatrixAccumulators.h:982-1345): 45 upper-tri entries x 4
// SSE lanes, shiftUp every >1000 updates into the 1k tier, >1000 of those into 1M.
namespace {
struct Acc9f {
  float d[45][4], d1k[45][4], d1m[45][4];
  float numIn1, numIn1k, numIn1m;
  void init() { std::memset(this, 0, sizeof(*this)); }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) {
      for (int i = 0; i < 45; ++i) for (int l = 0; l < 4; ++l) { d1k[i][l] = d[i][l] + d1k[i][l]; d[i][l] = 0; }
      numIn1k += numIn1; numIn1 = 0;
    }
    if (numIn1k > 1000 || force) {
      for (int i = 0; i < 45; ++i) for (int l = 0; l < 4; ++l) { d1m[i][l] = d1k[i][l] + d1m[i][l]; d1k[i][l] = 0; }
      numIn1m += numIn1k; numIn1k = 0;
    }
  }
  // updateSSE_eighted, :1091-1166 : entry(r,c) += (J[r]*w)*J[c]
  void update4(const float J[9][4], const float w[4]) {
    int idx = 0;
    for (int r = 0; r < 9; ++r) {
      float Jw[4];
      for (int l = 0; l < 4; ++l) Jw[l] = J[r][l] * w[l];
      for (int c = r; c < 9; ++c, ++idx)
        for (int l = 0; l < 4; ++l) d[idx][l] = d[idx][l] + Jw[l] * J[c][l];
    }
    numIn1++;
    shiftUp(false);
  }
  void finish(float H[81]) {  // :1001-1017
    shiftUp(true);
    int idx = 0;
    for (int r = 0; r < 9; ++r)
      for (int c = r; c < 9; ++c, ++idx) {
        float v = d1m[idx][0] + d1m[idx][1] + d1m[idx][2] + d1m[idx][3];
        H[r * 9 + c] = H[c * 9 + r] = v;
      }
  }
};
}  // namespace

static void scale_Hb(const hslam_params* prm, double* H, double* b) {
  // CoarseTracker.cpp:312-323 (cols then rows; idx 0-2 x SCALE_XI_ROT, 3-5 x SCALE_XI_TRANS, 6 x A, 7 x B)
  double s[8] = {prm->scale_xi_rot, prm->scale_xi_rot, prm->scale_xi_rot, prm->scale_xi_trans,
                 prm->scale_xi_trans, prm->scale_xi_trans, prm->scale_a, prm->scale_b};
  for (int c = 0; c < 8; ++c)
    for (int r = 0; r < 8; ++r) H[r + 8 * c] *= s[c];
  for (int r = 0; r < 8; ++r)
    for (int c = 0; c < 8; ++c) H[r + 8 * c] *= s[r];
  for (int r = 0; r < 8; ++r) b[r] *= s[r];
}

template <class T>
static int calc_gs_impl(const orc_trk_level* L, const orc_trk_frame_pair* fp, const hslam_params* prm, double aff_a,
                        double aff_b, const T* const* warped, int n_unpadded, double* H, double* b);

template <>
int calc_gs_impl<float>(const orc_trk_level* L, const orc_trk_frame_pair* fp, const hslam_params* prm, double aff_a,
                        double aff_b, const float* const* W, int n_unpadded, double* H, double* b) {
  static thread_local Acc9f acc;
  acc.init();
  float fxl = L->fx, fyl = L->fy;
  float b0 = (float)fp->ref_aff_b;
  double aff2[2];
  aff_from_to(fp->ref_exposure, fp->new_exposure, fp->ref_aff_a, fp->ref_aff_b, aff_a, aff_b, aff2);
  float a = (float)aff2[0];
  int n = (n_unpadded + 3) & ~3;
  for (int i = 0; i < n; i += 4) {
    float J[9][4], w[4];
    for (int l = 0; l < 4; ++l) {
      float dx = W[3][i + l] * fxl, dy = W[4][i + l] * fyl;
      float u = W[1][i + l], v = W[2][i + l], id = W[0][i + l];
      J[0][l] = id * dx;
      J[1][l] = id * dy;
      J[2][l] = 0.0f - id * (u * dx + v * dy);
      J[3][l] = 0.0f - ((u * v) * dx + dy * (1.0f + v * v));
      J[4][l] = (u * v) * dy + dx * (1.0f + u * u);
      J[5][l] = u * dy - v * dx;
      J[6][l] = a * (b0 - W[7][i + l]);
      J[7][l] = -1.0f;
      J[8][l] = W[5][i + l];
      w[l] = W[6][i + l];
    }
    acc.update4(J, w);
  }
  float Hf[81];
  acc.finish(Hf);
  float inv = 1.0f / n;  // CoarseTracker.cpp:309-310 (float reciprocal, promoted)
  for (int r = 0; r < 8; ++r) {
    for (int c = 0; c < 8; ++c) H[r + 8 * c] = (double)Hf[r * 9 + c] * inv;
    b[r] = (double)Hf[r * 9 + 8] * inv;
  }
  scale_Hb(prm, H, b);
  return 0;
}

template <>
int calc_gs_impl<double>(const orc_trk_level* L, const orc_trk_frame_pair* fp, const hslam_params* prm, double aff_a,
                         double aff_b, const double* const* W, int n_unpadded, double* H, double* b) {
  double acc[9][9];
  std::memset(acc, 0, sizeof(acc));
  double fxl = L->fx, fyl = L->fy;
  double b0 = (double)(float)fp->ref_aff_b;
  double aff2[2];
  aff_from_to(fp->ref_exposure, fp->new_exposure, fp->ref_aff_a, fp->ref_aff_b, aff_a, aff_b, aff2);
  double a = (double)(float)aff2[0];
  int n = (n_unpadded + 3) & ~3;
  for (int i = 0; i < n_unpadded; ++i) {
    double dx = W[3][i] * fxl, dy = W[4][i] * fyl, u = W[1][i], v = W[2][i], id = W[0][i];
    double J[9] = {id * dx, id * dy, -id * (u * dx + v * dy), -(u * v * dx + dy * (1 + v * v)),
                   u * v * dy + dx * (1 + u * u), u * dy - v * dx, a * (b0 - W[7][i]), -1.0, W[5][i]};
    double w = W[6][i];
    for (int r = 0; r < 9; ++r)
      for (int c = r; c < 9; ++c) acc[r][c] += J[r] * w * J[c];
  }
  double inv = (double)(1.0f / n);
  for (int r = 0; r < 8; ++r) {
    for (int c = 0; c < 8; ++c) H[r + 8 * c] = (r <= c ? acc[r][c] : acc[c][r]) * inv;
    b[r] = acc[r][8] * inv;
  }
  scale_Hb(prm, H, b);
  return 0;
}

// ------------------------------------------------------------------ trackNewestCoarse
template <class T>
static int track_impl(const orc_trk_level* Ls, int n_levels, const float* dI_new, const orc_trk_frame_pair* fp,
                      const hslam_params* prm, double* refToNew_io, double* aff_io, int coarsestLvl,
                      const double* minResForAbort, double* lastResiduals, double* lastFlow, int* ok,
                      int* iterations, double* trace, int trace_cap, int* trace_n) {
  if (coarsestLvl >= 5 || coarsestLvl >= n_levels) return -1;  // assert at :511
  for (int i = 0; i < 5; ++i) lastResiduals[i] = NAN;
  for (int i = 0; i < 3; ++i) lastFlow[i] = 1000;
  if (iterations) for (int i = 0; i < 5; ++i) iterations[i] = 0;
  int tn = 0;
  auto tr = [&](int lvl, double e, double acc, double lam) {
    if (trace && tn < trace_cap) { trace[4 * tn] = lvl; trace[4 * tn + 1] = e; trace[4 * tn + 2] = acc; trace[4 * tn + 3] = lam; }
    tn++;
  };
  const float lambdaExtrapolationLimit = 0.001f;
  double cur[12];
  std::memcpy(cur, refToNew_io, sizeof(cur));
  double aff_cur[2] = {aff_io[0], aff_io[1]};
  bool haveRepeated = false;
  int w0 = Ls[0].w, h0 = Ls[0].h;
  size_t maxn = 0;
  for (int l = 0; l < n_levels; ++l) maxn = std::max(maxn, (size_t)Ls[l].n);
  std::vector<T> wbuf(8 * (maxn + 8));
  T* warped[8];
  for (int a = 0; a < 8; ++a) warped[a] = wbuf.data() + a * (maxn + 8);
  *ok = 0;
  for (int lvl = coarsestLvl; lvl >= 0; lvl--) {
    const float* dIl = dI_new + 3 * orc_pyramid_offset(w0, h0, lvl);
    double H[64], b[8];
    float levelCutoffRepeat = 1;
    double resOld[6], resNew[6];
    int nw = 0;
    calc_res_impl<T>(Ls, lvl, dIl, fp, prm, cur, aff_cur[0], aff_cur[1], prm->coarseCutoffTH * levelCutoffRepeat,
                     resOld, warped, &nw, nullptr);
    while (resOld[5] > 0.6 && levelCutoffRepeat < 50) {
      levelCutoffRepeat *= 2;
      calc_res_impl<T>(Ls, lvl, dIl, fp, prm, cur, aff_cur[0], aff_cur[1], prm->coarseCutoffTH * levelCutoffRepeat,
                       resOld, warped, &nw, nullptr);
    }
    calc_gs_impl<T>(&Ls[lvl], fp, prm, aff_cur[0], aff_cur[1], warped, nw, H, b);
    tr(lvl, resOld[0] / resOld[1], -1, 0.01);
    float lambda = 0.01f;
    for (int iteration = 0; iteration < prm->max_tracker_iterations[lvl]; iteration++) {
      if (iterations) iterations[lvl]++;
      double Hl[64];
      std::memcpy(Hl, H, sizeof(Hl));
      for (int i = 0; i < 8; ++i) Hl[i + 8 * i] *= (1 + lambda);
      double nb[8], inc[8];
      for (int i = 0; i < 8; ++i) nb[i] = -b[i];
      ldlt_solve(Hl, nb, 8, inc);
      bool fixA = prm->affineOptModeA < 0, fixB = prm->affineOptModeB < 0;
      if (fixA && fixB) {  // :565-569
        double H6[36], x6[6];
        for (int c = 0; c < 6; ++c) for (int r = 0; r < 6; ++r) H6[r + 6 * c] = Hl[r + 8 * c];
        ldlt_solve(H6, nb, 6, x6);
        for (int i = 0; i < 6; ++i) inc[i] = x6[i];
        inc[6] = inc[7] = 0;
      }
      if (!fixA && fixB) {  // :570-574
        double H7[49], x7[7];
        for (int c = 0; c < 7; ++c) for (int r = 0; r < 7; ++r) H7[r + 7 * c] = Hl[r + 8 * c];
        ldlt_solve(H7, nb, 7, x7);
        for (int i = 0; i < 7; ++i) inc[i] = x7[i];
        inc[7] = 0;
      }
      if (fixA && !fixB) {  // :575-587
        double Hs[64], bs[8], H7[49], nb7[7], x7[7];
        std::memcpy(Hs, Hl, sizeof(Hs));
        std::memcpy(bs, b, sizeof(bs));
        for (int r = 0; r < 8; ++r) Hs[r + 8 * 6] = Hs[r + 8 * 7];
        for (int c = 0; c < 8; ++c) Hs[6 + 8 * c] = Hs[7 + 8 * c];
        bs[6] = bs[7];
        for (int c = 0; c < 7; ++c) for (int r = 0; r < 7; ++r) H7[r + 7 * c] = Hs[r + 8 * c];
        for (int i = 0; i < 7; ++i) nb7[i] = -bs[i];
        ldlt_solve(H7, nb7, 7, x7);
        for (int i = 0; i < 8; ++i) inc[i] = 0;
        for (int i = 0; i < 6; ++i) inc[i] = x7[i];
        inc[7] = x7[6];
      }
      float extrapFac = 1;
      if (lambda < lambdaExtrapolationLimit) extrapFac = (float)std::sqrt(std::sqrt(lambdaExtrapolationLimit / lambda));
      for (int i = 0; i < 8; ++i) inc[i] *= extrapFac;
      double incScaled[8];
      for (int i = 0; i < 3; ++i) incScaled[i] = inc[i] * prm->scale_xi_rot;
      for (int i = 3; i < 6; ++i) incScaled[i] = inc[i] * prm->scale_xi_trans;
      incScaled[6] = inc[6] * prm->scale_a;
      incScaled[7] = inc[7] * prm->scale_b;
      double s = 0;
      for (int i = 0; i < 8; ++i) s += incScaled[i];
      if (!std::isfinite(s)) for (int i = 0; i < 8; ++i) incScaled[i] = 0;
      double Texp[12], newT[12];
      se3_exp(incScaled, Texp);
      se3_mul(Texp, cur, newT);
      double aff_new[2] = {aff_cur[0] + incScaled[6], aff_cur[1] + incScaled[7]};
      int nw_new = 0;
      calc_res_impl<T>(Ls, lvl, dIl, fp, prm, newT, aff_new[0], aff_new[1], prm->coarseCutoffTH * levelCutoffRepeat,
                       resNew, warped, &nw_new, nullptr);
      bool accept = (resNew[0] / resNew[1]) < (resOld[0] / resOld[1]);
      tr(lvl, resNew[0] / resNew[1], accept ? 1 : 0, lambda);
      if (accept) {
        calc_gs_impl<T>(&Ls[lvl], fp, prm, aff_new[0], aff_new[1], warped, nw_new, H, b);
        std::memcpy(resOld, resNew, sizeof(resOld));
        aff_cur[0] = aff_new[0]; aff_cur[1] = aff_new[1];
        std::memcpy(cur, newT, sizeof(cur));
        lambda *= 0.5;
      } else {
        lambda *= 4;
        if (lambda < lambdaExtrapolationLimit) lambda = lambdaExtrapolationLimit;
      }
      double nrm = 0;
      for (int i = 0; i < 8; ++i) nrm += inc[i] * inc[i];
      nrm = std::sqrt(nrm);
      if (!(nrm > 1e-3)) break;
    }
    lastResiduals[lvl] = sqrtf((float)(resOld[0] / resOld[1]));
    lastFlow[0] = resOld[2]; lastFlow[1] = resOld[3]; lastFlow[2] = resOld[4];
    if (lastResiduals[lvl] > 1.5 * minResForAbort[lvl]) { if (trace_n) *trace_n = tn; return 0; }
    if (levelCutoffRepeat > 1 && !haveRepeated) { lvl++; haveRepeated = true; }
  }
  std::memcpy(refToNew_io, cur, sizeof(cur));
  aff_io[0] = aff_cur[0]; aff_io[1] = aff_cur[1];
  if (trace_n) *trace_n = tn;
  if ((prm->affineOptModeA != 0 && (fabsf((float)aff_io[0]) > 1.2)) ||
      (prm->affineOptModeB != 0 && (fabsf((float)aff_io[1]) > 200)))
    return 0;
  double rel[2];
  aff_from_to(fp->ref_exposure, fp->new_exposure, fp->ref_aff_a, fp->ref_aff_b, aff_io[0], aff_io[1], rel);
  if ((prm->affineOptModeA == 0 && (fabsf(logf((float)rel[0])) > 1.5)) ||
      (prm->affineOptModeB == 0 && (fabsf((float)rel[1]) > 200)))
    return 0;
  if (prm->affineOptModeA < 0) aff_io[0] = 0;
  if (prm->affineOptModeB < 0) aff_io[1] = 0;
  *ok = 1;
  return 0;
}

extern "C" {
int orc_trk_calc_res_f(const orc_trk_level* L, int lvl, const float* dI_new, const orc_trk_frame_pair* fp,
                       const hslam_params* prm, const double refToNew[12], double aff_a, double aff_b,
                       float cutoffTH, double out6[6], float* const* warped, int* n_warped, uint8_t* mask) {
  return calc_res_impl<float>(L, lvl, dI_new, fp, prm, refToNew, aff_a, aff_b, cutoffTH, out6, warped, n_warped, mask);
}
int orc_trk_calc_res_d(const orc_trk_level* L, int lvl, const float* dI_new, const orc_trk_frame_pair* fp,
                       const hslam_params* prm, const double refToNew[12], double aff_a, double aff_b,
                       float cutoffTH, double out6[6], double* const* warped, int* n_warped, uint8_t* mask) {
  return calc_res_impl<double>(L, lvl, dI_new, fp, prm, refToNew, aff_a, aff_b, cutoffTH, out6, warped, n_warped, mask);
}
int orc_trk_calc_gs_f(const orc_trk_level* L, const orc_trk_frame_pair* fp, const hslam_params* prm, double aff_a,
                      double aff_b, const float* const* warped, int n, double H[64], double b[8]) {
  return calc_gs_impl<float>(L, fp, prm, aff_a, aff_b, warped, n, H, b);
}
int orc_trk_calc_gs_d(const orc_trk_level* L, const orc_trk_frame_pair* fp, const hslam_params* prm, double aff_a,
                      double aff_b, const double* const* warped, int n, double H[64], double b[8]) {
  return calc_gs_impl<double>(L, fp, prm, aff_a, aff_b, warped, n, H, b);
}
int orc_trk_track_f(const orc_trk_level* levels, int n_levels, const float* dI_new, const orc_trk_frame_pair* fp,
                    const hslam_params* prm, double refToNew_io[12], double aff_io[2], int coarsestLvl,
                    const double minResForAbort[5], double lastResiduals[5], double lastFlow[3], int* ok,
                    int* iterations, double* trace, int trace_cap, int* trace_n) {
  return track_impl<float>(levels, n_levels, dI_new, fp, prm, refToNew_io, aff_io, coarsestLvl, minResForAbort,
                           lastResiduals, lastFlow, ok, iterations, trace, trace_cap, trace_n);
}
int orc_trk_track_d(const orc_trk_level* levels, int n_levels, const float* dI_new, const orc_trk_frame_pair* fp,
                    const hslam_params* prm, double refToNew_io[12], double aff_io[2], int coarsestLvl,
                    const double minResForAbort[5], double lastResiduals[5], double lastFlow[3], int* ok,
                    int* iterations, double* trace, int trace_cap, int* trace_n) {
  return track_impl<double>(levels, n_levels, dI_new, fp, prm, refToNew_io, aff_io, coarsestLvl, minResForAbort,
                            lastResiduals, lastFlow, ok, iterations, trace, trace_cap, trace_n);
}
}
