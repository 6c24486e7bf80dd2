// oracle_math.hpp — tiny fixed-size linear algebra for the CPU oracle (TEST INFRASTRUCTURE ONLY).
// No Eigen in this image; the few dense operations the reference takes from
// Eigen/Sophus are restated here (SURVEY.md §8c "Where does the arithmetic live").
#pragma once
#include <cmath>
#include <cstring>
#include <algorithm>

namespace orc {

// a0 + (a1 + a2): the association Eigen 3.3's redux_novec_unroller<.,.,0,3>
// produces for a 3-term coefficient-based product (Length/2 = 1 | 2).
template <class T> static inline T sum3(T a0, T a1, T a2) { return a0 + (a1 + a2); }

// 3x3 (row-major) times 3-vector, Eigen lazy-product coefficient order.
template <class T> static inline void mat3_vec(const T* M, T x, T y, T z, T* o) {
  for (int i = 0; i < 3; ++i) o[i] = sum3<T>(M[3 * i + 0] * x, M[3 * i + 1] * y, M[3 * i + 2] * z);
}
template <class T> static inline void mat3_mul(const T* A, const T* B, T* C) {
  T tmp[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      tmp[3 * i + j] = sum3<T>(A[3 * i + 0] * B[0 + j], A[3 * i + 1] * B[3 + j], A[3 * i + 2] * B[6 + j]);
  std::memcpy(C, tmp, sizeof(tmp));
}

// Eigen's compute_inverse_size3 (cofactor formula), row-major in/out.
template <class T> static inline T cof3(const T* m, int i, int j) {
  int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m[3 * i1 + j1] * m[3 * i2 + j2] - m[3 * i1 + j2] * m[3 * i2 + j1];
}
template <class T> static inline void inv3_eigen(const T* m, T* r) {
  T c0 = cof3(m, 0, 0), c1 = cof3(m, 1, 0), c2 = cof3(m, 2, 0);
  T det = sum3<T>(c0 * m[0], c1 * m[3], c2 * m[6]);
  T invdet = T(1) / det;
  r[0] = c0 * invdet; r[1] = c1 * invdet; r[2] = c2 * invdet;
  r[3] = cof3(m, 0, 1) * invdet; r[4] = cof3(m, 1, 1) * invdet; r[5] = cof3(m, 2, 1) * invdet;
  r[6] = cof3(m, 0, 2) * invdet; r[7] = cof3(m, 1, 2) * invdet; r[8] = cof3(m, 2, 2) * invdet;
}

// ---- SE3 in double: T = [R row-major (9) | t (3)] ----
static inline void hat(const double* w, double* O) {
  O[0] = 0; O[1] = -w[2]; O[2] = w[1];
  O[3] = w[2]; O[4] = 0; O[5] = -w[0];
  O[6] = -w[1]; O[7] = w[0]; O[8] = 0;
}
static inline void matmul3d(const double* A, const double* B, double* C) {
  double t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  std::memcpy(C, t, sizeof(t));
}
// Sophus SE3Group::exp (se3.hpp:407-427) with SO3 via the unit quaternion of
// so3.hpp:343-369 converted to a rotation matrix.
static inline void se3_exp(const double* xi, double* T) {
  const double* ups = xi;
  const double* om = xi + 3;
  double theta_sq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  double theta = std::sqrt(theta_sq);
  double half = 0.5 * theta, imag, real;
  const double eps = 1e-10;  // SophusConstants<double>::epsilon()
  if (theta < eps) {
    double t4 = theta_sq * theta_sq;
    imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * t4;
    real = 1.0 - 0.5 * theta_sq + (1.0 / 384.0) * t4;
  } else {
    imag = std::sin(half) / theta;
    real = std::cos(half);
  }
  double qw = real, qx = imag * om[0], qy = imag * om[1], qz = imag * om[2];
  // Eigen::Quaternion::toRotationMatrix
  double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  double* R = T;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
  double O[9], O2[9], V[9];
  hat(om, O);
  matmul3d(O, O, O2);
  if (theta < eps) {
    std::memcpy(V, R, sizeof(V));
  } else {
    double a = (1 - std::cos(theta)) / theta_sq;
    double b = (theta - std::sin(theta)) / (theta_sq * theta);
    for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0 ? 1.0 : 0.0) + a * O[i] + b * O2[i];
  }
  for (int i = 0; i < 3; ++i) T[9 + i] = V[3 * i] * ups[0] + V[3 * i + 1] * ups[1] + V[3 * i + 2] * ups[2];
}
static inline void se3_mul(const double* A, const double* B, double* C) {
  double R[9], t[3];
  matmul3d(A, B, R);
  for (int i = 0; i < 3; ++i) t[i] = A[3 * i] * B[9] + A[3 * i + 1] * B[10] + A[3 * i + 2] * B[11] + A[9 + i];
  std::memcpy(C, R, sizeof(R));
  std::memcpy(C + 9, t, sizeof(t));
}
static inline void se3_inv(const double* A, double* Ai) {
  double R[9], t[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[3 * i + j] = A[3 * j + i];
  for (int i = 0; i < 3; ++i) t[i] = -(R[3 * i] * A[9] + R[3 * i + 1] * A[10] + R[3 * i + 2] * A[11]);
  std::memcpy(Ai, R, sizeof(R));
  std::memcpy(Ai + 9, t, sizeof(t));
}
// Sophus SE3Group::Adj (se3.hpp:131-139): [[R, hat(t) R],[0, R]], row-major 6x6.
static inline void se3_adj(const double* T, double* Ad) {
  double Ht[9], HR[9];
  hat(T + 9, Ht);
  matmul3d(Ht, T, HR);
  for (int i = 0; i < 36; ++i) Ad[i] = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Ad[6 * i + j] = T[3 * i + j];
      Ad[6 * (i + 3) + (j + 3)] = T[3 * i + j];
      Ad[6 * i + (j + 3)] = HR[3 * i + j];
    }
}
// SO3 log via quaternion (so3.hpp logAndTheta) and SE3 log (se3.hpp log).
static inline void se3_log(const double* T, double* xi) {
  const double* R = T;
  // rotation matrix -> quaternion (Eigen)
  double q[4];  // w,x,y,z
  double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    double s = std::sqrt(tr + 1.0);
    q[0] = 0.5 * s; s = 0.5 / s;
    q[1] = (R[7] - R[5]) * s; q[2] = (R[2] - R[6]) * s; q[3] = (R[3] - R[1]) * s;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[1 + i] = 0.5 * s; s = 0.5 / s;
    q[0] = (R[3 * k + j] - R[3 * j + k]) * s;
    q[1 + j] = (R[3 * j + i] + R[3 * i + j]) * s;
    q[1 + k] = (R[3 * k + i] + R[3 * i + k]) * s;
  }
  double n2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  double n = std::sqrt(n2), w = q[0], two_atan;
  double theta;
  const double eps = 1e-10;
  if (n < eps) {
    two_atan = 2.0 / w - 2.0 * n2 / (w * w * w);
  } else if (std::fabs(w) < eps) {
    two_atan = (w > 0 ? M_PI : -M_PI) / n;
  } else {
    two_atan = 2.0 * std::atan(n / w) / n;
  }
  theta = two_atan * n;
  double om[3] = {two_atan * q[1], two_atan * q[2], two_atan * q[3]};
  double O[9], O2[9], Vi[9];
  hat(om, O);
  matmul3d(O, O, O2);
  if (theta < eps) {
    for (int i = 0; i < 9; ++i) Vi[i] = (i % 4 == 0 ? 1.0 : 0.0) - 0.5 * O[i] + (1.0 / 12.0) * O2[i];
  } else {
    double c = (1 - theta / (2 * std::tan(theta / 2))) / (theta * theta);
    for (int i = 0; i < 9; ++i) Vi[i] = (i % 4 == 0 ? 1.0 : 0.0) - 0.5 * O[i] + c * O2[i];
  }
  for (int i = 0; i < 3; ++i) xi[i] = Vi[3 * i] * T[9] + Vi[3 * i + 1] * T[10] + Vi[3 * i + 2] * T[11];
  xi[3] = om[0]; xi[4] = om[1]; xi[5] = om[2];
}

// LDLT with symmetric diagonal pivoting (the scheme of Eigen::LDLT), A col-major n x n.
static inline void ldlt_solve(const double* Ain, const double* rhs, int n, double* x) {
  double A[16 * 16];
  int perm[16];
  for (int i = 0; i < n * n; ++i) A[i] = Ain[i];
  for (int i = 0; i < n; ++i) perm[i] = i;
  auto at = [&](int r, int c) -> double& { return A[r + c * n]; };
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = std::fabs(at(k, k));
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(at(i, i)) > best) { best = std::fabs(at(i, i)); p = i; }
    if (p != k) {
      for (int c = 0; c < n; ++c) std::swap(at(k, c), at(p, c));
      for (int r = 0; r < n; ++r) std::swap(at(r, k), at(r, p));
      std::swap(perm[k], perm[p]);
    }
    double d = at(k, k);
    if (d == 0.0) continue;
    for (int i = k + 1; i < n; ++i) at(i, k) /= d;
    for (int j = k + 1; j < n; ++j)
      for (int i = j; i < n; ++i) {
        at(i, j) -= at(i, k) * d * at(j, k);
        at(j, i) = at(i, j);
      }
  }
  double y[16];
  for (int i = 0; i < n; ++i) y[i] = rhs[perm[i]];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) y[i] -= at(i, j) * y[j];
  for (int i = 0; i < n; ++i) y[i] = (at(i, i) != 0.0) ? y[i] / at(i, i) : 0.0;
  for (int i = n - 1; i >= 0; --i)
    for (int j = i + 1; j < n; ++j) y[i] -= at(j, i) * y[j];
  for (int i = 0; i < n; ++i) x[perm[i]] = y[i];
}

// AffLight::fromToVecExposure (Include/GlobalTypes.h:334-346)
static inline void aff_from_to(float eF, float eT, double aF, double bF, double aT, double bT, double* out) {
  if (eF == 0 || eT == 0) { eT = eF = 1; }
  double a = std::exp(aT - aF) * eT / eF;
  double b = bT - a * bF;
  out[0] = a; out[1] = b;
}

}  // namespace orc
