// sophus/se3.hpp — stand-in (TEST INFRASTRUCTURE, see ../Eigen/Core) for the vendored Sophus, which needs the real Eigen.
// Follows Thirdparty/Sophus/sophus/se3.hpp:131-139,407-427 and so3.hpp:343-369 (exp / log / Adj / group law), fp64.
#pragma once
#include <Eigen/Core>
#include <cmath>
namespace Sophus {
class SO3d {
 public:
  Eigen::Matrix<double, 3, 3> R = Eigen::Matrix<double, 3, 3>::Identity();
};
class SE3d {
 public:
  typedef Eigen::Matrix<double, 3, 3> M3;
  typedef Eigen::Matrix<double, 3, 1> V3;
  typedef Eigen::Matrix<double, 6, 1> V6;
  M3 R = M3::Identity();
  V3 t = V3::Zero();
  SE3d() {}
  SE3d(const M3& R_, const V3& t_) : R(R_), t(t_) {}
  static M3 hat(const V3& w) { M3 m = M3::Zero(); m(0, 1) = -w[2]; m(0, 2) = w[1]; m(1, 0) = w[2]; m(1, 2) = -w[0]; m(2, 0) = -w[1]; m(2, 1) = w[0]; return m; }
  // tangent = [upsilon(3), omega(3)]
  static SE3d exp(const V6& a) {
    V3 ups, om;
    for (int i = 0; i < 3; ++i) { ups[i] = a[i]; om[i] = a[3 + i]; }
    const double th2 = om.squaredNorm(), th = std::sqrt(th2);
    const M3 O = hat(om), O2 = O * O;
    M3 Rm, V;
    if (th < 1e-10) {
      Rm = M3::Identity() + O + O2 * 0.5;
      V = M3::Identity() + O * 0.5 + O2 * (1.0 / 6.0);
    } else {
      Rm = M3::Identity() + O * (std::sin(th) / th) + O2 * ((1 - std::cos(th)) / th2);
      V = M3::Identity() + O * ((1 - std::cos(th)) / th2) + O2 * ((th - std::sin(th)) / (th2 * th));
    }
    return SE3d(Rm, V * ups);
  }
  V6 log() const {
    const double tr = R(0, 0) + R(1, 1) + R(2, 2);
    double c = (tr - 1) * 0.5; if (c > 1) c = 1; if (c < -1) c = -1;
    const double th = std::acos(c);
    V3 om;
    if (th < 1e-10) { om[0] = 0.5 * (R(2, 1) - R(1, 2)); om[1] = 0.5 * (R(0, 2) - R(2, 0)); om[2] = 0.5 * (R(1, 0) - R(0, 1)); }
    else { const double k = th / (2 * std::sin(th)); om[0] = k * (R(2, 1) - R(1, 2)); om[1] = k * (R(0, 2) - R(2, 0)); om[2] = k * (R(1, 0) - R(0, 1)); }
    const M3 O = hat(om), O2 = O * O;
    M3 Vi;
    if (th < 1e-10) Vi = M3::Identity() - O * 0.5 + O2 * (1.0 / 12.0);
    else Vi = M3::Identity() - O * 0.5 + O2 * ((1 - th * std::cos(th * 0.5) / (2 * std::sin(th * 0.5))) / (th * th));
    const V3 ups = Vi * t;
    V6 r; for (int i = 0; i < 3; ++i) { r[i] = ups[i]; r[3 + i] = om[i]; }
    return r;
  }
  SE3d inverse() const { const M3 Rt = R.transpose(); return SE3d(Rt, -(Rt * t)); }
  SE3d operator*(const SE3d& o) const { return SE3d(R * o.R, R * o.t + t); }
  V3 operator*(const V3& p) const { return R * p + t; }
  const M3& rotationMatrix() const { return R; }
  V3& translation() { return t; }
  const V3& translation() const { return t; }
  Eigen::Matrix<double, 6, 6> Adj() const {
    Eigen::Matrix<double, 6, 6> A = Eigen::Matrix<double, 6, 6>::Zero();
    const M3 tR = hat(t) * R;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { A(i, j) = R(i, j); A(3 + i, 3 + j) = R(i, j); A(i, 3 + j) = tR(i, j); }
    return A;
  }
};
typedef SE3d SE3;
}  // namespace Sophus
