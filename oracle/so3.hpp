// ORACLE (test infrastructure, NOT product code).  CPU fp64 restatement of the
// Lie-group leaf math the Ctrl-VIO hot path uses.  Parity unpinned: the
// reference ships no tests/golden vectors and cannot be built here (needs
// Eigen + Ceres); this file is audited line-by-line against the cited sources
// and self-validated by tests/test_oracle_*.py (finite differences, group
// identities, view-vs-plain-spline equality).
//
// Restates:
//   sophus_lib/so3.hpp:220-261   logAndTheta  (atan-based log, eps branches)
//   sophus_lib/so3.hpp:338-354   operator*=   (first-order renormalisation)
//   sophus_lib/so3.hpp:534-568   expAndTheta  (Taylor branch theta < 1e-10)
//   sophus_lib/so3.hpp:283,618   matrix(), hat()
//   utils/sophus_utils.hpp:165-199  rightJacobianSO3
//   utils/sophus_utils.hpp:209-242  rightJacobianInvSO3
//   utils/sophus_utils.hpp:251-329  left Jacobians (cross-check evaluators only)
//   Eigen::Quaternion product / _transformVector / toRotationMatrix semantics
// Quaternion storage order is [x, y, z, w] (Eigen coeffs order, so3.hpp:196).
#pragma once
#include <cmath>
#include <cstdint>

namespace ctvio_oracle {

constexpr double kEps = 1e-10;  // Sophus::Constants<double>::epsilon(), common.hpp:144

struct Vec3 {
  double x, y, z;
  Vec3() : x(0), y(0), z(0) {}
  Vec3(double a, double b, double c) : x(a), y(b), z(c) {}
  double& operator[](int i) { return (&x)[i]; }
  double operator[](int i) const { return (&x)[i]; }
};
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator-(const Vec3& a) { return {-a.x, -a.y, -a.z}; }
inline Vec3 operator*(double s, const Vec3& a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 operator*(const Vec3& a, double s) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(const Vec3& a, const Vec3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Row-major 3x3.
struct Mat3 {
  double m[9];
  double& operator()(int r, int c) { return m[3 * r + c]; }
  double operator()(int r, int c) const { return m[3 * r + c]; }
  static Mat3 Zero() {
    Mat3 a;
    for (double& v : a.m) v = 0;
    return a;
  }
  static Mat3 Identity() {
    Mat3 a = Zero();
    a.m[0] = a.m[4] = a.m[8] = 1;
    return a;
  }
};
inline Mat3 operator*(const Mat3& a, const Mat3& b) {
  Mat3 c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += a(i, k) * b(k, j);
      c(i, j) = s;
    }
  return c;
}
inline Vec3 operator*(const Mat3& a, const Vec3& v) {
  return {a(0, 0) * v.x + a(0, 1) * v.y + a(0, 2) * v.z,
          a(1, 0) * v.x + a(1, 1) * v.y + a(1, 2) * v.z,
          a(2, 0) * v.x + a(2, 1) * v.y + a(2, 2) * v.z};
}
inline Mat3 operator*(double s, const Mat3& a) {
  Mat3 c;
  for (int i = 0; i < 9; ++i) c.m[i] = s * a.m[i];
  return c;
}
inline Mat3 operator+(const Mat3& a, const Mat3& b) {
  Mat3 c;
  for (int i = 0; i < 9; ++i) c.m[i] = a.m[i] + b.m[i];
  return c;
}
inline Mat3 operator-(const Mat3& a, const Mat3& b) {
  Mat3 c;
  for (int i = 0; i < 9; ++i) c.m[i] = a.m[i] - b.m[i];
  return c;
}
inline Mat3 transpose(const Mat3& a) {
  Mat3 c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c(i, j) = a(j, i);
  return c;
}
// SO3::hat, so3.hpp:618
inline Mat3 hat(const Vec3& w) {
  Mat3 a = Mat3::Zero();
  a(0, 1) = -w.z; a(0, 2) = w.y;
  a(1, 0) = w.z;  a(1, 2) = -w.x;
  a(2, 0) = -w.y; a(2, 1) = w.x;
  return a;
}

// Unit quaternion, coefficient order x,y,z,w.
struct Quat {
  double x, y, z, w;
  Quat() : x(0), y(0), z(0), w(1) {}
  Quat(double x_, double y_, double z_, double w_) : x(x_), y(y_), z(z_), w(w_) {}
  static Quat fromPtr(const double* p) { return {p[0], p[1], p[2], p[3]}; }
  void toPtr(double* p) const { p[0] = x; p[1] = y; p[2] = z; p[3] = w; }
  Quat conj() const { return {-x, -y, -z, w}; }  // SO3::inverse(), so3.hpp:202
  double sqnorm() const { return x * x + y * y + z * z + w * w; }
};

// Eigen quaternion product a*b (no normalisation).
inline Quat qmul_raw(const Quat& a, const Quat& b) {
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
          a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}

// SO3 group product with Sophus' first-order renormalisation, so3.hpp:338-354.
inline Quat so3_mul(const Quat& a, const Quat& b) {
  Quat q = qmul_raw(a, b);
  const double sn = q.sqnorm();
  if (sn != 1.0) {
    const double s = 2.0 / (1.0 + sn);
    q.x *= s; q.y *= s; q.z *= s; q.w *= s;
  }
  return q;
}

// SO3 * point == Eigen::Quaternion::_transformVector, so3.hpp:318-320.
inline Vec3 so3_rotate(const Quat& q, const Vec3& v) {
  const Vec3 qv(q.x, q.y, q.z);
  Vec3 uv = cross(qv, v);
  uv = uv + uv;
  return v + q.w * uv + cross(qv, uv);
}

// SO3::matrix() == Eigen::Quaternion::toRotationMatrix, so3.hpp:283.
inline Mat3 so3_matrix(const Quat& q) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  Mat3 r;
  r(0, 0) = 1 - (tyy + tzz); r(0, 1) = txy - twz;       r(0, 2) = txz + twy;
  r(1, 0) = txy + twz;       r(1, 1) = 1 - (txx + tzz); r(1, 2) = tyz - twx;
  r(2, 0) = txz - twy;       r(2, 1) = tyz + twx;       r(2, 2) = 1 - (txx + tyy);
  return r;
}

// SO3::exp, so3.hpp:534-568.
inline Quat so3_exp(const Vec3& omega) {
  const double theta_sq = dot(omega, omega);
  const double theta = std::sqrt(theta_sq);
  const double half_theta = 0.5 * theta;
  double imag, real;
  if (theta < kEps) {
    const double theta_po4 = theta_sq * theta_sq;
    imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    imag = std::sin(half_theta) / theta;
    real = std::cos(half_theta);
  }
  return {imag * omega.x, imag * omega.y, imag * omega.z, real};
}

// SO3::log, so3.hpp:220-261.
inline Vec3 so3_log(const Quat& q) {
  const double squared_n = q.x * q.x + q.y * q.y + q.z * q.z;
  const double n = std::sqrt(squared_n);
  const double w = q.w;
  double two_atan_nbyw_by_n;
  if (n < kEps) {
    const double squared_w = w * w;
    two_atan_nbyw_by_n = 2.0 / w - 2.0 * squared_n / (w * squared_w);
  } else {
    if (std::fabs(w) < kEps) {
      two_atan_nbyw_by_n = (w > 0.0) ? M_PI / n : -M_PI / n;
    } else {
      two_atan_nbyw_by_n = 2.0 * std::atan(n / w) / n;
    }
  }
  return {two_atan_nbyw_by_n * q.x, two_atan_nbyw_by_n * q.y, two_atan_nbyw_by_n * q.z};
}

// sophus_utils.hpp:165-199
inline Mat3 rightJacobianSO3(const Vec3& phi) {
  const double n2 = dot(phi, phi);
  const Mat3 ph = hat(phi);
  const Mat3 ph2 = ph * ph;
  Mat3 J = Mat3::Identity();
  if (n2 > kEps) {
    const double n = std::sqrt(n2);
    const double n3 = n2 * n;
    J = J - ((1 - std::cos(n)) / n2) * ph;
    J = J + ((n - std::sin(n)) / n3) * ph2;
  } else {
    J = J - 0.5 * ph;
    J = J + (1.0 / 6.0) * ph2;
  }
  return J;
}

// sophus_utils.hpp:209-242
inline Mat3 rightJacobianInvSO3(const Vec3& phi) {
  const double n2 = dot(phi, phi);
  const Mat3 ph = hat(phi);
  const Mat3 ph2 = ph * ph;
  Mat3 J = Mat3::Identity();
  J = J + 0.5 * ph;
  if (n2 > kEps) {
    const double n = std::sqrt(n2);
    J = J + (1 / n2 - (1 + std::cos(n)) / (2 * n * std::sin(n))) * ph2;
  } else {
    J = J + (1.0 / 12.0) * ph2;
  }
  return J;
}

// sophus_utils.hpp:251-286 (only used by the plain-spline cross-check evaluator)
inline Mat3 leftJacobianSO3(const Vec3& phi) {
  const double n2 = dot(phi, phi);
  const Mat3 ph = hat(phi);
  const Mat3 ph2 = ph * ph;
  Mat3 J = Mat3::Identity();
  if (n2 > kEps) {
    const double n = std::sqrt(n2);
    const double n3 = n2 * n;
    J = J + ((1 - std::cos(n)) / n2) * ph;
    J = J + ((n - std::sin(n)) / n3) * ph2;
  } else {
    J = J + 0.5 * ph;
    J = J + (1.0 / 6.0) * ph2;
  }
  return J;
}

// sophus_utils.hpp:295-329
inline Mat3 leftJacobianInvSO3(const Vec3& phi) {
  const double n2 = dot(phi, phi);
  const Mat3 ph = hat(phi);
  const Mat3 ph2 = ph * ph;
  Mat3 J = Mat3::Identity();
  J = J - 0.5 * ph;
  if (n2 > kEps) {
    const double n = std::sqrt(n2);
    J = J + (1 / n2 - (1 + std::cos(n)) / (2 * n * std::sin(n))) * ph2;
  } else {
    J = J + (1.0 / 12.0) * ph2;
  }
  return J;
}

}  // namespace ctvio_oracle
