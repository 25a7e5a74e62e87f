// SO(3) / small-matrix leaf math of the CUDA engine (fp64).
//
// Behavioural contract = the reference's Sophus/Eigen semantics, including the
// small-angle branches (cited per function).  Written as host+device inline
// functions: the kernels use them on the GPU, and tests/emu compiles the same
// header with g++ to check the per-thread math against the oracle without a GPU.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define CTVIO_HD __host__ __device__ __forceinline__
#else
#define CTVIO_HD inline
#endif

namespace ctvio {

constexpr double kSo3Eps = 1e-10;  // Sophus::Constants<double>::epsilon (sophus_lib/common.hpp:144)

struct V3 {
  double x, y, z;
};
struct Q4 {  // [x,y,z,w], Eigen coeff order (sophus_lib/so3.hpp:196)
  double x, y, z, w;
};
struct M3 {  // row-major
  double m[9];
};

CTVIO_HD V3 v3(double x, double y, double z) { return V3{x, y, z}; }
CTVIO_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
CTVIO_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
CTVIO_HD V3 operator*(double s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
CTVIO_HD V3 neg(V3 a) { return V3{-a.x, -a.y, -a.z}; }
CTVIO_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
CTVIO_HD V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

CTVIO_HD M3 m3_identity() {
  M3 r;
  r.m[0] = 1; r.m[1] = 0; r.m[2] = 0; r.m[3] = 0; r.m[4] = 1; r.m[5] = 0; r.m[6] = 0; r.m[7] = 0; r.m[8] = 1;
  return r;
}
CTVIO_HD M3 m3_mul(const M3& a, const M3& b) {
  M3 c;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
  return c;
}
// a * b^T
CTVIO_HD M3 m3_mul_bt(const M3& a, const M3& b) {
  M3 c;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c.m[3 * i + j] = a.m[3 * i] * b.m[3 * j] + a.m[3 * i + 1] * b.m[3 * j + 1] + a.m[3 * i + 2] * b.m[3 * j + 2];
  return c;
}
CTVIO_HD M3 m3_transpose(const M3& a) {
  M3 c;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) c.m[3 * i + j] = a.m[3 * j + i];
  return c;
}
CTVIO_HD M3 m3_scale(double s, const M3& a) {
  M3 c;
#pragma unroll
  for (int i = 0; i < 9; ++i) c.m[i] = s * a.m[i];
  return c;
}
CTVIO_HD M3 m3_sub(const M3& a, const M3& b) {
  M3 c;
#pragma unroll
  for (int i = 0; i < 9; ++i) c.m[i] = a.m[i] - b.m[i];
  return c;
}
CTVIO_HD V3 m3_vec(const M3& a, V3 v) {
  return V3{a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z,
            a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z};
}
// a^T v
CTVIO_HD V3 m3_tvec(const M3& a, V3 v) {
  return V3{a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z,
            a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z};
}
// a * hat(v)   (columns: a * (e_k x ...)): (a hat(v))_{ij} = sum_k a_ik hat(v)_kj
CTVIO_HD M3 m3_mul_hat(const M3& a, V3 v) {
  M3 c;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double a0 = a.m[3 * i], a1 = a.m[3 * i + 1], a2 = a.m[3 * i + 2];
    c.m[3 * i] = a1 * v.z - a2 * v.y;
    c.m[3 * i + 1] = a2 * v.x - a0 * v.z;
    c.m[3 * i + 2] = a0 * v.y - a1 * v.x;
  }
  return c;
}

CTVIO_HD Q4 q_conj(Q4 a) { return Q4{-a.x, -a.y, -a.z, a.w}; }  // SO3::inverse, so3.hpp:202

// SO3 group product = Eigen quaternion product + Sophus first-order renormalisation (so3.hpp:338-354)
CTVIO_HD Q4 so3_mul(Q4 a, Q4 b) {
  Q4 q;
  q.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  q.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  q.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  q.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  const double sn = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  if (sn != 1.0) {
    const double s = 2.0 / (1.0 + sn);
    q.x *= s; q.y *= s; q.z *= s; q.w *= s;
  }
  return q;
}

// SO3 * point (Eigen::Quaternion::_transformVector, so3.hpp:318)
CTVIO_HD V3 so3_rotate(Q4 q, V3 v) {
  const V3 qv = V3{q.x, q.y, q.z};
  V3 uv = cross(qv, v);
  uv = uv + uv;
  return v + q.w * uv + cross(qv, uv);
}

// SO3::matrix (Eigen::Quaternion::toRotationMatrix, so3.hpp:283)
CTVIO_HD M3 so3_matrix(Q4 q) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  M3 r;
  r.m[0] = 1 - (tyy + tzz); r.m[1] = txy - twz;       r.m[2] = txz + twy;
  r.m[3] = txy + twz;       r.m[4] = 1 - (txx + tzz); r.m[5] = tyz - twx;
  r.m[6] = txz - twy;       r.m[7] = tyz + twx;       r.m[8] = 1 - (txx + tyy);
  return r;
}

// SO3::exp (so3.hpp:534-568).  theta is passed in by callers that already have it.
CTVIO_HD Q4 so3_exp_theta(V3 omega, double theta_sq, double theta) {
  double imag, real;
  if (theta < kSo3Eps) {
    const double theta_po4 = theta_sq * theta_sq;
    imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    double sh, ch;
#if defined(__CUDA_ARCH__)
    sincos(0.5 * theta, &sh, &ch);
#else
    sh = std::sin(0.5 * theta);
    ch = std::cos(0.5 * theta);
#endif
    imag = sh / theta;
    real = ch;
  }
  return Q4{imag * omega.x, imag * omega.y, imag * omega.z, real};
}
CTVIO_HD Q4 so3_exp(V3 omega) {
  const double tsq = dot(omega, omega);
  return so3_exp_theta(omega, tsq, sqrt(tsq));
}

// SO3::log (so3.hpp:220-261)
CTVIO_HD V3 so3_log(Q4 q) {
  const double squared_n = q.x * q.x + q.y * q.y + q.z * q.z;
  const double n = sqrt(squared_n);
  const double w = q.w;
  double f;
  if (n < kSo3Eps) {
    f = 2.0 / w - 2.0 * squared_n / (w * (w * w));
  } else if (fabs(w) < kSo3Eps) {
    f = (w > 0.0 ? 3.14159265358979323846 : -3.14159265358979323846) / n;
  } else {
    f = 2.0 * atan(n / w) / n;
  }
  return V3{f * q.x, f * q.y, f * q.z};
}

// I + a*hat(phi) + b*hat(phi)^2, hat(phi)^2 = phi phi^T - |phi|^2 I
CTVIO_HD M3 rodrigues_like(V3 p, double n2, double a, double b) {
  M3 J;
  J.m[0] = 1 + b * (p.x * p.x - n2); J.m[1] = -a * p.z + b * p.x * p.y;   J.m[2] = a * p.y + b * p.x * p.z;
  J.m[3] = a * p.z + b * p.x * p.y;  J.m[4] = 1 + b * (p.y * p.y - n2);   J.m[5] = -a * p.x + b * p.y * p.z;
  J.m[6] = -a * p.y + b * p.x * p.z; J.m[7] = a * p.x + b * p.y * p.z;    J.m[8] = 1 + b * (p.z * p.z - n2);
  return J;
}

// rightJacobianSO3 (utils/sophus_utils.hpp:165-199): I - (1-cos)/n2 hat + (n - sin)/n3 hat^2
CTVIO_HD M3 right_jacobian(V3 phi) {
  const double n2 = dot(phi, phi);
  if (n2 > kSo3Eps) {
    const double n = sqrt(n2);
    double sn, cn;
#if defined(__CUDA_ARCH__)
    sincos(n, &sn, &cn);
#else
    sn = std::sin(n);
    cn = std::cos(n);
#endif
    return rodrigues_like(phi, n2, -(1 - cn) / n2, (n - sn) / (n2 * n));
  }
  return rodrigues_like(phi, n2, -0.5, 1.0 / 6.0);
}

// rightJacobianInvSO3 (utils/sophus_utils.hpp:209-242)
CTVIO_HD M3 right_jacobian_inv(V3 phi) {
  const double n2 = dot(phi, phi);
  if (n2 > kSo3Eps) {
    const double n = sqrt(n2);
    double sn, cn;
#if defined(__CUDA_ARCH__)
    sincos(n, &sn, &cn);
#else
    sn = std::sin(n);
    cn = std::cos(n);
#endif
    return rodrigues_like(phi, n2, 0.5, 1 / n2 - (1 + cn) / (2 * n * sn));
  }
  return rodrigues_like(phi, n2, 0.5, 1.0 / 12.0);
}

// Eigen::Quaternion(Matrix3) (used by SE3d(rot_diff, tran_diff), trajectory_manager.cpp:508)
CTVIO_HD Q4 quat_from_matrix(const M3& m) {
  double t = m.m[0] + m.m[4] + m.m[8];
  double qv[4];
  if (t > 0) {
    t = sqrt(t + 1.0);
    qv[3] = 0.5 * t;
    t = 0.5 / t;
    qv[0] = (m.m[7] - m.m[5]) * t;
    qv[1] = (m.m[2] - m.m[6]) * t;
    qv[2] = (m.m[3] - m.m[1]) * t;
  } else {
    int i = 0;
    if (m.m[4] > m.m[0]) i = 1;
    if (m.m[8] > m.m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m.m[4 * i] - m.m[4 * j] - m.m[4 * k] + 1.0);
    qv[i] = 0.5 * t;
    t = 0.5 / t;
    qv[3] = (m.m[3 * k + j] - m.m[3 * j + k]) * t;
    qv[j] = (m.m[3 * j + i] + m.m[3 * i + j]) * t;
    qv[k] = (m.m[3 * k + i] + m.m[3 * i + k]) * t;
  }
  return Q4{qv[0], qv[1], qv[2], qv[3]};
}

}  // namespace ctvio
