// ORACLE (test infrastructure, NOT product code).  Parity unpinned (see so3.hpp).
// CPU fp64 restatement of the order-4 uniform cumulative B-spline evaluators
// of the Ctrl-VIO hot path.  Knots are addressed by GLOBAL index into flat
// arrays q[N][4] (x,y,z,w) / p[N][3]; the reference's per-factor segment
// bookkeeping (SplineMeta) resolves to the same global start index
// s = (t - t0) / dt (spline_segment.h:72-88, se3_spline.h:463-503).
//
// Restates:
//   spline/spline_common.h:76-153          blending matrices / base coefficients
//   spline/spline_segment.h:72-88,103-114  computeTIndexNs (int64 div/mod)
//   factor/analytic_diff/so3_spline_view.h:69-126   EvaluateRotation
//                                          :136-198  EvaluateRp
//                                          :208-276  EvaluateRTp
//                                          :356-426  VelocityBody
//   factor/analytic_diff/rd_spline_view.h:63-94     evaluate<D>
//   factor/analytic_diff/split_spline_view.h:67-214 SplitSpineView::Evaluate
//   spline/so3_spline.h:240-367, spline/rd_spline.h:229-259  (plain evaluators,
//        independent second implementation used as cross-check in tests)
#pragma once
#include <cstdint>
#include <cstdio>

#include "so3.hpp"

namespace ctvio_oracle {

constexpr int kN = 4;    // SplineOrder, spline_common.h:47
constexpr int kDEG = 3;

// spline_common.h:76-115 evaluated for N=4 (entries are exact integers / 6).
// Row k = coefficient of knot k against [1,u,u^2,u^3].
struct Blend {
  double M[4][4];   // plain (RdSplineView::blending_matrix_)
  double Mc[4][4];  // cumulative (So3SplineView::blending_matrix_)
  Blend() {
    const double m[4][4] = {{1, -3, 3, -1}, {4, 0, -6, 3}, {1, 3, 3, -3}, {0, 0, 0, 1}};
    const double c[4][4] = {{6, 0, 0, 0}, {5, 3, -3, 1}, {1, 3, 3, -2}, {0, 0, 0, 1}};
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        M[i][j] = m[i][j] / 6.0;
        Mc[i][j] = c[i][j] / 6.0;
      }
  }
};
inline const Blend& blend() {
  static const Blend b;
  return b;
}

// base_coefficients_ (spline_common.h:134-153): rows [1,1,1,1],[0,1,2,3],[0,0,2,6].
// baseCoeffsWithTime<D> (so3_spline_view.h:438-459 / rd_spline_view.h:124-145).
template <int D>
inline void baseCoeffsWithTime(double res[4], double t) {
  static const double base[3][4] = {{1, 1, 1, 1}, {0, 1, 2, 3}, {0, 0, 2, 6}};
  for (int i = 0; i < 4; ++i) res[i] = 0;
  res[D] = base[D][D];
  double _t = t;
  for (int j = D + 1; j < 4; ++j) {
    res[j] = base[D][j] * _t;
    _t = _t * t;
  }
}

inline void matvec4(const double A[4][4], const double p[4], double out[4]) {
  for (int i = 0; i < 4; ++i) out[i] = A[i][0] * p[0] + A[i][1] * p[1] + A[i][2] * p[2] + A[i][3] * p[3];
}

struct SplineGrid {
  int64_t t0_ns;  // time of the first valid instant of knot 0's interval
  int64_t dt_ns;  // knot spacing
  int n_knots;
  double pow_inv_dt[4];  // spline_segment.h:54-64
  SplineGrid() : t0_ns(0), dt_ns(1), n_knots(0) { set(0, 1, 0); }
  void set(int64_t t0, int64_t dt, int n) {
    t0_ns = t0; dt_ns = dt; n_knots = n;
    pow_inv_dt[0] = 1.0;
    pow_inv_dt[1] = 1e9 / double(dt_ns);
    for (int i = 2; i < 4; ++i) pow_inv_dt[i] = pow_inv_dt[i - 1] * pow_inv_dt[1];
  }
  int64_t maxTimeNs() const { return t0_ns + int64_t(n_knots - kDEG) * dt_ns; }
  // spline_segment.h:72-88.  Returns false when t is outside [min,max) (the
  // reference asserts there).
  bool computeTIndexNs(int64_t time_ns, int64_t& s, double& u) const {
    if (time_ns < t0_ns || time_ns >= maxTimeNs()) return false;
    const int64_t st_ns = time_ns - t0_ns;
    s = st_ns / dt_ns;
    u = double(st_ns % dt_ns) / double(dt_ns);
    return true;
  }
};

struct So3Jacobian {
  int64_t start_idx;
  Mat3 d_val_d_knot[4];
};
struct RdJacobian {
  int64_t start_idx;
  double d_val_d_knot[4];
};

inline Quat knotQ(const double* q, int64_t i) { return Quat::fromPtr(q + 4 * i); }
inline Vec3 knotP(const double* p, int64_t i) { return Vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }

// so3_spline_view.h:136-198.  Returns R(t); J (if given) with the left factor
// -Lhs*R(t)*hat(rhs) omitted.
inline Quat EvaluateRp(const SplineGrid& g, const double* q, int64_t time_ns, So3Jacobian* J) {
  int64_t s; double u;
  g.computeTIndexNs(time_ns, s, u);
  double pw[4], coeff[4];
  baseCoeffsWithTime<0>(pw, u);
  matvec4(blend().Mc, pw, coeff);
  if (J) J->start_idx = s;

  Quat A_accum_inv;  // identity
  Mat3 A_post_inv[kDEG + 1];
  Mat3 Jr_inv_delta[kDEG], Jr_kdelta[kDEG];
  A_post_inv[kDEG] = so3_matrix(A_accum_inv);
  for (int i = kDEG - 1; i >= 0; --i) {
    const Quat R0 = knotQ(q, s + i), R1 = knotQ(q, s + i + 1);
    const Vec3 delta = so3_log(so3_mul(R0.conj(), R1));
    const Vec3 kdelta = delta * coeff[i + 1];
    A_accum_inv = so3_mul(A_accum_inv, so3_exp(-kdelta));
    if (J) {
      Jr_inv_delta[i] = rightJacobianInvSO3(delta);
      Jr_kdelta[i] = rightJacobianSO3(kdelta);
      A_post_inv[i] = so3_matrix(A_accum_inv);
    }
  }
  const Quat Ri = knotQ(q, s);
  const Quat res = so3_mul(Ri, A_accum_inv.conj());
  if (J) {
    Mat3 J_helper = A_post_inv[0];
    J->d_val_d_knot[0] = J_helper;
    for (int i = 0; i < kDEG; ++i) {
      J_helper = coeff[i + 1] * (A_post_inv[i + 1] * Jr_kdelta[i]);
      J->d_val_d_knot[i] = J->d_val_d_knot[i] - J_helper * transpose(Jr_inv_delta[i]);
      J->d_val_d_knot[i + 1] = J_helper * Jr_inv_delta[i];
    }
  }
  return res;
}

// so3_spline_view.h:208-276.  Returns R(t)^T; J with the left factor
// Lhs*R(t)^T*hat(rhs) omitted.
inline Quat EvaluateRTp(const SplineGrid& g, const double* q, int64_t time_ns, So3Jacobian* J) {
  int64_t s; double u;
  g.computeTIndexNs(time_ns, s, u);
  double pw[4], coeff[4];
  baseCoeffsWithTime<0>(pw, u);
  matvec4(blend().Mc, pw, coeff);
  if (J) J->start_idx = s;

  Quat Si_A_pre[kDEG + 1];
  Mat3 Ri_A_pre[kDEG + 1];
  Mat3 Jr_inv_delta[kDEG], Jr_kdelta[kDEG];
  Si_A_pre[0] = knotQ(q, s);
  for (int i = 0; i < kDEG; ++i) {
    const Quat R0 = knotQ(q, s + i), R1 = knotQ(q, s + i + 1);
    const Vec3 delta = so3_log(so3_mul(R0.conj(), R1));
    const Vec3 kdelta = delta * coeff[i + 1];
    Si_A_pre[i + 1] = so3_mul(Si_A_pre[i], so3_exp(kdelta));
    if (J) {
      Jr_inv_delta[i] = rightJacobianInvSO3(delta);
      Jr_kdelta[i] = rightJacobianSO3(-kdelta);
    }
  }
  for (int i = 0; i < kDEG + 1; ++i) Ri_A_pre[i] = so3_matrix(Si_A_pre[i]);
  const Quat res = Si_A_pre[kDEG].conj();
  if (J) {
    Mat3 J_helper = Ri_A_pre[0];
    J->d_val_d_knot[0] = J_helper;
    for (int i = 0; i < kDEG; ++i) {
      J_helper = coeff[i + 1] * (Ri_A_pre[i] * Jr_kdelta[i]);
      J->d_val_d_knot[i] = J->d_val_d_knot[i] - J_helper * transpose(Jr_inv_delta[i]);
      J->d_val_d_knot[i + 1] = J_helper * Jr_inv_delta[i];
    }
  }
  return res;
}

// so3_spline_view.h:69-126 (value + right-tangent Jacobian; unused by the
// factors, kept for the cross-check tests).
inline Quat EvaluateRotation(const SplineGrid& g, const double* q, int64_t time_ns, So3Jacobian* J) {
  int64_t s; double u;
  g.computeTIndexNs(time_ns, s, u);
  double pw[4], coeff[4];
  baseCoeffsWithTime<0>(pw, u);
  matvec4(blend().Mc, pw, coeff);
  Quat res = knotQ(q, s);
  Mat3 J_helper;
  if (J) {
    J->start_idx = s;
    J_helper = so3_matrix(res);
  }
  Mat3 R_tmp[kDEG], Jr_inv_delta[kDEG], Jr_kdelta[kDEG];
  for (int i = 0; i < kDEG; ++i) {
    const Quat p0 = knotQ(q, s + i), p1 = knotQ(q, s + i + 1);
    const Vec3 delta = so3_log(so3_mul(p0.conj(), p1));
    const Vec3 kdelta = delta * coeff[i + 1];
    res = so3_mul(res, so3_exp(kdelta));
    Jr_inv_delta[i] = rightJacobianInvSO3(delta);
    Jr_kdelta[i] = rightJacobianSO3(kdelta);
    R_tmp[i] = so3_matrix(res);
  }
  if (J) {
    J->d_val_d_knot[0] = J_helper;
    for (int i = 0; i < kDEG; ++i) {
      J_helper = coeff[i + 1] * (R_tmp[i] * Jr_kdelta[i]);
      J->d_val_d_knot[i] = J->d_val_d_knot[i] - J_helper * transpose(Jr_inv_delta[i]);
      J->d_val_d_knot[i + 1] = J_helper * Jr_inv_delta[i];
    }
  }
  return res;
}

// so3_spline_view.h:356-426.  Body angular velocity w(t) (+ Jacobian w.r.t. the
// right perturbation of the 4 knots when J != nullptr).
inline Vec3 VelocityBody(const SplineGrid& g, const double* q, int64_t time_ns, So3Jacobian* J) {
  int64_t s; double u;
  g.computeTIndexNs(time_ns, s, u);
  double pw[4], coeff[4], dcoeff[4];
  baseCoeffsWithTime<0>(pw, u);
  matvec4(blend().Mc, pw, coeff);
  baseCoeffsWithTime<1>(pw, u);
  matvec4(blend().Mc, pw, dcoeff);
  for (int i = 0; i < 4; ++i) dcoeff[i] = g.pow_inv_dt[1] * dcoeff[i];

  Vec3 delta_vec[kDEG];
  Mat3 R_tmp[kDEG];
  Quat accum;
  Quat exp_k_delta[kDEG];
  Mat3 Jr_delta_inv[kDEG], Jr_kdelta[kDEG];
  for (int i = kDEG - 1; i >= 0; --i) {
    const Quat p0 = knotQ(q, s + i), p1 = knotQ(q, s + i + 1);
    delta_vec[i] = so3_log(so3_mul(p0.conj(), p1));
    Jr_delta_inv[i] = rightJacobianInvSO3(delta_vec[i]) * so3_matrix(p1.conj());
    const Vec3 k_delta = coeff[i + 1] * delta_vec[i];
    Jr_kdelta[i] = rightJacobianSO3(-k_delta);
    R_tmp[i] = so3_matrix(accum);
    exp_k_delta[i] = so3_exp(-k_delta);
    accum = so3_mul(accum, exp_k_delta[i]);
  }
  Mat3 d_vel_d_delta[kDEG];
  d_vel_d_delta[0] = dcoeff[1] * (R_tmp[0] * Jr_delta_inv[0]);
  Vec3 rot_vel = delta_vec[0] * dcoeff[1];
  for (int i = 1; i < kDEG; ++i) {
    d_vel_d_delta[i] = coeff[i + 1] * (R_tmp[i - 1] * hat(rot_vel) * Jr_kdelta[i]) + dcoeff[i + 1] * R_tmp[i];
    d_vel_d_delta[i] = d_vel_d_delta[i] * Jr_delta_inv[i];
    rot_vel = so3_rotate(exp_k_delta[i], rot_vel) + delta_vec[i] * dcoeff[i + 1];
  }
  if (J) {
    J->start_idx = s;
    for (int i = 0; i < kN; ++i) J->d_val_d_knot[i] = Mat3::Zero();
    for (int i = 0; i < kDEG; ++i) {
      J->d_val_d_knot[i] = J->d_val_d_knot[i] - d_vel_d_delta[i];
      J->d_val_d_knot[i + 1] = J->d_val_d_knot[i + 1] + d_vel_d_delta[i];
    }
  }
  return rot_vel;
}

// rd_spline_view.h:63-94.
template <int D>
inline Vec3 RdEvaluate(const SplineGrid& g, const double* p, int64_t time_ns, RdJacobian* J) {
  int64_t s; double u;
  g.computeTIndexNs(time_ns, s, u);
  double pw[4], coeff[4];
  baseCoeffsWithTime<D>(pw, u);
  matvec4(blend().M, pw, coeff);
  for (int i = 0; i < 4; ++i) coeff[i] = g.pow_inv_dt[D] * coeff[i];
  Vec3 res;
  for (int i = 0; i < kN; ++i) {
    res = res + coeff[i] * knotP(p, s + i);
    if (J) J->d_val_d_knot[i] = coeff[i];
  }
  if (J) J->start_idx = s;
  return res;
}

// split_spline_view.h:56-63
struct SplineIMUData {
  int64_t time_ns;
  Vec3 gyro;
  Vec3 accel;
  Quat R_inv;
  int64_t start_idx;
};

// split_spline_view.h:67-214.  (R_accum is sized DEG here; the reference
// declares DEG-1 and indexes one past the end)
inline SplineIMUData SplitEvaluate(const SplineGrid& g, const double* q, const double* p, int64_t time_ns,
                                   const Vec3& gravity, So3Jacobian* J_rot_w, So3Jacobian* J_rot_a,
                                   RdJacobian* J_pos) {
  SplineIMUData out;
  out.time_ns = time_ns;
  int64_t s; double u;
  g.computeTIndexNs(time_ns, s, u);
  out.start_idx = s;

  double Up[4], lambda_a[4], Ur[4], lambda_R[4], Uw[4], lambda_w[4];
  baseCoeffsWithTime<2>(Up, u);
  matvec4(blend().M, Up, lambda_a);
  for (int i = 0; i < 4; ++i) lambda_a[i] = g.pow_inv_dt[2] * lambda_a[i];
  baseCoeffsWithTime<0>(Ur, u);
  matvec4(blend().Mc, Ur, lambda_R);
  baseCoeffsWithTime<1>(Uw, u);
  matvec4(blend().Mc, Uw, lambda_w);
  for (int i = 0; i < 4; ++i) lambda_w[i] = g.pow_inv_dt[1] * lambda_w[i];

  Vec3 accelerate;
  if (J_pos) J_pos->start_idx = s;
  for (int i = 0; i < kN; ++i) {
    accelerate = accelerate + lambda_a[i] * knotP(p, s + i);
    if (J_pos) J_pos->d_val_d_knot[i] = lambda_a[i];
  }

  Vec3 d_vec[kDEG];
  Quat A_rot_inv[kDEG];
  Quat A_accum_inv;
  Mat3 A_post_inv[kN];
  Mat3 Jr_dvec_inv[kDEG], Jr_kdelta[kDEG];
  A_post_inv[kN - 1] = so3_matrix(A_accum_inv);
  for (int i = kDEG - 1; i >= 0; --i) {
    const Quat R0 = knotQ(q, s + i), R1 = knotQ(q, s + i + 1);
    d_vec[i] = so3_log(so3_mul(R0.conj(), R1));
    const Vec3 k_delta = lambda_R[i + 1] * d_vec[i];
    A_rot_inv[i] = so3_exp(-k_delta);
    A_accum_inv = so3_mul(A_accum_inv, A_rot_inv[i]);
    if (J_rot_w || J_rot_a) {
      A_post_inv[i] = so3_matrix(A_accum_inv);
      Jr_dvec_inv[i] = rightJacobianInvSO3(d_vec[i]);
      Jr_kdelta[i] = rightJacobianSO3(-k_delta);
    }
  }

  Vec3 omega[kN];
  for (int i = 0; i < kDEG; ++i) omega[i + 1] = so3_rotate(A_rot_inv[i], omega[i]) + lambda_w[i + 1] * d_vec[i];
  out.gyro = omega[3];
  const Quat Ri = knotQ(q, s);
  const Quat R_inv = so3_mul(A_accum_inv, Ri.conj());
  out.accel = so3_rotate(R_inv, accelerate + gravity);
  out.R_inv = R_inv;

  if (J_rot_w) {
    J_rot_w->start_idx = s;
    for (int i = 0; i < kN; ++i) J_rot_w->d_val_d_knot[i] = Mat3::Zero();
    Mat3 d_omega_d_delta[kDEG];
    d_omega_d_delta[0] = lambda_w[1] * A_post_inv[1];
    for (int i = 1; i < kDEG; ++i) {
      d_omega_d_delta[i] = lambda_R[i + 1] * (A_post_inv[i] * hat(omega[i]) * Jr_kdelta[i]) +
                           lambda_w[i + 1] * A_post_inv[i + 1];
    }
    for (int i = 0; i < kDEG; ++i) {
      J_rot_w->d_val_d_knot[i] = J_rot_w->d_val_d_knot[i] - d_omega_d_delta[i] * transpose(Jr_dvec_inv[i]);
      J_rot_w->d_val_d_knot[i + 1] = J_rot_w->d_val_d_knot[i + 1] + d_omega_d_delta[i] * Jr_dvec_inv[i];
    }
  }
  if (J_rot_a) {
    Mat3 R_accum[kDEG];  // R_i, R_i*A_1, R_i*A_1*A_2
    R_accum[0] = so3_matrix(Ri);
    for (int i = 1; i < kDEG; ++i) R_accum[i] = R_accum[i - 1] * transpose(so3_matrix(A_rot_inv[i - 1]));
    J_rot_a->start_idx = s;
    for (int i = 0; i < kN; ++i) J_rot_a->d_val_d_knot[i] = Mat3::Zero();
    const Mat3 lhs = so3_matrix(R_inv) * hat(accelerate + gravity);
    J_rot_a->d_val_d_knot[0] = J_rot_a->d_val_d_knot[0] + lhs * R_accum[0];
    for (int i = 0; i < kDEG; ++i) {
      const Mat3 d_a_d_delta = lambda_R[i + 1] * (lhs * R_accum[i] * Jr_kdelta[i]);
      J_rot_a->d_val_d_knot[i] = J_rot_a->d_val_d_knot[i] - d_a_d_delta * transpose(Jr_dvec_inv[i]);
      J_rot_a->d_val_d_knot[i + 1] = J_rot_a->d_val_d_knot[i + 1] + d_a_d_delta * Jr_dvec_inv[i];
    }
  }
  return out;
}

// ---------------------------------------------------------------------------
// Plain-spline evaluators: the reference's independent second implementation
// (value only), used by tests to cross-check the View evaluators above.
// so3_spline.h:240-292 evaluate, :294-321 velocityBody, :323-367 accelerationBody.
inline Quat PlainSo3Evaluate(const SplineGrid& g, const double* q, int64_t time_ns) {
  int64_t s; double u;
  g.computeTIndexNs(time_ns, s, u);
  double pw[4], coeff[4];
  baseCoeffsWithTime<0>(pw, u);
  matvec4(blend().Mc, pw, coeff);
  Quat res = knotQ(q, s);
  for (int i = 0; i < kDEG; ++i) {
    const Vec3 delta = so3_log(so3_mul(knotQ(q, s + i).conj(), knotQ(q, s + i + 1)));
    res = so3_mul(res, so3_exp(delta * coeff[i + 1]));
  }
  return res;
}
inline Vec3 PlainSo3VelocityBody(const SplineGrid& g, const double* q, int64_t time_ns) {
  int64_t s; double u;
  g.computeTIndexNs(time_ns, s, u);
  double pw[4], coeff[4], dcoeff[4];
  baseCoeffsWithTime<0>(pw, u);
  matvec4(blend().Mc, pw, coeff);
  baseCoeffsWithTime<1>(pw, u);
  matvec4(blend().Mc, pw, dcoeff);
  for (int i = 0; i < 4; ++i) dcoeff[i] *= g.pow_inv_dt[1];
  Vec3 rot_vel;
  for (int i = 0; i < kDEG; ++i) {
    const Vec3 delta = so3_log(so3_mul(knotQ(q, s + i).conj(), knotQ(q, s + i + 1)));
    rot_vel = so3_rotate(so3_exp(-(delta * coeff[i + 1])), rot_vel);
    rot_vel = rot_vel + delta * dcoeff[i + 1];
  }
  return rot_vel;
}
inline Vec3 PlainSo3AccelerationBody(const SplineGrid& g, const double* q, int64_t time_ns) {
  int64_t s; double u;
  g.computeTIndexNs(time_ns, s, u);
  double pw[4], coeff[4], dcoeff[4], ddcoeff[4];
  baseCoeffsWithTime<0>(pw, u);
  matvec4(blend().Mc, pw, coeff);
  baseCoeffsWithTime<1>(pw, u);
  matvec4(blend().Mc, pw, dcoeff);
  baseCoeffsWithTime<2>(pw, u);
  matvec4(blend().Mc, pw, ddcoeff);
  for (int i = 0; i < 4; ++i) {
    dcoeff[i] *= g.pow_inv_dt[1];
    ddcoeff[i] *= g.pow_inv_dt[2];
  }
  Vec3 rot_vel, rot_accel;
  for (int i = 0; i < kDEG; ++i) {
    const Vec3 delta = so3_log(so3_mul(knotQ(q, s + i).conj(), knotQ(q, s + i + 1)));
    const Quat rot = so3_exp(-(delta * coeff[i + 1]));
    rot_vel = so3_rotate(rot, rot_vel);
    const Vec3 vel_current = dcoeff[i + 1] * delta;
    rot_vel = rot_vel + vel_current;
    rot_accel = so3_rotate(rot, rot_accel);
    rot_accel = rot_accel + ddcoeff[i + 1] * delta + cross(rot_vel, vel_current);
  }
  return rot_accel;
}

}  // namespace ctvio_oracle
