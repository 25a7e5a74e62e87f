// TEST INFRASTRUCTURE: compiles the CUDA engine's per-thread math headers
// (ctrl-vio_b200/csrc/device_math.cuh, spline_eval.cuh) with plain g++ so the
// lane-level arithmetic can be compared with the oracle on a machine without a
// GPU.  Not linked into the product library; nothing in the product calls it.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../ctrl-vio_b200/csrc/spline_eval.cuh"

using namespace ctvio;

extern "C" {

// J layout == ctvio_eval_image_factors: [side][k][rot 2x3 | pos 2x3] (96) + rho (2) + ld (2)
int emu_eval_image(int64_t t0_ns, int64_t dt_ns, int n_knots, const double* q, const double* p, const double* q_CI,
                   const double* p_CI, double w_img, double rho, double ld, int64_t ti, int rowi, const double* pi,
                   int64_t tj, int rowj, const double* pj, double cauchy, int want_jac, double* r, int* s, double* J,
                   double* cost) {
  SplineParams sp{t0_ns, dt_ns, n_knots, 1e9 / double(dt_ns)};
  std::vector<KnotPair> tab(n_knots - 1);
  for (int k = 0; k < n_knots - 1; ++k) make_knot_pair(q, k, tab[k]);
  RigParams rig;
  rig.R_CI = so3_matrix(Q4{q_CI[0], q_CI[1], q_CI[2], q_CI[3]});
  rig.p_CI = V3{p_CI[0], p_CI[1], p_CI[2]};
  rig.w_img = w_img;
  const int64_t ld_ns = int64_t(ld * 1e9);
  int32_t si, sj;
  double ui, uj;
  if (!spline_index(sp, ti + int64_t(rowi) * ld_ns, si, ui)) return -1;
  if (!spline_index(sp, tj + int64_t(rowj) * ld_ns, sj, uj)) return -1;
  SideEval a, b;
  if (want_jac) {
    eval_side<true, 3>(sp, q, p, tab.data(), si, ui, a);
    eval_side<true, 3>(sp, q, p, tab.data(), sj, uj, b);
  } else {
    eval_side<false, 3>(sp, q, p, tab.data(), si, ui, a);
    eval_side<false, 3>(sp, q, p, tab.data(), sj, uj, b);
  }
  ImageCommon cm;
  image_common(rig, pi, pj, rho, a.R, a.p, b.R, b.p, cauchy, cm);
  r[0] = cm.r[0]; r[1] = cm.r[1];
  s[0] = si; s[1] = sj;
  *cost = cm.cost;
  if (!want_jac) return 0;
  double rot[4][6], pos[4][6];
  image_side_blocks(0, cm, a, rot, pos);
  for (int k = 0; k < 4; ++k) {
    std::memcpy(J + k * 12, rot[k], 6 * sizeof(double));
    std::memcpy(J + k * 12 + 6, pos[k], 6 * sizeof(double));
  }
  image_side_blocks(1, cm, b, rot, pos);
  for (int k = 0; k < 4; ++k) {
    std::memcpy(J + 48 + k * 12, rot[k], 6 * sizeof(double));
    std::memcpy(J + 48 + k * 12 + 6, pos[k], 6 * sizeof(double));
  }
  image_jrho(rig, cm, a.R, rho, J + 96);
  image_jld(rig, cm, rowi, rowj, a.R, a.omega, a.vel, b.R, b.omega, b.vel, J + 98);
  return 0;
}

// J layout == ctvio_eval_imu_factors: [k][rot 6x3 | pos 6x3] (144) + bg diag (6) + ba diag (6)
int emu_eval_imu(int64_t t0_ns, int64_t dt_ns, int n_knots, const double* q, const double* p, const double* gravity,
                 const double* imu_info, int64_t t, const double* gyro, const double* accel, const double* bias,
                 int want_jac, double* r, int* s, double* J, double* cost) {
  SplineParams sp{t0_ns, dt_ns, n_knots, 1e9 / double(dt_ns)};
  std::vector<KnotPair> tab(n_knots - 1);
  for (int k = 0; k < n_knots - 1; ++k) make_knot_pair(q, k, tab[k]);
  RigParams rig;
  rig.gravity = V3{gravity[0], gravity[1], gravity[2]};
  for (int k = 0; k < 6; ++k) rig.imu_info[k] = imu_info[k];
  int32_t si;
  double u;
  if (!spline_index(sp, t, si, u)) return -1;
  ImuEvalOut o;
  if (want_jac) eval_imu<true, 3>(sp, rig, q, p, tab.data(), si, u, gyro, accel, bias, o);
  else eval_imu<false, 3>(sp, rig, q, p, tab.data(), si, u, gyro, accel, bias, o);
  std::memcpy(r, o.r, sizeof(o.r));
  *s = si;
  *cost = o.cost;
  if (!want_jac) return 0;
  for (int k = 0; k < 4; ++k) {
    std::memcpy(J + k * 36, o.Jrot[k], 18 * sizeof(double));
    std::memcpy(J + k * 36 + 18, o.Jpos[k], 18 * sizeof(double));
  }
  for (int k = 0; k < 3; ++k) {
    J[144 + k] = imu_info[k]; J[147 + k] = 0; J[150 + k] = 0; J[153 + k] = imu_info[3 + k];
  }
  return 0;
}

// two-stage pose / Jacobian path used by the fused visual kernel vs the one-shot eval_side.
// out: max abs difference over R, p, omega, vel, c and the four Jacobian blocks
double emu_two_stage_maxdiff(int64_t t0_ns, int64_t dt_ns, int n_knots, const double* q, const double* p, int64_t t) {
  SplineParams sp{t0_ns, dt_ns, n_knots, 1e9 / double(dt_ns)};
  std::vector<KnotPair> tab(n_knots - 1);
  for (int k = 0; k < n_knots - 1; ++k) make_knot_pair(q, k, tab[k]);
  int32_t s;
  double u;
  if (!spline_index(sp, t, s, u)) return -1.0;
  SideEval a;
  eval_side<true, 3>(sp, q, p, tab.data(), s, u, a);
  PoseStage b;
  pose_stage<true, 3>(sp, q, p, tab.data(), s, u, b);
  double m = 0;
  auto upd = [&](double x, double y) { m = std::max(m, std::fabs(x - y)); };
  for (int e = 0; e < 9; ++e) upd(a.R.m[e], b.R.m[e]);
  upd(a.p.x, b.p.x); upd(a.p.y, b.p.y); upd(a.p.z, b.p.z);
  upd(a.omega.x, b.omega.x); upd(a.omega.y, b.omega.y); upd(a.omega.z, b.omega.z);
  upd(a.vel.x, b.vel.x); upd(a.vel.y, b.vel.y); upd(a.vel.z, b.vel.z);
  for (int k = 0; k < 4; ++k) upd(a.c[k], b.c[k]);
  jacobian_stage(tab.data(), b, [&](int k, const M3& Jk) { for (int e = 0; e < 9; ++e) upd(a.J[k].m[e], Jk.m[e]); });
  return m;
}

void emu_quat_from_matrix(const double* m, double* q) {
  M3 a;
  std::memcpy(a.m, m, sizeof(a.m));
  Q4 r = quat_from_matrix(a);
  q[0] = r.x; q[1] = r.y; q[2] = r.z; q[3] = r.w;
}
}
