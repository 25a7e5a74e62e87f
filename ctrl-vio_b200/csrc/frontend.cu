// Front-end data formats either side of the hot path.
//
//   triangulate_kernel   replaces FeatureManager::triangulate (visual_odometry/feature_manager.cpp:173-223 and the
//                        camera-extrinsic overload :230-275): per-landmark DLT, depth = V(2)/V(3) of the right singular
//                        vector of the smallest singular value of the 2m x 4 system, INIT_DEPTH fallback below 0.1.
//   unpack_cloud_kernel  replaces FeatureMsg2Image (visual_odometry/visual_struct.h:98-121) on the tracker's message
//                        (visual_feature/feature_tracker_node.cpp:146-184): sensor_msgs::PointCloud arrives as packed
//                        float32 triples + five float32 channels and is converted ON THE DEVICE into the resident
//                        per-frame feature table (id, bearing xy, pixel row).
//   unpack_imu_kernel    IMUData (utils/parameter_struct.h:58-65) records -> {t, gyro, accel} table.
//   gather_factors_kernel builds the sorted SoA image-factor arrays of K1 from the resident tables and an 8-byte
//                        (slot_i, slot_j) descriptor per factor: the payload never passes through host marshalling.
//
// The reference runs Eigen::JacobiSVD on the tall matrix (QR preconditioner + two-sided Jacobi on R).  Here one thread
// per landmark streams the rows through a Givens QR (R stays in registers, any number of frames) and then runs a
// one-sided Jacobi SVD on the 4x4 R: same conditioning as the reference (no A'A squaring), no local-memory arrays.
#include "frontend.h"

#include <algorithm>

namespace ctvio {

namespace {

struct R4 {
  double r[4][4];  // upper triangular
};

// fold one row a[4] into R with 4 Givens rotations
__device__ __forceinline__ void qr_push_row(R4& R, double a0, double a1, double a2, double a3) {
  double a[4] = {a0, a1, a2, a3};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double x = R.r[c][c], y = a[c];
    if (y == 0.0) continue;
    const double h = hypot(x, y);
    const double cs = x / h, sn = y / h;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < c) continue;
      const double rk = R.r[c][k], ak = a[k];
      R.r[c][k] = cs * rk + sn * ak;
      a[k] = -sn * rk + cs * ak;
    }
  }
}

// right singular vector of the smallest singular value of the upper triangular R (one-sided Jacobi, Hestenes)
__device__ __forceinline__ void smallest_right_singular_vector(const R4& R, double v_out[4]) {
  double G[4][4], V[4][4];  // column-major use: G[row][col]
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      G[i][j] = j >= i ? R.r[i][j] : 0.0;
      V[i][j] = i == j ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 40; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int q = p + 1; q < 4; ++q) {
        double al = 0, be = 0, ga = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          al = fma(G[i][p], G[i][p], al);
          be = fma(G[i][q], G[i][q], be);
          ga = fma(G[i][p], G[i][q], ga);
        }
        if (ga == 0.0 || fabs(ga) <= 1e-300 || fabs(ga) <= 2.3e-16 * sqrt(al * be)) continue;
        rotated = true;
        const double zeta = (be - al) / (2.0 * ga);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double gp = G[i][p], gq = G[i][q];
          G[i][p] = c * gp - s * gq;
          G[i][q] = s * gp + c * gq;
          const double vp = V[i][p], vq = V[i][q];
          V[i][p] = c * vp - s * vq;
          V[i][q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  int best = 0;
  double best_n = 1e300;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double nj = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) nj = fma(G[i][j], G[i][j], nj);
    if (nj < best_n) { best_n = nj; best = j; }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    // select without dynamic register indexing
    v_out[i] = best == 0 ? V[i][0] : best == 1 ? V[i][1] : best == 2 ? V[i][2] : V[i][3];
  }
}

__global__ void triangulate_kernel(TriangulateArgs a) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= a.n_landmarks) return;
  const int o0 = a.obs_offset[l], used = a.obs_offset[l + 1] - o0;
  const int start = a.start_frame[l];
  // feature_manager.cpp:236-240: candidates only, already-initialised depths are kept
  if (!(used >= 2 && start < a.window_size - 2)) return;
  if (a.depth[l] > 0.0) return;
  if (start < 0 || start + used > a.n_frames) { a.depth[l] = a.init_depth; return; }
  auto cam_pose = [&](int f, M3& Rc, V3& tc) {
    M3 Rf;
#pragma unroll
    for (int e = 0; e < 9; ++e) Rf.m[e] = a.Rs[9 * f + e];
    const V3 Pf{a.Ps[3 * f], a.Ps[3 * f + 1], a.Ps[3 * f + 2]};
    Rc = m3_mul(Rf, a.ric);          // R0 = Rs[i] * ric        (:246)
    tc = Pf + m3_vec(Rf, a.tic);     // t0 = Ps[i] + Rs[i] tic  (:245)
  };
  M3 R0;
  V3 t0;
  cam_pose(start, R0, t0);
  R4 Rq;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Rq.r[i][j] = 0.0;
  for (int k = 0; k < used; ++k) {
    M3 R1;
    V3 t1;
    cam_pose(start + k, R1, t1);
    const V3 t = m3_tvec(R0, t1 - t0);
    const M3 R = m3_mul(m3_transpose(R0), R1);
    // P = [R' | -R' t]   (:255-257)
    const M3 Rt = m3_transpose(R);
    const V3 pt = neg(m3_vec(Rt, t));
    V3 f{a.obs_point[3 * (o0 + k)], a.obs_point[3 * (o0 + k) + 1], a.obs_point[3 * (o0 + k) + 2]};
    const double fn = sqrt(f.x * f.x + f.y * f.y + f.z * f.z);
    f = (1.0 / fn) * f;
    const double P0[4] = {Rt.m[0], Rt.m[1], Rt.m[2], pt.x};
    const double P1[4] = {Rt.m[3], Rt.m[4], Rt.m[5], pt.y};
    const double P2[4] = {Rt.m[6], Rt.m[7], Rt.m[8], pt.z};
    qr_push_row(Rq, f.x * P2[0] - f.z * P0[0], f.x * P2[1] - f.z * P0[1], f.x * P2[2] - f.z * P0[2],
                f.x * P2[3] - f.z * P0[3]);
    qr_push_row(Rq, f.y * P2[0] - f.z * P1[0], f.y * P2[1] - f.z * P1[1], f.y * P2[2] - f.z * P1[2],
                f.y * P2[3] - f.z * P1[3]);
  }
  double v[4];
  smallest_right_singular_vector(Rq, v);
  double d = v[2] / v[3];
  if (!(d >= 0.1)) d = a.init_depth;  // :268-271 (NaN / inf from v[3] == 0 also fall back)
  if (!isfinite(d)) d = a.init_depth;
  a.depth[l] = d;
}

// ---- wire formats -> resident tables ----------------------------------------------------------------

__global__ void unpack_cloud_kernel(UnpackCloudArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  // FeatureMsg2Image: id = int(channels[0] + 0.5); x, y from points[i] (z == 1); p_v = channels[2]
  const int id = int(double(a.ch_id[i]) + 0.5);
  const double x = double(a.points[3 * i]), y = double(a.points[3 * i + 1]);
  FrameFeature f;
  f.x = x; f.y = y;
  f.id = id;
  f.row = int(round(double(a.ch_v[i])));  // std::round(uv(1)) at the Add* call site (trajectory_manager.cpp:366, 374)
  a.out[i] = f;
}

__global__ void unpack_imu_kernel(UnpackImuArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const unsigned char* rec = a.raw + size_t(i) * a.stride;
  // IMUData: int64 timestamp @0, Vector3d gyro, Vector3d accel (offsets given by the caller)
  const int64_t t = *reinterpret_cast<const int64_t*>(rec);
  const double* g = reinterpret_cast<const double*>(rec + a.off_gyro);
  const double* ac = reinterpret_cast<const double*>(rec + a.off_accel);
  // bias node of the sample (trajectory_manager.cpp:383-403): 0 before kf 0, last at / after the newest kf,
  // else the interval [kf_{i-1}, kf_i) -> i - 1
  int node = 0;
  if (a.n_kf > 0) {
    if (t < a.kf_t[0]) node = 0;
    else if (t >= a.kf_t[a.n_kf - 1]) node = a.n_kf - 1;
    else {
      for (int k = 1; k < a.n_kf; ++k)
        if (t >= a.kf_t[k - 1] && t < a.kf_t[k]) { node = k - 1; break; }
    }
  }
  a.t_node[a.dst0 + i] = make_longlong2(t, node);
  a.ga[3 * size_t(a.dst0 + i)] = make_double2(g[0], g[1]);
  a.ga[3 * size_t(a.dst0 + i) + 1] = make_double2(g[2], ac[0]);
  a.ga[3 * size_t(a.dst0 + i) + 2] = make_double2(ac[1], ac[2]);
}

__global__ void gather_factors_kernel(GatherFactorsArgs a) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.n) return;
  const FactorDesc d = a.desc[k];
  const FrameFeature fi = a.table[d.slot_i], fj = a.table[d.slot_j];
  a.t[k] = make_longlong2(a.frame_t[d.slot_i / a.frame_cap], a.frame_t[d.slot_j / a.frame_cap]);
  a.pi[k] = make_double2(fi.x, fi.y);
  a.pj[k] = make_double2(fj.x, fj.y);
  a.meta[k] = make_int4(fi.row, fj.row, d.lm, d.marg);
}

// ---- device-resident window bookkeeping: nothing below touches the host ---------------------------
__global__ void extend_knots_kernel(StatePtrs st, int old_n, int new_n) {
  const int k = old_n + blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= new_n) return;
  // trajectory_manager.cpp:114-115: extendKnotsTo(max_time, last_knot)
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    st.q[4 * k + c] = st.q[4 * (old_n - 1) + c];
    st.p[kPStride * k + c] = st.p[kPStride * (old_n - 1) + c];
  }
}
__global__ void slide_copy_kernel(StatePtrs st, int nK, int nB, double* tmp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4 * nK) tmp[i] = st.q[i];
  if (i < kPStride * nK) tmp[4 * nK + i] = st.p[i];
  if (i < 6 * nB) tmp[8 * nK + i] = st.bias[i];
}
__global__ void slide_shift_kernel(StatePtrs st, int nK, int nB, int dk, int db, int new_bias, const double* tmp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4 * (nK - dk)) st.q[i] = tmp[4 * dk + i];
  if (i < kPStride * (nK - dk)) st.p[i] = tmp[4 * nK + kPStride * dk + i];
  const int keepb = nB - db;
  if (i < 6 * (keepb + new_bias)) {
    const int b = i / 6, c = i - 6 * b;
    // kept nodes move down; appended nodes start from the newest estimate (Bgs_[WINDOW_SIZE] after slideWindow)
    const int srcb = b < keepb ? b + db : nB - 1;
    st.bias[i] = nB > 0 ? tmp[8 * nK + 6 * srcb + c] : 0.0;
  }
}
__global__ void remap_rho_kernel(const double* old_rho, const int32_t* old_index, const double* init_rho, int n, double* out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n) return;
  out[l] = old_index[l] >= 0 ? old_rho[old_index[l]] : init_rho[l];
}
__global__ void shift_imu_copy_kernel(const longlong2* t, const double2* ga, int from, int count, double* tmp) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  longlong2* tt = reinterpret_cast<longlong2*>(tmp);
  double2* tg = reinterpret_cast<double2*>(tmp + 2 * size_t(count));
  tt[k] = t[from + k];
#pragma unroll
  for (int c = 0; c < 3; ++c) tg[3 * k + c] = ga[3 * size_t(from + k) + c];
}
__global__ void shift_imu_back_kernel(longlong2* t, double2* ga, int count, const double* tmp) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const longlong2* tt = reinterpret_cast<const longlong2*>(tmp);
  const double2* tg = reinterpret_cast<const double2*>(tmp + 2 * size_t(count));
  t[k] = tt[k];
#pragma unroll
  for (int c = 0; c < 3; ++c) ga[3 * size_t(k) + c] = tg[3 * k + c];
}
__global__ void gather_imu_kernel(const int2* src, int n, const longlong2* tab_t, const double2* tab_ga, longlong2* out_t,
                                  double2* out_ga) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int2 s = src[k];
  out_t[k] = make_longlong2(tab_t[s.x].x, s.y);
#pragma unroll
  for (int c = 0; c < 3; ++c) out_ga[3 * size_t(k) + c] = tab_ga[3 * size_t(s.x) + c];
}
// linearisation point of the kept blocks of a fresh prior (keep_block_data, marginalization_factor.cpp:223-236)
__global__ void prior_x0_kernel(StatePtrs st, const int32_t* type, const int32_t* index, int nb, double* x0) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const int t = type[b], i = index[b];
  double v[4] = {0, 0, 0, 0};
  if (t == 0) { for (int c = 0; c < 4; ++c) v[c] = st.q[4 * i + c]; }
  else if (t == 1) { for (int c = 0; c < 3; ++c) v[c] = st.p[kPStride * i + c]; }
  else if (t == 2) { for (int c = 0; c < 3; ++c) v[c] = st.bias[6 * i + c]; }
  else if (t == 3) { for (int c = 0; c < 3; ++c) v[c] = st.bias[6 * i + 3 + c]; }
  else if (t == 4) v[0] = *st.ld;
  for (int c = 0; c < 4; ++c) x0[4 * b + c] = v[c];
}

}  // namespace

int launch_triangulate(const TriangulateArgs& a, cudaStream_t s) {
  if (a.n_landmarks <= 0) return 0;
  triangulate_kernel<<<(a.n_landmarks + 127) / 128, 128, 0, s>>>(a);
  return 1;
}
int launch_unpack_cloud(const UnpackCloudArgs& a, cudaStream_t s) {
  if (a.n <= 0) return 0;
  unpack_cloud_kernel<<<(a.n + 127) / 128, 128, 0, s>>>(a);
  return 1;
}
int launch_unpack_imu(const UnpackImuArgs& a, cudaStream_t s) {
  if (a.n <= 0) return 0;
  unpack_imu_kernel<<<(a.n + 127) / 128, 128, 0, s>>>(a);
  return 1;
}
int launch_gather_factors(const GatherFactorsArgs& a, cudaStream_t s) {
  if (a.n <= 0) return 0;
  gather_factors_kernel<<<(a.n + 127) / 128, 128, 0, s>>>(a);
  return 1;
}

int launch_extend_knots(const StatePtrs& st, int old_n, int new_n, cudaStream_t s) {
  if (new_n <= old_n) return 0;
  extend_knots_kernel<<<(new_n - old_n + 63) / 64, 64, 0, s>>>(st, old_n, new_n);
  return 1;
}
int launch_slide_state(const StatePtrs& st, int nK, int nB, int dk, int db, int new_bias, double* tmp, cudaStream_t s) {
  const int work = std::max(8 * nK, 6 * (nB + new_bias)) + 1;
  slide_copy_kernel<<<(work + 127) / 128, 128, 0, s>>>(st, nK, nB, tmp);
  slide_shift_kernel<<<(work + 127) / 128, 128, 0, s>>>(st, nK, nB, dk, db, new_bias, tmp);
  return 2;
}
int launch_remap_rho(const double* old_rho, const int32_t* old_index, const double* init_rho, int n, double* out, cudaStream_t s) {
  if (n <= 0) return 0;
  remap_rho_kernel<<<(n + 127) / 128, 128, 0, s>>>(old_rho, old_index, init_rho, n, out);
  return 1;
}
int launch_shift_imu_table(longlong2* t, double2* ga, int from, int count, double* tmp, cudaStream_t s) {
  if (count <= 0 || from <= 0) return 0;
  shift_imu_copy_kernel<<<(count + 127) / 128, 128, 0, s>>>(t, ga, from, count, tmp);
  shift_imu_back_kernel<<<(count + 127) / 128, 128, 0, s>>>(t, ga, count, tmp);
  return 2;
}
int launch_gather_imu(const int2* src, int n, const longlong2* tab_t, const double2* tab_ga, longlong2* out_t, double2* out_ga,
                      cudaStream_t s) {
  if (n <= 0) return 0;
  gather_imu_kernel<<<(n + 127) / 128, 128, 0, s>>>(src, n, tab_t, tab_ga, out_t, out_ga);
  return 1;
}
int launch_prior_x0(const StatePtrs& st, const int32_t* type, const int32_t* index, int nb, double* x0, cudaStream_t s) {
  if (nb <= 0) return 0;
  prior_x0_kernel<<<(nb + 63) / 64, 64, 0, s>>>(st, type, index, nb, x0);
  return 1;
}

}  // namespace ctvio
