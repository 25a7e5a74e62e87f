// Linear-algebra / LM-step kernels of the CUDA engine (sm_100a, fp64).
//
// Replaces what Ceres does inside ceres::Solve for one trust-region step
// (trajectory_estimator.cpp:367-408 -> Ceres 1.14 levenberg_marquardt_strategy.cc +
// SPARSE_NORMAL_CHOLESKY; Ceres is not under /root/reference):
//   K4  scale_copy + schur     reduced camera system  M = S A S + D^2 - sum_l w_l w_l' / h_l
//   K5  chol_panel/chol_update blocked right-looking Cholesky (NB = 64) + block triangular solves
//   K6  backsub / quad / apply landmark back-substitution, model cost change, x (+) delta, norms
// The Schur complement is a grouped SYRK: landmarks are batched by knot range on the host so one CTA
// reduces a batch in a register-tiled 8x8 SYRK and flushes once.
#include "kernels.h"

namespace ctvio {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// atomic max for non-negative doubles (bit pattern order == value order)
__device__ __forceinline__ void atomic_max_pos(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), static_cast<unsigned long long>(__double_as_longlong(v)));
}

constexpr double kMinLmDiag = 1e-6, kMaxLmDiag = 1e32;  // Ceres min/max_lm_diagonal

// ------------------------------------------------------------------------------------------------
__global__ void jacobi_scale_kernel(LinearLaunch a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int np = a.dims.np;
  if (i < np) a.sc[i] = 1.0 / (1.0 + sqrt(a.ne.A[size_t(i) * np + i]));
  if (i < a.dims.nL) a.sl[i] = 1.0 / (1.0 + sqrt(a.ne.hl[i]));
}
int launch_jacobi_scale(const LinearLaunch& a, cudaStream_t s) {
  const int n = max(a.dims.np, a.dims.nL);
  jacobi_scale_kernel<<<(n + 255) / 256, 256, 0, s>>>(a);
  return 1;
}

// M = S A S + clamp(diag)/radius on the full (padded, symmetric) matrix; rhs = S g
__global__ void scale_copy_kernel(LinearLaunch a, double radius) {
  const int npad = a.npad, np = a.dims.np;
  const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= size_t(npad) * npad) return;
  const int i = int(idx / npad), j = int(idx % npad);
  double m;
  const double ident = a.sharded ? 0.0 : 1.0;  // sharded: identity rows are set after the all-reduce
  if (i < np && j < np) {
    if (a.cmask[i] || a.cmask[j]) {
      m = (i == j) ? ident : 0.0;
    } else {
      const double v = (i <= j) ? a.ne.A[size_t(i) * np + j] : a.ne.A[size_t(j) * np + i];
      m = a.sc[i] * v * a.sc[j];
      if (i == j) {
        if (a.sharded) a.diagA[i] = v;  // unscaled local diagonal, summed over ranks with M
        else m += fmin(fmax(m, kMinLmDiag), kMaxLmDiag) / radius;
      }
    }
  } else {
    m = (i == j) ? ident : 0.0;
    if (a.sharded && i == j) a.diagA[i] = 0.0;
  }
  if (a.sharded && i == j && i < np && a.cmask[i]) a.diagA[i] = 0.0;
  a.M[idx] = m;
  if (j == 0) {
    a.rhs[i] = (i < np && !a.cmask[i]) ? a.sc[i] * a.ne.gc[i] : 0.0;
    if (i == 0) {
      a.scal->gd = 0.0;
      a.scal->dHd = 0.0;
      a.scal->dir_max = 0.0;
      a.scal->chol_fail = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4: Schur complement of one landmark batch
__global__ void __launch_bounds__(256) schur_batch_kernel(LinearLaunch a, double radius) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  double* V = reinterpret_cast<double*>(dyn_smem);
  const SchurBatch b = a.batches[blockIdx.x];
  const int U = b.uhi - b.ulo;
  const int LD = ((U + 2 + 7) / 8) * 8;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ild = a.dims.idx_ld;
  // prologue: v_l = sl * W_l o sc / sqrt(hh_l), extra columns: line delay, scaled landmark gradient
  for (int r = warp; r < b.count; r += 8) {
    const int l = a.schur_order[b.first + r];
    const double sl = a.sl[l];
    const double hs = sl * sl * a.ne.hl[l];
    const double hh = hs + fmin(fmax(hs, kMinLmDiag), kMaxLmDiag) / radius;
    if (lane == 0) a.hh[l] = hh;
    const double is = rsqrt(hh) * sl;
    const int lo = a.lm.lo[l], hi = a.lm.hi[l];
    const double* Wl = a.ne.W + a.lm.woff[l] - lo;
    double* row = V + size_t(r) * LD;
    for (int c = lane; c < LD; c += 32) {
      const int g = b.ulo + c;
      double v = 0.0;
      if (c < U) {
        if (g >= lo && g < hi) v = is * Wl[g] * a.sc[g];
      } else if (c == U) {
        v = is * a.ne.wld[l] * a.sc[ild];
      } else if (c == U + 1) {
        v = is * a.ne.gl[l];
      }
      row[c] = v;
    }
  }
  __syncthreads();
  const int nt = LD / 8;
  const int ntiles = nt * (nt + 1) / 2;
  const int npad = a.npad;
  for (int t = tid; t < ntiles; t += 256) {
    int ti = 0, rem = t;
    while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
    const int tj = ti + rem;
    double acc[64];
#pragma unroll
    for (int e = 0; e < 64; ++e) acc[e] = 0.0;
    for (int r = 0; r < b.count; ++r) {
      const double2* ra = reinterpret_cast<const double2*>(V + size_t(r) * LD + ti * 8);
      const double2* rb = reinterpret_cast<const double2*>(V + size_t(r) * LD + tj * 8);
      double av[8], bv[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const double2 x = ra[e], y = rb[e];
        av[2 * e] = x.x; av[2 * e + 1] = x.y;
        bv[2 * e] = y.x; bv[2 * e + 1] = y.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i * 8 + j] = fma(av[i], bv[j], acc[i * 8 + j]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double val = acc[i * 8 + j];
        const int la = ti * 8 + i, lb = tj * 8 + j;
        if (val == 0.0 || la > lb || la > U || lb > U + 1) continue;
        const int ga = la < U ? b.ulo + la : ild;
        if (lb == U + 1) {
          atomicAdd(a.rhs + ga, -val);
          continue;
        }
        const int gb = lb < U ? b.ulo + lb : ild;
        atomicAdd(a.M + size_t(ga) * npad + gb, -val);
        if (ga != gb) atomicAdd(a.M + size_t(gb) * npad + ga, -val);
      }
  }
}

// slow path for landmarks whose own knot range is wider than kSchurMaxDim: one warp per landmark
__global__ void schur_wide_kernel(LinearLaunch a, double radius) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= a.n_wide) return;
  const int l = a.wide_lms[w];
  const double sl = a.sl[l];
  const double hs = sl * sl * a.ne.hl[l];
  const double hh = hs + fmin(fmax(hs, kMinLmDiag), kMaxLmDiag) / radius;
  if (lane == 0) a.hh[l] = hh;
  const double inv = sl * sl / hh;
  const int lo = a.lm.lo[l], hi = a.lm.hi[l], n = hi - lo + 1, ild = a.dims.idx_ld, npad = a.npad;
  const double* Wl = a.ne.W + a.lm.woff[l] - lo;
  for (int e = lane; e < n * n; e += 32) {
    const int ia = e / n, ib = e % n;
    const int ga = ia < n - 1 ? lo + ia : ild, gb = ib < n - 1 ? lo + ib : ild;
    const double wa = (ia < n - 1 ? Wl[ga] : a.ne.wld[l]) * a.sc[ga];
    const double wb = (ib < n - 1 ? Wl[gb] : a.ne.wld[l]) * a.sc[gb];
    const double v = wa * wb * inv;
    if (v != 0.0) atomicAdd(a.M + size_t(ga) * npad + gb, -v);
    if (ib == 0) {
      const double g = wa * a.ne.gl[l] * inv;
      if (g != 0.0) atomicAdd(a.rhs + ga, -g);
    }
  }
}

// K5 (blocked Cholesky + triangular solves) lives in chol_coop.cu

// ------------------------------------------------------------------------------------------------
// K6
// camera part of the step: dc = -sc o y ; gd += gc.dc ; dHd += dc' A dc ; dir_max
__global__ void __launch_bounds__(256) camera_step_kernel(LinearLaunch a) {
  __shared__ double red[3][8];
  const int np = a.dims.np;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);  // one warp per row
  const int lane = threadIdx.x & 31;
  double gd = 0, dHd = 0, dmax = 0;
  if (i < np) {
    const double di = a.cmask[i] ? 0.0 : -a.sc[i] * a.y[i];
    if (lane == 0) {
      a.dc[i] = di;
      gd = a.ne.gc[i] * di;
      dmax = fabs(di);
      if (!isfinite(di)) a.scal->chol_fail = 1;
    }
    if (di != 0.0) {
      double s = 0;
      for (int j = i + lane; j < np; j += 32) {
        const double dj = a.cmask[j] ? 0.0 : -a.sc[j] * a.y[j];
        s = fma(a.ne.A[size_t(i) * np + j] * (j == i ? 0.5 : 1.0), dj, s);
      }
      dHd = 2.0 * di * s;
    }
  }
  dHd = warp_sum_d(dHd);
  if (lane == 0) { red[0][threadIdx.x >> 5] = gd; red[1][threadIdx.x >> 5] = dHd; red[2][threadIdx.x >> 5] = dmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double g = 0, h = 0, m = 0;
    for (int w = 0; w < 8; ++w) { g += red[0][w]; h += red[1][w]; m = fmax(m, red[2][w]); }
    atomicAdd(&a.scal->gd, g);
    atomicAdd(&a.scal->dHd, h);
    atomic_max_pos(&a.scal->dir_max, m);
  }
}

// landmark back-substitution + landmark parts of gd / dHd: one WARP per landmark (coalesced reads of
// its coupling row), reads y (not dc) so that it can run concurrently with camera_step_kernel.
__global__ void __launch_bounds__(256) landmark_step_kernel(LinearLaunch a) {
  __shared__ double red[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l = blockIdx.x * 8 + warp;
  double gd = 0, dHd = 0, dmax = 0;
  if (l < a.dims.nL) {
    const double sl = a.sl[l];
    const int lo = a.lm.lo[l], hi = a.lm.hi[l], ild = a.dims.idx_ld;
    const double* Wl = a.ne.W + a.lm.woff[l] - lo;
    double wy = 0;  // W_l . (sc o y)
    for (int g = lo + lane; g < hi; g += 32) wy = fma(Wl[g], a.cmask[g] ? 0.0 : a.sc[g] * a.y[g], wy);
    wy = warp_sum_d(wy);
    if (lane == 0) {
      wy = fma(a.ne.wld[l], a.cmask[ild] ? 0.0 : a.sc[ild] * a.y[ild], wy);
      const double hh = a.hh[l];
      const double yl = hh > 0.0 ? (sl * a.ne.gl[l] - sl * wy) / hh : 0.0;
      const double d = -sl * yl;
      a.dl[l] = d;
      gd = a.ne.gl[l] * d;
      dHd = 2.0 * d * (-wy) + a.ne.hl[l] * d * d;  // W_l . dc = -wy
      dmax = fabs(d);
      if (!isfinite(d)) a.scal->chol_fail = 1;
    }
  }
  if (lane == 0) { red[0][warp] = gd; red[1][warp] = dHd; red[2][warp] = dmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double g = 0, h = 0, m = 0;
    for (int w = 0; w < 8; ++w) { g += red[0][w]; h += red[1][w]; m = fmax(m, red[2][w]); }
    if (g != 0.0) atomicAdd(&a.scal->gd, g);
    if (h != 0.0) atomicAdd(&a.scal->dHd, h);
    if (m > 0.0) atomic_max_pos(&a.scal->dir_max, m);
  }
}

// after the all-reduce of [M | rhs | diagA]
__global__ void add_damping_kernel(LinearLaunch a, double radius) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.npad) return;
  double* d = a.M + size_t(i) * a.npad + i;
  if (i < a.dims.np && !a.cmask[i]) {
    const double sd = a.sc[i] * a.sc[i] * a.diagA[i];
    *d += fmin(fmax(sd, kMinLmDiag), kMaxLmDiag) / radius;
  } else {
    *d = 1.0;
    a.rhs[i] = 0.0;
  }
}
int launch_add_damping(const LinearLaunch& a, double radius, cudaStream_t s) {
  add_damping_kernel<<<(a.npad + 255) / 256, 256, 0, s>>>(a, radius);
  return 1;
}
__global__ void extract_diag_kernel(LinearLaunch a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.npad) return;
  a.diagA[i] = i < a.dims.np ? a.ne.A[size_t(i) * a.dims.np + i] : 0.0;
}
int launch_extract_diag(const LinearLaunch& a, cudaStream_t s) {
  extract_diag_kernel<<<(a.npad + 255) / 256, 256, 0, s>>>(a);
  return 1;
}
__global__ void jacobi_scale_from_diag_kernel(LinearLaunch a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.dims.np) a.sc[i] = 1.0 / (1.0 + sqrt(a.diagA[i]));
  if (i < a.dims.nL) a.sl[i] = 1.0 / (1.0 + sqrt(a.ne.hl[i]));
}
int launch_jacobi_scale_from_diag(const LinearLaunch& a, cudaStream_t s) {
  const int n = max(a.dims.np, a.dims.nL);
  jacobi_scale_from_diag_kernel<<<(n + 255) / 256, 256, 0, s>>>(a);
  return 1;
}

int launch_reduced_system(const LinearLaunch& a, double radius, cudaStream_t s) {
  int launches = 0;
  const size_t total = size_t(a.npad) * a.npad;
  scale_copy_kernel<<<unsigned((total + 255) / 256), 256, 0, s>>>(a, radius);
  ++launches;
  if (a.n_batches > 0) {
    const size_t smem = size_t(kSchurBatch) * (kSchurMaxDim + 8) * sizeof(double);
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(schur_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      attr_set = true;
    }
    schur_batch_kernel<<<a.n_batches, 256, smem, s>>>(a, radius);
    ++launches;
  }
  if (a.n_wide > 0) {
    schur_wide_kernel<<<(a.n_wide * 32 + 255) / 256, 256, 0, s>>>(a, radius);
    ++launches;
  }
  return launches;
}

int launch_step_vectors(const LinearLaunch& a, cudaStream_t s) {
  int launches = 0;
  camera_step_kernel<<<(a.dims.np + 7) / 8, 256, 0, s>>>(a);
  ++launches;
  if (a.dims.nL > 0) {
    landmark_step_kernel<<<(a.dims.nL + 7) / 8, 256, 0, s>>>(a);
    ++launches;
  }
  return launches;
}

int launch_lm_step(const LinearLaunch& a, double radius, cudaStream_t s) {
  return launch_reduced_system(a, radius, s) + launch_factor_solve(a, s) + launch_step_vectors(a, s);
}

// max-norm of the (bounds-projected) gradient over the active parameters
__global__ void gradient_norm_kernel(LinearLaunch a, StatePtrs st, int fix_ld, double ld_lower, double ld_upper) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int np = a.dims.np, nL = a.dims.nL;
  double v = 0.0;
  if (i < np) {
    if (a.active[i]) {
      v = fabs(a.ne.gc[i]);
      if (i == a.dims.idx_ld && !fix_ld) {
        const double ld = *st.ld;
        v = fabs(ld - fmin(fmax(ld - a.ne.gc[i], ld_lower), ld_upper));
      }
    }
  } else if (i < np + nL) {
    if (a.active[i]) v = fabs(a.ne.gl[i - np]);
  }
  v = warp_max_d(v);
  if ((threadIdx.x & 31) == 0 && v > 0.0) atomic_max_pos(&a.scal->gmax, v);
}
int launch_gradient_norm(const LinearLaunch& a, const StatePtrs& st, int fix_ld, double ld_lower, double ld_upper,
                         cudaStream_t s) {
  const int n = a.dims.np + a.dims.nL;
  cudaMemsetAsync(&a.scal->gmax, 0, sizeof(double), s);
  gradient_norm_kernel<<<(n + 255) / 256, 256, 0, s>>>(a, st, fix_ld, ld_lower, ld_upper);
  return 1;
}

// x+ = x (+) alpha*delta  (SO(3): q * exp(delta), ceres_local_param.h:137-145; box projection of the
// line delay, Ceres parameter_block.h Plus) and the ambient norms |x|^2, |x - x+|^2 over active blocks
__global__ void __launch_bounds__(256) apply_step_kernel(ApplyLaunch a) {
  __shared__ double red[2][8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nK = a.dims.nK, nB = a.dims.nB, nL = a.dims.nL;
  double xn = 0, sn = 0;
  if (i < nK) {
    const double* d = a.dc + 6 * i;
    const Q4 q = load_q(a.x.q, i);
    Q4 qn = q;
    if (d[0] != 0.0 || d[1] != 0.0 || d[2] != 0.0) qn = so3_mul(q, so3_exp(V3{a.alpha * d[0], a.alpha * d[1], a.alpha * d[2]}));
    a.xc.q[4 * i] = qn.x; a.xc.q[4 * i + 1] = qn.y; a.xc.q[4 * i + 2] = qn.z; a.xc.q[4 * i + 3] = qn.w;
    if (a.count_camera && a.active[6 * i]) {
      xn += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
      sn += (q.x - qn.x) * (q.x - qn.x) + (q.y - qn.y) * (q.y - qn.y) + (q.z - qn.z) * (q.z - qn.z) + (q.w - qn.w) * (q.w - qn.w);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double p = a.x.p[kPStride * i + c];
      const double pn = p + a.alpha * d[3 + c];
      a.xc.p[kPStride * i + c] = pn;
      if (a.count_camera && a.active[6 * i + 3 + c]) { xn += p * p; sn += (p - pn) * (p - pn); }
    }
    a.xc.p[kPStride * i + 3] = 0.0;
  } else if (i < nK + 6 * nB) {
    const int b = i - nK;
    const double v = a.x.bias[b];
    const double vn = v + a.alpha * a.dc[a.dims.idx_bias0 + b];
    a.xc.bias[b] = vn;
    if (a.count_camera && a.active[a.dims.idx_bias0 + b]) { xn += v * v; sn += (v - vn) * (v - vn); }
  } else if (i == nK + 6 * nB) {
    const double v = *a.x.ld;
    double vn = v + a.alpha * a.dc[a.dims.idx_ld];
    if (a.clamp_ld) vn = fmin(fmax(vn, a.ld_lower), a.ld_upper);
    *a.xc.ld = vn;
    a.scal->ld_value = vn;
    if (a.count_camera && a.active[a.dims.idx_ld]) { xn += v * v; sn += (v - vn) * (v - vn); }
  } else if (i < nK + 6 * nB + 1 + nL) {
    const int l = i - (nK + 6 * nB + 1);
    const double v = a.x.rho[l];
    const double vn = v + a.alpha * a.dl[l];
    a.xc.rho[l] = vn;
    if (a.active[a.dims.np + l]) { xn += v * v; sn += (v - vn) * (v - vn); }
  }
  xn = warp_sum_d(xn); sn = warp_sum_d(sn);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = xn; red[1][threadIdx.x >> 5] = sn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double x = 0, s = 0;
    for (int w = 0; w < 8; ++w) { x += red[0][w]; s += red[1][w]; }
    if (x != 0.0) atomicAdd(&a.scal->x_norm2, x);
    if (s != 0.0) atomicAdd(&a.scal->step_norm2, s);
  }
}

int launch_apply_step(const ApplyLaunch& a, cudaStream_t s) {
  const int n = a.dims.nK + 6 * a.dims.nB + 1 + a.dims.nL;
  cudaMemsetAsync(&a.scal->step_norm2, 0, 2 * sizeof(double), s);  // step_norm2, x_norm2 are adjacent
  apply_step_kernel<<<(n + 255) / 256, 256, 0, s>>>(a);
  int launches = 1;
  launches += launch_knot_table(a.xc, a.dims.nK, s);
  return launches;
}

// ------------------------------------------------------------------------------------------------
// TrajectoryManager::double2vector (trajectory_manager.cpp:485-516) + R2ypr / ypr2R (eigen_utils.hpp:114-150)
__device__ void r2ypr_deg(const M3& R, double ypr[3]) {
  const double nx = R.m[0], ny = R.m[3], nz = R.m[6];
  const double ox = R.m[1], oy = R.m[4];
  const double ax = R.m[2], ay = R.m[5];
  const double y = atan2(ny, nx);
  const double p = atan2(-nz, nx * cos(y) + ny * sin(y));
  const double r = atan2(ax * sin(y) - ay * cos(y), -ox * sin(y) + oy * cos(y));
  const double k = 180.0 / 3.14159265358979323846;
  ypr[0] = y * k; ypr[1] = p * k; ypr[2] = r * k;
}

__global__ void gauge_realign_kernel(StatePtrs st, int nK, int min_idx, const double* R0t0) {
  __shared__ double T[16];  // qd (4), tran_diff (3)
  if (threadIdx.x == 0) {
    M3 R0;
    for (int e = 0; e < 9; ++e) R0.m[e] = R0t0[e];
    const V3 t0 = V3{R0t0[9], R0t0[10], R0t0[11]};
    const M3 R00 = so3_matrix(load_q(st.q, min_idx));
    const V3 t00 = load_p<kPStride>(st.p, min_idx);
    double e0[3], e00[3];
    r2ypr_deg(R0, e0);
    r2ypr_deg(R00, e00);
    const double y = (e0[0] - e00[0]) / 180.0 * 3.14159265358979323846;
    M3 rd = m3_identity();
    rd.m[0] = cos(y); rd.m[1] = -sin(y); rd.m[3] = sin(y); rd.m[4] = cos(y);
    if (fabs(fabs(e0[1]) - 90.0) < 1.0 || fabs(fabs(e00[1]) - 90.0) < 1.0) rd = m3_mul_bt(R0, R00);
    const V3 td = t0 - m3_vec(rd, t00);
    const Q4 qd = quat_from_matrix(rd);
    T[0] = qd.x; T[1] = qd.y; T[2] = qd.z; T[3] = qd.w; T[4] = td.x; T[5] = td.y; T[6] = td.z;
  }
  __syncthreads();
  const Q4 qd = Q4{T[0], T[1], T[2], T[3]};
  const V3 td = V3{T[4], T[5], T[6]};
  for (int i = min_idx + threadIdx.x; i < nK; i += blockDim.x) {
    const Q4 qn = so3_mul(qd, load_q(st.q, i));
    const V3 pn = so3_rotate(qd, load_p<kPStride>(st.p, i)) + td;
    st.q[4 * i] = qn.x; st.q[4 * i + 1] = qn.y; st.q[4 * i + 2] = qn.z; st.q[4 * i + 3] = qn.w;
    st.p[kPStride * i] = pn.x; st.p[kPStride * i + 1] = pn.y; st.p[kPStride * i + 2] = pn.z;
  }
}

int launch_gauge_realign(const StatePtrs& st, int nK, int min_idx, const double* R0_t0_dev, cudaStream_t s) {
  gauge_realign_kernel<<<1, 256, 0, s>>>(st, nK, min_idx, R0_t0_dev);
  return 1 + launch_knot_table(st, nK, s);
}

}  // namespace ctvio
