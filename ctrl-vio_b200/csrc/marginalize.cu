// K7: marginalization on the GPU.  Replaces MarginalizationInfo::preMarginalize / marginalize
// (factor/analytic_diff/marginalization_factor.cpp:106-265): evaluate every recorded factor at the current
// state (image factors with the loss corrector), build the dense A = sum J'J, b = sum J'r over
// [dropped | kept] parameter positions, Schur-complement the dropped block through its eigen-decomposition
// (pseudo-inverse with eps = 1e-30, marginalization_factor.h:129) and factor the result into
// (J_lin, r_lin) by a second eigen-decomposition.
// The reference sums A on 4 pthreads (ThreadsConstructA, :141-176); here one thread per factor reduces with
// fp64 atomics, the two eigen-decompositions run as a parallel (round-robin ordered) cyclic Jacobi solver in
// one CTA, and the dense products are plain tiled kernels — this runs once per window, not per LM step.
#include "marginalize.h"
#include <cstdlib>
#include <cstring>

namespace ctvio {

// ------------------------------------------------------------------------------------------------
// recorded factors -> rows of the compressed Jacobian  Jrow[R][ldj]  (columns = positions in the [dropped | kept]
// ordering, last used column P = the residual), then  [A | b] = Jrow' Jrow  by ONE kernel whose threads each own an
// output entry and run over the rows in a FIXED order: the accumulation is bit-reproducible run to run (the r1 version
// used ~2 500 unordered fp64 atomics per factor).  Every thread below writes only its own rows.

__global__ void marg_image_kernel(MargImageArgs a) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= a.n_marg) return;
  const int n = a.marg_index[m];
  const longlong2 tt = a.obs.t[n];
  const double2 pi = a.obs.pi[n], pj = a.obs.pj[n];
  const int4 meta = a.obs.meta[n];
  const double rho = a.st.rho[meta.z];
  const int64_t ld_ns = int64_t(*a.st.ld * 1e9);
  int32_t si, sj;
  double ui, uj;
  if (!spline_index(a.sp, tt.x + int64_t(meta.x) * ld_ns, si, ui) ||
      !spline_index(a.sp, tt.y + int64_t(meta.y) * ld_ns, sj, uj)) {
    atomicOr(&a.scal->error_flags, 1);
    return;
  }
  SideEval ea, eb;
  eval_side<true, kPStride>(a.sp, a.st.q, a.st.p, a.st.tab, si, ui, ea);
  eval_side<true, kPStride>(a.sp, a.st.q, a.st.p, a.st.tab, sj, uj, eb);
  ImageCommon cm;
  const double pixy[2] = {pi.x, pi.y}, pjxy[2] = {pj.x, pj.y};
  image_common(a.rig, pixy, pjxy, rho, ea.R, ea.p, eb.R, eb.p, a.cauchy, cm);
  double* row0 = a.Jrow + size_t(a.row0 + 2 * m) * a.ldj;
  double* row1 = row0 + a.ldj;
  // the two knot windows may overlap (frames closer than the spline support): contributions ADD in the thread's own rows
  auto put = [&](int pos, double v0, double v1) {
    if (pos < 0) return;
    row0[pos] += v0;
    row1[pos] += v1;
  };
  {
    double rot[4][6], posb[4][6];
    image_side_blocks(0, cm, ea, rot, posb);
    for (int k = 0; k < 4; ++k)
      for (int c = 0; c < 3; ++c) {
        put(a.pos_cam[6 * (si + k) + c], rot[k][c], rot[k][3 + c]);
        put(a.pos_cam[6 * (si + k) + 3 + c], posb[k][c], posb[k][3 + c]);
      }
    image_side_blocks(1, cm, eb, rot, posb);
    for (int k = 0; k < 4; ++k)
      for (int c = 0; c < 3; ++c) {
        put(a.pos_cam[6 * (sj + k) + c], rot[k][c], rot[k][3 + c]);
        put(a.pos_cam[6 * (sj + k) + 3 + c], posb[k][c], posb[k][3 + c]);
      }
  }
  double t2[2];
  image_jrho(a.rig, cm, ea.R, rho, t2);
  put(a.pos_lm[meta.z], t2[0], t2[1]);
  image_jld(a.rig, cm, meta.x, meta.y, ea.R, ea.omega, ea.vel, eb.R, eb.omega, eb.vel, t2);
  put(a.pos_cam[a.idx_ld], t2[0], t2[1]);
  row0[a.P] = cm.r[0];
  row1[a.P] = cm.r[1];
}

__global__ void marg_imu_kernel(MargImuArgs a) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= a.n_marg) return;
  const int n = a.marg_index[m];
  const longlong2 tn = a.obs.t_node[n];
  const double2 g0 = a.obs.ga[3 * n], g1 = a.obs.ga[3 * n + 1], g2 = a.obs.ga[3 * n + 2];
  const double gyro[3] = {g0.x, g0.y, g1.x}, accel[3] = {g1.y, g2.x, g2.y};
  const int node = int(tn.y);
  double bias[6];
  for (int c = 0; c < 6; ++c) bias[c] = a.st.bias[6 * node + c];
  int32_t s;
  double u;
  if (!spline_index(a.sp, tn.x, s, u)) {
    atomicOr(&a.scal->error_flags, 1);
    return;
  }
  ImuEvalOut o;
  eval_imu<true, kPStride>(a.sp, a.rig, a.st.q, a.st.p, a.st.tab, s, u, gyro, accel, bias, o);
  // 6 rows x 30 local columns: 4 knots x (rot 3 | pos 3), bg 3, ba 3
  for (int r = 0; r < 6; ++r) {
    double* row = a.Jrow + size_t(a.row0 + 6 * m + r) * a.ldj;
    for (int k = 0; k < 4; ++k)
      for (int c = 0; c < 3; ++c) {
        const int pr = a.pos_cam[6 * (s + k) + c], pp = a.pos_cam[6 * (s + k) + 3 + c];
        if (pr >= 0) row[pr] = o.Jrot[k][3 * r + c];
        if (pp >= 0) row[pp] = o.Jpos[k][3 * r + c];
      }
    const int pb = a.pos_cam[a.idx_bias0 + 6 * node + r];
    if (pb >= 0) row[pb] = a.rig.imu_info[r];
    row[a.P] = o.r[r];
  }
}

// bias factors flagged marg + the old prior (if recorded); one CTA
__global__ void marg_small_kernel(MargSmallArgs a) {
  const int tid = threadIdx.x;
  for (int n = tid; n < a.n_bias; n += blockDim.x) {
    const int2 ij = a.bf_ij[n];
    for (int k = 0; k < 6; ++k) {
      const double s = a.bf_s[6 * n + k];
      const double r = s * (a.st.bias[6 * ij.y + k] - a.st.bias[6 * ij.x + k]);
      const int pi = a.pos_cam[a.idx_bias0 + 6 * ij.x + k], pj = a.pos_cam[a.idx_bias0 + 6 * ij.y + k];
      double* row = a.Jrow + size_t(a.row0_bias + 6 * n + k) * a.ldj;
      if (pi >= 0) row[pi] = -s;
      if (pj >= 0) row[pj] = s;
      row[a.P] = r;
    }
  }
  const int n = a.prior.n;
  if (n <= 0 || !a.use_prior) return;
  // residual of the old prior at the current state (dx / res scratch as in small_factors_kernel)
  for (int b = tid; b < a.prior.n_blocks; b += blockDim.x) {
    const int type = a.prior.type[b];
    const int index = a.prior.index[b];
    const double* x = type == 0 ? a.st.q + 4 * index : type == 1 ? a.st.p + kPStride * index
                    : type == 2 ? a.st.bias + 6 * index : type == 3 ? a.st.bias + 6 * index + 3 : a.st.ld;
    const double* x0 = a.prior.x0 + 4 * b;
    double* dx = a.prior.dx + a.prior.col[b];
    if (type == 0) {
      const double n2 = x0[0] * x0[0] + x0[1] * x0[1] + x0[2] * x0[2] + x0[3] * x0[3];
      const double ax = -x0[0] / n2, ay = -x0[1] / n2, az = -x0[2] / n2, aw = x0[3] / n2;
      const double qx = aw * x[0] + ax * x[3] + ay * x[2] - az * x[1];
      const double qy = aw * x[1] + ay * x[3] + az * x[0] - ax * x[2];
      const double qz = aw * x[2] + az * x[3] + ax * x[1] - ay * x[0];
      const double qw = aw * x[3] - ax * x[0] - ay * x[1] - az * x[2];
      const double sg = (qw >= 0) ? 2.0 : -2.0;
      dx[0] = sg * qx; dx[1] = sg * qy; dx[2] = sg * qz;
    } else {
      const int sz = type == 4 ? 1 : 3;
      for (int d = 0; d < sz; ++d) dx[d] = x[d] - x0[d];
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    double s = a.prior.r[i];
    for (int j = 0; j < n; ++j) s = fma(a.prior.J[size_t(i) * n + j], a.prior.dx[j], s);
    a.Jrow[size_t(a.row0_prior + i) * a.ldj + a.P] = s;
  }
  // prior_pos[j]: position of prior column j in the new ordering (distinct columns -> distinct positions)
  for (int e = tid; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e % n;
    const int pj = a.prior_pos[j];
    if (pj >= 0) a.Jrow[size_t(a.row0_prior + i) * a.ldj + pj] = a.prior.J[e];
  }
}

// [A | b] = Jrow' Jrow over columns 0..P (column P = residual): C[i][j] = sum_r Jrow[r][i] Jrow[r][j], r ascending.
// 16x16 outputs per CTA, 32 rows staged per step; the sum of every entry runs in one thread in a fixed order.
__global__ void __launch_bounds__(256) marg_syrk_kernel(const double* __restrict__ Jrow, int R, int ldj, int P, double* A,
                                                        double* b) {
  __shared__ double Si[32][17], Sj[32][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
  if (j0 + 15 < i0) return;  // strictly lower blocks are mirrored from the upper ones
  const int i = i0 + ty, j = j0 + tx;
  double acc = 0.0;
  for (int r0 = 0; r0 < R; r0 += 32) {
    for (int e = threadIdx.x; e < 32 * 16; e += 256) {
      const int rr = e >> 4, c = e & 15;
      const bool ok = r0 + rr < R;
      Si[rr][c] = (ok && i0 + c <= P) ? Jrow[size_t(r0 + rr) * ldj + i0 + c] : 0.0;
      Sj[rr][c] = (ok && j0 + c <= P) ? Jrow[size_t(r0 + rr) * ldj + j0 + c] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < 32; ++rr) acc = fma(Si[rr][ty], Sj[rr][tx], acc);
    __syncthreads();
  }
  if (i > P || j > P || i == P) return;
  if (j == P) { b[i] = acc; return; }
  if (j >= i) {
    A[size_t(i) * P + j] = acc;
    A[size_t(j) * P + i] = acc;
  }
}

int launch_marg_image(const MargImageArgs& a, cudaStream_t s) {
  if (a.n_marg <= 0) return 0;
  marg_image_kernel<<<(a.n_marg + 63) / 64, 64, 0, s>>>(a);
  return 1;
}
int launch_marg_imu(const MargImuArgs& a, cudaStream_t s) {
  if (a.n_marg <= 0) return 0;
  marg_imu_kernel<<<(a.n_marg + 63) / 64, 64, 0, s>>>(a);
  return 1;
}
int launch_marg_small(const MargSmallArgs& a, cudaStream_t s) {
  if (a.n_bias <= 0 && !(a.use_prior && a.prior.n > 0)) return 0;
  marg_small_kernel<<<1, 256, 0, s>>>(a);
  return 1;
}
int launch_marg_syrk(const double* Jrow, int R, int ldj, int P, double* A, double* b, cudaStream_t s) {
  if (P <= 0) return 0;
  const int nb = (P + 1 + 15) / 16;
  marg_syrk_kernel<<<dim3(nb, nb), 256, 0, s>>>(Jrow, R, ldj, P, A, b);
  return 1;
}

// ------------------------------------------------------------------------------------------------
// symmetric eigen-decomposition: parallel cyclic Jacobi, one CTA of 1024 threads, matrix in L2
//   A (n x n, row-major, symmetric, overwritten), V (n x n) <- eigenvectors in columns, ev[n] <- diagonal
// Pairs of one round come from the round-robin tournament schedule: n/2 disjoint rotations at once.

__global__ void __launch_bounds__(1024) jacobi_eig_kernel(double* A, double* V, double* ev, int n, int max_sweeps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* cs = reinterpret_cast<double*>(smem_raw);           // [npairs][2]
  int* pp = reinterpret_cast<int*>(cs + 2 * ((n + 1) / 2));   // [npairs][2]
  __shared__ double s_off, s_diag, s_prev;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int ne = (n + 1) & ~1;  // even player count (a dummy player if n is odd)
  const int npairs = ne / 2;
  for (int e = tid; e < n * n; e += nt) V[e] = (e / n == e % n) ? 1.0 : 0.0;
  __syncthreads();
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    // convergence: off-diagonal mass
    if (tid == 0) { s_off = 0.0; s_diag = 0.0; }
    __syncthreads();
    double off = 0, dg = 0;
    for (int e = tid; e < n * n; e += nt) {
      const int i = e / n, j = e % n;
      const double v = A[e];
      if (i == j) dg += v * v; else off += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) {
      off += __shfl_xor_sync(0xffffffffu, off, o);
      dg += __shfl_xor_sync(0xffffffffu, dg, o);
    }
    if ((tid & 31) == 0) { atomicAdd(&s_off, off); atomicAdd(&s_diag, dg); }
    __syncthreads();
    // converged, or stagnating at the rounding floor (the off-diagonal mass no longer halves per sweep once it is
    // below ~(n eps)^2 of the diagonal mass): more sweeps only shuffle noise
    // converged, or stagnating at the rounding floor (the off-diagonal mass no longer halves per sweep once it is
    // below ~(n eps)^2 of the diagonal mass): more sweeps only shuffle noise.  (Stopping earlier, e.g. one sweep after
    // 1e-20, is NOT an option: the eps = 1e-30 pseudo-inverse of the reference inverts the smallest eigenvalues, so
    // their RELATIVE accuracy matters - measured: 5e-3 cost drift over a few windows.)
    const double prev = sweep > 0 ? s_prev : 1e300;
    if (s_off <= 1e-60 || s_off <= 1e-30 * s_diag || (s_off <= 1e-24 * s_diag && s_off > 0.5 * prev)) break;
    __syncthreads();
    if (tid == 0) s_prev = s_off;
    for (int round = 0; round < ne - 1; ++round) {
      // tournament pairing: player ne-1 fixed, the others rotate
      for (int k = tid; k < npairs; k += nt) {
        int p, q;
        if (k == 0) { p = ne - 1; q = round % (ne - 1); }
        else { p = (round + k) % (ne - 1); q = (round + ne - 1 - k) % (ne - 1); }
        if (p > q) { const int t = p; p = q; q = t; }
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = A[size_t(p) * n + q];
          if (apq != 0.0) {
            const double app = A[size_t(p) * n + p], aqq = A[size_t(q) * n + q];
            const double theta = (aqq - app) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            c = 1.0 / sqrt(t * t + 1.0);
            s = t * c;
          }
        }
        cs[2 * k] = c; cs[2 * k + 1] = s;
        pp[2 * k] = p; pp[2 * k + 1] = q;
      }
      __syncthreads();
      // columns p,q of A and V:  X[:,p] = c X[:,p] - s X[:,q] ;  X[:,q] = s X[:,p] + c X[:,q]
      for (int e = tid; e < npairs * n; e += nt) {
        const int k = e / n, r = e % n;
        const int p = pp[2 * k], q = pp[2 * k + 1];
        if (q >= n) continue;
        const double c = cs[2 * k], s = cs[2 * k + 1];
        if (s == 0.0) continue;
        const double ap = A[size_t(r) * n + p], aq = A[size_t(r) * n + q];
        A[size_t(r) * n + p] = c * ap - s * aq;
        A[size_t(r) * n + q] = s * ap + c * aq;
        const double vp = V[size_t(r) * n + p], vq = V[size_t(r) * n + q];
        V[size_t(r) * n + p] = c * vp - s * vq;
        V[size_t(r) * n + q] = s * vp + c * vq;
      }
      __syncthreads();
      // rows p,q of A
      for (int e = tid; e < npairs * n; e += nt) {
        const int k = e / n, col = e % n;
        const int p = pp[2 * k], q = pp[2 * k + 1];
        if (q >= n) continue;
        const double c = cs[2 * k], s = cs[2 * k + 1];
        if (s == 0.0) continue;
        const double ap = A[size_t(p) * n + col], aq = A[size_t(q) * n + col];
        A[size_t(p) * n + col] = c * ap - s * aq;
        A[size_t(q) * n + col] = s * ap + c * aq;
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < n; i += nt) ev[i] = A[size_t(i) * n + i];
}

// Same algorithm with A and V resident in shared memory (odd row stride: column accesses spread over the banks);
// used whenever both fit (n <= ~117), i.e. for every prior of a TUM-RSVI-scale window: the three phases of a round
// are then bound by shared-memory latency (~0.3 us per round) instead of L2 round trips (~4 us).
__global__ void __launch_bounds__(1024) jacobi_eig_smem_kernel(double* Ag, double* Vg, double* ev, int n, int max_sweeps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ld = n | 1;  // odd row stride (in doubles): a column walk visits 16 distinct bank pairs
  double* A = reinterpret_cast<double*>(smem_raw);
  double* V = A + size_t(n) * ld;
  double* cs = V + size_t(n) * ld;                             // [npairs][2]
  int* pp = reinterpret_cast<int*>(cs + 2 * ((n + 1) / 2));   // [npairs][2]
  __shared__ double s_off, s_diag, s_prev;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int ne = (n + 1) & ~1;
  const int npairs = ne / 2;
  for (int e = tid; e < n * n; e += nt) {
    const int i = e / n, j = e % n;
    A[i * ld + j] = Ag[e];
    V[i * ld + j] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    if (tid == 0) { s_off = 0.0; s_diag = 0.0; }
    __syncthreads();
    double off = 0, dg = 0;
    for (int e = tid; e < n * n; e += nt) {
      const int i = e / n, j = e % n;
      const double v = A[i * ld + j];
      if (i == j) dg += v * v; else off += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) {
      off += __shfl_xor_sync(0xffffffffu, off, o);
      dg += __shfl_xor_sync(0xffffffffu, dg, o);
    }
    if ((tid & 31) == 0) { atomicAdd(&s_off, off); atomicAdd(&s_diag, dg); }
    __syncthreads();
    // converged, or stagnating at the rounding floor (the off-diagonal mass no longer halves per sweep once it is
    // below ~(n eps)^2 of the diagonal mass): more sweeps only shuffle noise
    // converged, or stagnating at the rounding floor (the off-diagonal mass no longer halves per sweep once it is
    // below ~(n eps)^2 of the diagonal mass): more sweeps only shuffle noise.  (Stopping earlier, e.g. one sweep after
    // 1e-20, is NOT an option: the eps = 1e-30 pseudo-inverse of the reference inverts the smallest eigenvalues, so
    // their RELATIVE accuracy matters - measured: 5e-3 cost drift over a few windows.)
    const double prev = sweep > 0 ? s_prev : 1e300;
    if (s_off <= 1e-60 || s_off <= 1e-30 * s_diag || (s_off <= 1e-24 * s_diag && s_off > 0.5 * prev)) break;
    __syncthreads();
    if (tid == 0) s_prev = s_off;
    for (int round = 0; round < ne - 1; ++round) {
      for (int k = tid; k < npairs; k += nt) {
        int p, q;
        if (k == 0) { p = ne - 1; q = round % (ne - 1); }
        else { p = (round + k) % (ne - 1); q = (round + ne - 1 - k) % (ne - 1); }
        if (p > q) { const int t = p; p = q; q = t; }
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = A[p * ld + q];
          if (apq != 0.0) {
            const double app = A[p * ld + p], aqq = A[q * ld + q];
            const double theta = (aqq - app) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            c = 1.0 / sqrt(t * t + 1.0);
            s = t * c;
          }
        }
        cs[2 * k] = c; cs[2 * k + 1] = s;
        pp[2 * k] = p; pp[2 * k + 1] = q;
      }
      __syncthreads();
      // warp = pair (k, k + 32, ...), lanes stride over the rows / columns: no integer divisions in the hot loops
      const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
      for (int k = warp; k < npairs; k += nwarps) {
        const int p = pp[2 * k], q = pp[2 * k + 1];
        if (q >= n) continue;
        const double c = cs[2 * k], s = cs[2 * k + 1];
        if (s == 0.0) continue;
        for (int r = lane; r < n; r += 32) {
          const double ap = A[r * ld + p], aq = A[r * ld + q];
          A[r * ld + p] = c * ap - s * aq;
          A[r * ld + q] = s * ap + c * aq;
          const double vp = V[r * ld + p], vq = V[r * ld + q];
          V[r * ld + p] = c * vp - s * vq;
          V[r * ld + q] = s * vp + c * vq;
        }
      }
      __syncthreads();
      for (int k = warp; k < npairs; k += nwarps) {
        const int p = pp[2 * k], q = pp[2 * k + 1];
        if (q >= n) continue;
        const double c = cs[2 * k], s = cs[2 * k + 1];
        if (s == 0.0) continue;
        for (int col = lane; col < n; col += 32) {
          const double ap = A[p * ld + col], aq = A[q * ld + col];
          A[p * ld + col] = c * ap - s * aq;
          A[q * ld + col] = s * ap + c * aq;
        }
      }
      __syncthreads();
    }
  }
  for (int e = tid; e < n * n; e += nt) {
    const int i = e / n, j = e % n;
    Ag[e] = A[i * ld + j];
    Vg[e] = V[i * ld + j];
  }
  for (int i = tid; i < n; i += nt) ev[i] = A[i * ld + i];
}

// Two-sided parallel Jacobi, block form.  The n/2 disjoint rotations of a round act on A as  A <- J' A J : the 2x2
// block of rows {p_k, q_k} and columns {p_l, q_l} becomes  R_k' B R_l  and depends on NO other entry, so one thread
// owns one (k, l) block for the whole round, in place, rows and columns in ONE phase: 2 barriers per round
// (rotations | update) instead of the 3 + separate row pass of the kernels above, and 4 loads / 4 stores per 12 FMAs.
// V <- V J rides in the same phase.  The (k, l) -> thread assignment does not depend on the round, so the index
// arithmetic is hoisted out of the sweep loop.  Rotation sequence and every sum are fixed: bit-reproducible.
// A_SMEM / V_SMEM: the matrix / the eigenvectors live in shared memory (odd row stride) or stay in global memory (L2).
__device__ int g_jacobi_dbg[8];  // [0] sweeps, [1] n of the last eigen-decomposition (tools/marg_timing.py)
extern "C" int ctvio_debug_jacobi(int* out8) { return cudaMemcpyFromSymbol(out8, g_jacobi_dbg, sizeof(g_jacobi_dbg)) == cudaSuccess ? 0 : -1; }

// full-precision 1/x and 1/sqrt(x) from the hardware seeds + Newton steps: the rotation parameters sit on the serial
// part of every round (43 threads compute, 1000 wait), IEEE division / sqrt sequences are 3x longer
__device__ __forceinline__ double rcp_fast(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
__device__ __forceinline__ double rsqrt_fast(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-(y * y), x, 1.0);
  y = fma(fma(e, 0.375, 0.5), y * e, y);
  e = fma(-(y * y), x, 1.0);
  return fma(fma(e, 0.375, 0.5), y * e, y);
}

// Rotation log entry of one pair of one round: V <- V J is NOT done inside the eigenvalue kernel.  Every row of V
// transforms independently of the others (row r: (v_p, v_q) <- (c v_p - s v_q, s v_p + c v_q) for the pairs of the round),
// so the eigenvalue kernel (one CTA, A only: half the shared-memory traffic and instructions per round) just logs
// (p, q, c, s), and a second kernel with one CTA PER ROW of V replays the log on a row of the identity - n SMs instead of
// one for the eigenvector half of the work, ~50 us for the n = 85 prior of a streaming window.
struct JacobiRot {
  int p, q;
  double c, s;
};

template <bool A_SMEM>
__global__ void __launch_bounds__(1024) jacobi_eig_block_kernel(double* Ag, double* ev, JacobiRot* log, int* log_rounds, int n,
                                                                int max_sweeps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int ne = (n + 1) & ~1, npairs = ne / 2;
  const int lda = A_SMEM ? (n | 1) : n;
  double* sm = reinterpret_cast<double*>(smem_raw);
  double* A = A_SMEM ? sm : Ag;
  double* cs = sm + (A_SMEM ? size_t(n) * lda : 0);   // [npairs][2]
  int* pq = reinterpret_cast<int*>(cs + 2 * npairs);  // [npairs][2]
  __shared__ double s_off, s_diag, s_prev;
  int sweeps_done = 0, rounds_done = 0;
  if (A_SMEM)
    for (int e = tid; e < n * n; e += nt) {
      const int i = e / n, j = e - i * n;
      A[i * lda + j] = Ag[e];
    }
  // fixed work assignment: A blocks (k, l), k <= l (the mirror block is written by the same thread)
  constexpr int kMaxBlk = 12;
  const int nblk_total = npairs * (npairs + 1) / 2;
  short bk[kMaxBlk], bl[kMaxBlk];
  int nblk = 0;
  for (int b = tid; b < nblk_total && nblk < kMaxBlk; b += nt) {
    int k = 0, rem = b;
    while (rem >= npairs - k) { rem -= npairs - k; ++k; }
    bk[nblk] = short(k); bl[nblk] = short(k + rem);
    ++nblk;
  }
  __syncthreads();
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    if (tid == 0) { s_off = 0.0; s_diag = 0.0; }
    __syncthreads();
    double off = 0, dg = 0;
    for (int e = tid; e < n * n; e += nt) {
      const int i = e / n, j = e - i * n;
      const double v = A[i * lda + j];
      if (i == j) dg += v * v; else off += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) {
      off += __shfl_xor_sync(0xffffffffu, off, o);
      dg += __shfl_xor_sync(0xffffffffu, dg, o);
    }
    if ((tid & 31) == 0) { atomicAdd(&s_off, off); atomicAdd(&s_diag, dg); }
    __syncthreads();
    // converged, or stagnating at the rounding floor (the off-diagonal mass no longer halves per sweep once it is below
    // ~(n eps)^2 of the diagonal mass).  Stopping earlier is NOT an option: the reference's eps = 1e-30 pseudo-inverse
    // inverts the smallest eigenvalues, so their relative accuracy matters.  (shared-memory atomics of 32 warp sums:
    // the order can differ between runs, but s_off only steers the loop count through comparisons far from any tie)
    const double prev = sweep > 0 ? s_prev : 1e300;
    if (s_off <= 1e-60 || s_off <= 1e-30 * s_diag || (s_off <= 1e-24 * s_diag && s_off > 0.5 * prev)) break;
    __syncthreads();
    if (tid == 0) s_prev = s_off;
    ++sweeps_done;
    for (int round = 0; round < ne - 1; ++round) {
      // ---- phase R: the round's pairs (round-robin tournament), their rotations, the log entry ----
      for (int k = tid; k < npairs; k += nt) {
        int p, q;
        if (k == 0) { p = ne - 1; q = round % (ne - 1); }
        else { p = (round + k) % (ne - 1); q = (round + ne - 1 - k) % (ne - 1); }
        if (p > q) { const int t = p; p = q; q = t; }
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = A[p * lda + q];
          if (apq != 0.0) {
            const double app = A[p * lda + p], aqq = A[q * lda + q];
            const double theta = 0.5 * (aqq - app) * rcp_fast(apq);
            const double h = fma(theta, theta, 1.0);
            if (isfinite(h)) {
              const double t = copysign(1.0, theta) * rcp_fast(fabs(theta) + h * rsqrt_fast(h));
              c = rsqrt_fast(fma(t, t, 1.0));
              s = t * c;
            }  // |theta| ~ 1e154+: the rotation is the identity to working precision
          }
        } else {
          q = -1;  // dummy player of an odd n: the pair is idle
        }
        cs[2 * k] = c; cs[2 * k + 1] = s;
        pq[2 * k] = p; pq[2 * k + 1] = q;
        log[size_t(rounds_done) * npairs + k] = JacobiRot{p, q, c, s};
      }
      ++rounds_done;
      __syncthreads();
      // ---- phase A: every 2x2 block B(k,l) <- R_k' B R_l  (and its mirror) ----
#pragma unroll 1
      for (int w = 0; w < nblk; ++w) {
        const int k = bk[w], l = bl[w];
        const int pk = pq[2 * k], qk = pq[2 * k + 1], pl = pq[2 * l], ql = pq[2 * l + 1];
        const double ck = cs[2 * k], sk = cs[2 * k + 1], cl = cs[2 * l], sl = cs[2 * l + 1];
        if (qk < 0 && ql < 0) continue;
        const bool hk = qk >= 0, hl = ql >= 0;  // an idle (dummy) partner: only the real row / column exists
        const double b00 = A[pk * lda + pl];
        const double b01 = hl ? A[pk * lda + ql] : 0.0;
        const double b10 = hk ? A[qk * lda + pl] : 0.0;
        const double b11 = (hk && hl) ? A[qk * lda + ql] : 0.0;
        // columns: [x_p x_q] <- [c x_p - s x_q , s x_p + c x_q] with (cl, sl); rows likewise with (ck, sk)
        const double t00 = cl * b00 - sl * b01, t01 = sl * b00 + cl * b01;
        const double t10 = cl * b10 - sl * b11, t11 = sl * b10 + cl * b11;
        const double r00 = ck * t00 - sk * t10, r10 = sk * t00 + ck * t10;
        const double r01 = ck * t01 - sk * t11, r11 = sk * t01 + ck * t11;
        A[pk * lda + pl] = r00;
        if (hl) A[pk * lda + ql] = r01;
        if (hk) A[qk * lda + pl] = r10;
        if (hk && hl) A[qk * lda + ql] = r11;
        if (k != l) {  // mirror block (l, k) = transpose
          A[pl * lda + pk] = r00;
          if (hl) A[ql * lda + pk] = r01;
          if (hk) A[pl * lda + qk] = r10;
          if (hk && hl) A[ql * lda + qk] = r11;
        }
      }
      __syncthreads();
    }
  }
  if (A_SMEM)
    for (int e = tid; e < n * n; e += nt) {
      const int i = e / n, j = e - i * n;
      Ag[e] = A[i * lda + j];
    }
  for (int i = tid; i < n; i += nt) ev[i] = A[i * lda + i];
  if (tid == 0) {
    *log_rounds = rounds_done;
    g_jacobi_dbg[0] = sweeps_done; g_jacobi_dbg[1] = n; g_jacobi_dbg[2] = int(-log10(fmax(s_off / fmax(s_diag, 1e-300), 1e-300)));
  }
}

// V = product of the logged rotations, one CTA per row of V (= row of the identity pushed through the log): thread k
// applies pair k of every round to the row held in shared memory; the pairs of a round are disjoint.  The log is staged
// through shared memory in chunks of kLogChunk rounds (coalesced bulk loads: one L2 round trip per chunk, not per round).
constexpr int kLogChunk = 16;
__global__ void __launch_bounds__(128) jacobi_apply_log_kernel(const JacobiRot* __restrict__ log, const int* log_rounds, int n,
                                                               double* Vg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int npairs = ((n + 1) & ~1) / 2;
  double* row = reinterpret_cast<double*>(smem_raw);                                    // [n]
  JacobiRot* stage = reinterpret_cast<JacobiRot*>(row + ((n + 1) & ~1));                // [kLogChunk][npairs]
  const int r = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  for (int j = tid; j < n; j += nt) row[j] = j == r ? 1.0 : 0.0;
  const int rounds = *log_rounds;
  for (int t0 = 0; t0 < rounds; t0 += kLogChunk) {
    const int nr = min(kLogChunk, rounds - t0);
    __syncthreads();
    {  // JacobiRot = 24 bytes = 3 doubles: straight copy as doubles
      const double* src = reinterpret_cast<const double*>(log + size_t(t0) * npairs);
      double* dst = reinterpret_cast<double*>(stage);
      for (int e = tid; e < nr * npairs * 3; e += nt) dst[e] = src[e];
    }
    __syncthreads();
    for (int t = 0; t < nr; ++t) {
      for (int k = tid; k < npairs; k += nt) {
        const JacobiRot g = stage[t * npairs + k];
        if (g.q >= 0 && g.s != 0.0) {
          const double vp = row[g.p], vq = row[g.q];
          row[g.p] = g.c * vp - g.s * vq;
          row[g.q] = g.s * vp + g.c * vq;
        }
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < n; j += nt) Vg[size_t(r) * n + j] = row[j];
}

size_t jacobi_log_bytes(int n, int max_sweeps) {
  const int ne = (n + 1) & ~1;
  const size_t elementwise = size_t(max_sweeps) * (ne - 1) * (ne / 2) * sizeof(JacobiRot) + 64;
  const size_t blocked = jacobi_blocked_log_bytes(n, max_sweeps);
  return elementwise > blocked ? elementwise : blocked;
}

int launch_jacobi_eig(double* A, double* V, double* ev, int n, void* log_buf, cudaStream_t s) {
  if (n <= 0) return 0;
  constexpr int kMaxSweeps = 40;
  // streaming-window sizes: blocked solver (jacobi_blocked.cu); CTVIO_JACOBI=elementwise keeps the solver below
  if (log_buf && jacobi_blocked_fits(n)) {
    const char* v = std::getenv("CTVIO_JACOBI");
    if (!(v && std::strcmp(v, "elementwise") == 0)) return launch_jacobi_blocked(A, V, ev, n, log_buf, kMaxSweeps, s);
  }
  const int ne = (n + 1) & ~1, npairs = ne / 2;
  const size_t pairs = size_t(npairs) * (2 * sizeof(double) + 2 * sizeof(int));
  const size_t mat = size_t(n) * (n | 1) * sizeof(double);
  const size_t limit = 224 * 1024;
  // per-thread block slots: npairs (npairs + 1) / 2 blocks over 1024 threads, at most 12 each (n <= ~310); 2 pairs per
  // thread in the replay kernel (n <= 512)
  if (size_t(npairs) * (npairs + 1) / 2 > size_t(12) * 1024 || !log_buf) {
    jacobi_eig_kernel<<<1, 1024, pairs, s>>>(A, V, ev, n, 60);
    return 1;
  }
  static PerDeviceOnce once;
  if (once.first())
    cudaFuncSetAttribute(jacobi_eig_block_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(limit));
  int* rounds = reinterpret_cast<int*>(log_buf);
  JacobiRot* log = reinterpret_cast<JacobiRot*>(reinterpret_cast<unsigned char*>(log_buf) + 64);
  if (mat + pairs <= limit) jacobi_eig_block_kernel<true><<<1, 1024, mat + pairs, s>>>(A, ev, log, rounds, n, kMaxSweeps);
  else jacobi_eig_block_kernel<false><<<1, 1024, pairs, s>>>(A, ev, log, rounds, n, kMaxSweeps);
  jacobi_apply_log_kernel<<<n, 128, size_t((n + 1) & ~1) * sizeof(double) + size_t(kLogChunk) * npairs * sizeof(JacobiRot), s>>>(
      log, rounds, n, V);
  return 2;
}

// ------------------------------------------------------------------------------------------------
// small dense helpers (naive tiles; sizes are a few hundred)

// C (m x n) = alpha * op(A) * op(B) + beta * C ; row-major with leading dimensions
__global__ void dense_gemm_kernel(int m, int n, int k, double alpha, const double* A, int lda, int ta, const double* B,
                                  int ldb, int tb, double beta, double* C, int ldc) {
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
  if (i >= m || j >= n) return;
  double s = 0;
  for (int x = 0; x < k; ++x) {
    const double av = ta ? A[size_t(x) * lda + i] : A[size_t(i) * lda + x];
    const double bv = tb ? B[size_t(j) * ldb + x] : B[size_t(x) * ldb + j];
    s = fma(av, bv, s);
  }
  C[size_t(i) * ldc + j] = alpha * s + (beta != 0.0 ? beta * C[size_t(i) * ldc + j] : 0.0);
}
int launch_dense_gemm(int m, int n, int k, double alpha, const double* A, int lda, bool ta, const double* B, int ldb,
                      bool tb, double beta, double* C, int ldc, cudaStream_t s) {
  if (m <= 0 || n <= 0) return 0;
  dim3 grid((n + 15) / 16, (m + 15) / 16), block(16, 16);
  dense_gemm_kernel<<<grid, block, 0, s>>>(m, n, k, alpha, A, lda, ta ? 1 : 0, B, ldb, tb ? 1 : 0, beta, C, ldc);
  return 1;
}

// mode 0: Amm = 0.5 (A[0:m,0:m] + A[0:m,0:m]')           (marginalization_factor.cpp:240)
// mode 1: scale columns of V (m x m) by 1/ev where ev > eps, else 0  -> Vs   (:243-244)
// mode 2: symmetrise from the lower triangle (Eigen's solver reads the lower triangle, :254)
// mode 3: J_lin[k][i] = sqrt(S_k) V[i][k] ; r_lin[k] = sqrt(1/S_k) (V' b)[k] with S_k = ev_k > eps ? ev_k : 0  (:255-263)
__global__ void marg_elementwise_kernel(int mode, int n, int ld, const double* src, double* dst, const double* ev,
                                        const double* vb, double* rlin, double eps) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  const int i = idx / n, j = idx % n;
  if (mode == 0) {
    dst[idx] = 0.5 * (src[size_t(i) * ld + j] + src[size_t(j) * ld + i]);
  } else if (mode == 1) {
    dst[idx] = ev[j] > eps ? src[idx] / ev[j] : 0.0;
  } else if (mode == 2) {
    dst[idx] = i >= j ? src[idx] : src[size_t(j) * n + i];
  } else {
    const int k = i;  // row of J_lin = eigen index
    const double S = ev[k] > eps ? ev[k] : 0.0;
    dst[idx] = sqrt(S) * src[size_t(j) * n + k];
    if (j == 0) rlin[k] = S > 0.0 ? sqrt(1.0 / S) * vb[k] : 0.0;
  }
}
int launch_marg_elementwise(int mode, int n, int ld, const double* src, double* dst, const double* ev, const double* vb,
                            double* rlin, double eps, cudaStream_t s) {
  if (n <= 0) return 0;
  marg_elementwise_kernel<<<(n * n + 255) / 256, 256, 0, s>>>(mode, n, ld, src, dst, ev, vb, rlin, eps);
  return 1;
}

}  // namespace ctvio
