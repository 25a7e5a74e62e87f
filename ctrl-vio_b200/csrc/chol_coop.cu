// K5: blocked right-looking Cholesky of the reduced camera system + both triangular solves in ONE
// cooperative persistent kernel (grid-wide barriers instead of two launches per block column).
// Replaces the factor/solve half of Ceres' SPARSE_NORMAL_CHOLESKY step (trajectory_estimator.cpp:374;
// Ceres is not under /root/reference) on the Schur-reduced system.
//
//   per block column k (NB = 64):
//     phase P  every CTA that owns a panel slab loads the (already updated) diagonal block, factors it
//              and inverts the factor redundantly in shared memory, turns its slabs' TRSM into a GEMM
//              with that inverse, and folds the forward substitution of the right-hand side in;
//     phase U  the trailing tiles are spread over all CTAs (64^3 register-tiled GEMM each).
//   afterwards CTA 0 runs the backward substitution with the stored block inverses.
// The 64x64 diagonal factorisation is the serial critical path (measured: profiles/r1/c_chol_phase_timing.txt),
// so it is latency-optimised: right-looking on 4x4 register blocks, the next diagonal block is factored
// by its owner thread while everybody else is still applying the rank-4 update (look-ahead), panels use
// substitution with the 4 reciprocal pivots (no inverse on the path), 4x4 inverses + recursive merges
// afterwards, both the inverse and its transpose are produced so that no shared-memory transpose is needed.
// M: lower triangle used; strictly-lower panels are overwritten with L, diagonal blocks are left
// untouched (only their inverses, Linv, are kept).
#include <cooperative_groups.h>

#include <algorithm>

#include "kernels.h"

namespace ctvio {

namespace cg = cooperative_groups;

// optional phase timing of CTA 0 (debug): build with CTVIO_EXTRA_NVCC_FLAGS=-DCTVIO_CHOL_TIMING
#ifdef CTVIO_CHOL_TIMING
__device__ unsigned long long g_chol_stamps[4096];
__device__ __forceinline__ void stamp(int& n) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (n < 4096) g_chol_stamps[n] = t;
  }
  ++n;
}
#define STAMP() stamp(stamp_n)
#else
#define STAMP()
#endif

constexpr int kTS = kCholNB + 2;  // shared tile row stride (doubles): rows stay 16-B aligned
constexpr int kTile = kCholNB * kTS;
constexpr size_t kCholCoopSmem = (5 * size_t(kTile) + 4 * kCholNB) * sizeof(double);

// acc[4][4] += A * B^T for 64x64 operands, BOTH stored transposed in smem: At[c][i] = A[i][c], Bt[c][j] = B[j][c].
// 256 threads, thread (ty, tx) owns rows 4ty.., cols 4tx..; four 16-byte shared loads feed 16 FMAs per k.
__device__ __forceinline__ void tile_gemm_tt(const double* At, const double* Bt, double acc[4][4], int ty, int tx) {
#pragma unroll 8
  for (int c = 0; c < kCholNB; ++c) {
    const double2 a01 = *reinterpret_cast<const double2*>(At + c * kTS + 4 * ty);
    const double2 a23 = *reinterpret_cast<const double2*>(At + c * kTS + 4 * ty + 2);
    const double2 b01 = *reinterpret_cast<const double2*>(Bt + c * kTS + 4 * tx);
    const double2 b23 = *reinterpret_cast<const double2*>(Bt + c * kTS + 4 * tx + 2);
    const double av[4] = {a01.x, a01.y, a23.x, a23.y};
    const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
  }
}

// smem tile <- TRANSPOSE of the 64x64 global block at M[r0.., c0..]: dst[c][r] = M[r0 + r][c0 + c]
// (coalesced 16-byte global reads along c)
__device__ __forceinline__ void load_tile_transposed(double* dst, const double* M, int npad, int r0, int c0, int tid) {
  for (int e = tid; e < kCholNB * kCholNB / 2; e += 256) {
    const int r = e >> 5, c = (e & 31) * 2;
    const double2 v = *reinterpret_cast<const double2*>(M + size_t(r0 + r) * npad + c0 + c);
    dst[c * kTS + r] = v.x;
    dst[(c + 1) * kTS + r] = v.y;
  }
}

// 4x4 lower Cholesky of a (registers), reciprocal pivots rd.  Returns false on a bad pivot.
__device__ __forceinline__ bool chol4(const double a[4][4], double l[4][4], double rd[4]) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double v = a[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) v = fma(-l[j][k], l[j][k], v);
    if (!(v > 0.0) || !isfinite(v)) { ok = false; v = 1.0; }
    rd[j] = rsqrt(v);
    l[j][j] = v * rd[j];
#pragma unroll
    for (int i = j + 1; i < 4; ++i) {
      double w = a[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) w = fma(-l[i][k], l[j][k], w);
      l[i][j] = w * rd[j];
    }
  }
  return ok;
}

// merge step of the triangular inverse for block size h (one thread per 4x4 output tile; a 4-lane k-split
// with shuffles and a fully unrolled templated variant were both measured slower):
//   mode 0: T     <- L21 * X11        mode 1: X21 <- -(X22 * T), written to Xi and (transposed) to XiT
__device__ __forceinline__ void merge_gemm(int mode, int h, const double* D, double* Xi, double* XiT, double* T, int tid) {
  const int npair = kCholNB / (2 * h), tpb = (h / 4) * (h / 4);
  if (tid < npair * tpb) {
    const int pb = tid / tpb, t = tid % tpb;
    const int r0 = 4 * (t / (h / 4)), c0 = 4 * (t % (h / 4));
    const int o = 2 * h * pb;
    const double* A = mode == 0 ? D + (o + h) * kTS + o : Xi + (o + h) * kTS + o + h;
    const double* B = mode == 0 ? Xi + o * kTS + o : T + (o + h) * kTS + o;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 4
    for (int m = 0; m < h; ++m) {
      double av[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = A[(r0 + i) * kTS + m];
      const double2 b01 = *reinterpret_cast<const double2*>(B + m * kTS + c0);
      const double2 b23 = *reinterpret_cast<const double2*>(B + m * kTS + c0 + 2);
      const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    if (mode == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        *reinterpret_cast<double2*>(T + (o + h + r0 + i) * kTS + o + c0) = make_double2(acc[i][0], acc[i][1]);
        *reinterpret_cast<double2*>(T + (o + h + r0 + i) * kTS + o + c0 + 2) = make_double2(acc[i][2], acc[i][3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        *reinterpret_cast<double2*>(Xi + (o + h + r0 + i) * kTS + o + c0) = make_double2(-acc[i][0], -acc[i][1]);
        *reinterpret_cast<double2*>(Xi + (o + h + r0 + i) * kTS + o + c0 + 2) = make_double2(-acc[i][2], -acc[i][3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) XiT[(o + c0 + j) * kTS + o + h + r0 + i] = -acc[i][j];
      }
    }
  }
}

// In-place lower Cholesky of the 64x64 block D (row stride kTS) by 256 threads, then Xi = D^-1 (lower
// triangular) and XiT = Xi^T.  T is a scratch tile, rdiag[64] receives 1/L_jj.
#ifdef CTVIO_CHOL_TIMING
#define FSTAMP() stamp(*pn)
#else
#define FSTAMP()
#endif
__device__ bool factor_and_invert_64(double* D, double* Xi, double* XiT, double* T, double* rdiag, int* s_bad, int* pn) {
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  if (tid == 0) *s_bad = 0;
  double a[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double2 v01 = *reinterpret_cast<const double2*>(D + (4 * ty + i) * kTS + 4 * tx);
    const double2 v23 = *reinterpret_cast<const double2*>(D + (4 * ty + i) * kTS + 4 * tx + 2);
    a[i][0] = v01.x; a[i][1] = v01.y; a[i][2] = v23.x; a[i][3] = v23.y;
  }
  for (int e = tid; e < kTile; e += 256) { Xi[e] = 0.0; XiT[e] = 0.0; }
  __syncthreads();
  if (tid == 0) {  // diagonal block 0
    double l[4][4], rd[4];
    if (!chol4(a, l, rd)) *s_bad = 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      rdiag[i] = rd[i];
#pragma unroll
      for (int j = 0; j < 4; ++j) D[i * kTS + j] = j <= i ? l[i][j] : 0.0;
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int jb = 0; jb < 16; ++jb) {
    if (tx == jb && ty > jb) {
      // panel block by substitution: x[r][c] = (a[r][c] - sum_{m<c} x[r][m] l[c][m]) / l[c][c]
      double l[4][4], rd[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        rd[c] = rdiag[4 * jb + c];
#pragma unroll
        for (int m = 0; m < 4; ++m) l[c][m] = D[(4 * jb + c) * kTS + 4 * jb + m];
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        double x[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double t = a[r][c];
#pragma unroll
          for (int m = 0; m < c; ++m) t = fma(-x[m], l[c][m], t);
          x[c] = t * rd[c];
        }
        *reinterpret_cast<double2*>(D + (4 * ty + r) * kTS + 4 * jb) = make_double2(x[0], x[1]);
        *reinterpret_cast<double2*>(D + (4 * ty + r) * kTS + 4 * jb + 2) = make_double2(x[2], x[3]);
      }
    }
    __syncthreads();
    if (tx > jb && ty >= tx) {
      double lr[4][4], lc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double2 r01 = *reinterpret_cast<const double2*>(D + (4 * ty + i) * kTS + 4 * jb);
        const double2 r23 = *reinterpret_cast<const double2*>(D + (4 * ty + i) * kTS + 4 * jb + 2);
        const double2 c01 = *reinterpret_cast<const double2*>(D + (4 * tx + i) * kTS + 4 * jb);
        const double2 c23 = *reinterpret_cast<const double2*>(D + (4 * tx + i) * kTS + 4 * jb + 2);
        lr[i][0] = r01.x; lr[i][1] = r01.y; lr[i][2] = r23.x; lr[i][3] = r23.y;
        lc[i][0] = c01.x; lc[i][1] = c01.y; lc[i][2] = c23.x; lc[i][3] = c23.y;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int m = 0; m < 4; ++m) a[i][j] = fma(-lr[i][m], lc[j][m], a[i][j]);
      if (ty == jb + 1 && tx == jb + 1) {  // look-ahead: factor the next diagonal block right away
        double l[4][4], rd[4];
        if (!chol4(a, l, rd)) *s_bad = 1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          rdiag[4 * tx + i] = rd[i];
#pragma unroll
          for (int j = 0; j < 4; ++j) D[(4 * tx + i) * kTS + 4 * tx + j] = j <= i ? l[i][j] : 0.0;
        }
      }
    }
    __syncthreads();
  }
  FSTAMP();  // main loop done
  // 4x4 inverses of the 16 diagonal blocks (off the critical path, all in parallel)
  if (ty == tx) {
    double l[4][4], li[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) l[i][j] = D[(4 * ty + i) * kTS + 4 * ty + j];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (r < c) { li[r][c] = 0.0; continue; }
        double t = (r == c) ? 1.0 : 0.0;
#pragma unroll
        for (int m = c; m < r; ++m) t = fma(-l[r][m], li[m][c], t);
        li[r][c] = t * rdiag[4 * ty + r];
      }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Xi[(4 * ty + i) * kTS + 4 * ty + j] = li[i][j];
        XiT[(4 * ty + j) * kTS + 4 * ty + i] = li[i][j];
      }
  }
  __syncthreads();
  FSTAMP();  // 4x4 inverses done
#pragma unroll 1
  for (int h = 4; h < kCholNB; h *= 2) {
    merge_gemm(0, h, D, Xi, XiT, T, tid);
    __syncthreads();
    merge_gemm(1, h, D, Xi, XiT, T, tid);
    __syncthreads();
    FSTAMP();  // merge level done
  }
  return *s_bad == 0;
}

__global__ void __launch_bounds__(256, 1)
chol_coop_kernel(double* __restrict__ M, int npad, double* __restrict__ Linv, const double* __restrict__ rhs,
                 double* __restrict__ y, double* __restrict__ yf, LmScalars* scal) {
  extern __shared__ __align__(16) unsigned char chol_smem[];
  double* D = reinterpret_cast<double*>(chol_smem);  // diagonal block -> its factor
  double* Xi = D + kTile;                             // inverse of the factor (lower)
  double* XiT = Xi + kTile;                           // its transpose (B operand of the slab GEMM)
  double* S1 = XiT + kTile;                           // operand A, transposed
  double* S2 = S1 + kTile;                            // operand B, transposed / merge scratch / slab result
  double* rdiag = S2 + kTile;                         // [64]
  double* xk = rdiag + kCholNB;                       // [64]
  double* red = xk + kCholNB;                         // [2][64] partial sums
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int G = gridDim.x, cta = blockIdx.x;
  const int nb = npad / kCholNB;
  cg::grid_group grid = cg::this_grid();
#ifdef CTVIO_CHOL_TIMING
  int stamp_n = 0;
#endif
  STAMP();
  for (int r = cta * 256 + tid; r < npad; r += G * 256) y[r] = rhs[r];
  grid.sync();
  STAMP();

#pragma unroll 1
  for (int k = 0; k < nb; ++k) {
    const int nslab = nb - k - 1;
    const int d0 = k * kCholNB;
    // ---------------- phase P: diagonal block, panel slabs, forward substitution ----------------
    if (cta == 0 || cta < nslab) {
      for (int e = tid; e < kCholNB * kCholNB / 2; e += 256) {
        const int r = e >> 5, c = (e & 31) * 2;
        *reinterpret_cast<double2*>(D + r * kTS + c) = *reinterpret_cast<const double2*>(M + size_t(d0 + r) * npad + d0 + c);
      }
      __syncthreads();
      STAMP();  // diag block loaded
      #ifdef CTVIO_CHOL_TIMING
      const bool ok = factor_and_invert_64(D, Xi, XiT, S2, rdiag, &s_bad, &stamp_n);
#else
      const bool ok = factor_and_invert_64(D, Xi, XiT, S2, rdiag, &s_bad, nullptr);
#endif
      if (!ok && cta == 0 && tid == 0) scal->chol_fail = 1;
      STAMP();  // factored + inverted
      // x_k = Xi * y_k  (4 lanes per row)
      {
        const int r = tid >> 2, pt = tid & 3;
        double s = 0.0;
        for (int c = pt; c <= r; c += 4) s = fma(Xi[r * kTS + c], y[d0 + c], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (pt == 0) xk[r] = s;
      }
      if (cta == 0) {  // publish Linv_k
        double* Li = Linv + size_t(k) * kCholNB * kCholNB;
        for (int e = tid; e < kCholNB * kCholNB / 2; e += 256) {
          const int r = e >> 5, c = (e & 31) * 2;
          *reinterpret_cast<double2*>(Li + r * kCholNB + c) = *reinterpret_cast<const double2*>(Xi + r * kTS + c);
        }
      }
      __syncthreads();
      if (cta == 0 && tid < kCholNB) yf[d0 + tid] = xk[tid];  // forward-solved block
      for (int b = cta; b < nslab; b += G) {
        const int r0 = (k + 1 + b) * kCholNB;
        load_tile_transposed(S1, M, npad, r0, d0, tid);
        __syncthreads();
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
        tile_gemm_tt(S1, XiT, acc, ty, tx);  // X = A_slab * Xi^T
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          *reinterpret_cast<double2*>(S2 + (4 * ty + i) * kTS + 4 * tx) = make_double2(acc[i][0], acc[i][1]);
          *reinterpret_cast<double2*>(S2 + (4 * ty + i) * kTS + 4 * tx + 2) = make_double2(acc[i][2], acc[i][3]);
          *reinterpret_cast<double2*>(M + size_t(r0 + 4 * ty + i) * npad + d0 + 4 * tx) = make_double2(acc[i][0], acc[i][1]);
          *reinterpret_cast<double2*>(M + size_t(r0 + 4 * ty + i) * npad + d0 + 4 * tx + 2) = make_double2(acc[i][2], acc[i][3]);
        }
        __syncthreads();
        {  // y_slab -= X * x_k
          const int r = tid >> 2, pt = tid & 3;
          double s = 0.0;
#pragma unroll 4
          for (int c = pt; c < kCholNB; c += 4) s = fma(S2[r * kTS + c], xk[c], s);
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          if (pt == 0) y[r0 + r] -= s;
        }
        __syncthreads();
      }
    }
    STAMP();  // phase P done
    if (nslab == 0) break;
    grid.sync();
    STAMP();  // sync 1
    // ---------------- phase U: trailing update, tiles (bi >= bj) spread over the grid ----------------
    const int ntiles = nslab * (nslab + 1) / 2;
    for (int t = cta; t < ntiles; t += G) {
      int bi = 0, rem = t;
      while (rem > bi) { rem -= bi + 1; ++bi; }
      const int bj = rem;
      const int ri = (k + 1 + bi) * kCholNB, rj = (k + 1 + bj) * kCholNB;
      load_tile_transposed(S1, M, npad, ri, d0, tid);
      if (bi != bj) load_tile_transposed(S2, M, npad, rj, d0, tid);
      __syncthreads();
      double acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
      tile_gemm_tt(S1, bi != bj ? S2 : S1, acc, ty, tx);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double2* p01 = reinterpret_cast<double2*>(M + size_t(ri + 4 * ty + i) * npad + rj + 4 * tx);
        double2 v01 = p01[0], v23 = p01[1];
        v01.x -= acc[i][0]; v01.y -= acc[i][1]; v23.x -= acc[i][2]; v23.y -= acc[i][3];
        p01[0] = v01; p01[1] = v23;
      }
      __syncthreads();
    }
    STAMP();  // phase U done
    grid.sync();
    STAMP();  // sync 2
  }
  if (cta != 0) return;
  // ---------------- backward substitution  L^T x = yf  (CTA 0) ----------------
  __syncthreads();
#pragma unroll 1
  for (int k = nb - 1; k >= 0; --k) {
    const int d0 = k * kCholNB;
    const int c = tid & 63, pt = tid >> 6;
    // t_c = yf_c - sum_{r below block k} L[r][d0 + c] x_r ; 4 row phases x 4 independent accumulators
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int r = d0 + kCholNB + pt;
    for (; r + 12 < npad; r += 16) {
      s0 = fma(M[size_t(r) * npad + d0 + c], y[r], s0);
      s1 = fma(M[size_t(r + 4) * npad + d0 + c], y[r + 4], s1);
      s2 = fma(M[size_t(r + 8) * npad + d0 + c], y[r + 8], s2);
      s3 = fma(M[size_t(r + 12) * npad + d0 + c], y[r + 12], s3);
    }
    for (; r < npad; r += 4) s0 = fma(M[size_t(r) * npad + d0 + c], y[r], s0);
    const double s = (s0 + s1) + (s2 + s3);
    if (pt < 2) red[pt * kCholNB + c] = s;
    __syncthreads();
    if (pt >= 2) red[(pt - 2) * kCholNB + c] += s;
    __syncthreads();
    if (tid < kCholNB) xk[tid] = yf[d0 + tid] - red[tid] - red[kCholNB + tid];
    __syncthreads();
    // x = Linv_k^T * t
    const double* Li = Linv + size_t(k) * kCholNB * kCholNB;
    double u = 0.0;
    for (int rr = c + ((pt - c) & 3); rr < kCholNB; rr += 4) u = fma(Li[rr * kCholNB + c], xk[rr], u);
    if (pt < 2) red[pt * kCholNB + c] = u;
    __syncthreads();
    if (pt >= 2) red[(pt - 2) * kCholNB + c] += u;
    __syncthreads();
    if (tid < kCholNB) y[d0 + tid] = red[tid] + red[kCholNB + tid];
    __syncthreads();
  }
  STAMP();  // backward substitution done
}

#ifdef CTVIO_CHOL_TIMING
extern "C" int ctvio_debug_chol_stamps(unsigned long long* out, int n) {
  return cudaMemcpyFromSymbol(out, g_chol_stamps, sizeof(unsigned long long) * n) == cudaSuccess ? 0 : -1;
}
#endif

int launch_factor_solve(const LinearLaunch& a, cudaStream_t s) {
  static int n_sm = 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(chol_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kCholCoopSmem));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    attr_set = true;
  }
  const int nb = a.npad / kCholNB;
  const int t0 = nb - 1;
  int grid = std::max(1, std::min(n_sm, std::max(t0, t0 * (t0 + 1) / 2)));
  double* M = a.M;
  int npad = a.npad;
  double* Linv = a.Linv;
  const double* rhs = a.rhs;
  double* y = a.y;
  double* yf = a.yf;
  LmScalars* scal = a.scal;
  void* args[] = {&M, &npad, &Linv, &rhs, &y, &yf, &scal};
  cudaLaunchCooperativeKernel(reinterpret_cast<void*>(chol_coop_kernel), dim3(grid), dim3(256), args, kCholCoopSmem, s);
  return 1;
}

}  // namespace ctvio
