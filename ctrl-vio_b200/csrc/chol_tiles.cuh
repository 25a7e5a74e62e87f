// Shared device helpers of the dense Cholesky kernels (K5): 64x64 fp64 tiles in shared memory, 256 threads,
// thread (ty, tx) = (tid >> 4, tid & 15) owns the 4x4 register block rows 4ty.., cols 4tx...
#pragma once
#include "kernels.h"

namespace ctvio {

constexpr int kTS = kCholNB + 4;  // shared tile row stride (doubles): rows stay 16-B aligned; 68 = 4 mod 8 keeps the
                                  // m8n8k4 fragment loads ([4 k-rows][8 consecutive]) of a half-warp on distinct banks
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

// In-place lower Cholesky of the 64x64 block D (row stride kTS) by 256 threads, then Xi = D^-1 (lower
// triangular) and XiT = Xi^T.  T is a scratch area (>= 272 doubles), rdiag[64] receives 1/L_jj.
// Right-looking on 4x4 register blocks; the next diagonal block is factored by its owner thread while everybody
// else is still applying the rank-4 update (look-ahead); the inverse is produced in the same sweep by the threads
// the elimination front has retired (forward substitution of the identity), so there is no separate inversion.
__device__ __forceinline__ bool factor_and_invert_64(double* D, double* Xi, double* XiT, double* T, double* rdiag, int* s_bad) {
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  double* Pt = T;          // [4][64]  current block column of L, transposed: Pt[m][row]  (bank-conflict-free operand)
  double* Ldd = T + 256;   // [4][4]   current diagonal block of L
  if (tid == 0) *s_bad = 0;
  // a: block (ty, tx) of A while tx is ahead of the elimination front, afterwards (ty > tx) the running sum
  //    W(ty, tx) = -sum_m L(ty, m) X(m, tx) of the inverse  X = L^-1  (forward substitution of the identity,
  //    carried by the threads the trailing update has already retired).
  double a[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double2 v01 = *reinterpret_cast<const double2*>(D + (4 * ty + i) * kTS + 4 * tx);
    const double2 v23 = *reinterpret_cast<const double2*>(D + (4 * ty + i) * kTS + 4 * tx + 2);
    a[i][0] = v01.x; a[i][1] = v01.y; a[i][2] = v23.x; a[i][3] = v23.y;
  }
  for (int e = tid; e < kTile; e += 256) Xi[e] = 0.0;
  if (tid == 0) {  // diagonal block 0
    double l[4][4], rd[4];
    if (!chol4(a, l, rd)) *s_bad = 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      rdiag[i] = rd[i];
#pragma unroll
      for (int j = 0; j < 4; ++j) Ldd[i * 4 + j] = j <= i ? l[i][j] : 0.0;
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int jb = 0; jb < 16; ++jb) {
    // ---- phase 1: block column jb of L, block row jb of X ----
    if ((tx == jb && ty > jb) || (ty == jb && tx <= jb)) {
      double l[4][4], rd[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        rd[c] = rdiag[4 * jb + c];
        const double2 l01 = *reinterpret_cast<const double2*>(Ldd + 4 * c);
        const double2 l23 = *reinterpret_cast<const double2*>(Ldd + 4 * c + 2);
        l[c][0] = l01.x; l[c][1] = l01.y; l[c][2] = l23.x; l[c][3] = l23.y;
      }
      if (ty > jb) {
        // panel block by substitution: x[r][c] = (a[r][c] - sum_{m<c} x[r][m] l[c][m]) / l[c][c]
        double x[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            double t = a[r][c];
#pragma unroll
            for (int m = 0; m < c; ++m) t = fma(-x[r][m], l[c][m], t);
            x[r][c] = t * rd[c];
            a[r][c] = 0.0;  // from now on this thread accumulates W(ty, jb)
          }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          *reinterpret_cast<double2*>(Pt + c * 64 + 4 * ty) = make_double2(x[0][c], x[1][c]);
          *reinterpret_cast<double2*>(Pt + c * 64 + 4 * ty + 2) = make_double2(x[2][c], x[3][c]);
        }
      } else {
        // X(jb, tx) = Ldd^-1 * W   (W = identity on the diagonal block): forward substitution per column
        if (tx == jb) {
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) a[r][c] = r == c ? 1.0 : 0.0;
        }
        double x[4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            double t = a[r][c];
#pragma unroll
            for (int m = 0; m < r; ++m) t = fma(-l[r][m], x[m][c], t);
            x[r][c] = t * rd[r];
          }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          *reinterpret_cast<double2*>(Xi + (4 * jb + r) * kTS + 4 * tx) = make_double2(x[r][0], x[r][1]);
          *reinterpret_cast<double2*>(Xi + (4 * jb + r) * kTS + 4 * tx + 2) = make_double2(x[r][2], x[r][3]);
        }
      }
    }
    __syncthreads();
    // ---- phase 2: rank-4 update of the trailing blocks (tx > jb) and of the inverse sums (tx <= jb) ----
    if (ty > jb && (tx <= jb || ty >= tx)) {
      const bool inv = tx <= jb;
      double lr[4][4], lc[4][4];  // lr[i][m] = L(4ty+i, m), lc[j][m]: B^T operand
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const double2 r01 = *reinterpret_cast<const double2*>(Pt + m * 64 + 4 * ty);
        const double2 r23 = *reinterpret_cast<const double2*>(Pt + m * 64 + 4 * ty + 2);
        lr[0][m] = r01.x; lr[1][m] = r01.y; lr[2][m] = r23.x; lr[3][m] = r23.y;
        const double* Bsrc = inv ? Xi + (4 * jb + m) * kTS + 4 * tx : Pt + m * 64 + 4 * tx;
        const double2 c01 = *reinterpret_cast<const double2*>(Bsrc);
        const double2 c23 = *reinterpret_cast<const double2*>(Bsrc + 2);
        lc[0][m] = c01.x; lc[1][m] = c01.y; lc[2][m] = c23.x; lc[3][m] = c23.y;
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
          for (int j = 0; j < 4; ++j) Ldd[i * 4 + j] = j <= i ? l[i][j] : 0.0;
        }
      }
    }
    __syncthreads();
  }
  // XiT = Xi^T (B operand of the slab GEMM)
  for (int e = tid; e < kCholNB * kCholNB; e += 256) {
    const int r = e >> 6, c = e & 63;
    XiT[c * kTS + r] = Xi[r * kTS + c];
  }
  __syncthreads();
  return *s_bad == 0;
}

}  // namespace ctvio
