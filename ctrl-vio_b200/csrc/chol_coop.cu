// K5: blocked right-looking Cholesky of the reduced camera system + both triangular solves in ONE
// cooperative persistent kernel (grid-wide barriers instead of two launches per block column).
// Replaces the factor/solve half of Ceres' SPARSE_NORMAL_CHOLESKY step (trajectory_estimator.cpp:374;
// Ceres is not under /root/reference) on the Schur-reduced system.
//
//   per block column k (NB = 64):
//     phase P  every CTA that owns a panel slab loads the (already updated) diagonal block, factors
//              it and inverts the factor redundantly in shared memory (latency-optimised: 4-way split
//              dot products, rsqrt pivots, recursive 16-block inversion), turns its slabs' TRSM into
//              a GEMM with that inverse, and folds the forward substitution of the right-hand side in;
//     phase U  the trailing tiles are spread over all CTAs (64^3 register-tiled GEMM each).
//   afterwards CTA 0 runs the backward substitution with the stored block inverses.
// M: lower triangle used; strictly-lower panels are overwritten with L, diagonal blocks are left
// untouched (only their inverses, Linv, are kept).
#include <cooperative_groups.h>

#include <algorithm>

#include "kernels.h"

namespace ctvio {

namespace cg = cooperative_groups;

// optional phase timing of CTA 0 (debug): CTVIO_CHOL_TIMING=1 at build time; stamps go to g_chol_stamps
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
constexpr size_t kCholCoopSmem = (4 * size_t(kCholNB) * kTS + 4 * kCholNB) * sizeof(double);

// acc[4][4] += A[64x64] * B^T, A row-major in smem (As[i][c]), B TRANSPOSED in smem (Bt[c][j]); 256 threads
__device__ __forceinline__ void tile_gemm_abt(const double* As, const double* Bt, double acc[4][4], int ty, int tx) {
#pragma unroll 4
  for (int c = 0; c < kCholNB; ++c) {
    double av[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) av[i] = As[(4 * ty + i) * kTS + c];
    const double2 b01 = *reinterpret_cast<const double2*>(Bt + c * kTS + 4 * tx);
    const double2 b23 = *reinterpret_cast<const double2*>(Bt + c * kTS + 4 * tx + 2);
    const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
  }
}

// Register-tiled product for the triangular-inverse merge: for every pair of adjacent h-blocks on the
// diagonal, C = alpha * A * B on h x h operands living in shared tiles (row stride kTS).  One thread per
// 4x4 output tile, so every k-step feeds 16 independent FMAs.
//   mode 0: C = T      <- L21 * X11       (A = D,  B = Xi)
//   mode 1: C = Xi21   <- -(X22 * T)      (A = Xi, B = T)
__device__ __forceinline__ void merge_gemm(int mode, int h, const double* D, double* Xi, double* T, int tid) {
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
    for (int m = 0; m < h; ++m) {
      double av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = A[(r0 + i) * kTS + m];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = B[m * kTS + c0 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    double* C = mode == 0 ? T + (o + h) * kTS + o : Xi + (o + h) * kTS + o;
    const double sg = mode == 0 ? 1.0 : -1.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) C[(r0 + i) * kTS + c0 + j] = sg * acc[i][j];
  }
}

// In-place lower Cholesky of the 64x64 block D (row stride kTS) by 256 threads, then Xi = D^-1 (lower
// triangular, full tile written).  Right-looking on 4x4 register blocks: thread (ty, tx) owns block
// (rows 4ty.., cols 4tx..); per block column jb: the diagonal owner factors + inverts its 4x4 block in
// registers, the 4x4 panel blocks below are multiplied by that inverse, everybody to the right applies the
// rank-4 update from shared memory.  The 64x64 inverse is then assembled from the 4x4 diagonal inverses by
// recursive merges [[L11,0],[L21,L22]]^-1 = [[X11,0],[-X22 L21 X11, X22]].
// T is a scratch tile.  Returns false (uniformly) when a pivot is not positive / finite.
__device__ bool factor_and_invert_64(double* D, double* Xi, double* T, int* s_bad) {
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  if (tid == 0) *s_bad = 0;
  double a[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = D[(4 * ty + i) * kTS + 4 * tx + j];
  for (int e = tid; e < kCholNB * kCholNB; e += 256) Xi[(e >> 6) * kTS + (e & 63)] = 0.0;
  __syncthreads();
  for (int jb = 0; jb < 16; ++jb) {
    if (ty == jb && tx == jb) {
      // 4x4 Cholesky + inverse in registers
      double l[4][4], li[4][4], rd[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double v = a[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) v = fma(-l[j][k], l[j][k], v);
        if (!(v > 0.0) || !isfinite(v)) { *s_bad = 1; v = 1.0; }
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
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (r < c) { li[r][c] = 0.0; continue; }
          double t = (r == c) ? 1.0 : 0.0;
#pragma unroll
          for (int m = c; m < r; ++m) t = fma(-l[r][m], li[m][c], t);
          li[r][c] = t * rd[r];
        }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          D[(4 * jb + i) * kTS + 4 * jb + j] = j <= i ? l[i][j] : 0.0;
          Xi[(4 * jb + i) * kTS + 4 * jb + j] = j <= i ? li[i][j] : 0.0;
        }
    }
    __syncthreads();
    if (tx == jb && ty > jb) {
      // panel block: x = a * li^T   (li lower: x[r][c] = sum_{m<=c} a[r][m] li[c][m])
      double li[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) li[i][j] = Xi[(4 * jb + i) * kTS + 4 * jb + j];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        double x[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double t = 0.0;
#pragma unroll
          for (int m = 0; m <= c; ++m) t = fma(a[r][m], li[c][m], t);
          x[c] = t;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) D[(4 * ty + r) * kTS + 4 * jb + c] = x[c];
      }
    }
    __syncthreads();
    if (tx > jb && ty >= tx) {
      double lr[4][4], lc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          lr[i][m] = D[(4 * ty + i) * kTS + 4 * jb + m];
          lc[i][m] = D[(4 * tx + i) * kTS + 4 * jb + m];
        }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int m = 0; m < 4; ++m) a[i][j] = fma(-lr[i][m], lc[j][m], a[i][j]);
    }
  }
  // strictly-upper part of D is never read again by the callers' products except through Xi; zero it for safety
  if (ty < tx) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) D[(4 * ty + i) * kTS + 4 * tx + j] = 0.0;
  }
  __syncthreads();
  for (int h = 4; h < kCholNB; h *= 2) {
    merge_gemm(0, h, D, Xi, T, tid);
    __syncthreads();
    merge_gemm(1, h, D, Xi, T, tid);
    __syncthreads();
  }
  return *s_bad == 0;
}

__global__ void __launch_bounds__(256, 1)
chol_coop_kernel(double* __restrict__ M, int npad, double* __restrict__ Linv, const double* __restrict__ rhs,
                 double* __restrict__ y, double* __restrict__ yf, LmScalars* scal) {
  extern __shared__ __align__(16) unsigned char chol_smem[];
  double* D = reinterpret_cast<double*>(chol_smem);  // diagonal block -> its factor
  double* Xi = D + kCholNB * kTS;                     // inverse of the factor (lower)
  double* S1 = Xi + kCholNB * kTS;                    // operand A (row-major) / slab result
  double* S2 = S1 + kCholNB * kTS;                    // operand B transposed / merge scratch
  double* rdiag = S2 + kCholNB * kTS;                 // [64]
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

  for (int k = 0; k < nb; ++k) {
    const int nslab = nb - k - 1;
    const int d0 = k * kCholNB;
    // ---------------- phase P: diagonal block, panel slabs, forward substitution ----------------
    if (cta == 0 || cta < nslab) {
      for (int e = tid; e < kCholNB * kCholNB; e += 256) D[(e >> 6) * kTS + (e & 63)] = M[size_t(d0 + (e >> 6)) * npad + d0 + (e & 63)];
      __syncthreads();
      STAMP();  // diag block loaded
      const bool ok = factor_and_invert_64(D, Xi, S2, &s_bad);
      STAMP();  // factored + inverted
      if (!ok && cta == 0 && tid == 0) scal->chol_fail = 1;
      // x_k = Xi * y_k  (4 lanes per row)
      {
        const int r = tid >> 2, pt = tid & 3;
        double s = 0.0;
        for (int c = pt; c <= r; c += 4) s = fma(Xi[r * kTS + c], y[d0 + c], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (pt == 0) xk[r] = s;
      }
      // S2 <- Xi^T (B operand of the slab GEMM);  CTA 0 publishes Linv_k and the forward-solved block
      __syncthreads();
      for (int e = tid; e < kCholNB * kCholNB; e += 256) S2[(e & 63) * kTS + (e >> 6)] = Xi[(e >> 6) * kTS + (e & 63)];
      if (cta == 0) {
        double* Li = Linv + size_t(k) * kCholNB * kCholNB;
        for (int e = tid; e < kCholNB * kCholNB; e += 256) Li[e] = Xi[(e >> 6) * kTS + (e & 63)];
        if (tid < kCholNB) yf[d0 + tid] = xk[tid];
      }
      __syncthreads();
      for (int b = cta; b < nslab; b += G) {
        const int r0 = (k + 1 + b) * kCholNB;
        for (int e = tid; e < kCholNB * kCholNB; e += 256) S1[(e >> 6) * kTS + (e & 63)] = M[size_t(r0 + (e >> 6)) * npad + d0 + (e & 63)];
        __syncthreads();
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
        tile_gemm_abt(S1, S2, acc, ty, tx);  // X = A_slab * Xi^T
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            S1[(4 * ty + i) * kTS + 4 * tx + j] = acc[i][j];
            M[size_t(r0 + 4 * ty + i) * npad + d0 + 4 * tx + j] = acc[i][j];
          }
        __syncthreads();
        // y_slab -= X * x_k
        {
          const int r = tid >> 2, pt = tid & 3;
          double s = 0.0;
          for (int c = pt; c < kCholNB; c += 4) s = fma(S1[r * kTS + c], xk[c], s);
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
      for (int e = tid; e < kCholNB * kCholNB; e += 256) {
        const int r = e >> 6, c = e & 63;
        S1[r * kTS + c] = M[size_t(ri + r) * npad + d0 + c];
        S2[c * kTS + r] = M[size_t(rj + r) * npad + d0 + c];
      }
      __syncthreads();
      double acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
      tile_gemm_abt(S1, S2, acc, ty, tx);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) M[size_t(ri + 4 * ty + i) * npad + rj + 4 * tx + j] -= acc[i][j];
      __syncthreads();
    }
    STAMP();  // phase U done
    grid.sync();
    STAMP();  // sync 2
  }
  if (cta != 0) return;
  // ---------------- backward substitution  L^T x = yf  (CTA 0) ----------------
  __syncthreads();
  for (int k = nb - 1; k >= 0; --k) {
    const int d0 = k * kCholNB;
    const int c = tid & 63, pt = tid >> 6;
    double s = 0.0;
    for (int r = d0 + kCholNB + pt; r < npad; r += 4) s = fma(M[size_t(r) * npad + d0 + c], y[r], s);
    red[(pt & 1) * kCholNB + c] = 0.0;
    __syncthreads();
    // 4 partial sums per column: two rounds through the [2][64] buffer
    if (pt < 2) red[pt * kCholNB + c] = s;
    __syncthreads();
    if (pt >= 2) red[(pt - 2) * kCholNB + c] += s;
    __syncthreads();
    if (tid < kCholNB) xk[tid] = yf[d0 + tid] - red[tid] - red[kCholNB + tid];
    __syncthreads();
    // x = Linv_k^T * t
    const double* Li = Linv + size_t(k) * kCholNB * kCholNB;
    double u = 0.0;
    for (int r = c + ((pt - c) & 3); r < kCholNB; r += 4) u = fma(Li[r * kCholNB + c], xk[r], u);  // rows r >= c, r = pt mod 4
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
  double* yf = a.rhs + a.npad;  // rhs buffer is allocated with 2 * npad doubles
  LmScalars* scal = a.scal;
  void* args[] = {&M, &npad, &Linv, &rhs, &y, &yf, &scal};
  cudaLaunchCooperativeKernel(reinterpret_cast<void*>(chol_coop_kernel), dim3(grid), dim3(256), args, kCholCoopSmem, s);
  return 1;
}

}  // namespace ctvio
