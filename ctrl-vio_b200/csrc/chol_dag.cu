// K5 (fast path): tile-DAG Cholesky + both triangular solves of the reduced camera system in ONE persistent
// kernel WITHOUT grid-wide barriers.  Replaces the factor/solve half of Ceres' SPARSE_NORMAL_CHOLESKY step
// (estimator/trajectory_estimator.cpp:374; Ceres itself is not under /root/reference) on the Schur-reduced system.
//
// Owner computes: every 64x64 tile (i, j), i >= j, of the lower triangle has ONE owner CTA that keeps the tile
// in registers from the first to the last update, so the trailing matrix is never re-read or re-written in HBM:
//   * "column" CTA j owns the diagonal tile (j, j) AND the sub-diagonal tile (j, j-1) -- the two tiles on the
//     critical chain  factor(j-1) -> L(j,j-1) -> update of (j,j) -> factor(j)  stay on one SM;
//   * every other tile (i, j), i >= j + 2, has its own CTA.
// Dependencies are point-to-point flags in global memory (release/acquire, epoch valued so they never need
// clearing): tile_ready(i,k) "L(i,k) final and written", diag_ready(j) "Linv_j and the forward-solved x_j
// published", x_ready(r) / bwd_ready(r,k) for the backward sweep.  Owners apply the updates k = 0..j-1 in order,
// which is a topological order of the DAG, so no CTA ever waits on work queued behind its own (all CTAs are
// co-resident: cooperative launch).
// The forward substitution is folded in: the owner of (i,k) publishes L(i,k) x_k, the column CTA i sums those
// partials in a FIXED order (bit-reproducible, required by the replicated solve of the sharded mode); the
// backward sweep reuses the tiles still resident in shared memory: owner (r,k) publishes L(r,k)^T x_r.
// Needs  #tiles - (nb - 1) <= #SMs ; launch_factor_solve falls back to the barrier kernel (chol_coop.cu) otherwise.
#include <algorithm>

#include "chol_tiles.cuh"
#include "kernels.h"

namespace ctvio {

namespace {

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// all threads of the CTA: wait until *flag == epoch (thread 0 spins), then make the producer's data visible
__device__ __forceinline__ void wait_flag(const int* flag, int epoch) {
  if (threadIdx.x == 0) {
    while (ld_acquire(flag) != epoch) {}
  }
  __syncthreads();
}
// all threads have written their part of the payload; publish it
__device__ __forceinline__ void post_flag(int* flag, int epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    st_release(flag, epoch);
  }
}

// register block of the owned tile <- global (L2 path: the tile may have been produced by another SM)
__device__ __forceinline__ void load_block(double acc[4][4], const double* M, int npad, int r0, int c0, int ty, int tx) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double2* p = reinterpret_cast<const double2*>(M + size_t(r0 + 4 * ty + i) * npad + c0 + 4 * tx);
    const double2 v01 = __ldcg(p), v23 = __ldcg(p + 1);
    acc[i][0] = v01.x; acc[i][1] = v01.y; acc[i][2] = v23.x; acc[i][3] = v23.y;
  }
}
// smem tile <- TRANSPOSE of a 64x64 global block with row stride ld: dst[c][r] = src[r][c]
__device__ __forceinline__ void load_tile_t_cg(double* dst, const double* src, int ld, int tid) {
  for (int e = tid; e < kCholNB * kCholNB / 2; e += 256) {
    const int r = e >> 5, c = (e & 31) * 2;
    const double2 v = __ldcg(reinterpret_cast<const double2*>(src + size_t(r) * ld + c));
    dst[c * kTS + r] = v.x;
    dst[(c + 1) * kTS + r] = v.y;
  }
}
// acc -= A * B^T, operands transposed in smem (see tile_gemm_tt)
__device__ __forceinline__ void tile_gemm_tt_sub(const double* At, const double* Bt, double acc[4][4], int ty, int tx) {
#pragma unroll 8
  for (int c = 0; c < kCholNB; ++c) {
    const double2 a01 = *reinterpret_cast<const double2*>(At + c * kTS + 4 * ty);
    const double2 a23 = *reinterpret_cast<const double2*>(At + c * kTS + 4 * ty + 2);
    const double2 b01 = *reinterpret_cast<const double2*>(Bt + c * kTS + 4 * tx);
    const double2 b23 = *reinterpret_cast<const double2*>(Bt + c * kTS + 4 * tx + 2);
    const double av[4] = {-a01.x, -a01.y, -a23.x, -a23.y};
    const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
  }
}
// register block -> smem, transposed (dst[c][r]) or row-major (dst[r][c])
__device__ __forceinline__ void store_block_t(double* dst, const double acc[4][4], int ty, int tx) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    *reinterpret_cast<double2*>(dst + (4 * tx + j) * kTS + 4 * ty) = make_double2(acc[0][j], acc[1][j]);
    *reinterpret_cast<double2*>(dst + (4 * tx + j) * kTS + 4 * ty + 2) = make_double2(acc[2][j], acc[3][j]);
  }
}
__device__ __forceinline__ void store_block(double* dst, const double acc[4][4], int ty, int tx) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    *reinterpret_cast<double2*>(dst + (4 * ty + i) * kTS + 4 * tx) = make_double2(acc[i][0], acc[i][1]);
    *reinterpret_cast<double2*>(dst + (4 * ty + i) * kTS + 4 * tx + 2) = make_double2(acc[i][2], acc[i][3]);
  }
}
__device__ __forceinline__ void store_block_global(double* M, int npad, int r0, int c0, const double acc[4][4], int ty, int tx) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double2* p = reinterpret_cast<double2*>(M + size_t(r0 + 4 * ty + i) * npad + c0 + 4 * tx);
    p[0] = make_double2(acc[i][0], acc[i][1]);
    p[1] = make_double2(acc[i][2], acc[i][3]);
  }
}
// out[r] = sum_c Lrm[r][c] * v[c]  (Lrm row-major smem tile; 4 lanes per row); optional shared copy of the result
__device__ __forceinline__ void tile_matvec(const double* Lrm, const double* v, double* out_global, int tid,
                                            double* out_shared = nullptr) {
  const int r = tid >> 2, pt = tid & 3;
  double s = 0.0;
#pragma unroll 4
  for (int c = pt; c < kCholNB; c += 4) s = fma(Lrm[r * kTS + c], v[c], s);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  if (pt == 0) {
    out_global[r] = s;
    if (out_shared) out_shared[r] = s;
  }
}
// out[c] = sum_r Lrm[r][c] * v[r]  (4 row phases per column, reduced through `red` [4][64])
__device__ __forceinline__ void tile_matvec_t(const double* Lrm, const double* v, double* red, double* out_global, int tid) {
  const int c = tid & 63, pt = tid >> 6;
  double s = 0.0;
#pragma unroll 4
  for (int r = pt; r < kCholNB; r += 4) s = fma(Lrm[r * kTS + c], v[r], s);
  red[pt * kCholNB + c] = s;
  __syncthreads();
  if (tid < kCholNB) out_global[tid] = (red[tid] + red[kCholNB + tid]) + (red[2 * kCholNB + tid] + red[3 * kCholNB + tid]);
}

}  // namespace

struct CholDagArgs {
  double* M;          // [npad][npad], lower tiles overwritten with L (diagonal tiles untouched)
  int npad;
  double* Linv;       // [nb][64][64]
  const double* rhs;  // [npad]
  double* y;          // [npad] solution
  double* yf;         // [npad] forward-solved right-hand side
  double* part;       // [2][nb][nb][64] partial products of the forward / backward sweeps
  int* flags;         // tile_ready[nb*nb] | bwd_ready[nb*nb] | diag_ready[nb] | x_ready[nb]
  int epoch;
  LmScalars* scal;
};

constexpr size_t kCholDagSmem = (6 * size_t(kTile) + 8 * kCholNB) * sizeof(double);

__global__ void __launch_bounds__(256, 1) chol_dag_kernel(CholDagArgs a) {
  extern __shared__ __align__(16) unsigned char dag_smem[];
  double* D = reinterpret_cast<double*>(dag_smem);  // column CTA: diagonal tile (row-major) -> scratch of the factor
  double* Xi = D + kTile;                            // column CTA: Linv_j
  double* XiT = Xi + kTile;                          // column CTA: Linv_j^T
  double* S1 = XiT + kTile;                          // operand A (transposed)
  double* S2 = S1 + kTile;                           // operand B (transposed) / factor scratch
  double* Lrm = S2 + kTile;                          // the owned off-diagonal tile once final, row-major
  double* rdiag = Lrm + kTile;                       // [64]
  double* vec = rdiag + kCholNB;                     // [64]
  double* vec2 = vec + kCholNB;                      // [64]
  double* red = vec2 + kCholNB;                      // [4][64]
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int npad = a.npad, nb = npad / kCholNB, epoch = a.epoch;
  int* tile_ready = a.flags;
  int* bwd_ready = a.flags + nb * nb;
  int* diag_ready = a.flags + 2 * nb * nb;
  int* x_ready = diag_ready + nb;
  double* fwd_part = a.part;                          // [i][k][64] = L(i,k) x_k
  double* bwd_part = a.part + size_t(nb) * nb * kCholNB;  // [k][r][64] = L(r,k)^T x_r

  const int cta = blockIdx.x;
  if (cta >= nb) {
    // ======================= off-diagonal tile (i, j), i >= j + 2 =======================
    int t = cta - nb, j = 0;
    while (t >= nb - 2 - j) { t -= nb - 2 - j; ++j; }
    const int i = j + 2 + t;
    double acc[4][4];
    load_block(acc, a.M, npad, i * kCholNB, j * kCholNB, ty, tx);
    for (int k = 0; k < j; ++k) {
      wait_flag(tile_ready + i * nb + k, epoch);
      load_tile_t_cg(S1, a.M + size_t(i) * kCholNB * npad + k * kCholNB, npad, tid);
      wait_flag(tile_ready + j * nb + k, epoch);
      load_tile_t_cg(S2, a.M + size_t(j) * kCholNB * npad + k * kCholNB, npad, tid);
      __syncthreads();
      tile_gemm_tt_sub(S1, S2, acc, ty, tx);
      __syncthreads();
    }
    // L(i,j) = T * Linv_j^T
    store_block_t(S1, acc, ty, tx);
    wait_flag(diag_ready + j, epoch);
    load_tile_t_cg(S2, a.Linv + size_t(j) * kCholNB * kCholNB, kCholNB, tid);
    if (tid < kCholNB) vec[tid] = __ldcg(a.yf + j * kCholNB + tid);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
    tile_gemm_tt(S1, S2, acc, ty, tx);
    store_block_global(a.M, npad, i * kCholNB, j * kCholNB, acc, ty, tx);
    store_block(Lrm, acc, ty, tx);
    __syncthreads();
    tile_matvec(Lrm, vec, fwd_part + (size_t(i) * nb + j) * kCholNB, tid);
    post_flag(tile_ready + i * nb + j, epoch);
    // backward sweep: L(i,j)^T x_i
    wait_flag(x_ready + i, epoch);
    if (tid < kCholNB) vec[tid] = __ldcg(a.y + i * kCholNB + tid);
    __syncthreads();
    tile_matvec_t(Lrm, vec, red, bwd_part + (size_t(j) * nb + i) * kCholNB, tid);
    post_flag(bwd_ready + i * nb + j, epoch);
    return;
  }

  // ======================= column CTA j: tiles (j, j) and (j, j-1) =======================
  const int j = cta;
  double accD[4][4], accS[4][4];
  load_block(accD, a.M, npad, j * kCholNB, j * kCholNB, ty, tx);
  if (j >= 1) load_block(accS, a.M, npad, j * kCholNB, (j - 1) * kCholNB, ty, tx);
  for (int k = 0; k + 1 < j; ++k) {
    wait_flag(tile_ready + j * nb + k, epoch);
    load_tile_t_cg(S1, a.M + size_t(j) * kCholNB * npad + k * kCholNB, npad, tid);
    wait_flag(tile_ready + (j - 1) * nb + k, epoch);
    load_tile_t_cg(S2, a.M + size_t(j - 1) * kCholNB * npad + k * kCholNB, npad, tid);
    __syncthreads();
    tile_gemm_tt_sub(S1, S1, accD, ty, tx);
    tile_gemm_tt_sub(S1, S2, accS, ty, tx);
    __syncthreads();
  }
  if (j >= 1) {
    // L(j,j-1) = T * Linv_{j-1}^T, then the last update of the diagonal tile
    store_block_t(S1, accS, ty, tx);
    wait_flag(diag_ready + (j - 1), epoch);
    load_tile_t_cg(S2, a.Linv + size_t(j - 1) * kCholNB * kCholNB, kCholNB, tid);
    if (tid < kCholNB) vec[tid] = __ldcg(a.yf + (j - 1) * kCholNB + tid);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) accS[r][c] = 0.0;
    tile_gemm_tt(S1, S2, accS, ty, tx);
    __syncthreads();  // everybody is done reading S1
    store_block_global(a.M, npad, j * kCholNB, (j - 1) * kCholNB, accS, ty, tx);
    store_block(Lrm, accS, ty, tx);
    store_block_t(S1, accS, ty, tx);
    __syncthreads();
    tile_matvec(Lrm, vec, fwd_part + (size_t(j) * nb + (j - 1)) * kCholNB, tid);
    post_flag(tile_ready + j * nb + (j - 1), epoch);
    tile_gemm_tt_sub(S1, S1, accD, ty, tx);
  }
  store_block(D, accD, ty, tx);
  __syncthreads();
  if (!factor_and_invert_64(D, Xi, XiT, S2, rdiag, &s_bad) && tid == 0) a.scal->chol_fail = 1;
  // forward substitution of block j: x_j = Linv_j (rhs_j - sum_k L(j,k) x_k), partials summed in a fixed order
  if (tid < kCholNB) {
    double s = a.rhs[j * kCholNB + tid];
    for (int k = 0; k < j; ++k) s -= __ldcg(fwd_part + (size_t(j) * nb + k) * kCholNB + tid);
    vec[tid] = s;
  }
  for (int e = tid; e < kCholNB * kCholNB / 2; e += 256) {  // publish Linv_j
    const int r = e >> 5, c = (e & 31) * 2;
    *reinterpret_cast<double2*>(a.Linv + (size_t(j) * kCholNB + r) * kCholNB + c) = *reinterpret_cast<const double2*>(Xi + r * kTS + c);
  }
  __syncthreads();
  tile_matvec(Xi, vec, a.yf + j * kCholNB, tid, vec2);  // shared copy for the backward sweep
  post_flag(diag_ready + j, epoch);

  // backward sweep: x_j = Linv_j^T (yf_j - sum_{r > j} L(r,j)^T x_r)
  for (int r = j + 1; r < nb; ++r) wait_flag(bwd_ready + r * nb + j, epoch);
  if (tid < kCholNB) {
    double s = vec2[tid];
    for (int r = j + 1; r < nb; ++r) s -= __ldcg(bwd_part + (size_t(j) * nb + r) * kCholNB + tid);
    vec[tid] = s;
  }
  __syncthreads();
  tile_matvec(XiT, vec, a.y + j * kCholNB, tid, vec2);  // XiT row-major = Linv^T
  post_flag(x_ready + j, epoch);
  if (j >= 1) {
    // own sub-diagonal tile: L(j,j-1)^T x_j for column CTA j-1  (post_flag's barrier made vec2 visible)
    tile_matvec_t(Lrm, vec2, red, bwd_part + (size_t(j - 1) * nb + j) * kCholNB, tid);
    post_flag(bwd_ready + j * nb + (j - 1), epoch);
  }
}

// number of CTAs the DAG kernel needs for nb block columns
static int dag_grid(int nb) { return nb + (nb >= 3 ? (nb - 1) * (nb - 2) / 2 : 0); }

bool chol_dag_supported(int npad, int n_sm) { return dag_grid(npad / kCholNB) <= n_sm; }

size_t chol_dag_part_len(int npad) {
  const size_t nb = npad / kCholNB;
  return 2 * nb * nb * kCholNB;
}
size_t chol_dag_flags_len(int npad) {
  const size_t nb = npad / kCholNB;
  return 2 * nb * nb + 2 * nb;
}

int launch_chol_dag(const LinearLaunch& l, cudaStream_t s) {
  static bool attr_set = false;
  static int epoch = 0;
  if (!attr_set) {
    cudaFuncSetAttribute(chol_dag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kCholDagSmem));
    attr_set = true;
  }
  CholDagArgs a;
  a.M = l.M; a.npad = l.npad; a.Linv = l.Linv; a.rhs = l.rhs; a.y = l.y; a.yf = l.yf;
  a.part = l.chol_part; a.flags = l.chol_flags; a.scal = l.scal;
  epoch = epoch == 0x7fffffff ? 1 : epoch + 1;
  a.epoch = epoch;
  void* args[] = {&a};
  cudaLaunchCooperativeKernel(reinterpret_cast<void*>(chol_dag_kernel), dim3(dag_grid(l.npad / kCholNB)), dim3(256), args,
                              kCholDagSmem, s);
  return 1;
}

}  // namespace ctvio
