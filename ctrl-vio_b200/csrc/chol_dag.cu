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
#include "dmma_tiles.cuh"
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
// two flags at once: two polling threads in different warps, one barrier
__device__ __forceinline__ void wait_flags2(const int* f1, const int* f2, int epoch) {
  if (threadIdx.x == 0) {
    while (ld_acquire(f1) != epoch) {}
  } else if (threadIdx.x == 32) {
    while (ld_acquire(f2) != epoch) {}
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

// smem tile <- 64x64 global block with row stride ld, straight copy (16-byte accesses on both sides): final tiles
// and block inverses are PUBLISHED TRANSPOSED, i.e. already in the [k][row] operand layout
__device__ __forceinline__ void load_tile_cg(double* dst, const double* src, int ld, int tid) {
#pragma unroll
  for (int e = tid; e < kCholNB * kCholNB / 2; e += 256) {
    const int r = e >> 5, c = (e & 31) * 2;
    *reinterpret_cast<double2*>(dst + r * kTS + c) = __ldcg(reinterpret_cast<const double2*>(src + size_t(r) * ld + c));
  }
}
// smem tile (stride kTS) -> global 64x64 slot (row stride ld), straight copy with 16-byte accesses
__device__ __forceinline__ void store_tile_global(double* dst, int ld, const double* src, int tid) {
#pragma unroll
  for (int e = tid; e < kCholNB * kCholNB / 2; e += 256) {
    const int r = e >> 5, c = (e & 31) * 2;
    *reinterpret_cast<double2*>(dst + size_t(r) * ld + c) = *reinterpret_cast<const double2*>(src + r * kTS + c);
  }
}
// s[tid & 63] partial: sum over q = (tid >> 6), q + 4, ... < n of part[q * stride + (tid & 63)], combined over the
// four thread groups in a fixed order through red[4][64]; returns the total for tid < 64 (after a barrier)
__device__ __forceinline__ double sum_partials(const double* part, int n, size_t stride, double* red, int tid) {
  const int c = tid & 63, g = tid >> 6;
  double s = 0.0;
  int q = g;
  for (; q + 12 < n; q += 16) {
    const double v0 = __ldcg(part + size_t(q) * stride + c), v1 = __ldcg(part + size_t(q + 4) * stride + c);
    const double v2 = __ldcg(part + size_t(q + 8) * stride + c), v3 = __ldcg(part + size_t(q + 12) * stride + c);
    s += v0; s += v1; s += v2; s += v3;
  }
  for (; q < n; q += 4) s += __ldcg(part + size_t(q) * stride + c);
  red[g * kCholNB + c] = s;
  __syncthreads();
  return tid < kCholNB ? (red[tid] + red[kCholNB + tid]) + (red[2 * kCholNB + tid] + red[3 * kCholNB + tid]) : 0.0;
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
}  // namespace

struct CholDagArgs {
  double* M;          // [npad][npad], strictly-lower tiles overwritten with L(i,j)^T (transposed inside the tile slot)
  int npad;
  double* Linv;       // [nb][64][64]  TRANSPOSED block inverses (Linv_j^T), the B operand of the panel GEMMs
  const double* rhs;  // [npad]
  double* y;          // [npad] solution
  double* yf;         // [npad] forward-solved right-hand side
  double* part;       // [2][nb][nb][64] partial products of the forward / backward sweeps
  int* flags;         // tile_ready[nb*nb] | bwd_ready[nb*nb] | diag_ready[nb] | x_ready[nb] | fwd_ready[nb]
  int epoch;
  LmScalars* scal;
};

#ifdef CTVIO_CHOL_TIMING
__device__ unsigned long long g_dag_stamps[32 * 16];
#define DSTAMP(j, i) do { if (threadIdx.x == 0 && (j) < 32) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_dag_stamps[(j) * 16 + (i)] = t_; } } while (0)
extern "C" int ctvio_debug_fac_clk(long long* out) {
  return cudaMemcpyFromSymbol(out, g_fac_clk, sizeof(g_fac_clk)) == cudaSuccess ? 0 : -1;
}
extern "C" int ctvio_debug_dag_stamps(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_dag_stamps, sizeof(g_dag_stamps)) == cudaSuccess ? 0 : -1;
}
#else
#define DSTAMP(j, i)
#endif

constexpr size_t kCholDagSmem = (6 * size_t(kTile) + 8 * kCholNB) * sizeof(double);

__global__ void __launch_bounds__(256, 1) chol_dag_kernel(CholDagArgs a) {
  extern __shared__ __align__(16) unsigned char dag_smem[];
  double* D = reinterpret_cast<double*>(dag_smem);  // column CTA: diagonal tile (row-major) -> scratch of the factor
  double* Xi = D + kTile;                            // column CTA: Linv_j
  double* XiT = Xi + kTile;                          // column CTA: Linv_j^T
  double* S1 = XiT + kTile;                          // operand A ([k][row]); later the owned final tile, transposed
  double* S2 = S1 + kTile;                           // operand B ([k][row]) / factor scratch
  double* Lrm = S2 + kTile;                          // the owned off-diagonal tile once final, row-major
  double* rdiag = Lrm + kTile;                       // [64]
  double* vec = rdiag + kCholNB;                     // [64]
  double* vec2 = vec + kCholNB;                      // [64]
  double* red = vec2 + kCholNB;                      // [4][64]
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  const Lane L = lane_of(tid);
  const int npad = a.npad, nb = npad / kCholNB, epoch = a.epoch;
  int* tile_ready = a.flags;
  int* bwd_ready = a.flags + nb * nb;
  int* diag_ready = a.flags + 2 * nb * nb;
  int* x_ready = diag_ready + nb;
  int* fwd_ready = x_ready + nb;
  double* fwd_part = a.part;                          // [i][k][64] = L(i,k) x_k
  double* bwd_part = a.part + size_t(nb) * nb * kCholNB;  // [k][r][64] = L(r,k)^T x_r

  const int cta = blockIdx.x;
  if (cta >= nb) {
    // ======================= off-diagonal tile (i, j), i >= j + 2 =======================
    int t = cta - nb, j = 0;
    while (t >= nb - 2 - j) { t -= nb - 2 - j; ++j; }
    const int i = j + 2 + t;
    double* slot = a.M + size_t(i) * kCholNB * npad + j * kCholNB;
    Frag acc;
    frag_load_global(acc, slot, npad, L);
    for (int k = 0; k < j; ++k) {
      wait_flags2(tile_ready + i * nb + k, tile_ready + j * nb + k, epoch);
      load_tile_cg(S1, a.M + size_t(i) * kCholNB * npad + k * kCholNB, npad, tid);
      load_tile_cg(S2, a.M + size_t(j) * kCholNB * npad + k * kCholNB, npad, tid);
      __syncthreads();
      tile_gemm_dmma<true>(S1, S2, acc, L);
      __syncthreads();
    }
    // L(i,j) = T * Linv_j^T
    frag_store_t(S1, acc, L);
    wait_flag(diag_ready + j, epoch);
    load_tile_cg(S2, a.Linv + size_t(j) * kCholNB * kCholNB, kCholNB, tid);
    __syncthreads();
    frag_zero(acc);
    tile_gemm_dmma<false, kCholNB, kGemmLowerB>(S1, S2, acc, L);
    __syncthreads();  // everybody is done reading S1
    frag_store(Lrm, acc, L);
    frag_store_t(S1, acc, L);
    __syncthreads();
    store_tile_global(slot, npad, S1, tid);  // published transposed
    wait_flag(fwd_ready + j, epoch);
    if (tid < kCholNB) vec[tid] = __ldcg(a.yf + j * kCholNB + tid);
    __syncthreads();
    tile_matvec(Lrm, vec, fwd_part + (size_t(i) * nb + j) * kCholNB, tid);
    post_flag(tile_ready + i * nb + j, epoch);
    // backward sweep: L(i,j)^T x_i
    wait_flag(x_ready + i, epoch);
    if (tid < kCholNB) vec[tid] = __ldcg(a.y + i * kCholNB + tid);
    __syncthreads();
    tile_matvec(S1, vec, bwd_part + (size_t(j) * nb + i) * kCholNB, tid);
    post_flag(bwd_ready + i * nb + j, epoch);
    return;
  }

  // ======================= column CTA j: tiles (j, j) and (j, j-1) =======================
  const int j = cta;
  DSTAMP(j, 0);
  Frag accD, accS;
  frag_load_global(accD, a.M + size_t(j) * kCholNB * npad + j * kCholNB, npad, L);
  double* slot = a.M + size_t(j) * kCholNB * npad + (j >= 1 ? j - 1 : 0) * kCholNB;
  if (j >= 1) frag_load_global(accS, slot, npad, L);
  for (int k = 0; k + 1 < j; ++k) {
    wait_flags2(tile_ready + j * nb + k, tile_ready + (j - 1) * nb + k, epoch);
    load_tile_cg(S1, a.M + size_t(j) * kCholNB * npad + k * kCholNB, npad, tid);
    load_tile_cg(S2, a.M + size_t(j - 1) * kCholNB * npad + k * kCholNB, npad, tid);
    __syncthreads();
    tile_gemm_dmma<true, kCholNB, kGemmLowerOut>(S1, S1, accD, L);
    tile_gemm_dmma<true>(S1, S2, accS, L);
    __syncthreads();
  }
  DSTAMP(j, 1);
  {  // right-hand side of block j for the forward substitution: the partials of the other owners (all in by now), fixed
    // summation order; this CTA's own partial L(j,j-1) x_{j-1} is added after the factorisation (side job below)
    const double sp = sum_partials(fwd_part + size_t(j) * nb * kCholNB, j >= 1 ? j - 1 : 0, kCholNB, red, tid);
    if (tid < kCholNB) vec[tid] = a.rhs[j * kCholNB + tid] - sp;
  }
  if (j >= 1) {
    // L(j,j-1) = T * Linv_{j-1}^T, then the last update of the diagonal tile
    frag_store_t(S1, accS, L);
    wait_flag(diag_ready + (j - 1), epoch);
    DSTAMP(j, 2);
    load_tile_cg(S2, a.Linv + size_t(j - 1) * kCholNB * kCholNB, kCholNB, tid);
    __syncthreads();
    DSTAMP(j, 8);
    frag_zero(accS);
    tile_gemm_dmma<false, kCholNB, kGemmLowerB>(S1, S2, accS, L);
    __syncthreads();  // everybody is done reading S1
    DSTAMP(j, 9);
    frag_store(Lrm, accS, L);
    frag_store_t(S1, accS, L);
    __syncthreads();
    DSTAMP(j, 10);
    tile_gemm_dmma<true, kCholNB, kGemmLowerOut>(S1, S1, accD, L);
    DSTAMP(j, 3);
  }
  frag_store(D, accD, L);
  __syncthreads();
  DSTAMP(j, 4);
  // While warp 0 runs the pivot chain of the first two 16-column steps, warps 1..7 publish L(j,j-1) (tile store, fence,
  // flag) and form its forward partial: ~2 us that used to sit on the critical chain between two factorisations.
  double* own = red;  // [64] L(j,j-1) x_{j-1}
  auto side = [&](int step) {
    if (j == 0) return;
    const int tt = tid - 32;  // 0..223
    if (step == 0) {
      for (int e = tt; e < kCholNB * kCholNB / 2; e += 224) {
        const int r = e >> 5, c = (e & 31) * 2;
        *reinterpret_cast<double2*>(slot + size_t(r) * npad + c) = *reinterpret_cast<const double2*>(S1 + r * kTS + c);
      }
      asm volatile("bar.sync 1, 224;" ::: "memory");
      if (tt == 0) {
        __threadfence();
        st_release(tile_ready + j * nb + (j - 1), epoch);
      }
    } else if (step == 1 && tt < kCholNB) {
      // x_{j-1} was published (fwd_ready) a microsecond after Linv_{j-1}: long ago by now
      if (tt == 0) { while (ld_acquire(fwd_ready + (j - 1)) != epoch) {} }
      asm volatile("bar.sync 2, 64;" ::: "memory");
      vec2[tt] = __ldcg(a.yf + (j - 1) * kCholNB + tt);
    } else if (step == 2 && tt < 128) {
      const int r = tt >> 1, pt = tt & 1;
      double acc = 0.0;
#pragma unroll 8
      for (int c = pt; c < kCholNB; c += 2) acc = fma(Lrm[r * kTS + c], vec2[c], acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (pt == 0) own[r] = acc;
    }
  };
  if (!factor_and_invert_64(D, Xi, XiT, S2, rdiag, &s_bad, side) && tid == 0) a.scal->chol_fail = 1;
  DSTAMP(j, 5);
  // forward substitution of block j: x_j = Linv_j (rhs_j - sum_k L(j,k) x_k)
  if (j >= 1 && tid < kCholNB) vec[tid] -= own[tid];
  store_tile_global(a.Linv + size_t(j) * kCholNB * kCholNB, kCholNB, XiT, tid);  // publish Linv_j^T first: it is what
  post_flag(diag_ready + j, epoch);                                               // the next column's chain waits for
  DSTAMP(j, 6);
  tile_matvec(Xi, vec, a.yf + j * kCholNB, tid, vec2);  // x_j (forward); shared copy for the backward sweep
  post_flag(fwd_ready + j, epoch);

  // backward sweep: x_j = Linv_j^T (yf_j - sum_{r > j} L(r,j)^T x_r)
  // all partials of this block column: one polling thread per flag (the early ones cost a single L2 round trip in
  // parallel instead of nb - 1 - j sequential ones), then one barrier
  if (tid < nb - 1 - j) {
    const int* f = bwd_ready + (j + 1 + tid) * nb + j;
    while (ld_acquire(f) != epoch) {}
  }
  __syncthreads();
  {
    const double sp = sum_partials(bwd_part + (size_t(j) * nb + j + 1) * kCholNB, nb - 1 - j, kCholNB, red, tid);
    if (tid < kCholNB) vec[tid] = vec2[tid] - sp;
  }
  __syncthreads();
  tile_matvec(XiT, vec, a.y + j * kCholNB, tid, vec2);  // XiT row-major = Linv^T
  if (j >= 1) {
    // own sub-diagonal tile: L(j,j-1)^T x_j for column CTA j-1, published together with x_j (one fence for both:
    // the next column of the backward chain waits for exactly this partial)
    __syncthreads();
    tile_matvec(S1, vec2, bwd_part + (size_t(j - 1) * nb + j) * kCholNB, tid);  // S1 = L(j,j-1)^T
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (j >= 1) st_release(bwd_ready + j * nb + (j - 1), epoch);
    st_release(x_ready + j, epoch);
  }
  DSTAMP(j, 7);
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
  return 2 * nb * nb + 3 * nb;
}

int launch_chol_dag(const LinearLaunch& l, cudaStream_t s) {
  static PerDeviceOnce once;
  static std::atomic<unsigned> epoch_src{0};
  if (once.first()) cudaFuncSetAttribute(chol_dag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kCholDagSmem));
  CholDagArgs a;
  a.M = l.M; a.npad = l.npad; a.Linv = l.Linv; a.rhs = l.rhs; a.y = l.y; a.yf = l.yf;
  a.part = l.chol_part; a.flags = l.chol_flags; a.scal = l.scal;
  // process-wide unique, never 0 (flag buffers start zeroed); a wrap after 2^31 launches would need the flags of a
  // buffer to hold exactly the value 2^31 launches old: not a practical concern
  a.epoch = int(epoch_src.fetch_add(1, std::memory_order_relaxed) % 0x7ffffffeu) + 1;
  void* args[] = {&a};
  const cudaError_t err = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(chol_dag_kernel), dim3(dag_grid(l.npad / kCholNB)),
                                                      dim3(256), args, kCholDagSmem, s);
  if (err != cudaSuccess) {
    // the flag protocol needs every CTA resident; if the runtime cannot promise that (MIG slice, fewer usable SMs than
    // reported, ...) the barrier kernel with its own, smaller grid is the safe path
    cudaGetLastError();
    return launch_chol_coop(l, s);
  }
  return 1;
}

}  // namespace ctvio
