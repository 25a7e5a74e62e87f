// 64x64 fp64 tiles on the 5th-generation tensor cores' fp64 path (mma.sync.m8n8k4.f64; tcgen05 has no fp64 kind):
// shared by the dense Cholesky (K5) and the Schur complement (K4).
#pragma once
#include "chol_tiles.cuh"

namespace ctvio {

// ---- fp64 tensor-core tiles (mma.sync.m8n8k4.f64, measured 37 TFLOP/s = the DFMA peak, at 1/5 of the
// shared-memory operand traffic of a 4x4 register-tiled DFMA loop) ----
// Warp (wm, wn) = (warp & 3, warp >> 2) owns rows 16 wm.., cols 32 wn.. of the 64x64 tile as 2 x 4 m8n8 fragments;
// lane (g, q) = (lane >> 2, lane & 3) holds C[8 mt + g][8 nt + 2 q + {0, 1}] of each fragment.
struct Frag {
  double c[2][4][2];
};
struct Lane {
  int row0, col0, g, q;  // first row / column of the warp tile, lane coordinates
};
__device__ __forceinline__ Lane lane_of(int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  return Lane{16 * (warp & 3), 32 * (warp >> 2), lane >> 2, lane & 3};
}
__device__ __forceinline__ void frag_zero(Frag& f) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) f.c[mt][nt][0] = f.c[mt][nt][1] = 0.0;
}
// f (+|-)= A * B^T with both operands in [k][row] layout in smem (At[k][i] = A[i][k], Bt[k][j] = B[j][k])
template <bool SUB, int K = kCholNB>
__device__ __forceinline__ void tile_gemm_dmma(const double* At, const double* Bt, Frag& f, const Lane& L) {
  const double* pa = At + L.q * kTS + L.row0 + L.g;
  const double* pb = Bt + L.q * kTS + L.col0 + L.g;
#pragma unroll 4
  for (int k0 = 0; k0 < K; k0 += 4) {
    double av[2], bv[4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) av[mt] = SUB ? -pa[k0 * kTS + 8 * mt] : pa[k0 * kTS + 8 * mt];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) bv[nt] = pb[k0 * kTS + 8 * nt];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(f.c[mt][nt][0]), "+d"(f.c[mt][nt][1])
                     : "d"(av[mt]), "d"(bv[nt]));
  }
}
// fragments <- global tile (row-major, row stride npad; L2 path: another SM may have produced it)
__device__ __forceinline__ void frag_load_global(Frag& f, const double* tile, int npad, const Lane& L) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const double2 v = __ldcg(reinterpret_cast<const double2*>(tile + size_t(L.row0 + 8 * mt + L.g) * npad + L.col0 + 8 * nt + 2 * L.q));
      f.c[mt][nt][0] = v.x; f.c[mt][nt][1] = v.y;
    }
}
// fragments -> smem row-major (dst[r][c]) / transposed (dst[c][r])
__device__ __forceinline__ void frag_store(double* dst, const Frag& f, const Lane& L) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
      *reinterpret_cast<double2*>(dst + (L.row0 + 8 * mt + L.g) * kTS + L.col0 + 8 * nt + 2 * L.q) = make_double2(f.c[mt][nt][0], f.c[mt][nt][1]);
}
__device__ __forceinline__ void frag_store_t(double* dst, const Frag& f, const Lane& L) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) dst[(L.col0 + 8 * nt + 2 * L.q + e) * kTS + L.row0 + 8 * mt + L.g] = f.c[mt][nt][e];
}
}  // namespace ctvio
