// Gram matrix of the prior Jacobian (once per prior) and the line-search directional derivative.
#include "marginalize.h"

namespace ctvio {

__global__ void gram_kernel(const double* J, int rows, int cols, double* G) {
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
  if (i >= cols || j >= cols) return;
  double s = 0;
  for (int r = 0; r < rows; ++r) s = fma(J[size_t(r) * cols + i], J[size_t(r) * cols + j], s);
  G[size_t(i) * cols + j] = s;
}
int launch_gram(const double* J, int rows, int cols, double* G, cudaStream_t s) {
  if (cols <= 0) return 0;
  dim3 grid((cols + 15) / 16, (cols + 15) / 16), block(16, 16);
  gram_kernel<<<grid, block, 0, s>>>(J, rows, cols, G);
  return 1;
}

__global__ void dot_gradient_kernel(LinearLaunch a) {
  __shared__ double red[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int np = a.dims.np, nL = a.dims.nL;
  double v = 0;
  if (i < np) v = a.ne.gc[i] * a.dc[i];
  else if (i < np + nL) v = a.ne.gl[i - np] * a.dl[i - np];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < 8; ++w) s += red[w];
    if (s != 0.0) atomicAdd(&a.scal->gd, s);
  }
}
int launch_dot_gradient(const LinearLaunch& a, cudaStream_t s) {
  cudaMemsetAsync(&a.scal->gd, 0, sizeof(double), s);
  const int n = a.dims.np + a.dims.nL;
  dot_gradient_kernel<<<(n + 255) / 256, 256, 0, s>>>(a);
  return 1;
}

}  // namespace ctvio
