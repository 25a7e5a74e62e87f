// Small host/device helpers shared by engine.cu: prior bookkeeping, Gram kernel, gradient dot product,
// NCCL communicator wrappers.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "kernels.h"

namespace ctvio {

// MarginalizationInfo payload (factor/analytic_diff/marginalization_factor.h:96-131) in index-identity form
struct PriorHost {
  int n = 0;
  std::vector<double> J, r, x0;            // n x n row-major, n, n_blocks x 4
  std::vector<int32_t> type, index, col;   // per kept block
};

// first camera dim of a parameter block, -1 if it has none / is out of range
inline int prior_block_base(int type, int index, int nK, int nB) {
  switch (type) {
    case 0: return (index >= 0 && index < nK) ? 6 * index : -1;
    case 1: return (index >= 0 && index < nK) ? 6 * index + 3 : -1;
    case 2: return (index >= 0 && index < nB) ? 6 * nK + 6 * index : -1;
    case 3: return (index >= 0 && index < nB) ? 6 * nK + 6 * index + 3 : -1;
    case 4: return 6 * nK + 6 * nB;
    default: return -1;
  }
}

// G = J' J for a row-major rows x cols matrix (G: cols x cols)
int launch_gram(const double* J, int rows, int cols, double* G, cudaStream_t s);
// scal->gd = gc . dc + gl . dl  (directional derivative of the cost along the step, candidate buffers)
int launch_dot_gradient(const LinearLaunch& a, cudaStream_t s);

// measured fp64 FMA throughput of the current device (TFLOP/s); < 0 on error
double measure_fp64_tflops(cudaStream_t s);

bool comm_unique_id(uint8_t* id128, std::string* err);
void* comm_create(int rank, int world, const uint8_t* id128, std::string* err);
void comm_destroy(void* comm);
// in-place sum-allreduce of n doubles on the stream; returns false on error
bool comm_allreduce_sum(void* comm, double* buf, size_t n, cudaStream_t s, std::string* err);

}  // namespace ctvio
