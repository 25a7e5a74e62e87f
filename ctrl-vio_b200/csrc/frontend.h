// Front-end data formats either side of the hot path (frontend.cu): DLT triangulation, wire-format unpacking,
// device-side construction of the sorted image-factor arrays from the resident per-frame feature tables.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"

namespace ctvio {

struct TriangulateArgs {
  int32_t n_frames;
  const double* Rs;          // [n_frames][9] row-major body rotations
  const double* Ps;          // [n_frames][3]
  M3 ric;                    // camera -> body rotation
  V3 tic;
  int32_t n_landmarks;
  const int32_t* start_frame;  // [n_landmarks]
  const int32_t* obs_offset;   // [n_landmarks + 1] into obs_point; frame of observation k = start_frame + k
  const double* obs_point;     // [total][3]  FeaturePerFrame::point (x, y, 1)
  int32_t window_size;         // WINDOW_SIZE (candidate rule start_frame < WINDOW_SIZE - 2)
  double init_depth;           // INIT_DEPTH
  double* depth;               // [n_landmarks] in/out: > 0 is kept
};
int launch_triangulate(const TriangulateArgs& a, cudaStream_t s);

// one tracked feature of one frame in the resident table (32 B)
struct FrameFeature {
  double x, y;   // undistorted bearing (z == 1)
  int32_t id;    // tracker feature id
  int32_t row;   // rounded pixel row (rolling-shutter line)
  int64_t pad;
};
struct UnpackCloudArgs {
  int32_t n;
  const float* points;  // [n][3] geometry_msgs::Point32
  const float* ch_id;   // channels[0]
  const float* ch_v;    // channels[2]
  FrameFeature* out;    // [n] slice of the frame table
};
int launch_unpack_cloud(const UnpackCloudArgs& a, cudaStream_t s);

struct UnpackImuArgs {
  int32_t n;
  const unsigned char* raw;  // device copy of the IMUData records
  int32_t stride, off_gyro, off_accel;
  const int64_t* kf_t;       // [n_kf] keyframe timestamps (bias-node assignment)
  int32_t n_kf;
  int32_t dst0;              // first destination sample
  longlong2* t_node;
  double2* ga;
};
int launch_unpack_imu(const UnpackImuArgs& a, cudaStream_t s);

// image factor k = (anchor table slot, observation table slot, landmark, marg flag); slot = frame_slot * frame_cap + i
struct FactorDesc {
  int32_t slot_i, slot_j, lm, marg;
};
struct GatherFactorsArgs {
  int32_t n;
  const FactorDesc* desc;     // already in K1's sorted (frame-pair group) order
  const FrameFeature* table;  // [n_slots][frame_cap]
  const int64_t* frame_t;     // [n_slots]
  int32_t frame_cap;
  longlong2* t;
  double2* pi;
  double2* pj;
  int4* meta;
};
int launch_gather_factors(const GatherFactorsArgs& a, cudaStream_t s);

// device-resident window bookkeeping (all on the device)
int launch_extend_knots(const StatePtrs& st, int old_n, int new_n, cudaStream_t s);
int launch_slide_state(const StatePtrs& st, int nK, int nB, int dk, int db, int new_bias, double* tmp, cudaStream_t s);
int launch_remap_rho(const double* old_rho, const int32_t* old_index, const double* init_rho, int n, double* out, cudaStream_t s);
int launch_shift_imu_table(longlong2* t, double2* ga, int from, int count, double* tmp, cudaStream_t s);
int launch_gather_imu(const int2* src, int n, const longlong2* tab_t, const double2* tab_ga, longlong2* out_t, double2* out_ga,
                      cudaStream_t s);
int launch_prior_x0(const StatePtrs& st, const int32_t* type, const int32_t* index, int nb, double* x0, cudaStream_t s);

}  // namespace ctvio
