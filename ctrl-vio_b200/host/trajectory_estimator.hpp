// Host-side mirror of ctrlvio::TrajectoryEstimator (reference: src/estimator/trajectory_estimator.h:61-206,
// .cpp:97-408) on top of the C-ABI of include/ctvio.h.
//
// Same method names, argument order and meaning as the reference class so that
// TrajectoryManager::{UpdateTrajectory, UpdateVIOPrior, InitTrajectory}
// (src/estimator/trajectory_manager.cpp:122-483) keep their call sites; what changes is that Ceres parameter
// blocks (`double*` identity) become indices.  The adapter keeps the pointer -> index maps, so callers keep
// passing the same `double*` they pass today:
//   knot blocks      trajectory_->getKnotSO3(i).data() / getKnotPos(i).data()   -> knot index i
//   bias blocks      all_imu_bias_[t].gyro_bias.data() / accel_bias.data()      -> bias node index
//   para_Feature[k]                                                              -> landmark index k
//   &trajectory_->line_delay                                                     -> the line delay
// Header-only, C++17, no Eigen / Ceres / ROS: vectors and quaternions cross as plain arrays (Eigen maps bind to
// them without a copy: Eigen::Map<Eigen::Vector3d>(ptr)).
// Error behaviour: the reference aborts (assert / BASALT_ASSERT) on out-of-window times and otherwise ignores
// return values; here every failing C-ABI call throws ctvio::Error with ctvio_last_error().
#pragma once
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ctvio.h"

namespace ctvio_host {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
inline void check(int rc, const char* where) {
  if (rc != CTVIO_OK) throw Error(rc, std::string(where) + ": " + ctvio_last_error());
}

// Minimal stand-in for the parts of ctrlvio::Trajectory (src/spline/trajectory.h:38-117, se3_spline.h) the
// estimator touches: knot storage with stable addresses, time grid, extrinsics, line-delay settings.
struct Trajectory {
  int64_t t0_ns = 0, dt_ns = 50000000;
  std::vector<double> knot_q;  // [n][4] xyzw   (so3_spline.h:410 keeps Sophus::SO3d knots in a deque)
  std::vector<double> knot_p;  // [n][3]
  double q_CtoI[4] = {0, 0, 0, 1}, p_CinI[3] = {0, 0, 0};
  double line_delay = 0, ld_lower = 0, ld_upper = 0;
  bool fix_ld = true;
  size_t numKnots() const { return knot_q.size() / 4; }
  int64_t minTimeNs() const { return t0_ns; }
  int64_t maxTimeNs() const { return t0_ns + int64_t(numKnots() - 3) * dt_ns; }
  size_t knotIndex(int64_t t) const { return size_t((t - t0_ns) / dt_ns); }  // computeTIndexNs(t).second
  double* getKnotSO3(size_t i) { return &knot_q[4 * i]; }
  double* getKnotPos(size_t i) { return &knot_p[3 * i]; }
  void SetLineDelay(double init, bool fix, double lo, double hi) { line_delay = init; fix_ld = fix; ld_lower = lo; ld_upper = hi; }
};

// trajectory_estimator_options.h:34-68 (fields the hot path reads)
struct TrajectoryEstimatorOptions {
  bool lock_traj = false, lock_ab = true, lock_wb = true;
  bool is_marg_state = false;
  int ctrl_to_be_opt_now = 0, ctrl_to_be_opt_later = 0;
};

// MarginalizationInfo payload in index form (marginalization_factor.h:96-131)
struct MarginalizationInfo {
  using Ptr = std::shared_ptr<MarginalizationInfo>;
  int n = 0;
  std::vector<double> linearized_jacobians, linearized_residuals, keep_block_data;  // n x n, n, nb x 4
  std::vector<int32_t> keep_block_type, keep_block_index, keep_block_idx;
};

struct SolverSummary {  // what callers log from ceres::Solver::Summary::BriefReport()
  int iterations = 0, num_successful_steps = 0, num_unsuccessful_steps = 0, termination = 0;
  double initial_cost = 0, final_cost = 0, device_ms = 0;
};

class TrajectoryEstimator {
 public:
  using Ptr = std::shared_ptr<TrajectoryEstimator>;

  // reference: TrajectoryEstimator(Trajectory::Ptr, TrajectoryEstimatorOptions&)   (trajectory_estimator.cpp:97)
  // image_weight / imu_info / gravity are the statics the reference injects through InitFactorInfo
  // (trajectory_manager.cpp:51-62), OptWeight (opt_weight.h:124-126) and gravity_.
  TrajectoryEstimator(std::shared_ptr<Trajectory> trajectory, const TrajectoryEstimatorOptions& option, double image_weight,
                      const double imu_info[6], const double gravity[3], int device = 0)
      : trajectory_(std::move(trajectory)), options(option) {
    ctvio_config cfg{};
    cfg.t0_ns = trajectory_->t0_ns;
    cfg.dt_ns = trajectory_->dt_ns;
    for (int k = 0; k < 4; ++k) cfg.q_CtoI[k] = trajectory_->q_CtoI[k];
    for (int k = 0; k < 3; ++k) { cfg.p_CinI[k] = trajectory_->p_CinI[k]; cfg.gravity[k] = gravity[k]; }
    for (int k = 0; k < 6; ++k) cfg.imu_info[k] = imu_info[k];
    cfg.image_weight = image_weight;
    cfg.rs_padding_ns = 39000000;  // trajectory_estimator.cpp:299
    cfg.cauchy_solve = 2.0;        // :321
    cfg.cauchy_marg = 1.0;
    cfg.device = device;
    check(ctvio_create(&cfg, &h_), "ctvio_create");
  }
  ~TrajectoryEstimator() { if (h_) ctvio_destroy(h_); }
  TrajectoryEstimator(const TrajectoryEstimator&) = delete;
  TrajectoryEstimator& operator=(const TrajectoryEstimator&) = delete;

  void SetFixedIndex(int idx) { fixed_control_point_index_ = idx; }  // trajectory_estimator.h:90
  // Sliding the window: the Trajectory handed to the constructor holds only the window's slice of control points; when
  // the slice moves, shift its time origin (the reference instead keeps one growing spline and freezes old knots,
  // trajectory_manager.cpp:352-361).  Knot / bias indices of the prior are relative to the slice.
  void SetTimeOrigin(int64_t t0_ns) { check(ctvio_set_time_origin(h_, t0_ns), "ctvio_set_time_origin"); }

  // trajectory_estimator.cpp:219-263
  void AddIMUMeasurementAnalytic(int64_t timestamp, const double gyro[3], const double accel[3], double* gyro_bias,
                                 double* accel_bias, bool marg_this_factor = false) {
    imu_t_.push_back(timestamp);
    for (int k = 0; k < 3; ++k) { imu_gyro_.push_back(gyro[k]); imu_accel_.push_back(accel[k]); }
    imu_node_.push_back(biasNode(gyro_bias, accel_bias));
    imu_marg_.push_back(options.is_marg_state && marg_this_factor ? 1 : 0);
  }
  // trajectory_estimator.cpp:265-291 (sqrt_info is divided by sqrt(dt) like BiasFactor's constructor)
  void AddBiasFactor(double* bias_gyr_i, double* bias_gyr_j, double* bias_acc_i, double* bias_acc_j, double dt,
                     const double info_vec[6], bool marg_this_factor = false) {
    bf_i_.push_back(biasNode(bias_gyr_i, bias_acc_i));
    bf_j_.push_back(biasNode(bias_gyr_j, bias_acc_j));
    for (int k = 0; k < 6; ++k) bf_s_.push_back(info_vec[k] / std::sqrt(dt));
    bf_marg_.push_back(options.is_marg_state && marg_this_factor ? 1 : 0);
  }
  // trajectory_estimator.cpp:293-332; pi / pj are the undistorted normalised points (x, y, 1)
  void AddImageFeatureDelayAnalytic(int64_t ti, int rowi, const double pi[3], int64_t tj, int rowj, const double pj[3],
                                    double* inv_depth, double* line_delay, bool /*fixed_depth*/, bool marg_this_feature) {
    (void)line_delay;  // always &trajectory_->line_delay
    img_ti_.push_back(ti); img_tj_.push_back(tj); img_rowi_.push_back(rowi); img_rowj_.push_back(rowj);
    img_pi_.push_back(pi[0]); img_pi_.push_back(pi[1]); img_pj_.push_back(pj[0]); img_pj_.push_back(pj[1]);
    img_lm_.push_back(landmark(inv_depth));
    img_marg_.push_back(options.is_marg_state && marg_this_feature ? 1 : 0);
  }
  // trajectory_estimator.cpp:334-348
  void AddMarginalizationFactor(const MarginalizationInfo::Ptr& last) { prior_ = last; }

  // trajectory_estimator.cpp:367-408 — uploads state + factors, solves in HBM, writes every block back in place
  SolverSummary Solve(int max_iterations = 50, bool /*progress*/ = false, int /*num_threads*/ = -1) {
    upload();
    ctvio_summary s{};
    check(ctvio_solve(h_, max_iterations, &s), "ctvio_solve");
    download();
    SolverSummary out;
    out.iterations = s.iterations; out.num_successful_steps = s.num_successful_steps;
    out.num_unsuccessful_steps = s.num_unsuccessful_steps; out.termination = s.termination;
    out.initial_cost = s.initial_cost; out.final_cost = s.final_cost; out.device_ms = s.device_ms;
    return out;
  }

  // TrajectoryManager::double2vector (trajectory_manager.cpp:485-516): R0 row-major, t0; knots >= min_idx
  void GaugeRealign(int min_idx, const double R0[9], const double t0[3]) {
    check(ctvio_gauge_realign(h_, min_idx, R0, t0), "ctvio_gauge_realign");
    download();
  }

  // trajectory_estimator.cpp:184-204: every factor added with its marg flag (and the attached prior) is
  // recorded; returns nullptr when nothing can be kept
  MarginalizationInfo::Ptr SaveMarginalizationInfo() {
    upload();
    int32_t n = 0, nb = 0;
    check(ctvio_marginalize(h_, &n, &nb), "ctvio_marginalize");
    if (n <= 0) return nullptr;
    auto m = std::make_shared<MarginalizationInfo>();
    m->n = n;
    m->linearized_jacobians.resize(size_t(n) * n);
    m->linearized_residuals.resize(n);
    m->keep_block_type.resize(nb); m->keep_block_index.resize(nb); m->keep_block_idx.resize(nb);
    m->keep_block_data.resize(4 * size_t(nb));
    check(ctvio_get_prior(h_, m->linearized_jacobians.data(), m->linearized_residuals.data(), m->keep_block_type.data(),
                          m->keep_block_index.data(), m->keep_block_idx.data(), m->keep_block_data.data()),
          "ctvio_get_prior");
    return m;
  }

  // registration of the caller-owned blocks (the reference discovers them through AddParameterBlock)
  void RegisterBiasNodes(const std::vector<std::pair<double*, double*>>& bg_ba) { bias_nodes_ = bg_ba; }
  void RegisterLandmarks(const std::vector<double*>& inv_depths) { landmarks_ = inv_depths; }

  TrajectoryEstimatorOptions options;

 private:
  int biasNode(double* bg, double* ba) {
    for (size_t k = 0; k < bias_nodes_.size(); ++k)
      if (bias_nodes_[k].first == bg && bias_nodes_[k].second == ba) return int(k);
    bias_nodes_.emplace_back(bg, ba);
    return int(bias_nodes_.size()) - 1;
  }
  int landmark(double* inv_depth) {
    auto it = lm_index_.find(inv_depth);
    if (it != lm_index_.end()) return it->second;
    for (size_t k = 0; k < landmarks_.size(); ++k)
      if (landmarks_[k] == inv_depth) { lm_index_[inv_depth] = int(k); return int(k); }
    landmarks_.push_back(inv_depth);
    lm_index_[inv_depth] = int(landmarks_.size()) - 1;
    return int(landmarks_.size()) - 1;
  }
  void upload() {
    Trajectory& T = *trajectory_;
    ctvio_options o{};
    o.fixed_knot_index = fixed_control_point_index_;
    o.lock_traj = options.lock_traj; o.lock_wb = options.lock_wb; o.lock_ab = options.lock_ab;
    o.fix_ld = T.fix_ld; o.ld_lower = T.ld_lower; o.ld_upper = T.ld_upper;
    o.is_marg_state = options.is_marg_state;
    o.ctrl_to_be_opt_now = options.ctrl_to_be_opt_now; o.ctrl_to_be_opt_later = options.ctrl_to_be_opt_later;
    check(ctvio_set_options(h_, &o), "ctvio_set_options");
    check(ctvio_set_knots(h_, int32_t(T.numKnots()), T.knot_q.data(), T.knot_p.data()), "ctvio_set_knots");
    std::vector<double> b(6 * bias_nodes_.size()), r(landmarks_.size());
    for (size_t k = 0; k < bias_nodes_.size(); ++k)
      for (int c = 0; c < 3; ++c) { b[6 * k + c] = bias_nodes_[k].first[c]; b[6 * k + 3 + c] = bias_nodes_[k].second[c]; }
    for (size_t k = 0; k < landmarks_.size(); ++k) r[k] = *landmarks_[k];
    check(ctvio_set_biases(h_, int32_t(bias_nodes_.size()), b.data()), "ctvio_set_biases");
    check(ctvio_set_inv_depths(h_, int32_t(r.size()), r.data()), "ctvio_set_inv_depths");
    check(ctvio_set_line_delay(h_, T.line_delay), "ctvio_set_line_delay");
    check(ctvio_clear_factors(h_), "ctvio_clear_factors");
    if (!img_ti_.empty())
      check(ctvio_add_image_features(h_, int32_t(img_ti_.size()), img_ti_.data(), img_rowi_.data(), img_pi_.data(), img_tj_.data(),
                                     img_rowj_.data(), img_pj_.data(), img_lm_.data(), img_marg_.data()), "ctvio_add_image_features");
    if (!imu_t_.empty())
      check(ctvio_add_imu_measurements(h_, int32_t(imu_t_.size()), imu_t_.data(), imu_gyro_.data(), imu_accel_.data(),
                                       imu_node_.data(), imu_marg_.data()), "ctvio_add_imu_measurements");
    if (!bf_i_.empty())
      check(ctvio_add_bias_factors(h_, int32_t(bf_i_.size()), bf_i_.data(), bf_j_.data(), bf_s_.data(), bf_marg_.data()),
            "ctvio_add_bias_factors");
    if (prior_ && prior_->n > 0)
      check(ctvio_set_prior(h_, prior_->n, prior_->linearized_jacobians.data(), prior_->linearized_residuals.data(),
                            int32_t(prior_->keep_block_type.size()), prior_->keep_block_type.data(), prior_->keep_block_index.data(),
                            prior_->keep_block_idx.data(), prior_->keep_block_data.data()), "ctvio_set_prior");
    else
      check(ctvio_set_prior(h_, 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr), "ctvio_set_prior");
  }
  void download() {  // the reference's solver updates the caller's blocks in place
    Trajectory& T = *trajectory_;
    check(ctvio_get_knots(h_, T.knot_q.data(), T.knot_p.data()), "ctvio_get_knots");
    std::vector<double> b(6 * bias_nodes_.size()), r(landmarks_.size());
    if (!b.empty()) check(ctvio_get_biases(h_, b.data()), "ctvio_get_biases");
    if (!r.empty()) check(ctvio_get_inv_depths(h_, r.data()), "ctvio_get_inv_depths");
    for (size_t k = 0; k < bias_nodes_.size(); ++k)
      for (int c = 0; c < 3; ++c) { bias_nodes_[k].first[c] = b[6 * k + c]; bias_nodes_[k].second[c] = b[6 * k + 3 + c]; }
    for (size_t k = 0; k < landmarks_.size(); ++k) *landmarks_[k] = r[k];
    check(ctvio_get_line_delay(h_, &T.line_delay), "ctvio_get_line_delay");
  }

  std::shared_ptr<Trajectory> trajectory_;
  ctvio_handle h_ = nullptr;
  int fixed_control_point_index_ = -1;
  std::vector<std::pair<double*, double*>> bias_nodes_;
  std::vector<double*> landmarks_;
  std::map<double*, int> lm_index_;
  std::vector<int64_t> img_ti_, img_tj_, imu_t_;
  std::vector<int32_t> img_rowi_, img_rowj_, img_lm_, img_marg_, imu_node_, imu_marg_, bf_i_, bf_j_, bf_marg_;
  std::vector<double> img_pi_, img_pj_, imu_gyro_, imu_accel_, bf_s_;
  MarginalizationInfo::Ptr prior_;
};

}  // namespace ctvio_host
