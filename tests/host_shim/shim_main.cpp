// Test driver for the C++ host mirror (ctrl-vio_b200/host/trajectory_estimator.hpp): reads a window dumped by
// tests/test_host_shim.py, builds the problem through the reference-shaped TrajectoryEstimator methods exactly the
// way TrajectoryManager::UpdateTrajectory does (prior, image, IMU, bias factors; trajectory_manager.cpp:352-453),
// solves and writes the updated caller-owned blocks back out.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "../../ctrl-vio_b200/host/trajectory_estimator.hpp"

using namespace ctvio_host;

template <typename T>
static std::vector<T> readv(std::ifstream& f) {
  int64_t n = 0;
  f.read(reinterpret_cast<char*>(&n), sizeof(n));
  std::vector<T> v(n);
  f.read(reinterpret_cast<char*>(v.data()), n * sizeof(T));
  return v;
}

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: shim_main in.bin out.bin\n"); return 2; }
  std::ifstream f(argv[1], std::ios::binary);
  auto hdr = readv<int64_t>(f);  // t0, dt, n_iters, fix_ld
  auto q = readv<double>(f); auto p = readv<double>(f); auto bias = readv<double>(f); auto rho = readv<double>(f);
  auto misc = readv<double>(f);  // ld, ld_lower, ld_upper, image_weight, q_CI(4), p_CI(3), gravity(3), imu_info(6)
  auto ti = readv<int64_t>(f); auto tj = readv<int64_t>(f); auto rowi = readv<int32_t>(f); auto rowj = readv<int32_t>(f);
  auto pi = readv<double>(f); auto pj = readv<double>(f); auto lm = readv<int32_t>(f);
  auto imu_t = readv<int64_t>(f); auto gyro = readv<double>(f); auto accel = readv<double>(f); auto node = readv<int32_t>(f);
  auto bfi = readv<int32_t>(f); auto bfj = readv<int32_t>(f); auto bfs = readv<double>(f);

  auto traj = std::make_shared<Trajectory>();
  traj->t0_ns = hdr[0]; traj->dt_ns = hdr[1];
  traj->knot_q = q; traj->knot_p = p;
  traj->SetLineDelay(misc[0], hdr[3] != 0, misc[1], misc[2]);
  for (int k = 0; k < 4; ++k) traj->q_CtoI[k] = misc[4 + k];
  for (int k = 0; k < 3; ++k) traj->p_CinI[k] = misc[8 + k];
  TrajectoryEstimatorOptions option;
  option.lock_ab = false; option.lock_wb = false;   // trajectory_manager.cpp:346-348
  if (argc >= 4 && std::string(argv[3]) == "cycle") {
    // The reference's per-image sequence through the mirror with the reference's OWN argument shapes (Eigen-like values
    // with .data(), IMUData-like records): InitTrajectory -> UpdateTrajectory -> double2vector -> UpdateVIOPrior.
    auto cyc = readv<int64_t>(f);   // nowk, later, init_fixed_idx, init_t_min
    auto img_marg = readv<int32_t>(f); auto imu_marg = readv<int32_t>(f); auto bias_marg = readv<int32_t>(f);
    struct Vec3 { double v[3]; const double* data() const { return v; } };
    struct Vec6 { double v[6]; const double* data() const { return v; } };
    struct IMUData { int64_t timestamp; Vec3 gyro, accel; };
    const Vec3 gravity{{misc[11], misc[12], misc[13]}};
    Vec6 info{};
    for (int k = 0; k < 6; ++k) info.v[k] = misc[14 + k];
    std::vector<std::pair<double*, double*>> nodes;
    for (size_t k = 0; k < bias.size() / 6; ++k) nodes.emplace_back(&bias[6 * k], &bias[6 * k + 3]);
    std::vector<double*> feats;
    for (auto& r : rho) feats.push_back(&r);
    try {
      {  // ---- InitTrajectory (trajectory_manager.cpp:288-315) ----
        TrajectoryEstimatorOptions o1;  // lock_ab = lock_wb = true
        TrajectoryEstimator est(traj, o1, misc[3], &misc[14], &misc[11]);
        est.RegisterBiasNodes(nodes); est.RegisterLandmarks(feats);
        double* bg = &bias[6 * (bias.size() / 6 - 1)];
        for (size_t k = 0; k < imu_t.size(); ++k) {
          if (imu_t[k] < cyc[3]) continue;
          IMUData v{imu_t[k], {{gyro[3 * k], gyro[3 * k + 1], gyro[3 * k + 2]}}, {{accel[3 * k], accel[3 * k + 1], accel[3 * k + 2]}}};
          est.AddIMUMeasurementAnalytic(v, gravity, bg, bg + 3, info);
        }
        est.SetFixedIndex(int(cyc[2]));
        SolverSummary s1 = est.Solve(8, false);
        std::printf("init: iterations %d cost %.12g -> %.12g\n", s1.iterations, s1.initial_cost, s1.final_cost);
      }
      double R0[9], t0[3];
      {
        const double* qk = traj->getKnotSO3(size_t(cyc[0]));
        const double x = qk[0], y = qk[1], z = qk[2], w = qk[3];
        const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z),
                             2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
        for (int k = 0; k < 9; ++k) R0[k] = R[k];
        for (int k = 0; k < 3; ++k) t0[k] = traj->getKnotPos(size_t(cyc[0]))[k];
      }
      auto add_all = [&](TrajectoryEstimator& est, bool with_marg) {
        for (size_t k = 0; k < ti.size(); ++k) {
          const Vec3 a{{pi[2 * k], pi[2 * k + 1], 1.0}}, b{{pj[2 * k], pj[2 * k + 1], 1.0}};
          est.AddImageFeatureDelayAnalytic(ti[k], rowi[k], a, tj[k], rowj[k], b, &rho[lm[k]], &traj->line_delay, false,
                                           with_marg && img_marg[k] != 0);
        }
        for (size_t k = 0; k < imu_t.size(); ++k) {
          if (with_marg && !imu_marg[k]) continue;   // UpdateVIOPrior only adds the samples before keyframe 1 (:239-253)
          IMUData v{imu_t[k], {{gyro[3 * k], gyro[3 * k + 1], gyro[3 * k + 2]}}, {{accel[3 * k], accel[3 * k + 1], accel[3 * k + 2]}}};
          est.AddIMUMeasurementAnalytic(v, gravity, &bias[6 * node[k]], &bias[6 * node[k] + 3], info, with_marg);
        }
        for (size_t k = 0; k < bfi.size(); ++k) {
          if (with_marg && !bias_marg[k]) continue;  // only the first bias factor (:256-263)
          Vec6 sq{};
          for (int c = 0; c < 6; ++c) sq.v[c] = bfs[6 * k + c];
          est.AddBiasFactor(&bias[6 * bfi[k]], &bias[6 * bfj[k]], &bias[6 * bfi[k] + 3], &bias[6 * bfj[k] + 3], 1.0, sq, with_marg);
        }
      };
      {  // ---- UpdateTrajectory (:317-483) + double2vector (:485-516) ----
        TrajectoryEstimator est(traj, option, misc[3], &misc[14], &misc[11]);
        est.RegisterBiasNodes(nodes); est.RegisterLandmarks(feats);
        add_all(est, false);
        SolverSummary s2 = est.Solve(int(hdr[2]), false);
        est.GaugeRealign(int(cyc[0]), R0, t0);
        std::printf("update: iterations %d cost %.12g -> %.12g\n", s2.iterations, s2.initial_cost, s2.final_cost);
      }
      MarginalizationInfo::Ptr info_out;
      std::vector<double*> blocks_out;
      ResidualSummary rs;
      {  // ---- UpdateVIOPrior, MARGIN_OLD (:122-286) ----
        TrajectoryEstimatorOptions o3;
        o3.lock_ab = false; o3.lock_wb = false;
        o3.is_marg_state = true; o3.ctrl_to_be_opt_now = int(cyc[0]); o3.ctrl_to_be_opt_later = int(cyc[1]);
        TrajectoryEstimator est(traj, o3, misc[3], &misc[14], &misc[11]);
        est.RegisterBiasNodes(nodes); est.RegisterLandmarks(feats);
        add_all(est, true);
        rs = est.GetResidualSummary();
        est.SaveMarginalizationInfo(info_out, blocks_out);
      }
      std::ofstream o(argv[2], std::ios::binary);
      auto wr = [&](const std::vector<double>& v) { o.write(reinterpret_cast<const char*>(v.data()), v.size() * sizeof(double)); };
      wr(traj->knot_q); wr(traj->knot_p); wr(bias); wr(rho);
      o.write(reinterpret_cast<const char*>(&traj->line_delay), sizeof(double));
      const double n = info_out ? double(info_out->n) : 0.0, nblk = double(blocks_out.size());
      o.write(reinterpret_cast<const char*>(&n), sizeof(double));
      o.write(reinterpret_cast<const char*>(&nblk), sizeof(double));
      if (info_out) { wr(info_out->linearized_jacobians); wr(info_out->linearized_residuals); }
      std::vector<double> sum;
      for (int t : {int(RType_Image), int(RType_IMU), int(RType_Bias)}) {
        sum.push_back(double(rs.err_type_number[t]));
        for (double v : rs.err_type_sum[t]) sum.push_back(v);
      }
      wr(sum);
      // the block pointers handed back must be the caller's own blocks (pointer identity, like Ceres parameter blocks)
      int own = 0;
      for (double* ptr : blocks_out) {
        bool ok = ptr == &traj->line_delay;
        for (size_t k = 0; k < traj->numKnots() && !ok; ++k) ok = ptr == traj->getKnotSO3(k) || ptr == traj->getKnotPos(k);
        for (auto& nd : nodes) ok = ok || ptr == nd.first || ptr == nd.second;
        own += ok ? 1 : 0;
      }
      std::printf("prior: n %d blocks %zu (all caller-owned: %s)\n", int(n), blocks_out.size(), own == int(blocks_out.size()) ? "yes" : "NO");
      return own == int(blocks_out.size()) ? 0 : 5;
    } catch (const Error& e) {
      std::fprintf(stderr, "ctvio error %d: %s\n", e.code, e.what());
      return e.code == CTVIO_ERR_NO_DEVICE ? 3 : 4;
    }
  }
  try {
    TrajectoryEstimator est(traj, option, misc[3], &misc[14], &misc[11]);
    // caller-owned parameter blocks: bias nodes (all_imu_bias_) and para_Feature
    std::vector<std::pair<double*, double*>> nodes;
    for (size_t k = 0; k < bias.size() / 6; ++k) nodes.emplace_back(&bias[6 * k], &bias[6 * k + 3]);
    est.RegisterBiasNodes(nodes);
    std::vector<double*> feats;
    for (auto& r : rho) feats.push_back(&r);
    est.RegisterLandmarks(feats);
    for (size_t k = 0; k < ti.size(); ++k) {
      const double a[3] = {pi[2 * k], pi[2 * k + 1], 1.0}, b[3] = {pj[2 * k], pj[2 * k + 1], 1.0};
      est.AddImageFeatureDelayAnalytic(ti[k], rowi[k], a, tj[k], rowj[k], b, &rho[lm[k]], &traj->line_delay, false, false);
    }
    for (size_t k = 0; k < imu_t.size(); ++k)
      est.AddIMUMeasurementAnalytic(imu_t[k], &gyro[3 * k], &accel[3 * k], &bias[6 * node[k]], &bias[6 * node[k] + 3]);
    for (size_t k = 0; k < bfi.size(); ++k)
      est.AddBiasFactor(&bias[6 * bfi[k]], &bias[6 * bfj[k]], &bias[6 * bfi[k] + 3], &bias[6 * bfj[k] + 3], 1.0, &bfs[6 * k]);
    SolverSummary s = est.Solve(int(hdr[2]), false);
    std::printf("iterations %d cost %.12g -> %.12g\n", s.iterations, s.initial_cost, s.final_cost);
  } catch (const Error& e) {
    std::fprintf(stderr, "ctvio error %d: %s\n", e.code, e.what());
    return e.code == CTVIO_ERR_NO_DEVICE ? 3 : 4;
  }
  std::ofstream o(argv[2], std::ios::binary);
  auto wr = [&](const std::vector<double>& v) { o.write(reinterpret_cast<const char*>(v.data()), v.size() * sizeof(double)); };
  wr(traj->knot_q); wr(traj->knot_p); wr(bias); wr(rho);
  o.write(reinterpret_cast<const char*>(&traj->line_delay), sizeof(double));
  return 0;
}
