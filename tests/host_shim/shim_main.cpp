// Test driver for the C++ host mirror (ctrl-vio_b200/host/trajectory_estimator.hpp): reads a window dumped by
// tests/test_host_shim.py, builds the problem through the reference-shaped TrajectoryEstimator methods exactly the
// way TrajectoryManager::UpdateTrajectory does (prior, image, IMU, bias factors; trajectory_manager.cpp:352-453),
// solves and writes the updated caller-owned blocks back out.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

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
