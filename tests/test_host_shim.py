"""The C++ host mirror of ctrlvio::TrajectoryEstimator (ctrl-vio_b200/host/trajectory_estimator.hpp):
compiles against include/ctvio.h with plain g++ (no Eigen / Ceres / ROS), links the C-ABI library, and
 - without a GPU: fails loudly with CTVIO_ERR_NO_DEVICE (no CPU fallback),
 - on a B200 : reproduces the oracle's solve through the reference-shaped call sequence."""
import os
import subprocess

import numpy as np
import pytest

from helpers import pkg, rot_angle_between, syn

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "host_shim", "shim_main")


def build_shim():
    src = os.path.join(HERE, "host_shim", "shim_main.cpp")
    hdr = os.path.join(pkg.PKG_DIR, "host", "trajectory_estimator.hpp")
    pkg.load()
    if not os.path.exists(EXE) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(EXE):
        subprocess.run(["g++", "-std=c++17", "-O2", src, "-o", EXE, f"-L{pkg.CSRC_DIR}", "-lctvio_b200",
                        f"-Wl,-rpath,{pkg.CSRC_DIR}"], check=True)
    return EXE


def dump_window(w, path, iters):
    def wv(f, a, dt):
        a = np.ascontiguousarray(a, dt)
        np.array([a.size], np.int64).tofile(f); a.tofile(f)
    with open(path, "wb") as f:
        wv(f, [w.t0_ns, w.dt_ns, iters, int(w.fix_ld)], np.int64)
        wv(f, w.q0, np.float64); wv(f, w.p0, np.float64); wv(f, w.bias0, np.float64); wv(f, w.rho0, np.float64)
        misc = np.concatenate([[w.ld0, w.ld_lower, w.ld_upper, syn.IMAGE_WEIGHT], syn.Q_CtoI, syn.P_CinI, syn.GRAVITY,
                               [1 / syn.SIGMA_G] * 3 + [1 / syn.SIGMA_A] * 3])
        wv(f, misc, np.float64)
        wv(f, w.ti, np.int64); wv(f, w.tj, np.int64); wv(f, w.rowi, np.int32); wv(f, w.rowj, np.int32)
        wv(f, w.pi, np.float64); wv(f, w.pj, np.float64); wv(f, w.lm, np.int32)
        wv(f, w.imu_t, np.int64); wv(f, w.imu_gyro, np.float64); wv(f, w.imu_accel, np.float64); wv(f, w.imu_node, np.int32)
        wv(f, w.bf_i, np.int32); wv(f, w.bf_j, np.int32); wv(f, w.bf_sqrt_info, np.float64)


def test_host_shim_compiles_and_refuses_without_gpu(tmp_path):
    exe = build_shim()
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    w = syn.config_c1()
    dump_window(w, tmp_path / "in.bin", 3)
    r = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 3, (r.returncode, r.stderr)
    assert "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_host_shim_matches_oracle(oracle_lib, tmp_path):
    exe = build_shim()
    w = syn.config_c2(fix_ld=False)
    dump_window(w, tmp_path / "in.bin", 15)
    r = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(tmp_path / "out.bin", np.float64)
    n = w.n_knots
    q = out[:4 * n].reshape(n, 4); p = out[4 * n:7 * n].reshape(n, 3)
    o = pkg.setup_estimator(oracle_lib, w)
    o.Solve(15)
    qo, po = o.GetKnots()
    assert np.abs(p - po).max() / np.abs(po).max() < 1e-5
    assert rot_angle_between(qo, q).max() < 1e-4
    assert abs(out[-1] - o.GetLineDelay()) < 1e-9
