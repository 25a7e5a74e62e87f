"""Host emulation of the CUDA engine's per-thread math (csrc/*.cuh compiled with g++) vs the oracle.

No GPU needed: the same inline functions the kernels call are built into tests/emu/libemu.so and
checked against the oracle's factor evaluations on identical inputs.  fp64 with different (fused)
operation order => tolerance 1e-9 relative, far inside the 1e-5 / 1e-4 north-star tolerances.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import pkg, small_window, syn

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(HERE, "emu", "libemu.so")
    src = os.path.join(HERE, "emu", "emu.cpp")
    hdrs = [os.path.join(pkg.CSRC_DIR, f) for f in ("device_math.cuh", "spline_eval.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", src, "-o", so], check=True)
    return C.CDLL(so)


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("seed,ld", [(3, 0.0), (4, 20e-6), (5, 34.9e-6)])
def test_image_factor_math_matches_oracle(oracle_lib, emu, seed, ld):
    w = small_window(seed=seed, n_knots=9, n_kf=5, per_frame=8, fix_ld=False)
    e = pkg.setup_estimator(oracle_lib, w, state="init")
    e.SetLineDelay(ld)
    for cauchy in (0.0, 2.0):
        r_o, s_o, J_o, c_o = e.EvalImageFactors(True, cauchy)
        q = np.ascontiguousarray(w.q0); p = np.ascontiguousarray(w.p0)
        qci = np.ascontiguousarray(syn.Q_CtoI); pci = np.ascontiguousarray(syn.P_CinI)
        tot = 0.0
        for n in range(w.n_obs):
            r = np.zeros(2); s = np.zeros(2, np.int32); J = np.zeros(100); cost = C.c_double()
            pi = np.ascontiguousarray(w.pi[n]); pj = np.ascontiguousarray(w.pj[n])
            rc = emu.emu_eval_image(C.c_int64(w.t0_ns), C.c_int64(w.dt_ns), C.c_int(w.n_knots), dp(q), dp(p), dp(qci),
                                    dp(pci), C.c_double(syn.IMAGE_WEIGHT), C.c_double(w.rho0[w.lm[n]]), C.c_double(ld),
                                    C.c_int64(int(w.ti[n])), C.c_int(int(w.rowi[n])), dp(pi), C.c_int64(int(w.tj[n])),
                                    C.c_int(int(w.rowj[n])), dp(pj), C.c_double(cauchy), C.c_int(1), dp(r),
                                    s.ctypes.data_as(C.POINTER(C.c_int)), dp(J), C.byref(cost))
            assert rc == 0
            assert np.array_equal(s, s_o[n])
            assert np.allclose(r, r_o[n], rtol=1e-9, atol=1e-9)
            scale = np.abs(J_o[n]).max()
            assert np.allclose(J, J_o[n], rtol=1e-9, atol=1e-10 * scale), (n, np.abs(J - J_o[n]).max() / scale)
            tot += cost.value
        assert np.isclose(tot, c_o, rtol=1e-11)


def test_imu_factor_math_matches_oracle(oracle_lib, emu):
    w = small_window(seed=8, n_knots=9, n_kf=5, per_frame=4)
    e = pkg.setup_estimator(oracle_lib, w, state="init")
    bias = np.ascontiguousarray(np.random.default_rng(0).normal(0, 0.02, (e.n_bias, 6)))
    e.SetBiases(bias)
    r_o, s_o, J_o, c_o = e.EvalImuFactors(True)
    q = np.ascontiguousarray(w.q0); p = np.ascontiguousarray(w.p0)
    g = np.ascontiguousarray(syn.GRAVITY); info = np.array([1 / syn.SIGMA_G] * 3 + [1 / syn.SIGMA_A] * 3)
    tot = 0.0
    for n in range(len(w.imu_t)):
        r = np.zeros(6); s = C.c_int(); J = np.zeros(156); cost = C.c_double()
        gy = np.ascontiguousarray(w.imu_gyro[n]); ac = np.ascontiguousarray(w.imu_accel[n])
        b = np.ascontiguousarray(bias[w.imu_node[n]])
        rc = emu.emu_eval_imu(C.c_int64(w.t0_ns), C.c_int64(w.dt_ns), C.c_int(w.n_knots), dp(q), dp(p), dp(g), dp(info),
                              C.c_int64(int(w.imu_t[n])), dp(gy), dp(ac), dp(b), C.c_int(1), dp(r), C.byref(s), dp(J),
                              C.byref(cost))
        assert rc == 0 and s.value == s_o[n]
        assert np.allclose(r, r_o[n], rtol=1e-9, atol=1e-8)
        scale = np.abs(J_o[n]).max()
        assert np.allclose(J, J_o[n], rtol=1e-9, atol=1e-10 * scale)
        tot += cost.value
    assert np.isclose(tot, c_o, rtol=1e-11)


def test_quat_from_matrix(emu):
    rng = np.random.default_rng(1)
    for _ in range(50):
        q = syn.qexp(rng.normal(0, 2.0, (1, 3)))[0]
        R = syn.qrot(q[None], np.eye(3)).T.copy()
        out = np.zeros(4)
        emu.emu_quat_from_matrix(dp(np.ascontiguousarray(R)), dp(out))
        assert min(np.abs(out - q).max(), np.abs(out + q).max()) < 1e-12


def test_two_stage_evaluation_equals_one_shot(emu):
    """pose_stage + jacobian_stage (fused visual kernel) == eval_side (probe / query path)."""
    w = small_window(seed=14, n_knots=10, n_kf=5, per_frame=2)
    q = np.ascontiguousarray(w.q0); p = np.ascontiguousarray(w.p0)
    emu.emu_two_stage_maxdiff.restype = C.c_double
    rng = np.random.default_rng(2)
    for t in rng.integers(w.t0_ns, w.t0_ns + (w.n_knots - 3) * w.dt_ns - 1, 200):
        d = emu.emu_two_stage_maxdiff(C.c_int64(w.t0_ns), C.c_int64(w.dt_ns), C.c_int(w.n_knots), dp(q), dp(p),
                                      C.c_int64(int(t)))
        assert 0 <= d < 1e-12, d
    # identical consecutive knots (freshly extended trajectory, SURVEY C-16): Taylor branches
    q2 = q.copy(); q2[5:] = q2[5]
    for t in rng.integers(w.t0_ns + 3 * w.dt_ns, w.t0_ns + (w.n_knots - 3) * w.dt_ns - 1, 50):
        d = emu.emu_two_stage_maxdiff(C.c_int64(w.t0_ns), C.c_int64(w.dt_ns), C.c_int(w.n_knots), dp(q2), dp(p),
                                      C.c_int64(int(t)))
        assert 0 <= d < 1e-12, d
