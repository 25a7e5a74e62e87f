"""Generates tests/golden/small_window.npz from the ORACLE (oracle/liboracle.so).

The reference cannot run here (Eigen / Ceres / ROS absent) and ships no vectors, so these are NOT
reference outputs: they freeze the oracle's answers on a small rolling-shutter window (factor residuals + Jacobians,
normal equations, an LM solve, a marginalization prior as J'J / J'r, trajectory queries) so that
  * a change of the oracle that alters its arithmetic is caught on CPU (tests/test_golden.py, -m "not gpu"),
  * the CUDA path can be checked against committed numbers (-m gpu) independently of the oracle build.
Run from the repo root:  python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("ctrl-vio_b200")
syn = pkg.synthetic
from helpers import small_window  # noqa: E402


def golden_cases(lib):
    """Everything the golden file holds, computed through `lib` (oracle or CUDA engine)."""
    out = {}
    w = small_window(fix_ld=False)
    opt = pkg.make_options(fix_ld=False, ld_lower=0.0, ld_upper=syn.LD_UPPER)
    e = pkg.setup_estimator(lib, w, options=opt)
    r, s, J, cost = e.EvalImageFactors(True, 2.0)
    out.update(img_r=r, img_s=s, img_J=J, img_cost=cost)
    r, s, J, cost = e.EvalImuFactors(True)
    out.update(imu_r=r, imu_s=s, imu_J=J, imu_cost=cost)
    H, g, hl, gl, cost = e.NormalEquations()
    out.update(ne_H=np.triu(H), ne_g=g, ne_hl=hl, ne_gl=gl, ne_cost=cost)
    t = w.t0_ns + np.array([10_000_000, 55_500_000, 120_000_001, 170_999_999], np.int64)
    q, p, om, v, a = e.QueryTrajectory(t)
    out.update(query_t=t, query_q=q, query_p=p, query_omega=om, query_vel=v, query_acc=a)
    summ = e.Solve(10)
    q, p = e.GetKnots()
    out.update(solve_iterations=summ.iterations, solve_initial_cost=summ.initial_cost, solve_final_cost=summ.final_cost,
               solve_q=q, solve_p=p, solve_bias=e.GetBiases(), solve_rho=e.GetInvDepths(), solve_ld=e.GetLineDelay())
    # marginalization of the oldest keyframe at the solved state
    nowk = int((w.kf_times[0] - w.t0_ns) // w.dt_ns)
    later = int((w.kf_times[1] - w.t0_ns) // w.dt_ns)
    img_marg = (w.anchor_frame[w.lm] == 0).astype(np.int32)
    imu_marg = (w.imu_t < w.kf_times[1]).astype(np.int32)
    bias_marg = np.zeros(len(w.bf_i), np.int32); bias_marg[0] = 1
    em = pkg.setup_estimator(lib, w, image_marg=img_marg, imu_marg=imu_marg, bias_marg=bias_marg,
                             options=pkg.make_options(fix_ld=False, ld_lower=0.0, ld_upper=syn.LD_UPPER, is_marg_state=True,
                                                      ctrl_to_be_opt_now=nowk, ctrl_to_be_opt_later=later))
    em.SetKnots(out["solve_q"], out["solve_p"]); em.SetBiases(out["solve_bias"]); em.SetInvDepths(out["solve_rho"])
    em.SetLineDelay(out["solve_ld"])
    pr = em.SaveMarginalizationInfo()
    out.update(prior_n=pr.n, prior_JtJ=pr.J.T @ pr.J, prior_Jtr=pr.J.T @ pr.r, prior_blk_type=pr.blk_type,
               prior_blk_index=pr.blk_index, prior_blk_col=pr.blk_col, prior_x0=pr.blk_x0)
    return out


if __name__ == "__main__":
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    lib = pkg.CtvioLib(so, "ctvo_", optional=pkg.binding.DEVICE_ONLY_SYMBOLS)
    data = golden_cases(lib)
    path = os.path.join(ROOT, "tests", "golden", "small_window.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), "bytes,", len(data), "arrays")
