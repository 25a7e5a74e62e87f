"""debug: wall-clock split of the e2e path (host buffers through the C-ABI) for the C2 window."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("ctrl-vio_b200"); syn = pkg.synthetic
w = syn.config_c2()
e = pkg.Estimator(pkg.load(), pkg.make_config(device=0, **w.config_kwargs()))
e.SetOptions(pkg.make_options(fix_ld=w.fix_ld, ld_lower=w.ld_lower, ld_upper=w.ld_upper))
acc = {}
def tick(name, fn):
    t = time.perf_counter(); r = fn(); acc.setdefault(name, []).append(time.perf_counter() - t); return r
for it in range(12):
    tick("set_state", lambda: (e.SetKnots(w.q0, w.p0), e.SetBiases(w.bias0), e.SetInvDepths(w.rho0), e.SetLineDelay(w.ld0)))
    tick("clear", e.ClearFactors)
    tick("add_image", lambda: e.AddImageFeatureDelayAnalytic(w.ti, w.rowi, w.pi, w.tj, w.rowj, w.pj, w.lm))
    tick("add_imu_bias", lambda: (e.AddIMUMeasurementAnalytic(w.imu_t, w.imu_gyro, w.imu_accel, w.imu_node), e.AddBiasFactor(w.bf_i, w.bf_j, w.bf_sqrt_info)))
    s = tick("solve", lambda: e.Solve(15))
    acc.setdefault("solve_device", []).append(s.device_ms * 1e-3)
    tick("get_state", lambda: (e.GetKnots(), e.GetBiases(), e.GetInvDepths(), e.GetLineDelay()))
for k, v in acc.items():
    print(f"{k:14s} {1e3*np.median(v[2:]):7.3f} ms")
