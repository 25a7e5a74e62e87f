"""The step decision that gradient_norm_kernel takes on the device (kernels_linear.cu: LmDecision) restated in numpy
against the oracle's formula (Ceres 1.14 trust_region_minimizer.cc / trust_region_strategy: the oracle uses std::pow).
The device forms (2 rho - 1)^3 with two multiplications instead of pow(): the next radius can differ by a couple of ulp,
never more - far below anything the damping (diag / radius) can resolve.  No GPU needed."""
import numpy as np


def radius_pow(radius, rho, max_radius=1e16):
    return min(max_radius, radius / max(1.0 / 3.0, 1.0 - np.power(2.0 * rho - 1.0, 3)))


def radius_cube(radius, rho, max_radius=1e16):
    t = 2.0 * rho - 1.0
    return min(max_radius, radius / max(1.0 / 3.0, 1.0 - (t * t) * t))


def test_cube_and_pow_radius_updates_agree_to_a_few_ulp():
    rng = np.random.default_rng(11)
    worst = 0.0
    for _ in range(20000):
        rho = rng.uniform(1e-3, 1.5) if rng.random() < 0.8 else 10.0 ** rng.uniform(-3, 2)
        radius = 10.0 ** rng.uniform(-3, 12)
        a, b = radius_pow(radius, rho), radius_cube(radius, rho)
        worst = max(worst, abs(a - b) / np.spacing(a))
    assert worst <= 4.0, worst


def test_decision_thresholds():
    # accept iff the model decreases, the linear solve succeeded and rho > min_relative_decrease (1e-3)
    def decide(x_cost, cand_cost, gd, dHd, chol_fail=False):
        mcc = -gd - 0.5 * dHd
        valid = (not chol_fail) and np.isfinite(mcc) and mcc > 0.0
        rho = (x_cost - cand_cost) / mcc if valid else float("nan")
        return valid, bool(valid and rho > 1e-3)
    assert decide(10.0, 9.0, -2.0, 1.0) == (True, True)       # rho = 1 / 1.5
    assert decide(10.0, 9.9999, -2.0, 1.0) == (True, False)   # rho = 6.7e-5
    assert decide(10.0, 11.0, -2.0, 1.0) == (True, False)     # cost went up
    assert decide(10.0, 9.0, 2.0, 1.0) == (False, False)      # model does not decrease: invalid step
    assert decide(10.0, 9.0, -2.0, 1.0, chol_fail=True) == (False, False)
