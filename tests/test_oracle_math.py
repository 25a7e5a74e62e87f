"""Self-validation of the CPU oracle's leaf math and spline evaluators (no GPU).

The reference ships no tests or golden vectors, so the oracle is
pinned by: group identities, finite differences through the right perturbation
q <- q*exp(d) (ceres_local_param.h:137-145), agreement of the factor `*View`
evaluators with the reference's independent plain-spline evaluators
(so3_spline.h:240-367), and agreement with the numpy generator's third
implementation.
"""
import ctypes as C

import numpy as np
import pytest

from helpers import pkg, qexp, qlog, qmul, rot_angle_between, small_window, syn


def so3_probe(lib, kind, vin, nout):
    vin = np.ascontiguousarray(vin, float)
    out = np.zeros(nout)
    lib.raw("probe_so3")(C.c_int32(kind), vin.ctypes.data_as(C.POINTER(C.c_double)),
                         out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


@pytest.mark.parametrize("scale", [1e-12, 1e-6, 1e-2, 0.5, 2.0, 3.0])
def test_exp_log_roundtrip(oracle_lib, scale):
    rng = np.random.default_rng(1)
    for _ in range(20):
        w = rng.normal(size=3)
        w *= scale / np.linalg.norm(w)
        q = so3_probe(oracle_lib, 0, w, 4)
        assert abs(np.linalg.norm(q) - 1) < 1e-14
        w2 = so3_probe(oracle_lib, 1, q, 3)
        assert np.allclose(w, w2, rtol=1e-11, atol=1e-15)
        assert np.allclose(q, qexp(w[None])[0], atol=1e-15)


@pytest.mark.parametrize("scale", [1e-7, 1e-3, 0.3, 2.5])
def test_right_jacobians(oracle_lib, scale):
    rng = np.random.default_rng(2)
    w = rng.normal(size=3)
    w *= scale / np.linalg.norm(w)
    Jr = so3_probe(oracle_lib, 2, w, 9).reshape(3, 3)
    Jri = so3_probe(oracle_lib, 3, w, 9).reshape(3, 3)
    assert np.allclose(Jr @ Jri, np.eye(3), atol=1e-9)
    # exp(w + e) ~ exp(w) exp(Jr e)
    h = 1e-6
    fd = np.zeros((3, 3))
    for k in range(3):
        e = np.zeros(3); e[k] = h
        qp = so3_probe(oracle_lib, 0, w + e, 4); qm = so3_probe(oracle_lib, 0, w - e, 4)
        q0 = so3_probe(oracle_lib, 0, w, 4)
        dp = qlog(qmul(syn.qconj(q0)[None], qp[None]))[0]; dm = qlog(qmul(syn.qconj(q0)[None], qm[None]))[0]
        fd[:, k] = (dp - dm) / (2 * h)
    assert np.allclose(fd, Jr, atol=1e-7)


def test_taylor_branch_continuity(oracle_lib):
    # values just below / above the 1e-10 thresholds agree to first order
    w = np.array([1.0, -2.0, 0.5]); w /= np.linalg.norm(w)
    for kind, n in ((2, 9), (3, 9)):
        a = so3_probe(oracle_lib, kind, w * 0.99e-5, n)   # phi_norm2 just under 1e-10
        b = so3_probe(oracle_lib, kind, w * 1.01e-5, n)
        assert np.allclose(a, b, atol=1e-6)


def test_group_ops_against_matrices(oracle_lib):
    rng = np.random.default_rng(3)
    qa = qexp(rng.normal(size=(1, 3)))[0]; qb = qexp(rng.normal(size=(1, 3)))[0]
    Ra = so3_probe(oracle_lib, 6, qa, 9).reshape(3, 3); Rb = so3_probe(oracle_lib, 6, qb, 9).reshape(3, 3)
    qab = so3_probe(oracle_lib, 4, np.concatenate([qa, qb]), 4)
    Rab = so3_probe(oracle_lib, 6, qab, 9).reshape(3, 3)
    assert np.allclose(Rab, Ra @ Rb, atol=1e-14)
    assert np.allclose(Ra @ Ra.T, np.eye(3), atol=1e-14)
    v = rng.normal(size=3)
    assert np.allclose(so3_probe(oracle_lib, 5, np.concatenate([qa, v]), 3), Ra @ v, atol=1e-14)


@pytest.fixture(scope="module")
def est(oracle_lib):
    w = small_window(seed=11, n_knots=9)
    e = pkg.setup_estimator(oracle_lib, w, state="init")
    return e, w


def probe_view(lib, est, kind, t):
    val = np.zeros(4); J = np.zeros(36)
    s = lib.raw("probe_so3_view")(est.h, C.c_int32(kind), C.c_int64(int(t)), val.ctypes.data_as(C.POINTER(C.c_double)),
                                  J.ctypes.data_as(C.POINTER(C.c_double)))
    return val, J.reshape(4, 3, 3), s


def test_views_match_plain_spline_and_numpy(oracle_lib, est):
    e, w = est
    ts = np.linspace(w.t0_ns + 1, w.t0_ns + (w.n_knots - 3) * w.dt_ns - 1, 37).astype(np.int64)
    qn, pn = syn.spline_pose(w.q0, w.p0, ts, w.t0_ns, w.dt_ns)
    wn, an, vn = syn.spline_imu(w.q0, w.p0, ts, w.t0_ns, w.dt_ns)
    q_o, p_o, w_o, v_o, a_o = e.QueryTrajectory(ts)
    for n, t in enumerate(ts):
        q = np.zeros(4); om = np.zeros(3); al = np.zeros(3)
        oracle_lib.raw("probe_plain_so3")(e.h, C.c_int64(int(t)), q.ctypes.data_as(C.POINTER(C.c_double)),
                                          om.ctypes.data_as(C.POINTER(C.c_double)),
                                          al.ctypes.data_as(C.POINTER(C.c_double)))
        rp, _, _ = probe_view(oracle_lib, e, 0, t)
        rtp, _, _ = probe_view(oracle_lib, e, 1, t)
        rot, _, _ = probe_view(oracle_lib, e, 3, t)
        vb, _, _ = probe_view(oracle_lib, e, 2, t)
        assert rot_angle_between(q[None], rp[None])[0] < 1e-12
        assert rot_angle_between(q[None], rot[None])[0] < 1e-12
        assert rot_angle_between(q[None], syn.qconj(rtp)[None])[0] < 1e-12
        assert np.allclose(om, vb[:3], atol=1e-11)
        assert rot_angle_between(q[None], qn[n][None])[0] < 1e-12
        assert np.allclose(om, wn[n], atol=1e-10)
        assert rot_angle_between(q_o[n][None], q[None])[0] < 1e-12
    assert np.allclose(p_o, pn, atol=1e-12)
    assert np.allclose(v_o, vn, atol=1e-10)
    assert np.allclose(a_o, an, atol=1e-8)
    assert np.allclose(w_o, wn, atol=1e-10)


def _fd_rot(fn, q_knots, s, h=1e-6, left=False):
    """central FD of fn(q_knots) -> vec3 w.r.t. the right (q*exp(d), the estimator's local
    parameterisation) or left (exp(d)*q) perturbation of knots s..s+3."""
    out = np.zeros((4, 3, 3))
    for k in range(4):
        for c in range(3):
            d = np.zeros(3); d[c] = h
            qp = q_knots.copy(); qm = q_knots.copy()
            if left:
                qp[s + k] = qmul(qexp(d[None])[0], qp[s + k]); qm[s + k] = qmul(qexp(-d[None])[0], qm[s + k])
            else:
                qp[s + k] = qmul(qp[s + k], qexp(d[None])[0]); qm[s + k] = qmul(qm[s + k], qexp(-d[None])[0])
            out[k, :, c] = (fn(qp) - fn(qm)) / (2 * h)
    return out


def test_view_jacobians_finite_difference(oracle_lib, est):
    e, w = est
    rng = np.random.default_rng(5)
    q0 = w.q0.copy(); p0 = w.p0.copy()
    ts = rng.integers(w.t0_ns + 1, w.t0_ns + (w.n_knots - 3) * w.dt_ns - 1, 6)
    v = np.array([0.3, -1.2, 0.7])
    for t in ts:
        e.SetKnots(q0, p0)
        R, J, s = probe_view(oracle_lib, e, 0, t)       # EvaluateRp
        RT, JT, _ = probe_view(oracle_lib, e, 1, t)     # EvaluateRTp
        om, Jw, _ = probe_view(oracle_lib, e, 2, t)     # VelocityBody
        Rm = syn.qrot(R[None], np.eye(3)).T             # columns = R e_k  -> R matrix
        hatv = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])

        def f_rp(qk):
            e.SetKnots(qk, p0)
            r, _, _ = probe_view(oracle_lib, e, 0, t)
            return syn.qrot(r[None], v[None])[0]

        def f_rtp(qk):
            e.SetKnots(qk, p0)
            r, _, _ = probe_view(oracle_lib, e, 1, t)
            return syn.qrot(r[None], v[None])[0]

        def f_w(qk):
            e.SetKnots(qk, p0)
            r, _, _ = probe_view(oracle_lib, e, 2, t)
            return r[:3].copy()

        fd_rp = _fd_rot(f_rp, q0, s)
        fd_rtp = _fd_rot(f_rtp, q0, s)
        # VelocityBody's Jacobian (never requested by the factors, image_feature_factor.h:112,138) is
        # w.r.t. the LEFT perturbation: Jr_delta_inv *= p1.inverse().matrix() (so3_spline_view.h:388-389)
        fd_w = _fd_rot(f_w, q0, s, left=True)
        for k in range(4):
            # d(R v)/d delta_k = -R hat(v) J_k ; d(R^T v)/d delta_k = R^T hat(v) J_k   (Appendix A)
            assert np.allclose(fd_rp[k], -Rm @ hatv @ J[k], atol=2e-7)
            assert np.allclose(fd_rtp[k], Rm.T @ hatv @ JT[k], atol=2e-7)
            assert np.allclose(fd_w[k], Jw[k], atol=2e-5)  # omega ~ O(1) * inv_dt, FD noise scales with 20
    e.SetKnots(q0, p0)


def test_split_view_jacobians_finite_difference(oracle_lib, est):
    e, w = est
    q0 = w.q0.copy(); p0 = w.p0.copy()
    t = int(w.t0_ns + 2.37 * w.dt_ns)

    def split(qk, pk):
        e.SetKnots(qk, pk)
        g = np.zeros(3); a = np.zeros(3); Jw = np.zeros(36); Ja = np.zeros(36); Jp = np.zeros(4)
        dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
        s = oracle_lib.raw("probe_split")(e.h, C.c_int64(t), dp(g), dp(a), dp(Jw), dp(Ja), dp(Jp))
        return g, a, Jw.reshape(4, 3, 3), Ja.reshape(4, 3, 3), Jp, s

    g, a, Jw, Ja, Jp, s = split(q0, p0)
    fd_w = _fd_rot(lambda qk: split(qk, p0)[0], q0, s)
    fd_a = _fd_rot(lambda qk: split(qk, p0)[1], q0, s)
    for k in range(4):
        assert np.allclose(fd_w[k], Jw[k], atol=2e-5)
        assert np.allclose(fd_a[k], Ja[k], atol=2e-5)
    # d accel / d P_k = lambda_a[k] * R^T
    h = 1e-4
    q_t, *_ = e.QueryTrajectory(np.array([t]))
    Rm = syn.qrot(q_t[0][None], np.eye(3)).T
    for k in range(4):
        for c in range(3):
            pp = p0.copy(); pp[s + k, c] += h
            pm = p0.copy(); pm[s + k, c] -= h
            fd = (split(q0, pp)[1] - split(q0, pm)[1]) / (2 * h)
            assert np.allclose(fd, Jp[k] * Rm.T[:, c], atol=1e-5)
    e.SetKnots(q0, p0)
