"""Self-validation of the CPU oracle's factors, normal equations, LM loop and marginalization
(no GPU).  Checks (ii)-(iv) of SURVEY.md §8c:
  (ii)  finite-difference Jacobians of the image / IMU factors through the C-ABI,
  (iii) Schur-reduced LM step == full-system LM step (dense numpy restatement),
  (iv)  marginalization J'J, J'r == dense Schur complement of the recorded factors.
"""
import numpy as np
import pytest

from helpers import dense_jacobian, get_state, perturb_state, pkg, qexp, qmul, rot_angle_between, small_window, syn


@pytest.fixture(scope="module")
def small(oracle_lib):
    w = small_window(seed=21, n_knots=8, n_kf=5, per_frame=5, fix_ld=False)
    e = pkg.setup_estimator(oracle_lib, w, state="init")
    e.SetLineDelay(20e-6)
    return e, w


def test_image_factor_jacobians_fd(small):
    e, w = small
    base = get_state(e)
    r0, s0, J, _ = e.EvalImageFactors(True, 0.0)
    n_check = min(6, e.n_img)
    h = 1e-6
    for n in range(n_check):
        for side in range(2):
            for k in range(4):
                knot = int(s0[n, side] + k)
                blk = J[n, (side * 4 + k) * 12:(side * 4 + k) * 12 + 12]
                # when both sides touch the same knot the FD sees the SUM of both contributions
                tot_rot = np.zeros((2, 3)); tot_pos = np.zeros((2, 3))
                for s2 in range(2):
                    kk = knot - int(s0[n, s2])
                    if 0 <= kk <= 3:
                        b2 = J[n, (s2 * 4 + kk) * 12:(s2 * 4 + kk) * 12 + 12]
                        tot_rot += b2[:6].reshape(2, 3); tot_pos += b2[6:].reshape(2, 3)
                for c in range(3):
                    d = np.zeros(3); d[c] = h
                    perturb_state(e, w, knot=knot, rot=d, base=base); rp = e.EvalImageFactors(False, 0.0)[0][n]
                    perturb_state(e, w, knot=knot, rot=-d, base=base); rm = e.EvalImageFactors(False, 0.0)[0][n]
                    assert np.allclose((rp - rm) / (2 * h), tot_rot[:, c], rtol=1e-5, atol=2e-4), (n, side, k, c)
                    perturb_state(e, w, knot=knot, pos=d, base=base); rp = e.EvalImageFactors(False, 0.0)[0][n]
                    perturb_state(e, w, knot=knot, pos=-d, base=base); rm = e.EvalImageFactors(False, 0.0)[0][n]
                    assert np.allclose((rp - rm) / (2 * h), tot_pos[:, c], rtol=1e-5, atol=2e-4)
        l = int(w.lm[n])
        perturb_state(e, w, rho=(l, h), base=base); rp = e.EvalImageFactors(False, 0.0)[0][n]
        perturb_state(e, w, rho=(l, -h), base=base); rm = e.EvalImageFactors(False, 0.0)[0][n]
        assert np.allclose((rp - rm) / (2 * h), J[n, 96:98], rtol=1e-5, atol=2e-4)
        # line delay enters through int64 truncation to ns (image_feature_factor.h:72): use a step that is
        # an exact number of ns so the finite difference sees the intended +-200 ns
        hl = 200e-9
        perturb_state(e, w, ld=hl, base=base); rp = e.EvalImageFactors(False, 0.0)[0][n]
        perturb_state(e, w, ld=-hl, base=base); rm = e.EvalImageFactors(False, 0.0)[0][n]
        assert np.allclose((rp - rm) / (2 * hl), J[n, 98:100], rtol=2e-4, atol=5.0), (n, (rp - rm) / (2 * hl), J[n, 98:100])
    perturb_state(e, w, base=base)


def test_imu_factor_jacobians_fd(small):
    e, w = small
    base = get_state(e)
    r0, s0, J, _ = e.EvalImuFactors(True)
    h = 1e-6
    for n in (0, e.n_imu // 2, e.n_imu - 1):
        for k in range(4):
            knot = int(s0[n] + k)
            for c in range(3):
                d = np.zeros(3); d[c] = h
                perturb_state(e, w, knot=knot, rot=d, base=base); rp = e.EvalImuFactors(False)[0][n]
                perturb_state(e, w, knot=knot, rot=-d, base=base); rm = e.EvalImuFactors(False)[0][n]
                assert np.allclose((rp - rm) / (2 * h), J[n, k * 36:k * 36 + 18].reshape(6, 3)[:, c], rtol=1e-5, atol=5e-3)
                perturb_state(e, w, knot=knot, pos=d, base=base); rp = e.EvalImuFactors(False)[0][n]
                perturb_state(e, w, knot=knot, pos=-d, base=base); rm = e.EvalImuFactors(False)[0][n]
                assert np.allclose((rp - rm) / (2 * h), J[n, k * 36 + 18:k * 36 + 36].reshape(6, 3)[:, c], rtol=1e-5, atol=5e-3)
        node = int(w.imu_node[n])
        for c in range(6):
            d = np.zeros(6); d[c] = h
            perturb_state(e, w, bias=(node, d), base=base); rp = e.EvalImuFactors(False)[0][n]
            perturb_state(e, w, bias=(node, -d), base=base); rm = e.EvalImuFactors(False)[0][n]
            want = np.zeros(6); want[c] = J[n, 144 + c] if c < 3 else J[n, 150 + c]
            assert np.allclose((rp - rm) / (2 * h), want, atol=1e-5)
    perturb_state(e, w, base=base)


def test_cauchy_corrector(small):
    e, w = small
    r_raw, _, J_raw, c_raw = e.EvalImageFactors(True, 0.0)
    r_c, _, J_c, c_c = e.EvalImageFactors(True, 2.0)
    s = (r_raw ** 2).sum(1)
    rho1 = 1.0 / (1.0 + s / 4.0)
    assert np.allclose(r_c, r_raw * np.sqrt(rho1)[:, None], rtol=1e-13)
    assert np.allclose(J_c, J_raw * np.sqrt(rho1)[:, None], rtol=1e-13)
    assert np.isclose(c_c, 0.5 * (4.0 * np.log1p(s / 4.0)).sum(), rtol=1e-13)
    assert np.isclose(c_raw, 0.5 * s.sum(), rtol=1e-13)


def test_normal_equations_match_dense_jacobian(small):
    e, w = small
    Jd, rd = dense_jacobian(e, w, cauchy=2.0)
    npd = e.np_dim
    # fix_ld False here, nothing constant
    H, g, hl, gl, cost = e.NormalEquations()
    Hd = Jd.T @ Jd; gd = Jd.T @ rd
    scale = np.abs(Hd).max()
    assert np.allclose(H, Hd[:npd, :npd], atol=1e-10 * scale)
    assert np.allclose(g, gd[:npd], atol=1e-10 * np.abs(gd).max())
    assert np.allclose(hl, np.diag(Hd)[npd:], rtol=1e-11)
    assert np.allclose(gl, gd[npd:], atol=1e-10 * np.abs(gd).max())
    assert np.isclose(cost, e.EvalCost(), rtol=1e-13)


def _numpy_lm_step(Jd, rd, radius, npd):
    """One full-system (no Schur) LM step with Ceres' Jacobi scaling and diagonal clamping."""
    s = 1.0 / (1.0 + np.sqrt((Jd ** 2).sum(0)))
    Js = Jd * s
    A = Js.T @ Js
    D = np.clip(np.diag(A), 1e-6, 1e32) / radius
    y = np.linalg.solve(A + np.diag(D), Js.T @ rd)
    return -s * y


def test_schur_step_equals_full_system_step(oracle_lib):
    w = small_window(seed=33, n_knots=8, n_kf=5, per_frame=5, fix_ld=True)
    e = pkg.setup_estimator(oracle_lib, w, state="init")
    base = get_state(e)
    Jd, rd = dense_jacobian(e, w, cauchy=2.0)
    npd = e.np_dim
    Jd[:, npd - 1] = 0.0  # line delay constant
    delta = _numpy_lm_step(Jd, rd, 1e4, npd)
    summ = e.Solve(1)
    assert summ.iterations == 1 and summ.num_successful_steps == 2
    q1, p1, b1, r1, _ = get_state(e)
    q0, p0, b0, r0, _ = base
    nK = e.n_knots
    for k in range(nK):
        qe = qmul(q0[k], qexp(delta[6 * k:6 * k + 3][None])[0])
        assert rot_angle_between(qe[None], q1[k][None])[0] < 1e-9
        assert np.allclose(p1[k], p0[k] + delta[6 * k + 3:6 * k + 6], atol=1e-9)
    assert np.allclose(b1.ravel(), b0.ravel() + delta[6 * nK:6 * nK + 6 * e.n_bias], atol=1e-9)
    assert np.allclose(r1, r0 + delta[npd:], atol=1e-8)


@pytest.mark.parametrize("cfg", ["c1", "c2"])
def test_solve_converges_towards_truth(oracle_lib, cfg):
    w = syn.config_c1() if cfg == "c1" else syn.config_c2()
    e = pkg.setup_estimator(oracle_lib, w, state="init")
    c0 = e.EvalCost()
    c_gt = pkg.setup_estimator(oracle_lib, w, state="gt").EvalCost()  # noise floor of the synthetic data
    s = e.Solve(15)
    assert s.final_cost < c0 and s.final_cost < 1.02 * c_gt
    assert np.isclose(s.initial_cost, c0, rtol=1e-12)
    assert np.isclose(e.EvalCost(), s.final_cost, rtol=1e-12)
    assert s.iterations <= 15 and s.num_successful_steps >= 3
    if cfg == "c2":
        q, p = e.GetKnots()
        # IMU + vision: interior knots get pulled towards truth (gauge is unobservable by ~cm, so compare
        # relative motion between two interior knots)
        d_est = p[20] - p[8]; d_gt = w.p_gt[20] - w.p_gt[8]; d_0 = w.p0[20] - w.p0[8]
        assert np.linalg.norm(d_est - d_gt) < np.linalg.norm(d_0 - d_gt)
        b = e.GetBiases()
        assert np.abs(b[:, :3] - w.bias_gt[:, :3]).max() < 5e-3


def test_line_delay_bounds_and_line_search(oracle_lib):
    w = syn.config_c2(fix_ld=False)
    e = pkg.setup_estimator(oracle_lib, w, state="init")
    s = e.Solve(15)
    ld = e.GetLineDelay()
    assert 0.0 <= ld <= syn.LD_UPPER
    assert abs(ld - syn.LD_TRUE) < 6e-6, ld
    # every step of a bounds-constrained solve runs the Armijo search: >= 1 gradient pass per valid step
    assert s.num_jacobian_evals >= s.iterations + s.num_successful_steps - 1


def test_fixed_knots_and_locked_biases_imu_only(oracle_lib):
    """InitTrajectory-style problem (trajectory_manager.cpp:288-315): IMU only, biases locked,
    knots <= fixed index constant."""
    w = syn.config_c2()
    opt = pkg.make_options(fixed_knot_index=24, lock_wb=True, lock_ab=True, fix_ld=True)
    e = pkg.Estimator(oracle_lib, pkg.make_config(**w.config_kwargs()))
    e.SetOptions(opt)
    q0 = w.q_gt.copy(); p0 = w.p_gt.copy()
    q0[25:] = q0[24]; p0[25:] = p0[24]  # appended knots are copies of the last one (SURVEY C-16)
    e.SetKnots(q0, p0); e.SetBiases(w.bias_gt); e.SetInvDepths(w.rho_gt); e.SetLineDelay(w.ld_gt)
    tmin = w.t0_ns + 22 * w.dt_ns
    m = w.imu_t >= tmin
    e.AddIMUMeasurementAnalytic(w.imu_t[m], w.imu_gyro[m], w.imu_accel[m], w.imu_node[m])
    s = e.Solve(8)
    q, p = e.GetKnots()
    assert np.array_equal(q[:25], q0[:25]) and np.array_equal(p[:25], p0[:25])
    assert np.array_equal(e.GetBiases(), w.bias_gt)
    assert s.final_cost < s.initial_cost
    assert np.linalg.norm(p[25:] - w.p_gt[25:]) < np.linalg.norm(p0[25:] - w.p_gt[25:])


def _marg_setup(lib, with_prior=None):
    seq = syn.config_c3_sequence()
    wa = syn.subwindow(seq, 0, 10)
    later = int((wa.kf_times[1] - wa.t0_ns) // wa.dt_ns)
    nowk = int((wa.kf_times[0] - wa.t0_ns) // wa.dt_ns)
    img_marg = (wa.anchor_frame[wa.lm] == 0).astype(np.int32)
    imu_marg = (wa.imu_t < wa.kf_times[1]).astype(np.int32)
    bias_marg = np.zeros(len(wa.bf_i), np.int32); bias_marg[0] = 1
    opt = pkg.make_options(fix_ld=False, ld_lower=0.0, ld_upper=syn.LD_UPPER, is_marg_state=True,
                           ctrl_to_be_opt_now=nowk, ctrl_to_be_opt_later=later)
    e = pkg.setup_estimator(lib, wa, state="init", image_marg=img_marg, imu_marg=imu_marg, bias_marg=bias_marg,
                            options=opt)
    return e, wa, (img_marg, imu_marg, bias_marg, nowk, later)


def test_marginalization_equals_dense_schur(oracle_lib):
    e, wa, (img_marg, imu_marg, bias_marg, nowk, later) = _marg_setup(oracle_lib)
    e.Solve(4)
    pr = e.SaveMarginalizationInfo()
    assert pr is not None and pr.n > 0
    # dense restatement from the factor probes at the same state, Cauchy scale 1 for image factors
    Jd, rd = dense_jacobian(e, wa, cauchy=1.0)
    npd, nL = e.np_dim, e.n_lm
    rows = []
    r_i = 0
    keep_rows = np.zeros(Jd.shape[0], bool)
    keep_rows[:2 * e.n_img] = np.repeat(img_marg.astype(bool), 2)
    o = 2 * e.n_img
    keep_rows[o:o + 6 * e.n_imu] = np.repeat(imu_marg.astype(bool), 6)
    o += 6 * e.n_imu
    keep_rows[o:o + 6 * len(wa.bf_i)] = np.repeat(bias_marg.astype(bool), 6)
    J = Jd[keep_rows]; r = rd[keep_rows]
    A = J.T @ J; b = J.T @ r
    # dropped: knots < later, bias node 0, inverse depth of kf-0 landmarks
    drop = np.zeros(npd + nL, bool)
    drop[:6 * later] = True
    drop[6 * e.n_knots:6 * e.n_knots + 6] = True
    drop[npd + np.nonzero(wa.anchor_frame == 0)[0]] = True
    used = (np.abs(J).sum(0) > 0)
    # kept columns in the oracle's deterministic order: the prior's own block list
    col_of = {}
    for t, i, c in zip(pr.blk_type, pr.blk_index, pr.blk_col):
        base = {0: 6 * i, 1: 6 * i + 3, 2: 6 * e.n_knots + 6 * i, 3: 6 * e.n_knots + 6 * i + 3, 4: npd - 1}[int(t)]
        for d in range(1 if t == 4 else 3):
            col_of[base + d] = c + d
    keep_idx = np.array(sorted(col_of, key=lambda g: col_of[g]))
    assert not drop[keep_idx].any()
    assert set(np.nonzero(used & ~drop)[0]).issubset(set(keep_idx))
    di = np.nonzero(drop & used)[0]
    Amm = A[np.ix_(di, di)]; Amr = A[np.ix_(di, keep_idx)]; Arr = A[np.ix_(keep_idx, keep_idx)]
    Ainv = np.linalg.pinv(0.5 * (Amm + Amm.T), rcond=1e-15, hermitian=True)
    Ap = Arr - Amr.T @ Ainv @ Amr
    bp = b[keep_idx] - Amr.T @ Ainv @ b[di]
    JtJ = pr.J.T @ pr.J; Jtr = pr.J.T @ pr.r
    sc = np.abs(Ap).max()
    assert np.allclose(JtJ, Ap, atol=1e-7 * sc), np.abs(JtJ - Ap).max() / sc
    assert np.allclose(Jtr, bp, atol=1e-7 * np.abs(bp).max())
    # the line delay is a kept block (SURVEY C-13)
    assert (pr.blk_type == pkg.BLK_LD).sum() == 1
    assert pr.n == len(keep_idx)


def test_prior_factor_roundtrip(oracle_lib):
    """A prior re-attached at its own linearisation point reproduces r_lin / J_lin'J_lin, and solving window B
    with the prior keeps the cost finite and decreasing."""
    e, wa, (_, _, _, nowk, later) = _marg_setup(oracle_lib)
    e.Solve(6)
    pr = e.SaveMarginalizationInfo()
    seq = syn.config_c3_sequence()
    wb = syn.subwindow(seq, 1, 11)
    opt = pkg.make_options(fix_ld=False, ld_lower=0.0, ld_upper=syn.LD_UPPER)
    eb = pkg.setup_estimator(oracle_lib, wb, state="init", options=opt)
    q, p = e.GetKnots()
    eb.SetKnots(q, p); eb.SetLineDelay(e.GetLineDelay())
    b = np.zeros((11, 6)); b[:10] = e.GetBiases()[1:]; b[10] = b[9]
    eb.SetBiases(b)
    # bias node indices shift by one keyframe in window B (index identity == position in the window)
    pr_b = pkg.PriorData(n=pr.n, J=pr.J, r=pr.r, blk_type=pr.blk_type.copy(), blk_index=pr.blk_index.copy(),
                         blk_col=pr.blk_col, blk_x0=pr.blk_x0)
    isb = (pr_b.blk_type == pkg.BLK_BG) | (pr_b.blk_type == pkg.BLK_BA)
    pr_b.blk_index[isb] -= 1
    c_without = eb.EvalCost()
    eb.AddMarginalizationFactor(pr_b)
    c_with = eb.EvalCost()
    assert np.isclose(c_with - c_without, 0.5 * (pr.r ** 2).sum(), rtol=1e-9, atol=1e-9)
    s = eb.Solve(15)
    assert s.final_cost < s.initial_cost
    assert 0 <= eb.GetLineDelay() <= syn.LD_UPPER


def test_set_time_origin_slides_the_knot_slice(oracle_lib):
    """ctvio_set_time_origin (streaming, BASELINE config 5): dropping the first k control points and moving the origin
    by k*dt leaves every trajectory query unchanged; an origin off the knot grid is rejected."""
    w = small_window(seed=5, n_knots=8, n_kf=5, per_frame=4)
    est = pkg.setup_estimator(oracle_lib, w)
    t = w.t0_ns + np.array([1.2, 2.5, 3.7, 4.9], float) * w.dt_ns
    t = t.astype(np.int64) + 123
    ref = est.QueryTrajectory(t)
    k = 1
    est.SetTimeOrigin(w.t0_ns + k * w.dt_ns)
    est.SetKnots(w.q0[k:], w.p0[k:])
    got = est.QueryTrajectory(t)
    for a, b in zip(ref, got):
        assert np.allclose(a, b, rtol=0, atol=1e-12 * max(1.0, np.abs(a).max()))
    with pytest.raises(pkg.CtvioError):
        est.SetTimeOrigin(w.t0_ns + k * w.dt_ns + 7)
