"""Shared helpers for the parity tests (numpy only)."""
import ctypes as C
import importlib

import numpy as np

pkg = importlib.import_module("ctrl-vio_b200")
syn = pkg.synthetic


def qmul(a, b):
    return syn.qmul(np.asarray(a, float), np.asarray(b, float))


def qexp(w):
    return syn.qexp(np.asarray(w, float))


def qlog(q):
    return syn.qlog(np.asarray(q, float))


def rot_angle_between(qa, qb):
    """angle (rad) of qa^-1 * qb, per row."""
    d = syn.qmul(syn.qconj(qa), qb)
    return 2 * np.arctan2(np.linalg.norm(d[..., :3], axis=-1), np.abs(d[..., 3]))


def small_window(seed=7, n_knots=8, n_kf=5, per_frame=6, fix_ld=True, global_shutter=False, with_imu=True):
    """A small rolling-shutter window with IMU for exhaustive checks."""
    kf = syn.KF_OFFSET_NS + np.arange(n_kf, dtype=np.int64) * 40_000_000
    anchors = [per_frame, per_frame] + [0] * (n_kf - 2)
    return syn.make_window("small", n_knots, kf, anchors, n_kf, seed=seed, fix_ld=fix_ld,
                           global_shutter=global_shutter, with_imu=with_imu)


def perturb_state(est, w, *, knot=None, rot=None, pos=None, bias=None, rho=None, ld=None, base=None):
    """Set estimator state = base (+ a single right-multiplicative / additive perturbation)."""
    q, p, b, r, l = (np.array(x, float, copy=True) for x in base)
    if rot is not None:
        q[knot] = qmul(q[knot], qexp(np.asarray(rot, float)[None])[0])
    if pos is not None:
        p[knot] += pos
    if bias is not None:
        b[bias[0]] += bias[1]
    if rho is not None:
        r[rho[0]] += rho[1]
    if ld is not None:
        l = float(l) + ld
    est.SetKnots(q, p); est.SetBiases(b); est.SetInvDepths(r); est.SetLineDelay(float(l))


def get_state(est):
    q, p = est.GetKnots()
    return q, p, est.GetBiases(), est.GetInvDepths(), est.GetLineDelay()


def tangent_index(est):
    """(n_knots, n_bias) -> helper closures for camera-dim indices."""
    nK, nB = est.n_knots, est.n_bias
    return dict(knot=lambda k: 6 * k, bias=lambda b: 6 * nK + 6 * b, ld=6 * nK + 6 * nB, np=6 * nK + 6 * nB + 1)


def dense_jacobian(est, w, cauchy=2.0):
    """Full dense Jacobian / residual over [camera dims | landmarks] from the per-factor probes
    (image + IMU + bias factors; prior not included)."""
    ix = tangent_index(est)
    npd, nL = ix["np"], est.n_lm
    rows = []
    res = []
    ri, si, Ji, _ = est.EvalImageFactors(True, cauchy)
    for n in range(est.n_img):
        J = np.zeros((2, npd + nL))
        for side in range(2):
            for k in range(4):
                blk = Ji[n, (side * 4 + k) * 12:(side * 4 + k) * 12 + 12]
                g = 6 * (si[n, side] + k)
                J[:, g:g + 3] += blk[:6].reshape(2, 3)
                J[:, g + 3:g + 6] += blk[6:].reshape(2, 3)
        J[:, npd + w.lm[n]] += Ji[n, 96:98]
        J[:, ix["ld"]] += Ji[n, 98:100]
        rows.append(J); res.append(ri[n])
    rm, sm, Jm, _ = est.EvalImuFactors(True)
    for n in range(est.n_imu):
        J = np.zeros((6, npd + nL))
        for k in range(4):
            g = 6 * (sm[n] + k)
            J[:, g:g + 3] += Jm[n, k * 36:k * 36 + 18].reshape(6, 3)
            J[:, g + 3:g + 6] += Jm[n, k * 36 + 18:k * 36 + 36].reshape(6, 3)
        b = ix["bias"](int(w.imu_node[n]))
        J[0:3, b:b + 3] += np.diag(Jm[n, 144:147])
        J[3:6, b + 3:b + 6] += np.diag(Jm[n, 153:156])
        rows.append(J); res.append(rm[n])
    bias = est.GetBiases()
    for n in range(len(w.bf_i)):
        J = np.zeros((6, npd + nL))
        i, j = int(w.bf_i[n]), int(w.bf_j[n])
        s = w.bf_sqrt_info[n]
        J[:, ix["bias"](i):ix["bias"](i) + 6] = -np.diag(s)
        J[:, ix["bias"](j):ix["bias"](j) + 6] = np.diag(s)
        rows.append(J); res.append(s * (bias[j] - bias[i]))
    return np.concatenate(rows, 0), np.concatenate(res, 0)


def triangulation_case(seed=3, n_frames=11, n_lm=200, noise=1e-3, window_size=10):
    """Synthetic input of FeatureManager::triangulate: body poses of a window, camera extrinsics, landmarks observed
    over consecutive frames from their start frame.  Includes the reference's edge cases: tracks shorter than 2,
    start frames beyond the candidate range, already-initialised depths, a landmark BEHIND the camera (depth < 0.1 ->
    INIT_DEPTH) and a zero-parallax track."""
    rng = np.random.default_rng(seed)
    ang = np.linspace(0, 0.5, n_frames)
    Rs = np.stack([np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]]) for a in ang])
    Ps = np.stack([np.array([0.3 * k, 0.05 * np.sin(k), 0.02 * k]) for k in range(n_frames)]) + rng.normal(0, 0.01, (n_frames, 3))
    ric = syn.qrot(qexp(np.array([[0.02, -0.01, 0.03]])), np.eye(3)).T.copy()
    # camera looks along body +x: columns of ric = camera axes in the body frame
    base = np.array([[0, 0, 1.0], [-1, 0, 0], [0, -1, 0]])
    ric = ric @ base
    tic = np.array([0.007, -0.057, -0.042])
    start, offs, pts, truth = [], [0], [], []
    for l in range(n_lm):
        s = int(rng.integers(0, n_frames - 1))
        used = int(rng.integers(1, n_frames - s + 1))
        depth = float(rng.uniform(2.0, 20.0))
        if l % 17 == 0:
            depth = -3.0  # inconsistent track: triangulates behind the anchor camera
        R0 = Rs[s] @ ric; t0 = Ps[s] + Rs[s] @ tic
        bearing = np.array([rng.uniform(-0.4, 0.4), rng.uniform(-0.3, 0.3), 1.0])
        Pw = R0 @ (bearing * depth) + t0
        if l % 23 == 0:
            used = min(used, 3)
        for k in range(used):
            Rk = Rs[s + k] @ ric; tk = Ps[s + k] + Rs[s + k] @ tic
            pc = Rk.T @ (Pw - tk)
            if l % 29 == 0:
                pc = bearing * depth  # zero parallax: every frame sees the same bearing
            pts.append([pc[0] / pc[2] + rng.normal(0, noise), pc[1] / pc[2] + rng.normal(0, noise), 1.0])
        start.append(s); offs.append(offs[-1] + used); truth.append(depth)
    depth0 = np.full(n_lm, -1.0)
    depth0[::11] = 7.5  # already initialised: must be kept
    return dict(Rs=Rs.reshape(-1, 9), Ps=Ps, ric=ric.reshape(9), tic=tic, start_frame=np.array(start, np.int32),
                obs_offset=np.array(offs, np.int32), obs_point=np.array(pts), depth0=depth0, truth=np.array(truth),
                window_size=window_size)


def triangulate_numpy(c, init_depth=5.0):
    """numpy / LAPACK restatement of feature_manager.cpp:230-275 (independent of oracle/triangulate.hpp)."""
    Rs = c["Rs"].reshape(-1, 3, 3); Ps = c["Ps"]; ric = c["ric"].reshape(3, 3); tic = c["tic"]
    out = c["depth0"].copy()
    for l in range(len(out)):
        o0, o1 = c["obs_offset"][l], c["obs_offset"][l + 1]
        used, i = o1 - o0, c["start_frame"][l]
        if not (used >= 2 and i < c["window_size"] - 2) or out[l] > 0:
            continue
        R0 = Rs[i] @ ric; t0 = Ps[i] + Rs[i] @ tic
        A = np.zeros((2 * used, 4))
        for k in range(used):
            R1 = Rs[i + k] @ ric; t1 = Ps[i + k] + Rs[i + k] @ tic
            t = R0.T @ (t1 - t0); R = R0.T @ R1
            P = np.hstack([R.T, (-R.T @ t)[:, None]])
            f = c["obs_point"][o0 + k] / np.linalg.norm(c["obs_point"][o0 + k])
            A[2 * k] = f[0] * P[2] - f[2] * P[0]
            A[2 * k + 1] = f[1] * P[2] - f[2] * P[1]
        v = np.linalg.svd(A)[2][-1]
        d = v[2] / v[3] if v[3] != 0 else np.inf
        out[l] = d if (np.isfinite(d) and d >= 0.1) else init_depth
    return out


# ---- chained runs (solve -> marginalize -> next window) and the oracle's own summation-order sensitivity ----------

def c3_window_a(lib, perm_seed=None):
    import importlib
    return importlib.import_module("ctrl-vio_b200.streaming").c3_window_a(lib, perm_seed)


def run_c3_sequence(lib, perm_seed=None):
    """BASELINE config 3 chain; returns the quantities the parity / sensitivity assertions look at."""
    e, seq, wa, nowk = c3_window_a(lib, perm_seed)
    R0 = syn.qrot(wa.q0[nowk][None], np.eye(3)).T.copy(); t0 = wa.p0[nowk].copy()
    sa = e.Solve(15)
    e.GaugeRealign(nowk, R0, t0)
    pr = e.SaveMarginalizationInfo()
    assert pr is not None
    isb = (pr.blk_type == pkg.BLK_BG) | (pr.blk_type == pkg.BLK_BA)
    pr.blk_index[isb] -= 1   # bias node indices are window-relative: the window slides by one keyframe
    wb = syn.subwindow(seq, 1, 11)
    if perm_seed is not None:
        rng = np.random.default_rng(perm_seed + 100)
        pm = rng.permutation(wb.n_obs)
        for f in ("ti", "rowi", "pi", "tj", "rowj", "pj", "lm"):
            setattr(wb, f, np.ascontiguousarray(getattr(wb, f)[pm]))
    eb = pkg.setup_estimator(lib, wb, options=pkg.make_options(fix_ld=False, ld_lower=0.0, ld_upper=syn.LD_UPPER))
    q, p = e.GetKnots()
    b = np.zeros((11, 6)); b[:10] = e.GetBiases()[1:]; b[10] = b[9]
    rho = wb.rho0.copy()
    ra = e.GetInvDepths()
    ga = wa.meta["lm_global"]; gb = wb.meta["lm_global"]
    common = np.intersect1d(ga, gb)
    rho[np.searchsorted(gb, common)] = ra[np.searchsorted(ga, common)]
    eb.SetKnots(q, p); eb.SetBiases(b); eb.SetInvDepths(rho); eb.SetLineDelay(e.GetLineDelay())
    eb.AddMarginalizationFactor(pr)
    sb = eb.Solve(15)
    qb, pb = eb.GetKnots()
    return dict(iterations=[sa.iterations, sb.iterations], costs=[sa.final_cost, sb.final_cost],
                prior_consts=[0.0, 0.5 * float(np.dot(pr.r, pr.r))], q=qb, p=pb, ld=eb.GetLineDelay(), prior_n=pr.n)


def run_c5(lib, n_windows, second_new_every=0, perm_seed=None):
    import importlib
    st = importlib.import_module("ctrl-vio_b200.streaming")
    seq = st.config_c5_sequence(n_windows + 1)
    r = st.StreamingRunner(lib, seq, second_new_every=second_new_every, perm_seed=perm_seed)
    r.run(n_windows)
    return dict(iterations=[x["iterations"] for x in r.records], init_iterations=[x["init_iterations"] for x in r.records],
                costs=[x["final_cost"] for x in r.records], prior_consts=[x["prior_const"] for x in r.records], prior_dims=[x["prior_dim"] for x in r.records],
                marg_flags=[x["marg_flag"] for x in r.records], q=r.q[:r.ncp].copy(), p=r.p[:r.ncp].copy(), ld=r.ld,
                records=r.records)


def chain_difference(a, b):
    scale = max(np.abs(b["p"]).max(), 1e-12)
    # costs are compared with the prior's constant 0.5 |r_lin|^2 removed: along eigen-directions whose eigenvalue is
    # rounding noise the reference's r_lin = S^-1/2 V' b is noise / sqrt(noise) (marginalization_factor.cpp:255-263), a
    # constant offset of the cost that no two eigen-solvers agree on and that does not move the optimum
    ca = [x - c for x, c in zip(a["costs"], a["prior_consts"])]
    cb = [x - c for x, c in zip(b["costs"], b["prior_consts"])]
    return dict(cost_rel=max(abs(x - y) / abs(z) for x, y, z in zip(ca, cb, b["costs"])),
                cost_raw_rel=max(abs(x - y) / abs(y) for x, y in zip(a["costs"], b["costs"])),
                trans_rel=float(np.abs(a["p"] - b["p"]).max() / scale),
                rot_rad=float(rot_angle_between(a["q"], b["q"]).max()),
                ld_abs=abs(a["ld"] - b["ld"]))


def order_sensitivity(runs):
    """max pairwise difference of the first run against the others (the oracle against itself, factors shuffled)."""
    out = dict(cost_rel=0.0, cost_raw_rel=0.0, trans_rel=0.0, rot_rad=0.0, ld_abs=0.0)
    for r in runs[1:]:
        d = chain_difference(r, runs[0])
        for k in out:
            out[k] = max(out[k], d[k])
    return out
