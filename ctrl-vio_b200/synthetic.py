"""Deterministic synthetic sliding windows for the BASELINE.json configs (numpy only).

The generator is an independent (third) numpy restatement of the order-4
uniform cumulative SO(3)+R3 B-spline (src/spline/so3_spline.h:240-367,
src/spline/rd_spline.h:229-259 of the reference) used ONLY to fabricate inputs:
a ground-truth spline, rolling-shutter observations of random landmarks at their
per-row times, raw 200 Hz IMU samples, bias random-walk weights and a perturbed
initial guess.  It never touches the oracle or the CUDA engine.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

# config/ct_odometry_tumrs.yaml:13-28, tumrs/imu_tumrs.yaml:3, tumrs/cam_tumrs.yaml:9-19
DT_NS = 50_000_000
IMAGE_WEIGHT = 800.0
SIGMA_G, SIGMA_BG, SIGMA_A, SIGMA_BA = 4.0e-3, 2.0e-5, 8.0e-2, 4.0e-4
GRAVITY = np.array([0.0, 0.0, 9.80766])
P_CinI = np.array([0.00699407, -0.0570823, -0.0422772])
R_CtoI = np.array([[-0.00276873, -0.999936, -0.0110011],
                   [-0.999987, 0.00281495, -0.00418819],
                   [0.00421888, 0.0109894, -0.999931]])
FY, V0 = 739.1438452683457, 517.3370973594253
LD_TRUE = 29.4737e-6
LD_UPPER = 35e-6
RS_PADDING_NS = 39_000_000
IMU_DT_NS = 5_000_000

M_PLAIN = np.array([[1, -3, 3, -1], [4, 0, -6, 3], [1, 3, 3, -3], [0, 0, 0, 1]], float) / 6.0
M_CUM = np.array([[6, 0, 0, 0], [5, 3, -3, 1], [1, 3, 3, -2], [0, 0, 0, 1]], float) / 6.0


# ---------------------------------------------------------------------------
# quaternion helpers, [x, y, z, w]

def qmul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], -1)


def qconj(a):
    return a * np.array([-1.0, -1.0, -1.0, 1.0])


def qrot(q, v):
    qv = q[..., :3]
    uv = 2.0 * np.cross(qv, v)
    return v + q[..., 3:4] * uv + np.cross(qv, uv)


def qexp(w):
    th = np.linalg.norm(w, axis=-1, keepdims=True)
    small = th < 1e-10
    ths = np.where(small, 1.0, th)
    imag = np.where(small, 0.5 - th * th / 48.0, np.sin(0.5 * ths) / ths)
    real = np.where(small, 1.0 - th * th / 8.0, np.cos(0.5 * ths))
    return np.concatenate([imag * w, real], -1)


def qlog(q):
    n = np.linalg.norm(q[..., :3], axis=-1, keepdims=True)
    w = q[..., 3:4]
    small = n < 1e-10
    ns = np.where(small, 1.0, n)
    f = np.where(small, 2.0 / w, 2.0 * np.arctan(ns / w) / ns)
    return f * q[..., :3]


def qnormalize(q):
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def quat_from_matrix(R):
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[3] = (R[k, j] - R[j, k]) / s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
    return qnormalize(q)


Q_CtoI = quat_from_matrix(R_CtoI)


# ---------------------------------------------------------------------------
# spline evaluation (vectorised over times)

def _index(t_ns, t0_ns, dt_ns):
    st = np.asarray(t_ns, np.int64) - np.int64(t0_ns)
    s = st // np.int64(dt_ns)
    u = (st % np.int64(dt_ns)).astype(np.float64) / float(dt_ns)
    return s, u


def _coeffs(u, M, deriv, inv_dt):
    one = np.ones_like(u)
    zero = np.zeros_like(u)
    if deriv == 0:
        U = np.stack([one, u, u * u, u * u * u], -1)
    elif deriv == 1:
        U = np.stack([zero, one, 2 * u, 3 * u * u], -1)
    else:
        U = np.stack([zero, zero, 2 * one, 6 * u], -1)
    return (U @ M.T) * (inv_dt ** deriv)


def spline_pose(q_knots, p_knots, t_ns, t0_ns, dt_ns):
    """R(t) as quaternion and p(t)."""
    s, u = _index(t_ns, t0_ns, dt_ns)
    lam = _coeffs(u, M_CUM, 0, 1e9 / dt_ns)
    c = _coeffs(u, M_PLAIN, 0, 1e9 / dt_ns)
    q = q_knots[s]
    for j in range(3):
        d = qlog(qmul(qconj(q_knots[s + j]), q_knots[s + j + 1]))
        q = qmul(q, qexp(lam[:, j + 1:j + 2] * d))
    p = sum(c[:, k:k + 1] * p_knots[s + k] for k in range(4))
    return q, p


def spline_imu(q_knots, p_knots, t_ns, t0_ns, dt_ns):
    """body angular velocity w(t), world acceleration p''(t), world velocity p'(t)."""
    s, u = _index(t_ns, t0_ns, dt_ns)
    inv_dt = 1e9 / dt_ns
    lam = _coeffs(u, M_CUM, 0, inv_dt)
    dlam = _coeffs(u, M_CUM, 1, inv_dt)
    w = np.zeros((len(s), 3))
    for j in range(3):
        d = qlog(qmul(qconj(q_knots[s + j]), q_knots[s + j + 1]))
        w = qrot(qexp(-lam[:, j + 1:j + 2] * d), w) + dlam[:, j + 1:j + 2] * d
    c1 = _coeffs(u, M_PLAIN, 1, inv_dt)
    c2 = _coeffs(u, M_PLAIN, 2, inv_dt)
    v = sum(c1[:, k:k + 1] * p_knots[s + k] for k in range(4))
    a = sum(c2[:, k:k + 1] * p_knots[s + k] for k in range(4))
    return w, a, v


# ---------------------------------------------------------------------------

@dataclass
class Window:
    """One sliding-window problem in the flat layout the C-ABI takes."""

    name: str
    t0_ns: int
    dt_ns: int
    rs_padding_ns: int
    # state: ground truth and initial guess
    q_gt: np.ndarray
    p_gt: np.ndarray
    q0: np.ndarray
    p0: np.ndarray
    bias_gt: np.ndarray
    bias0: np.ndarray
    rho_gt: np.ndarray
    rho0: np.ndarray
    ld_gt: float
    ld0: float
    fix_ld: bool
    ld_lower: float
    ld_upper: float
    kf_times: np.ndarray
    # image factors
    ti: np.ndarray
    rowi: np.ndarray
    pi: np.ndarray
    tj: np.ndarray
    rowj: np.ndarray
    pj: np.ndarray
    lm: np.ndarray
    anchor_frame: np.ndarray  # per landmark
    obs_frame: np.ndarray     # per observation: target keyframe
    # imu
    imu_t: np.ndarray
    imu_gyro: np.ndarray
    imu_accel: np.ndarray
    imu_node: np.ndarray
    # bias factors
    bf_i: np.ndarray
    bf_j: np.ndarray
    bf_sqrt_info: np.ndarray
    meta: Dict = field(default_factory=dict)

    @property
    def n_knots(self):
        return self.q0.shape[0]

    @property
    def n_obs(self):
        return self.ti.shape[0]

    @property
    def n_residual_blocks(self):
        return self.n_obs + self.imu_t.shape[0] + self.bf_i.shape[0]

    def config_kwargs(self):
        return dict(t0_ns=self.t0_ns, dt_ns=self.dt_ns, q_CtoI=Q_CtoI, p_CinI=P_CinI, image_weight=IMAGE_WEIGHT,
                    gravity=GRAVITY, imu_info=np.array([1 / SIGMA_G] * 3 + [1 / SIGMA_A] * 3),
                    rs_padding_ns=self.rs_padding_ns, cauchy_solve=2.0, cauchy_marg=1.0)


def _truth_pose(t):
    """Analytic trajectory sampled for the ground-truth control points."""
    p = np.stack([2 * np.sin(0.8 * t), 2 * np.cos(0.6 * t), 0.5 * np.sin(1.1 * t)], -1)
    w = np.stack([0.3 * np.sin(0.7 * t), 0.3 * np.cos(0.5 * t), 0.4 * t], -1)
    return qexp(w), p


def _row_of(y):
    return np.clip(np.rint(FY * y + V0), 0, 1023).astype(np.int32)


def _project(qk, pk, t0_ns, dt_ns, pG, t_frame, ld_ns, global_shutter):
    """Rolling-shutter projection of world points into the camera at keyframe time t_frame.
    Fixed-point iteration on the row (row depends on y which depends on the row time)."""
    n = pG.shape[0]
    row = np.full(n, 512, np.int32) if not global_shutter else np.zeros(n, np.int32)
    xy = None
    z = None
    for _ in range(4):
        t = t_frame + row.astype(np.int64) * ld_ns
        q, p = spline_pose(qk, pk, t, t0_ns, dt_ns)
        pI = qrot(qconj(q), pG - p)
        pC = qrot(qconj(Q_CtoI)[None], pI - P_CinI)
        z = pC[:, 2]
        xy = pC[:, :2] / z[:, None]
        if global_shutter:
            break
        row = _row_of(xy[:, 1])
    return xy, z, row


def make_window(name: str, n_knots: int, kf_times_ns, anchors_per_frame, track_len: int, *, seed: int,
                global_shutter=False, with_imu=True, fix_ld=True, ld0=None, t0_ns=0, dt_ns=DT_NS,
                pixel_sigma=1.0 / 740.0, n_landmark_frames=None) -> Window:
    """anchors_per_frame[f] landmarks are anchored in keyframe f and observed in the next
    `track_len` keyframes (clipped to the window)."""
    rng = np.random.default_rng(seed)
    kf = np.asarray(kf_times_ns, np.int64)
    n_kf = len(kf)
    # ground-truth control points: control point k is centred on time t0 + (k-1) dt
    tk = (t0_ns + (np.arange(n_knots) - 1) * dt_ns) * 1e-9
    q_gt, p_gt = _truth_pose(tk)
    ld_gt = 0.0 if global_shutter else LD_TRUE
    ld_ns = np.int64(int(ld_gt * 1e9))
    max_t = t0_ns + (n_knots - 3) * dt_ns

    ti, rowi, pi, tj, rowj, pj, lm, obs_frame = [], [], [], [], [], [], [], []
    anchor_frame, rho_gt = [], []
    l_idx = 0
    for f, count in enumerate(anchors_per_frame):
        targets = [g for g in range(f + 1, min(f + 1 + track_len, n_kf))]
        if count == 0 or not targets:
            continue
        done = 0
        while done < count:
            m = max(2 * (count - done), 16)
            x = rng.uniform(-0.85, 0.85, m)
            y = rng.uniform(-0.69, 0.69, m)
            depth = rng.uniform(2.0, 10.0, m)
            r_a = np.zeros(m, np.int32) if global_shutter else _row_of(y)
            t_a = kf[f] + r_a.astype(np.int64) * ld_ns
            qa, pa = spline_pose(q_gt, p_gt, t_a, t0_ns, dt_ns)
            pC = np.stack([x, y, np.ones(m)], -1) * depth[:, None]
            pG = qrot(qa, qrot(Q_CtoI[None], pC) + P_CinI) + pa
            ok = np.ones(m, bool)
            obs = []
            for g in targets:
                xy, z, row = _project(q_gt, p_gt, t0_ns, dt_ns, pG, kf[g], ld_ns, global_shutter)
                ok &= (z > 0.5) & (np.abs(xy[:, 0]) < 1.5) & (np.abs(xy[:, 1]) < 1.2)
                obs.append((xy, row))
            sel = np.nonzero(ok)[0][: count - done]
            k = len(sel)
            if k == 0:
                continue
            # noisy anchor bearing (row recomputed from the noisy y like the tracker would)
            a_xy = np.stack([x[sel], y[sel]], -1) + rng.normal(0, pixel_sigma, (k, 2))
            a_row = np.zeros(k, np.int32) if global_shutter else _row_of(a_xy[:, 1])
            for (xy, row), g in zip(obs, targets):
                o_xy = xy[sel] + rng.normal(0, pixel_sigma, (k, 2))
                o_row = np.zeros(k, np.int32) if global_shutter else _row_of(o_xy[:, 1])
                ti.append(np.full(k, kf[f], np.int64)); rowi.append(a_row); pi.append(a_xy)
                tj.append(np.full(k, kf[g], np.int64)); rowj.append(o_row); pj.append(o_xy)
                lm.append(l_idx + np.arange(k, dtype=np.int32)); obs_frame.append(np.full(k, g, np.int32))
            anchor_frame.append(np.full(k, f, np.int32))
            rho_gt.append(1.0 / depth[sel])
            l_idx += k
            done += k
    cat = lambda xs, dt, shape=None: (np.concatenate(xs).astype(dt) if xs else np.zeros((0,) + (shape or ()), dt))
    ti = cat(ti, np.int64); tj = cat(tj, np.int64); rowi = cat(rowi, np.int32); rowj = cat(rowj, np.int32)
    pi = cat(pi, np.float64, (2,)); pj = cat(pj, np.float64, (2,)); lm = cat(lm, np.int32)
    obs_frame = cat(obs_frame, np.int32); anchor_frame = cat(anchor_frame, np.int32); rho_gt = cat(rho_gt, np.float64)
    # landmark-major order (all observations of a landmark adjacent), the order the reference's
    # feature loop produces (trajectory_manager.cpp:360-385)
    order = np.lexsort((obs_frame, lm))
    ti, tj, rowi, rowj, pi, pj, lm, obs_frame = (a[order] for a in (ti, tj, rowi, rowj, pi, pj, lm, obs_frame))

    # IMU (trajectory_manager.cpp:388-417): samples in [opt_min_time, maxTime)
    bias_true = np.array([0.01, -0.02, 0.005, 0.05, 0.02, -0.03])
    if with_imu:
        opt_min = t0_ns + ((kf[0] - t0_ns) // dt_ns) * dt_ns
        imu_t = np.arange(opt_min, max_t, IMU_DT_NS, dtype=np.int64)
        w, a, _ = spline_imu(q_gt, p_gt, imu_t, t0_ns, dt_ns)
        qi, _ = spline_pose(q_gt, p_gt, imu_t, t0_ns, dt_ns)
        acc = qrot(qconj(qi), a + GRAVITY)
        # per-sample noise equal to the sigma the reference's weights assume (imu_info = 1/sigma,
        # utils/opt_weight.h:124-126), so whitened IMU residuals have unit variance at the truth
        gyro = w + bias_true[:3] + rng.normal(0, SIGMA_G, w.shape)
        accel = acc + bias_true[3:] + rng.normal(0, SIGMA_A, acc.shape)
        node = np.clip(np.searchsorted(kf, imu_t, side="right") - 1, 0, n_kf - 1).astype(np.int32)
        # bias random-walk weights (trajectory_manager.cpp:420-450): cov = sum dt_k^2 sigma^2
        bf_i = np.arange(n_kf - 1, dtype=np.int32)
        bf_j = bf_i + 1
        si = bias_sqrt_info(imu_t, kf)
        n_nodes = n_kf
    else:
        imu_t = np.zeros(0, np.int64); gyro = np.zeros((0, 3)); accel = np.zeros((0, 3)); node = np.zeros(0, np.int32)
        bf_i = np.zeros(0, np.int32); bf_j = np.zeros(0, np.int32); si = np.zeros((0, 6))
        n_nodes = 1
    bias_gt = np.tile(bias_true, (n_nodes, 1))

    # initial guess: 1 cm / 0.3 deg knot noise, 10 % inverse-depth noise, zero biases
    p0 = p_gt + rng.normal(0, 0.01, p_gt.shape)
    q0 = qnormalize(qmul(q_gt, qexp(rng.normal(0, np.deg2rad(0.3), (n_knots, 3)))))
    rho0 = rho_gt * (1.0 + rng.normal(0, 0.1, rho_gt.shape))
    if ld0 is None:
        ld0 = ld_gt if fix_ld else 0.0
    return Window(name=name, t0_ns=t0_ns, dt_ns=dt_ns, rs_padding_ns=(0 if global_shutter else RS_PADDING_NS),
                  q_gt=q_gt, p_gt=p_gt, q0=q0, p0=p0, bias_gt=bias_gt, bias0=np.zeros_like(bias_gt), rho_gt=rho_gt,
                  rho0=rho0, ld_gt=ld_gt, ld0=float(ld0), fix_ld=fix_ld, ld_lower=0.0, ld_upper=LD_UPPER,
                  kf_times=kf, ti=ti, rowi=rowi, pi=pi, tj=tj, rowj=rowj, pj=pj, lm=lm, anchor_frame=anchor_frame,
                  obs_frame=obs_frame, imu_t=imu_t, imu_gyro=gyro, imu_accel=accel, imu_node=node, bf_i=bf_i,
                  bf_j=bf_j, bf_sqrt_info=si, meta=dict(seed=seed))


def bias_sqrt_info(imu_t, kf):
    """Bias random-walk weights between consecutive keyframes (trajectory_manager.cpp:420-450):
    cov = sum dt_k^2 sigma^2 over the IMU intervals [t_{k-1}, t_k] with t_{k-1} >= kf_i and t_k < kf_{i+1}."""
    kf = np.asarray(kf, np.int64)
    si = np.zeros((max(len(kf) - 1, 0), 6))
    if len(imu_t) < 2:
        return si
    dts = np.diff(imu_t) * 1e-9
    csum = np.concatenate([[0.0], np.cumsum(dts * dts)])      # csum[m] = sum of the first m intervals
    for i in range(len(kf) - 1):
        a = np.searchsorted(imu_t, kf[i], side="left")         # first sample >= kf_i  -> interval index a (ends at a+1)
        b = np.searchsorted(imu_t, kf[i + 1], side="left")     # first sample >= kf_{i+1}: intervals ending before it
        s2 = csum[max(b - 1, a)] - csum[a] if b - 1 > a else 0.0
        if s2 > 0:
            si[i, :3] = 1.0 / np.sqrt(s2 * SIGMA_BG ** 2)
            si[i, 3:] = 1.0 / np.sqrt(s2 * SIGMA_BA ** 2)
    return si


SEED0 = 0xC7A1
KF_OFFSET_NS = 31_000_000  # keyframes are not aligned with knot boundaries; high rows cross into the next interval


def config_c1(seed=SEED0 + 1):
    """C1: 4 knots, 50 landmarks, 200 global-shutter observations, no IMU (CPU plumbing case)."""
    kf = (np.array([0.05, 0.25, 0.45, 0.65, 0.85]) * DT_NS).astype(np.int64)
    return make_window("C1", 4, kf, [50, 0, 0, 0, 0], 4, seed=seed, global_shutter=True, with_imu=False, fix_ld=True)


def config_c2(seed=SEED0 + 2, fix_ld=True, n_kf=11, n_knots=30):
    """C2: 30 control points, 11 keyframes @100 ms, 300 landmarks, 2 700 rolling-shutter observations,
    270 IMU samples @200 Hz, 10 bias factors, line delay fixed at truth."""
    kf = KF_OFFSET_NS + np.arange(n_kf, dtype=np.int64) * 100_000_000
    anchors = [100, 100, 100] + [0] * (n_kf - 3)
    return make_window("C2" if fix_ld else "C2-ldfree", n_knots, kf, anchors, n_kf, seed=seed, fix_ld=fix_ld)


def config_c3_sequence(seed=SEED0 + 3):
    """C3 source sequence: 12 keyframes / 32 knots with the line delay free.  Window A = keyframes 0..10
    (solve, re-align, marginalize keyframe 0), window B = keyframes 1..11 with the resulting prior."""
    kf = KF_OFFSET_NS + np.arange(12, dtype=np.int64) * 100_000_000
    anchors = [100, 100, 100, 100] + [0] * 8
    return make_window("C3-seq", 32, kf, anchors, 12, seed=seed, fix_ld=False)


def config_c4(seed=SEED0 + 4, n_landmarks=10_000):
    """C4: 100 control points, 48 keyframes @100 ms, 10 000 landmarks anchored uniformly over frames 0..37
    and observed in the 10 following keyframes -> 100 000 observations, 970 IMU samples, 47 bias factors."""
    kf = KF_OFFSET_NS + np.arange(48, dtype=np.int64) * 100_000_000
    per = n_landmarks // 38
    anchors = [per + (1 if f < n_landmarks - per * 38 else 0) for f in range(38)] + [0] * 10
    return make_window("C4", 100, kf, anchors, 10, seed=seed, fix_ld=True)


def subwindow(w: Window, kf_first: int, kf_last: int, imu_min_ns=None, imu_max_ns=None) -> Window:
    """Restrict a sequence to keyframes [kf_first, kf_last] (same global knot array; bias nodes and
    landmark ids are re-indexed to the window like the reference's para_* arrays)."""
    import copy

    keep_lm = (w.anchor_frame >= kf_first) & (w.anchor_frame <= kf_last)
    o = keep_lm[w.lm] & (w.obs_frame >= kf_first) & (w.obs_frame <= kf_last)
    # landmarks need >= 1 surviving observation
    cnt = np.bincount(w.lm[o], minlength=len(w.rho_gt))
    keep_lm &= cnt > 0
    o &= keep_lm[w.lm]
    new_id = -np.ones(len(w.rho_gt), np.int64)
    new_id[keep_lm] = np.arange(keep_lm.sum())
    kf = w.kf_times[kf_first:kf_last + 1]
    opt_min = w.t0_ns + ((kf[0] - w.t0_ns) // w.dt_ns) * w.dt_ns
    lo = opt_min if imu_min_ns is None else imu_min_ns
    hi = (w.t0_ns + (w.n_knots - 3) * w.dt_ns) if imu_max_ns is None else imu_max_ns
    im = (w.imu_t >= lo) & (w.imu_t < hi)
    node = np.clip(np.searchsorted(kf, w.imu_t[im], side="right") - 1, 0, len(kf) - 1).astype(np.int32)
    out = copy.copy(w)
    out.name = f"{w.name}[{kf_first}:{kf_last}]"
    out.kf_times = kf
    out.ti, out.rowi, out.pi = w.ti[o], w.rowi[o], w.pi[o]
    out.tj, out.rowj, out.pj = w.tj[o], w.rowj[o], w.pj[o]
    out.lm = new_id[w.lm[o]].astype(np.int32)
    out.obs_frame = (w.obs_frame[o] - kf_first).astype(np.int32)
    out.anchor_frame = (w.anchor_frame[keep_lm] - kf_first).astype(np.int32)
    out.rho_gt, out.rho0 = w.rho_gt[keep_lm], w.rho0[keep_lm]
    out.imu_t, out.imu_gyro, out.imu_accel, out.imu_node = w.imu_t[im], w.imu_gyro[im], w.imu_accel[im], node
    out.bias_gt, out.bias0 = w.bias_gt[kf_first:kf_last + 1], w.bias0[kf_first:kf_last + 1]
    out.bf_i = np.arange(len(kf) - 1, dtype=np.int32)
    out.bf_j = out.bf_i + 1
    out.bf_sqrt_info = w.bf_sqrt_info[kf_first:kf_last]
    out.meta = dict(w.meta, lm_global=np.nonzero(keep_lm)[0], kf_first=kf_first)
    return out


def subwindow_frames(w: Window, frames, imu_min_ns=None, imu_max_ns=None, window_size=None) -> Window:
    """Restrict a sequence to an arbitrary ascending list of keyframes (the window after MARGIN_SECOND_NEW slides is
    not contiguous in the source sequence).  Landmarks keep their anchor; landmarks anchored in a frame outside the
    list disappear with all their observations.  window_size: apply FeatureManager::isLandmarkCandidate
    (feature_manager.h:58-65: used_num >= 2 && start_frame < WINDOW_SIZE - 2).  Bias nodes = positions in `frames`;
    bias random-walk weights are recomputed between consecutive listed frames."""
    import copy

    frames = np.asarray(frames, np.int64)
    n_src = len(w.kf_times)
    pos = -np.ones(n_src, np.int64)
    pos[frames] = np.arange(len(frames))
    keep_lm = pos[w.anchor_frame] >= 0
    if window_size is not None:
        keep_lm &= pos[w.anchor_frame] < window_size - 2
    o = keep_lm[w.lm] & (pos[w.obs_frame] >= 0)
    cnt = np.bincount(w.lm[o], minlength=len(w.rho_gt))
    keep_lm &= cnt > 0          # used_num = 1 + cnt >= 2
    o &= keep_lm[w.lm]
    new_id = -np.ones(len(w.rho_gt), np.int64)
    new_id[keep_lm] = np.arange(keep_lm.sum())
    kf = w.kf_times[frames]
    opt_min = w.t0_ns + ((kf[0] - w.t0_ns) // w.dt_ns) * w.dt_ns
    lo = opt_min if imu_min_ns is None else imu_min_ns
    hi = (w.t0_ns + (w.n_knots - 3) * w.dt_ns) if imu_max_ns is None else imu_max_ns
    im = (w.imu_t >= lo) & (w.imu_t < hi)
    node = np.clip(np.searchsorted(kf, w.imu_t[im], side="right") - 1, 0, len(kf) - 1).astype(np.int32)
    out = copy.copy(w)
    out.name = f"{w.name}{list(frames[[0, -1]])}"
    out.kf_times = kf
    out.ti, out.rowi, out.pi = w.ti[o], w.rowi[o], w.pi[o]
    out.tj, out.rowj, out.pj = w.tj[o], w.rowj[o], w.pj[o]
    out.lm = new_id[w.lm[o]].astype(np.int32)
    out.obs_frame = pos[w.obs_frame[o]].astype(np.int32)
    out.anchor_frame = pos[w.anchor_frame[keep_lm]].astype(np.int32)
    out.rho_gt, out.rho0 = w.rho_gt[keep_lm], w.rho0[keep_lm]
    out.imu_t, out.imu_gyro, out.imu_accel, out.imu_node = w.imu_t[im], w.imu_gyro[im], w.imu_accel[im], node
    out.bias_gt, out.bias0 = w.bias_gt[frames], w.bias0[frames]
    out.bf_i = np.arange(len(kf) - 1, dtype=np.int32)
    out.bf_j = out.bf_i + 1
    out.bf_sqrt_info = bias_sqrt_info(w.imu_t, kf)
    out.meta = dict(w.meta, lm_global=np.nonzero(keep_lm)[0], frames=frames)
    return out
