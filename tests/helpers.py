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
