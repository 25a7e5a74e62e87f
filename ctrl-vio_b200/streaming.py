"""BASELINE config 5: streaming sliding window over a long synthetic sequence.

Mirrors the reference's per-image cycle (TrajectoryManager::UpdateTrajectory -> double2vector re-alignment ->
UpdatePrior, src/estimator/trajectory_manager.cpp:130-516) through the public Estimator API: every window
  1. moves the time origin to the first control point the window touches and uploads the window's slice of control
     points, bias nodes, inverse depths (carried over from the previous window's solution; new ones from the tracker /
     initial guess),
  2. re-adds the window's image / IMU / bias factors (oldest keyframe flagged for marginalization) and the prior,
  3. solves, re-aligns the 4-DoF gauge, marginalizes the oldest keyframe into the next prior,
  4. reads the state back.
The host-side slicing of the synthetic sequence (the "feature tracker") is not part of the timed region; everything
that crosses the C-ABI is.  Used by tests (GPU vs oracle over a few windows) and by bench.py ("c5").
"""
from __future__ import annotations

import time

import numpy as np

from . import synthetic as syn
from .binding import BLK_BA, BLK_BG, BLK_POS, BLK_RHO, BLK_ROT, Estimator, PriorData

KF_DT_NS = 50_000_000  # 20 Hz keyframes
WIN_KF = 11            # keyframes per window (WINDOW_SIZE 10 + the newest one)


def config_c5_sequence(n_windows: int, seed=syn.SEED0 + 5, anchors=30, track_len=10):
    """n_windows + 10 keyframes at 20 Hz, `anchors` new landmarks per keyframe tracked over the next 10 keyframes,
    free line delay (online calibration)."""
    n_kf = n_windows + WIN_KF - 1
    kf = syn.KF_OFFSET_NS + np.arange(n_kf, dtype=np.int64) * KF_DT_NS
    n_knots = int((kf[-1] + 200_000_000) // syn.DT_NS) + 4
    per_frame = [anchors] * (n_kf - 1) + [0]
    return syn.make_window("C5-seq", n_knots, kf, per_frame, track_len, seed=seed, fix_ld=False)


class StreamingRunner:
    def __init__(self, lib, seq: "syn.Window", iters=8, device=0):
        from . import make_config, make_options
        self.lib, self.seq, self.iters = lib, seq, iters
        self.q = seq.q0.copy(); self.p = seq.p0.copy()          # global control points (solution so far / initial guess)
        self.bias = seq.bias0.copy()                             # per keyframe
        self.rho = seq.rho0.copy()                               # per landmark (global ids)
        self.ld = seq.ld0
        self.prior = None
        self.prev_ks = None
        self.prev_lm_global = None
        cfg = make_config(device=device, **seq.config_kwargs())
        self.est = Estimator(lib, cfg)
        self._make_options = make_options
        self.records = []

    def _layout(self, k):
        s = self.seq
        kf = s.kf_times[k:k + WIN_KF]
        idx = lambda t: int((t - s.t0_ns) // s.dt_ns)
        ks = max(0, idx(kf[0] - s.rs_padding_ns))
        last = idx(kf[-1] + 80_000_000) + 4
        return kf, ks, min(last, s.n_knots) - ks, idx(kf[0]) - ks, idx(kf[1]) - ks

    def step(self, k):
        s = self.seq
        kf, ks, nloc, nowk, later = self._layout(k)
        # ---- host-side "tracker": slice the sequence (not timed) ----
        w = syn.subwindow(s, k, k + WIN_KF - 1, imu_max_ns=int(kf[-1]))
        lm_global = w.meta["lm_global"]
        img_marg = (w.anchor_frame[w.lm] == 0).astype(np.int32)
        imu_marg = (w.imu_t < kf[1]).astype(np.int32)
        bias_marg = np.zeros(len(w.bf_i), np.int32); bias_marg[0] = 1
        q = np.ascontiguousarray(self.q[ks:ks + nloc]); p = np.ascontiguousarray(self.p[ks:ks + nloc])
        b = np.ascontiguousarray(self.bias[k:k + WIN_KF])
        rho = np.ascontiguousarray(self.rho[lm_global])
        prior = self._shift_prior(ks, lm_global)
        R0 = syn.qrot(q[nowk][None], np.eye(3)).T.copy(); t0 = p[nowk].copy()
        e = self.est
        # ---- timed region: everything that crosses the C-ABI ----
        t_start = time.perf_counter()
        e.SetTimeOrigin(s.t0_ns + ks * s.dt_ns)
        e.SetOptions(self._make_options(fix_ld=False, ld_lower=0.0, ld_upper=syn.LD_UPPER, is_marg_state=True,
                                        ctrl_to_be_opt_now=nowk, ctrl_to_be_opt_later=later))
        e.SetKnots(q, p); e.SetBiases(b); e.SetInvDepths(rho); e.SetLineDelay(self.ld)
        e.ClearFactors()
        e.AddImageFeatureDelayAnalytic(w.ti, w.rowi, w.pi, w.tj, w.rowj, w.pj, w.lm, img_marg)
        e.AddIMUMeasurementAnalytic(w.imu_t, w.imu_gyro, w.imu_accel, w.imu_node, imu_marg)
        e.AddBiasFactor(w.bf_i, w.bf_j, w.bf_sqrt_info, bias_marg)
        e.AddMarginalizationFactor(prior)
        summ = e.Solve(self.iters)
        e.GaugeRealign(nowk, R0, t0)
        new_prior = e.SaveMarginalizationInfo()
        qs, ps = e.GetKnots(); bs = e.GetBiases(); rs = e.GetInvDepths(); ld = e.GetLineDelay()
        ms = 1e3 * (time.perf_counter() - t_start)
        # ---- carry the solution over ----
        # control points beyond the support of the newest IMU sample are only touched by a few high-row features with
        # basis weights < 1e-2: they are not carried over, the front end re-initialises them (here: the generator's
        # initial guess, standing in for the reference's IMU-propagated InitTrajectory / extendKnotsTo)
        keep = int((kf[-1] - s.t0_ns) // s.dt_ns) - ks + 3
        self.q[ks:ks + keep] = qs[:keep]; self.p[ks:ks + keep] = ps[:keep]
        self.bias[k:k + WIN_KF] = bs
        if k + WIN_KF < len(self.bias):
            self.bias[k + WIN_KF] = bs[-1]          # the next keyframe's bias node starts from the newest estimate
        self.rho[lm_global] = rs
        self.ld = ld
        self.prior, self.prev_ks, self.prev_lm_global = new_prior, ks, lm_global
        rec = dict(window=k, ms=ms, iterations=summ.iterations, final_cost=summ.final_cost, n_obs=w.n_obs,
                   n_knots=nloc, device_ms=summ.device_ms, prior_dim=0 if new_prior is None else new_prior.n)
        self.records.append(rec)
        return rec

    def _shift_prior(self, ks, lm_global):
        """Block indices of the prior are relative to the window that produced it: re-index them for this window."""
        pr = self.prior
        if pr is None:
            return None
        out = PriorData(n=pr.n, J=pr.J, r=pr.r, blk_type=pr.blk_type.copy(), blk_index=pr.blk_index.copy(),
                        blk_col=pr.blk_col.copy(), blk_x0=pr.blk_x0)
        knots = (out.blk_type == BLK_ROT) | (out.blk_type == BLK_POS)
        out.blk_index[knots] -= ks - self.prev_ks
        biases = (out.blk_type == BLK_BG) | (out.blk_type == BLK_BA)
        out.blk_index[biases] -= 1
        isrho = out.blk_type == BLK_RHO
        if isrho.any():
            g = self.prev_lm_global[out.blk_index[isrho]]
            pos = np.searchsorted(lm_global, g)
            assert np.all(lm_global[np.clip(pos, 0, len(lm_global) - 1)] == g), "a landmark of the prior left the window"
            out.blk_index[isrho] = pos
        assert out.blk_index.min() >= 0
        return out

    def run(self, n_windows, first=0):
        for k in range(first, first + n_windows):
            self.step(k)
        return self.records
