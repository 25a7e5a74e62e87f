"""BASELINE config 5: streaming sliding window over a long synthetic sequence.

Mirrors the reference's per-image cycle (odometry_manager.cpp:253-281) through the public Estimator API:

  1. ExtendTrajectory(t_img + 40 ms)   trajectory_manager.cpp:108-120: new control points = copies of the last one
  2. InitTrajectory                     :288-315: IMU-only predictor over [max_bef_ns, maxTime), control points
                                        <= max_bef_idx fixed, biases locked, Solve(8)   (skipped for the first window)
  3. UpdateTrajectory(..., 15)          :317-483: prior + image + IMU + bias factors, Solve(15),
     double2vector                      :485-516: 4-DoF re-alignment to the pre-solve pose of the first control point
  4. UpdateVIOPrior(marg_flag)          :122-286: MARGIN_OLD marginalizes the oldest keyframe (its control points,
                                        bias node 0, the landmarks anchored in it) into the next prior;
                                        MARGIN_SECOND_NEW keeps the prior untouched
  5. SlideWindow                        drop the oldest / the second-newest frame

The host-side slicing of the synthetic sequence (the "feature tracker / feature manager") is not part of the timed
region; everything that crosses the C-ABI is.  The same class drives the CUDA engine and (tests / bench CPU leg) the
oracle.  Used by tests (GPU vs oracle over a few windows) and by bench.py ("c5").
"""
from __future__ import annotations

import time

import numpy as np

from . import synthetic as syn
from .binding import BLK_BA, BLK_BG, BLK_LD, BLK_POS, BLK_RHO, BLK_ROT, Estimator, PriorData

KF_DT_NS = 50_000_000       # 20 Hz keyframes
WINDOW_SIZE = 10            # visual_odometry/parameters.h:8
WIN_KF = WINDOW_SIZE + 1    # keyframes per window
EXTEND_NS = 40_000_000      # odometry_manager.cpp:246, 251
# keyframe phase inside a knot interval: (offset + 40 ms) mod 50 ms > 40 ms, so that the spline's end before the
# extension lies BEFORE the new image and InitTrajectory has IMU samples to work with (with 20 Hz images and 50 ms
# knots any other phase leaves the predictor without data)
C5_KF_OFFSET_NS = 7_000_000
MARGIN_OLD, MARGIN_SECOND_NEW = 0, 1


def config_c5_sequence(n_windows: int, seed=syn.SEED0 + 5, anchors=30, track_len=10):
    """n_windows + 10 keyframes at 20 Hz, `anchors` new landmarks per keyframe tracked over the next 10 keyframes,
    free line delay (online calibration)."""
    n_kf = n_windows + WIN_KF - 1
    kf = C5_KF_OFFSET_NS + np.arange(n_kf, dtype=np.int64) * KF_DT_NS
    n_knots = int((kf[-1] + 200_000_000) // syn.DT_NS) + 4
    per_frame = [anchors] * (n_kf - 1) + [0]
    return syn.make_window("C5-seq", n_knots, kf, per_frame, track_len, seed=seed, fix_ld=False)


class StreamingRunner:
    """One estimator engine driven through the reference's per-image cycle.

    perm_seed: shuffle the order in which factors are handed to the estimator (the summation order of the CPU
    oracle) -- used by the sensitivity tests; the window problem is mathematically unchanged.
    second_new_every: every N-th frame is treated as a non-keyframe: when it is the second-newest frame of the window
    the step takes the MARGIN_SECOND_NEW branch (no marginalization, the frame is dropped instead of the oldest).
    """

    def __init__(self, lib, seq: "syn.Window", iters=15, init_iters=8, device=0, second_new_every=0, perm_seed=None,
                 predictor=True):
        from . import make_config, make_options
        self.lib, self.seq, self.iters, self.init_iters = lib, seq, iters, init_iters
        self.predictor = predictor
        self.second_new_every = second_new_every
        self.rng = None if perm_seed is None else np.random.default_rng(perm_seed)
        s = seq
        self.frames = list(range(WIN_KF))                       # source keyframe ids in the window
        self.next_frame = WIN_KF
        # the spline so far: control points 0 .. ncp-1 (the initializer's output covers the first window)
        self.ncp = self._cp_needed(int(s.kf_times[WIN_KF - 1]) + EXTEND_NS)
        self.q = s.q0.copy(); self.p = s.p0.copy()
        self.bias = s.bias0.copy()                              # per source keyframe
        self.rho = s.rho0.copy()                                # per landmark (source ids)
        self.ld = s.ld0
        self.prior = None                                       # PriorData with GLOBAL knot / SOURCE frame / landmark ids
        cfg = make_config(device=device, **s.config_kwargs())
        self.est = Estimator(lib, cfg)
        self._make_options = make_options
        self.records = []
        self.step_index = 0

    # -- spline bookkeeping ------------------------------------------------------------------------
    def _cp_needed(self, t_ns):
        """smallest control-point count with maxTimeNs() >= t_ns (se3_spline.h:201-207)."""
        s = self.seq
        n = 4
        while s.t0_ns + (n - 3) * s.dt_ns < t_ns:
            n += 1
        return n

    def _knot_of(self, t_ns):
        return int((t_ns - self.seq.t0_ns) // self.seq.dt_ns)

    def _perm(self, n):
        return np.arange(n) if self.rng is None else self.rng.permutation(n)

    # -- prior re-indexing (index identity <-> window-relative indices) -----------------------------
    def _prior_to_window(self, ks, frames, lm_global):
        pr = self.prior
        if pr is None:
            return None
        out = PriorData(n=pr.n, J=pr.J, r=pr.r, blk_type=pr.blk_type.copy(), blk_index=pr.blk_index.copy(),
                        blk_col=pr.blk_col.copy(), blk_x0=pr.blk_x0)
        knots = (out.blk_type == BLK_ROT) | (out.blk_type == BLK_POS)
        out.blk_index[knots] -= ks
        biases = (out.blk_type == BLK_BG) | (out.blk_type == BLK_BA)
        if biases.any():
            pos = np.searchsorted(frames, out.blk_index[biases])
            assert np.all(np.asarray(frames)[np.clip(pos, 0, len(frames) - 1)] == out.blk_index[biases]), \
                "a bias node of the prior left the window"
            out.blk_index[biases] = pos
        isrho = out.blk_type == BLK_RHO
        if isrho.any():
            pos = np.searchsorted(lm_global, out.blk_index[isrho])
            assert np.all(lm_global[np.clip(pos, 0, len(lm_global) - 1)] == out.blk_index[isrho])
            out.blk_index[isrho] = pos
        assert out.blk_index.min() >= 0
        return out

    def _prior_to_global(self, pr, ks, frames, lm_global):
        if pr is None:
            return None
        knots = (pr.blk_type == BLK_ROT) | (pr.blk_type == BLK_POS)
        pr.blk_index[knots] += ks
        biases = (pr.blk_type == BLK_BG) | (pr.blk_type == BLK_BA)
        pr.blk_index[biases] = np.asarray(frames)[pr.blk_index[biases]]
        isrho = pr.blk_type == BLK_RHO
        pr.blk_index[isrho] = lm_global[pr.blk_index[isrho]]
        return pr

    # -- one image ------------------------------------------------------------------------------------
    def step(self, k=None):
        s = self.seq
        first = self.step_index == 0
        t_wall = 0.0
        e = self.est
        max_bef_ns = max_bef_idx = None
        if not first:
            # the new image joins the window (AddImageToWindow), then ExtendTrajectory
            self.frames.append(self.next_frame)
            self.bias[self.next_frame] = self.bias[self.frames[-2]]   # Bgs_[WINDOW_SIZE] starts from the newest estimate
            self.next_frame += 1
            t_img = int(s.kf_times[self.frames[-1]])
            max_bef_ns = s.t0_ns + (self.ncp - 3) * s.dt_ns
            max_bef_idx = self.ncp - 1
            ncp_new = self._cp_needed(t_img + EXTEND_NS)
            self.q[self.ncp:ncp_new] = self.q[self.ncp - 1]
            self.p[self.ncp:ncp_new] = self.p[self.ncp - 1]
            self.ncp = ncp_new
        frames = np.asarray(self.frames, np.int64)
        kf = s.kf_times[frames]
        t_newest = int(kf[-1])
        max_t = s.t0_ns + (self.ncp - 3) * s.dt_ns
        ks = self._knot_of(int(kf[0]))                # min_idx of UpdateTrajectory = first control point of the window
        nloc = self.ncp - ks
        nowk, later = 0, self._knot_of(int(kf[1])) - ks
        marg_flag = MARGIN_OLD
        if self.second_new_every and (self.frames[-2] % self.second_new_every) == self.second_new_every - 1:
            marg_flag = MARGIN_SECOND_NEW

        # ---- host-side "tracker / feature manager": slice the sequence (not timed) ----
        w = syn.subwindow_frames(s, frames, imu_max_ns=min(max_t, t_newest + 1), window_size=WINDOW_SIZE)
        lm_global = w.meta["lm_global"]
        rho = np.ascontiguousarray(self.rho[lm_global])
        img_marg = ((w.anchor_frame[w.lm] == 0) & (rho[w.lm] > 0)).astype(np.int32)   # :216-218
        imu_marg = (w.imu_t < kf[1]).astype(np.int32)                                  # :243-253
        bias_marg = np.zeros(len(w.bf_i), np.int32); bias_marg[0] = 1                 # :256-263
        if marg_flag != MARGIN_OLD:
            img_marg[:] = 0; imu_marg[:] = 0; bias_marg[:] = 0
        pi_, pm = self._perm(w.n_obs), self._perm(len(w.imu_t))
        q = np.ascontiguousarray(self.q[ks:self.ncp]); p = np.ascontiguousarray(self.p[ks:self.ncp])
        b = np.ascontiguousarray(self.bias[frames])
        prior = self._prior_to_window(ks, self.frames, lm_global)
        init_sel = None
        if not first and self.predictor:
            init_sel = np.nonzero((w.imu_t >= max_bef_ns) & (w.imu_t < max_t))[0]
        h2d = 0
        # the host buffers the front end hands over (factor order as the reference's containers would give it): built
        # here, outside the timed region, like the rest of the slicing
        img_args = tuple(np.ascontiguousarray(a[pi_]) for a in (w.ti, w.rowi, w.pi, w.tj, w.rowj, w.pj, w.lm, img_marg))
        imu_args = tuple(np.ascontiguousarray(a[pm]) for a in (w.imu_t, w.imu_gyro, w.imu_accel, w.imu_node, imu_marg))
        init_args = None
        if init_sel is not None and len(init_sel) > 0:
            init_args = (np.ascontiguousarray(w.imu_t[init_sel]), np.ascontiguousarray(w.imu_gyro[init_sel]),
                         np.ascontiguousarray(w.imu_accel[init_sel]), np.full(len(init_sel), len(frames) - 1, np.int32))
        opt_init = None if init_args is None else self._make_options(fixed_knot_index=max_bef_idx - ks, lock_wb=True, lock_ab=True,
                                                                    fix_ld=True)
        opt_main = self._make_options(fix_ld=False, ld_lower=0.0, ld_upper=syn.LD_UPPER, is_marg_state=(marg_flag == MARGIN_OLD),
                                      ctrl_to_be_opt_now=nowk, ctrl_to_be_opt_later=later)

        # ---- timed region: everything that crosses the C-ABI ----
        stats = e.lib.has("transfer_stats")
        if stats:
            e.TransferStats(reset=True)
        t_start = time.perf_counter()
        e.SetTimeOrigin(s.t0_ns + ks * s.dt_ns)
        e.SetKnots(q, p); e.SetBiases(b); e.SetInvDepths(rho); e.SetLineDelay(self.ld)
        h2d += q.nbytes + p.nbytes + b.nbytes + rho.nbytes + 8
        init_summary = None
        if init_args is not None:
            # InitTrajectory: IMU only, new control points only, biases locked at the newest keyframe's estimate
            e.SetOptions(opt_init)
            e.ClearFactors()
            e.AddMarginalizationFactor(None)
            e.AddIMUMeasurementAnalytic(*init_args)
            h2d += len(init_sel) * 64
            init_summary = e.Solve(self.init_iters)
        # UpdateTrajectory
        e.SetOptions(opt_main)
        e.ClearFactors()
        e.AddMarginalizationFactor(prior)
        e.AddImageFeatureDelayAnalytic(*img_args)
        e.AddIMUMeasurementAnalytic(*imu_args)
        e.AddBiasFactor(w.bf_i, w.bf_j, w.bf_sqrt_info, bias_marg)
        h2d += w.n_obs * 64 + len(w.imu_t) * 64 + len(w.bf_i) * 56 + (0 if prior is None else prior.n * prior.n * 8)
        # pre-solve pose of the window's first control point (R0, t0 of UpdateTrajectory:329-331) -- AFTER InitTrajectory
        q_pre, p_pre = (q[nowk], p[nowk])
        R0 = syn.qrot(q_pre[None], np.eye(3)).T.copy(); t0 = p_pre.copy()
        t_built = time.perf_counter()
        summ = e.Solve(self.iters)
        t_solved = time.perf_counter()
        e.GaugeRealign(nowk, R0, t0)
        new_prior = e.SaveMarginalizationInfo() if marg_flag == MARGIN_OLD else None
        t_marged = time.perf_counter()
        qs, ps = e.GetKnots(); bs = e.GetBiases(); rs = e.GetInvDepths(); ld = e.GetLineDelay()
        t_wall = time.perf_counter() - t_start
        d2h = qs.nbytes + ps.nbytes + bs.nbytes + rs.nbytes + 8 + (0 if new_prior is None else new_prior.J.nbytes)
        if stats:
            h2d, d2h = e.TransferStats(reset=True)  # the engine's own count (includes its index tables)

        # ---- carry the solution over (the reference updates the parameter blocks in place) ----
        self.q[ks:self.ncp] = qs; self.p[ks:self.ncp] = ps
        self.bias[frames] = bs
        self.rho[lm_global] = rs
        self.ld = ld
        if marg_flag == MARGIN_OLD:
            self.prior = self._prior_to_global(new_prior, ks, self.frames, lm_global)
            self.frames.pop(0)                      # slideWindowOld
        else:
            self.frames.pop(-2)                     # slideWindowNew: the prior stays as it is
        # 0.5 |r_lin|^2 of the prior this window was solved with: the part of the cost that is set by the eps = 1e-30
        # pseudo-inverse's noise eigen-directions (tests compare costs with this constant removed)
        prior_const = 0.0 if prior is None else 0.5 * float(np.dot(prior.r, prior.r))
        rec = dict(window=self.step_index, ms=1e3 * t_wall, prior_const=prior_const,
                   ms_build_and_predict=1e3 * (t_built - t_start), ms_solve=1e3 * (t_solved - t_built),
                   ms_realign_marginalize=1e3 * (t_marged - t_solved), ms_readback=1e3 * (t_start + t_wall - t_marged), iterations=summ.iterations, final_cost=summ.final_cost,
                   initial_cost=summ.initial_cost, termination=summ.termination, n_obs=w.n_obs, n_imu=len(w.imu_t),
                   n_knots=nloc, n_lm=len(lm_global), device_ms=summ.device_ms, marg_flag=marg_flag,
                   init_iterations=None if init_summary is None else init_summary.iterations,
                   init_n_imu=0 if init_sel is None else len(init_sel),
                   init_device_ms=0.0 if init_summary is None else init_summary.device_ms,
                   prior_dim=0 if self.prior is None else self.prior.n, h2d_bytes=h2d, d2h_bytes=d2h)
        self.records.append(rec)
        self.step_index += 1
        return rec

    def run(self, n_windows, first=0):
        for _ in range(n_windows):
            self.step()
        return self.records

    def state_error(self):
        """RMS translation error of the optimised part of the spline against the generator's truth (sanity metric)."""
        s = self.seq
        ks = self._knot_of(int(s.kf_times[self.frames[0]]))
        return float(np.sqrt(np.mean(np.sum((self.p[ks:self.ncp - 2] - s.p_gt[ks:self.ncp - 2]) ** 2, axis=1))))


def c3_window_a(lib, perm_seed=None, device=0):
    """BASELINE config 3, window A: the C3 source sequence restricted to keyframes 0..10 with the line delay free and
    the factors of keyframe 0 flagged for marginalization (trajectory_manager.cpp:206-263).  perm_seed shuffles the order
    in which factors are handed over (sensitivity tests).  Returns (estimator, sequence, window, first control point)."""
    from . import make_options, setup_estimator
    seq = syn.config_c3_sequence()
    wa = syn.subwindow(seq, 0, 10)
    later = int((wa.kf_times[1] - wa.t0_ns) // wa.dt_ns)
    nowk = int((wa.kf_times[0] - wa.t0_ns) // wa.dt_ns)
    img_marg = (wa.anchor_frame[wa.lm] == 0).astype(np.int32)
    imu_marg = (wa.imu_t < wa.kf_times[1]).astype(np.int32)
    bias_marg = np.zeros(len(wa.bf_i), np.int32); bias_marg[0] = 1
    if perm_seed is not None:
        rng = np.random.default_rng(perm_seed)
        pm = rng.permutation(wa.n_obs)
        for f in ("ti", "rowi", "pi", "tj", "rowj", "pj", "lm"):
            setattr(wa, f, np.ascontiguousarray(getattr(wa, f)[pm]))
        img_marg = img_marg[pm]
        pi = rng.permutation(len(wa.imu_t))
        for f in ("imu_t", "imu_gyro", "imu_accel", "imu_node"):
            setattr(wa, f, np.ascontiguousarray(getattr(wa, f)[pi]))
        imu_marg = imu_marg[pi]
    opt = make_options(fix_ld=False, ld_lower=0.0, ld_upper=syn.LD_UPPER, is_marg_state=True,
                       ctrl_to_be_opt_now=nowk, ctrl_to_be_opt_later=later)
    e = setup_estimator(lib, wa, image_marg=img_marg, imu_marg=imu_marg, bias_marg=bias_marg, options=opt, device=device)
    return e, seq, wa, nowk


# =====================================================================================================================
# Device-resident window fed by the wire formats (8f-4)

IMU_RECORD = np.dtype({"names": ["timestamp", "gyro", "accel", "orientation"],
                       "formats": [np.int64, (np.float64, 3), (np.float64, 3), (np.float64, 4)],
                       "offsets": [0, 8, 32, 64], "itemsize": 96})  # utils/parameter_struct.h:58-65 (Eigen alignment)


def quantize_wire(seq: "syn.Window"):
    """What survives the tracker's sensor_msgs::PointCloud: float32 bearings (rows are integers already).  Applied to
    the source sequence so that the classic (host-buffer) and the resident (wire-format) paths see identical numbers."""
    import copy
    s = copy.copy(seq)
    s.pi = seq.pi.astype(np.float32).astype(np.float64)
    s.pj = seq.pj.astype(np.float32).astype(np.float64)
    return s


class FrameClouds:
    """The tracker's per-frame messages rebuilt from the synthetic sequence: frame f carries the anchor observation of
    every landmark anchored in f followed by the observations of older landmarks seen in f (feature id = landmark id)."""

    def __init__(self, seq: "syn.Window"):
        self.seq = seq
        n_f = len(seq.kf_times)
        order_a = np.argsort(seq.anchor_frame, kind="stable")
        self.anchor_idx = np.empty(len(seq.anchor_frame), np.int32)     # landmark -> index inside its anchor frame's cloud
        counts_a = np.bincount(seq.anchor_frame, minlength=n_f)
        start_a = np.concatenate([[0], np.cumsum(counts_a)])
        self.anchor_idx[order_a] = np.arange(len(order_a)) - start_a[seq.anchor_frame[order_a]]
        order_o = np.argsort(seq.obs_frame, kind="stable")
        counts_o = np.bincount(seq.obs_frame, minlength=n_f)
        start_o = np.concatenate([[0], np.cumsum(counts_o)])
        self.obs_idx = np.empty(seq.n_obs, np.int32)                     # observation -> index inside its frame's cloud
        self.obs_idx[order_o] = np.arange(len(order_o)) - start_o[seq.obs_frame[order_o]] + counts_a[seq.obs_frame[order_o]]
        self._order_a, self._start_a, self._order_o, self._start_o = order_a, start_a, order_o, start_o
        # first observation of every landmark (its anchor bearing / row are replicated in every factor of the landmark)
        first = np.full(len(seq.anchor_frame), -1, np.int64)
        lms, idx = np.unique(seq.lm, return_index=True)
        first[lms] = idx
        self._first_obs = first

    def message(self, f):
        """(points float32 [n,3], id, u, v, vx, vy float32 [n]) of frame f."""
        s = self.seq
        la = self._order_a[self._start_a[f]:self._start_a[f + 1]]       # landmarks anchored here
        oo = self._order_o[self._start_o[f]:self._start_o[f + 1]]       # observations made here
        fo = self._first_obs[la]
        ok = fo >= 0
        xy = np.zeros((len(la), 2)); row = np.zeros(len(la))
        xy[ok] = s.pi[fo[ok]]; row[ok] = s.rowi[fo[ok]]
        xy = np.concatenate([xy, s.pj[oo]]); row = np.concatenate([row, s.rowj[oo].astype(np.float64)])
        ids = np.concatenate([la, s.lm[oo]]).astype(np.float32)
        n = len(ids)
        pts = np.ones((n, 3), np.float32); pts[:, :2] = xy
        z = np.zeros(n, np.float32)
        return pts, ids, (syn.FX * xy[:, 0] + syn.U0).astype(np.float32) if hasattr(syn, "FX") else z, row.astype(np.float32), z, z


class ResidentRunner(StreamingRunner):
    """The same per-image cycle with the window living in HBM: the new image's PointCloud and the new IMUData records go
    up as they are, control points are extended / dropped on the device, inverse depths are re-indexed on the device,
    the prior is handed over device-to-device, and the factor payload is gathered from the resident tables (only index
    tables cross the boundary).  MARGIN_OLD only (every frame a keyframe, the C5 configuration)."""

    def __init__(self, lib, seq, **kw):
        assert not kw.get("second_new_every"), "the resident runner implements the MARGIN_OLD slide only"
        super().__init__(lib, seq, **kw)
        self.clouds = FrameClouds(seq)
        self.n_slots = 16
        self.imu_sent = 0          # samples of the source sequence already ingested
        self.prev_lm_global = None
        self.prev_ks = None
        self.readback = None

    def _imu_records(self, lo, hi):
        s = self.seq
        rec = np.zeros(hi - lo, IMU_RECORD)
        rec["timestamp"] = s.imu_t[lo:hi]; rec["gyro"] = s.imu_gyro[lo:hi]; rec["accel"] = s.imu_accel[lo:hi]
        rec["orientation"][:, 3] = 1.0
        return rec

    def step(self, k=None):
        s = self.seq
        e = self.est
        first = self.step_index == 0
        t_push = 0.0
        if first:
            # the initializer's window: state, the 11 clouds and the IMU samples so far go up once
            e.SetTimeOrigin(s.t0_ns)
            e.SetKnots(self.q[:self.ncp], self.p[:self.ncp]); e.SetBiases(self.bias[self.frames]); e.SetLineDelay(self.ld)
            for f in self.frames:
                e.IngestFeatureCloud(f % self.n_slots, int(s.kf_times[f]), *self.clouds.message(f))
            self.base_knot = 0      # global index of the engine's knot 0
        else:
            self.frames.append(self.next_frame)
            self.next_frame += 1
        frames = np.asarray(self.frames, np.int64)
        kf = s.kf_times[frames]
        t_newest = int(kf[-1])
        t0 = time.perf_counter()
        max_bef_ns = max_bef_idx = None
        if not first:
            f = self.frames[-1]
            e.IngestFeatureCloud(f % self.n_slots, t_newest, *self.clouds.message(f))
            max_bef_ns = s.t0_ns + (self.ncp - 3) * s.dt_ns
            max_bef_idx = self.ncp - 1
            self.ncp = e.ExtendKnotsTo(t_newest + EXTEND_NS) + self.base_knot
        hi = int(np.searchsorted(s.imu_t, t_newest, side="right"))
        if hi > self.imu_sent:
            opt_min = s.t0_ns + self._knot_of(int(kf[0])) * s.dt_ns
            e.IngestImu(self._imu_records(self.imu_sent, hi), 8, 32, drop_before_ns=opt_min)
            self.imu_sent = hi
        t_push = time.perf_counter() - t0
        max_t = s.t0_ns + (self.ncp - 3) * s.dt_ns
        ks = self._knot_of(int(kf[0]))
        assert ks == self.base_knot, (ks, self.base_knot)
        nloc = self.ncp - ks
        nowk, later = 0, self._knot_of(int(kf[1])) - ks

        # ---- host-side index work of the "feature manager" (ids only, not timed like the classic runner's slicing) ----
        w = syn.subwindow_frames(s, frames, imu_max_ns=min(max_t, t_newest + 1), window_size=WINDOW_SIZE)
        lm_global = w.meta["lm_global"]
        if self.prev_lm_global is None:
            old_index = np.full(len(lm_global), -1, np.int32)
        else:
            pos = np.searchsorted(self.prev_lm_global, lm_global)
            pos = np.clip(pos, 0, len(self.prev_lm_global) - 1)
            old_index = np.where(self.prev_lm_global[pos] == lm_global, pos, -1).astype(np.int32)
        init_rho = s.rho0[lm_global]
        img_marg = (w.anchor_frame[w.lm] == 0).astype(np.int32)  # (inverse depths are positive in the synthetic sequences)
        bias_marg = np.zeros(len(w.bf_i), np.int32); bias_marg[0] = 1
        # factor -> (frame slot, index in that frame's cloud) of its two observations
        sel = self._factor_selection(frames, lm_global)
        g_lm = lm_global[w.lm]
        slot_i = (frames[w.anchor_frame[w.lm]] % self.n_slots).astype(np.int32)
        idx_i = self.clouds.anchor_idx[g_lm]
        slot_j = (frames[w.obs_frame] % self.n_slots).astype(np.int32)
        idx_j = self.clouds.obs_idx[sel]
        R0 = t0_ = None
        if self.readback is not None:
            qn, pn = self.readback[0][ks - self.prev_ks], self.readback[1][ks - self.prev_ks]
        else:
            qn, pn = self.q[ks], self.p[ks]
        R0 = syn.qrot(qn[None], np.eye(3)).T.copy(); t0_ = np.array(pn, float)

        # ---- timed region ----
        e.TransferStats(reset=True) if not first else None
        t_start = time.perf_counter()
        e.RemapLandmarks(old_index, init_rho)
        init_summary = None
        if not first and self.predictor:
            e.SetOptions(self._make_options(fixed_knot_index=max_bef_idx - ks, lock_wb=True, lock_ab=True, fix_ld=True))
            e.ClearFactors()
            e.EnablePrior(False)
            n_init = e.AddImuFromTable(max_bef_ns, max_t, fixed_node=len(frames) - 1)
            if n_init > 0:
                init_summary = e.Solve(self.init_iters)
        e.SetOptions(self._make_options(fix_ld=False, ld_lower=0.0, ld_upper=syn.LD_UPPER, is_marg_state=True,
                                        ctrl_to_be_opt_now=nowk, ctrl_to_be_opt_later=later))
        e.ClearFactors()
        e.EnablePrior(True)
        e.AddImageFeaturesFromSlots(slot_i, idx_i, slot_j, idx_j, w.lm, img_marg)
        opt_min = s.t0_ns + ks * s.dt_ns
        e.AddImuFromTable(opt_min, min(max_t, t_newest + 1), kf_times=kf, marg_before_ns=int(kf[1]))
        e.AddBiasFactor(w.bf_i, w.bf_j, w.bf_sqrt_info, bias_marg)
        t_built = time.perf_counter()
        summ = e.Solve(self.iters)
        t_solved = time.perf_counter()
        e.GaugeRealign(nowk, R0, t0_)
        n_out = C_int32(); nb_out = C_int32()
        e.lib.call("marginalize", e.h, byref(n_out), byref(nb_out))
        if n_out.value > 0:
            e.AdoptPrior()           # device-to-device
        t_marged = time.perf_counter()
        qs, ps = e.GetKnots(); ld = e.GetLineDelay()   # the trajectory is the product the caller publishes
        drop_knots = later
        e.SlideWindow(drop_knots, 1, 1)                # slideWindowOld: oldest frame's control points and bias node leave
        t_wall = time.perf_counter() - t_start
        h2d, d2h = (0, 0) if first else e.TransferStats(reset=True)

        self.q[ks:self.ncp] = qs; self.p[ks:self.ncp] = ps
        self.ld = ld
        self.readback, self.prev_ks, self.prev_lm_global = (qs, ps), ks, lm_global
        self.base_knot = ks + drop_knots
        self.frames.pop(0)
        rec = dict(window=self.step_index, ms=1e3 * (t_wall + t_push), prior_const=0.0,
                   ms_build_and_predict=1e3 * (t_built - t_start + t_push), ms_solve=1e3 * (t_solved - t_built),
                   ms_realign_marginalize=1e3 * (t_marged - t_solved), ms_readback=1e3 * (t_start + t_wall - t_marged),
                   iterations=summ.iterations, final_cost=summ.final_cost, initial_cost=summ.initial_cost,
                   termination=summ.termination, n_obs=w.n_obs, n_imu=len(w.imu_t), n_knots=nloc, n_lm=len(lm_global),
                   device_ms=summ.device_ms, marg_flag=MARGIN_OLD,
                   init_iterations=None if init_summary is None else init_summary.iterations, init_n_imu=0,
                   init_device_ms=0.0 if init_summary is None else init_summary.device_ms, prior_dim=n_out.value,
                   h2d_bytes=h2d, d2h_bytes=d2h)
        self.records.append(rec)
        self.step_index += 1
        return rec

    def _factor_selection(self, frames, lm_global):
        """indices (into the source sequence's observation arrays) of the window's factors, in subwindow_frames order"""
        s = self.seq
        pos = -np.ones(len(s.kf_times), np.int64); pos[frames] = np.arange(len(frames))
        keep = np.zeros(len(s.rho_gt), bool); keep[lm_global] = True
        return np.nonzero(keep[s.lm] & (pos[s.obs_frame] >= 0))[0]

    def sync_state_to_host(self):
        """full state read-back (tests): biases of the window's frames and inverse depths of its landmarks.
        Call right after step(): the engine's window has already slid by one keyframe."""
        e = self.est
        b = e.GetBiases()
        self.bias[np.asarray(self.frames[:len(b) - 1])] = b[:-1]
        return b


from ctypes import byref, c_int32 as C_int32  # noqa: E402  (used by ResidentRunner.step)
