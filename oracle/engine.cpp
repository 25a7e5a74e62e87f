// ORACLE (test infrastructure, NOT product code).  Parity unpinned (see so3.hpp).
// See engine.hpp for the list of reference sources restated here.
#include "engine.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <thread>

#include "linalg.hpp"

namespace ctvio_oracle {

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------------------
// structure

void Window::knotWindow(int64_t t, int& first, int& last) const {
  // se3_spline.h:463-503 with times {t, t + rs_padding}: knots i1 .. i2 + N - 1.
  int64_t s1 = 0, s2 = 0;
  double u;
  const int smax = nK() - kN;
  if (!grid.computeTIndexNs(t, s1, u)) s1 = std::min<int64_t>(std::max<int64_t>((t - grid.t0_ns) / grid.dt_ns, 0), smax);
  if (!grid.computeTIndexNs(t + opt.rs_padding_ns, s2, u))
    s2 = std::min<int64_t>(std::max<int64_t>((t + opt.rs_padding_ns - grid.t0_ns) / grid.dt_ns, s1), smax);
  first = int(s1);
  last = int(s2) + kN - 1;
}

void Window::buildStructure(NormalEq* ne) const {
  ne->nK = nK(); ne->nB = nB(); ne->nL = nL(); ne->np = np();
  ne->lo.assign(ne->nL, std::numeric_limits<int>::max());
  ne->hi.assign(ne->nL, 0);
  for (const ImageObs& o : img) {
    int f, l;
    knotWindow(o.ti, f, l);
    ne->lo[o.lm] = std::min(ne->lo[o.lm], 6 * f);
    ne->hi[o.lm] = std::max(ne->hi[o.lm], 6 * (l + 1));
    knotWindow(o.tj, f, l);
    ne->lo[o.lm] = std::min(ne->lo[o.lm], 6 * f);
    ne->hi[o.lm] = std::max(ne->hi[o.lm], 6 * (l + 1));
  }
  ne->woff.assign(ne->nL + 1, 0);
  for (int l = 0; l < ne->nL; ++l) {
    if (ne->hi[l] == 0) ne->lo[l] = 0;
    ne->woff[l + 1] = ne->woff[l] + size_t(ne->hi[l] - ne->lo[l]);
  }
  ne->Hcc.assign(size_t(ne->np) * ne->np, 0.0);
  ne->gc.assign(ne->np, 0.0);
  ne->hl.assign(ne->nL, 0.0);
  ne->gl.assign(ne->nL, 0.0);
  ne->wld.assign(ne->nL, 0.0);
  ne->W.assign(ne->woff[ne->nL], 0.0);
}

std::vector<char> Window::constMask() const {
  std::vector<char> c(np(), 0);
  for (int k = 0; k < nK(); ++k)
    if (opt.lock_traj || (opt.fixed_knot_index >= 0 && k <= opt.fixed_knot_index))
      for (int d = 0; d < 6; ++d) c[idxKnot(k) + d] = 1;  // trajectory_estimator.cpp:134-138
  for (int b = 0; b < nB(); ++b)
    for (int d = 0; d < 3; ++d) {
      if (opt.lock_wb) c[idxBias(b) + d] = 1;      // :236-240
      if (opt.lock_ab) c[idxBias(b) + 3 + d] = 1;  // :241-245
    }
  if (opt.fix_ld) c[idxLd()] = 1;  // :312-313
  return c;
}

std::vector<char> Window::touchedMask() const {
  std::vector<char> t(np() + nL(), 0);
  auto markKnots = [&](int f, int l) {
    for (int k = f; k <= l; ++k)
      for (int d = 0; d < 6; ++d) t[idxKnot(k) + d] = 1;
  };
  for (const ImageObs& o : img) {
    int f, l;
    knotWindow(o.ti, f, l); markKnots(f, l);
    knotWindow(o.tj, f, l); markKnots(f, l);
    t[idxLd()] = 1;
    t[np() + o.lm] = 1;
  }
  for (const ImuObs& o : imu) {
    int64_t s; double u;
    if (!grid.computeTIndexNs(o.t, s, u)) continue;
    markKnots(int(s), int(s) + 3);
    for (int d = 0; d < 6; ++d) t[idxBias(o.bias_idx) + d] = 1;
  }
  for (const BiasObs& o : biasf)
    for (int d = 0; d < 6; ++d) t[idxBias(o.i) + d] = t[idxBias(o.j) + d] = 1;
  if (prior.valid())
    for (const PriorBlock& b : prior.blocks) {
      const int ls = blockLocalSize(b.type);
      int base = 0;
      switch (b.type) {
        case kBlkRot: base = idxKnot(b.index); break;
        case kBlkPos: base = idxKnot(b.index) + 3; break;
        case kBlkBg: base = idxBias(b.index); break;
        case kBlkBa: base = idxBias(b.index) + 3; break;
        case kBlkLd: base = idxLd(); break;
        default: base = np() + b.index; break;
      }
      for (int d = 0; d < ls; ++d) t[base + d] = 1;
    }
  return t;
}

// ---------------------------------------------------------------------------
// prior factor (marginalization_factor.cpp:326-373)

static int priorBlockBase(const Window& w, const PriorBlock& b) {
  switch (b.type) {
    case kBlkRot: return w.idxKnot(b.index);
    case kBlkPos: return w.idxKnot(b.index) + 3;
    case kBlkBg: return w.idxBias(b.index);
    case kBlkBa: return w.idxBias(b.index) + 3;
    case kBlkLd: return w.idxLd();
    default: return -1;  // inverse depths never survive into a prior (SURVEY C-13)
  }
}

static const double* blockData(const Window& w, int type, int index) {
  switch (type) {
    case kBlkRot: return &w.q[4 * index];
    case kBlkPos: return &w.p[3 * index];
    case kBlkBg: return &w.bias[6 * index];
    case kBlkBa: return &w.bias[6 * index + 3];
    case kBlkLd: return &w.ld;
    default: return &w.rho[index];
  }
}

// dx of one kept block, marginalization_factor.cpp:334-352.
static void priorBlockDx(const Window& w, const PriorBlock& b, double* dx) {
  const double* x = blockData(w, b.type, b.index);
  if (b.type != kBlkRot) {
    const int sz = blockGlobalSize(b.type);
    for (int d = 0; d < sz; ++d) dx[d] = x[d] - b.x0[d];
    return;
  }
  // q0_inv = Quaterniond(x0).inverse() == conj / squaredNorm
  const double n2 = b.x0[0] * b.x0[0] + b.x0[1] * b.x0[1] + b.x0[2] * b.x0[2] + b.x0[3] * b.x0[3];
  Quat q0_inv(-b.x0[0] / n2, -b.x0[1] / n2, -b.x0[2] / n2, b.x0[3] / n2);
  Quat dq = qmul_raw(q0_inv, Quat::fromPtr(x));
  double sgn = (dq.w >= 0) ? 2.0 : -2.0;  // :346-350 (positify is the identity)
  dx[0] = sgn * dq.x; dx[1] = sgn * dq.y; dx[2] = sgn * dq.z;
}

static void priorResidual(const Window& w, const Prior& pr, std::vector<double>& dx, std::vector<double>& r) {
  dx.assign(pr.n, 0.0);
  for (const PriorBlock& b : pr.blocks) priorBlockDx(w, b, &dx[b.col]);
  r.assign(pr.n, 0.0);
  for (int i = 0; i < pr.n; ++i) {
    double s = pr.r[i];
    const double* Ji = &pr.J[size_t(i) * pr.n];
    for (int j = 0; j < pr.n; ++j) s += Ji[j] * dx[j];
    r[i] = s;
  }
}

// ---------------------------------------------------------------------------
// assembly

namespace {
struct LocalJ {  // merged per-factor Jacobian over unique camera indices
  int n = 0;
  int idx[64];
  double J[6][64];
  void clear() { n = 0; }
  int slot(int gi, int rows) {
    for (int a = 0; a < n; ++a)
      if (idx[a] == gi) return a;
    idx[n] = gi;
    for (int r = 0; r < rows; ++r) J[r][n] = 0.0;
    return n++;
  }
};
}  // namespace

static void accumulateCamera(NormalEq* ne, const LocalJ& L, int rows, const double* r, bool full) {
  const int np = ne->np;
  for (int a = 0; a < L.n; ++a) {
    double g = 0;
    for (int k = 0; k < rows; ++k) g += L.J[k][a] * r[k];
    ne->gc[L.idx[a]] += g;
  }
  if (!full) return;
  for (int a = 0; a < L.n; ++a)
    for (int b = a; b < L.n; ++b) {
      double h = 0;
      for (int k = 0; k < rows; ++k) h += L.J[k][a] * L.J[k][b];
      const int ia = std::min(L.idx[a], L.idx[b]), ib = std::max(L.idx[a], L.idx[b]);
      ne->Hcc[size_t(ia) * np + ib] += h;
    }
}

static double assembleRange(const Window& w, Window::Mode mode, NormalEq* ne, const std::vector<char>& cmask,
                            size_t img0, size_t img1, size_t imu0, size_t imu1, bool do_small) {
  const bool want_jac = mode != Window::kCost;
  const bool full = mode == Window::kFull;
  double cost = 0;
  const double* q = w.q.data();
  const double* p = w.p.data();
  LocalJ L;
  // [2] image factors (trajectory_manager.cpp:360-385 -> trajectory_estimator.cpp:293-332)
  for (size_t n = img0; n < img1; ++n) {
    const ImageObs& o = w.img[n];
    ImageEval e;
    EvaluateImage(w.grid, w.cal, q, p, w.rho[o.lm], w.ld, o, want_jac, e);
    if (!e.ok) continue;
    cost += ApplyLossImage(w.opt.cauchy_solve, e, want_jac);
    if (!want_jac) continue;
    L.clear();
    for (int side = 0; side < 2; ++side)
      for (int k = 0; k < 4; ++k)
        for (int c = 0; c < 3; ++c) {
          const int gr = w.idxKnot(int(e.s[side]) + k) + c, gp = gr + 3;
          int a = L.slot(gr, 2);
          if (!cmask[gr]) { L.J[0][a] += e.Jrot[side][k][c]; L.J[1][a] += e.Jrot[side][k][3 + c]; }
          a = L.slot(gp, 2);
          if (!cmask[gp]) { L.J[0][a] += e.Jpos[side][k][c]; L.J[1][a] += e.Jpos[side][k][3 + c]; }
        }
    const int nknot = L.n;
    {
      const int a = L.slot(w.idxLd(), 2);
      if (!cmask[w.idxLd()]) { L.J[0][a] += e.Jld[0]; L.J[1][a] += e.Jld[1]; }
    }
    accumulateCamera(ne, L, 2, e.r, full);
    const int l = o.lm;
    ne->gl[l] += e.Jrho[0] * e.r[0] + e.Jrho[1] * e.r[1];
    if (full) {
      ne->hl[l] += e.Jrho[0] * e.Jrho[0] + e.Jrho[1] * e.Jrho[1];
      double* Wl = &ne->W[ne->woff[l]];
      for (int a = 0; a < nknot; ++a) Wl[L.idx[a] - ne->lo[l]] += L.J[0][a] * e.Jrho[0] + L.J[1][a] * e.Jrho[1];
      ne->wld[l] += L.J[0][nknot] * e.Jrho[0] + L.J[1][nknot] * e.Jrho[1];
    }
  }
  // [3] IMU factors (trajectory_manager.cpp:388-417)
  for (size_t n = imu0; n < imu1; ++n) {
    const ImuObs& o = w.imu[n];
    ImuEval e;
    EvaluateImu(w.grid, w.cal, q, p, &w.bias[6 * o.bias_idx], &w.bias[6 * o.bias_idx + 3], o, want_jac, e);
    if (!e.ok) continue;
    double s = 0;
    for (int k = 0; k < 6; ++k) s += e.r[k] * e.r[k];
    cost += 0.5 * s;
    if (!want_jac) continue;
    L.clear();
    for (int k = 0; k < 4; ++k)
      for (int c = 0; c < 3; ++c) {
        const int gr = w.idxKnot(int(e.s) + k) + c, gp = gr + 3;
        int a = L.slot(gr, 6);
        if (!cmask[gr]) for (int r = 0; r < 6; ++r) L.J[r][a] += e.Jrot[k][3 * r + c];
        a = L.slot(gp, 6);
        if (!cmask[gp]) for (int r = 0; r < 6; ++r) L.J[r][a] += e.Jpos[k][3 * r + c];
      }
    for (int c = 0; c < 3; ++c) {
      const int gg = w.idxBias(o.bias_idx) + c, ga = gg + 3;
      int a = L.slot(gg, 6);
      if (!cmask[gg]) L.J[c][a] += e.Jbg[c];
      a = L.slot(ga, 6);
      if (!cmask[ga]) L.J[3 + c][a] += e.Jba[3 + c];
    }
    accumulateCamera(ne, L, 6, e.r, full);
  }
  if (!do_small) return cost;
  // [4] bias random-walk factors (trajectory_manager.cpp:420-451)
  for (const BiasObs& o : w.biasf) {
    double r[6];
    EvaluateBias(&w.bias[6 * o.i], &w.bias[6 * o.j], o, r);
    double s = 0;
    for (int k = 0; k < 6; ++k) s += r[k] * r[k];
    cost += 0.5 * s;
    if (!want_jac) continue;
    for (int k = 0; k < 6; ++k) {
      const int gi = w.idxBias(o.i) + k, gj = w.idxBias(o.j) + k;
      const double ji = cmask[gi] ? 0.0 : -o.sqrt_info[k], jj = cmask[gj] ? 0.0 : o.sqrt_info[k];
      ne->gc[gi] += ji * r[k];
      ne->gc[gj] += jj * r[k];
      if (full) {
        const int np = ne->np;
        ne->Hcc[size_t(gi) * np + gi] += ji * ji;
        ne->Hcc[size_t(gj) * np + gj] += jj * jj;
        ne->Hcc[size_t(std::min(gi, gj)) * np + std::max(gi, gj)] += ji * jj;
      }
    }
  }
  // [1] prior (trajectory_manager.cpp:353-357)
  if (w.prior.valid()) {
    const Prior& pr = w.prior;
    std::vector<double> dx, r;
    priorResidual(w, pr, dx, r);
    double s = 0;
    for (int i = 0; i < pr.n; ++i) s += r[i] * r[i];
    cost += 0.5 * s;
    if (want_jac) {
      std::vector<int> col2g(pr.n, -1);
      for (const PriorBlock& b : pr.blocks) {
        const int base = priorBlockBase(w, b);
        for (int d = 0; d < blockLocalSize(b.type); ++d)
          if (base >= 0 && !cmask[base + d]) col2g[b.col + d] = base + d;
      }
      for (int j = 0; j < pr.n; ++j) {
        if (col2g[j] < 0) continue;
        double g = 0;
        for (int i = 0; i < pr.n; ++i) g += pr.J[size_t(i) * pr.n + j] * r[i];
        ne->gc[col2g[j]] += g;
      }
      if (full) {
        const int np = ne->np;
        for (int a = 0; a < pr.n; ++a) {
          if (col2g[a] < 0) continue;
          for (int b = a; b < pr.n; ++b) {
            if (col2g[b] < 0) continue;
            double h = 0;
            for (int i = 0; i < pr.n; ++i) h += pr.J[size_t(i) * pr.n + a] * pr.J[size_t(i) * pr.n + b];
            const int ia = std::min(col2g[a], col2g[b]), ib = std::max(col2g[a], col2g[b]);
            ne->Hcc[size_t(ia) * np + ib] += h;
          }
        }
      }
    }
  }
  return cost;
}

double Window::assemble(Mode mode, NormalEq* ne) const {
  const std::vector<char> cmask = constMask();
  if (mode != kCost) {
    std::fill(ne->gc.begin(), ne->gc.end(), 0.0);
    std::fill(ne->gl.begin(), ne->gl.end(), 0.0);
    if (mode == kFull) {
      std::fill(ne->Hcc.begin(), ne->Hcc.end(), 0.0);
      std::fill(ne->hl.begin(), ne->hl.end(), 0.0);
      std::fill(ne->wld.begin(), ne->wld.end(), 0.0);
      std::fill(ne->W.begin(), ne->W.end(), 0.0);
    }
  }
  double cost = 0;
  int nt = std::max(1, opt.num_threads);
  if (nt == 1) {
    cost = assembleRange(*this, mode, ne, cmask, 0, img.size(), 0, imu.size(), true);
  } else {
    // Landmark-contiguous chunks: observations are expected sorted by landmark,
    // so per-landmark accumulators of different chunks rarely collide; camera
    // blocks are reduced from per-thread copies.
    std::vector<NormalEq> part(nt);
    std::vector<double> costs(nt, 0.0);
    auto worker = [&](int t) {
      NormalEq& me = part[t];
      me.nK = ne->nK; me.nB = ne->nB; me.nL = ne->nL; me.np = ne->np;
      me.lo = ne->lo; me.hi = ne->hi; me.woff = ne->woff;
      if (mode != kCost) {
        me.gc.assign(ne->np, 0.0);
        me.gl.assign(ne->nL, 0.0);
        if (mode == kFull) {
          me.Hcc.assign(size_t(ne->np) * ne->np, 0.0);
          me.hl.assign(ne->nL, 0.0);
          me.wld.assign(ne->nL, 0.0);
          me.W.assign(ne->W.size(), 0.0);
        }
      }
      const size_t i0 = img.size() * t / nt, i1 = img.size() * (t + 1) / nt;
      const size_t m0 = imu.size() * t / nt, m1 = imu.size() * (t + 1) / nt;
      costs[t] = assembleRange(*this, mode, &me, cmask, i0, i1, m0, m1, t == 0);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto& th : pool) th.join();
    for (int t = 0; t < nt; ++t) {
      cost += costs[t];
      if (mode == kCost) continue;
      for (size_t i = 0; i < ne->gc.size(); ++i) ne->gc[i] += part[t].gc[i];
      for (size_t i = 0; i < ne->gl.size(); ++i) ne->gl[i] += part[t].gl[i];
      if (mode != kFull) continue;
      for (size_t i = 0; i < ne->Hcc.size(); ++i) ne->Hcc[i] += part[t].Hcc[i];
      for (size_t i = 0; i < ne->hl.size(); ++i) ne->hl[i] += part[t].hl[i];
      for (size_t i = 0; i < ne->wld.size(); ++i) ne->wld[i] += part[t].wld[i];
      for (size_t i = 0; i < ne->W.size(); ++i) ne->W[i] += part[t].W[i];
    }
  }
  ne->cost = cost;
  return cost;
}

// ---------------------------------------------------------------------------
// LM (Ceres 1.14 trust_region_minimizer.cc / levenberg_marquardt_strategy.cc)

namespace {

struct StateVec {
  std::vector<double> q, p, bias, rho;
  double ld;
};

StateVec getState(const Window& w) { return {w.q, w.p, w.bias, w.rho, w.ld}; }
void setState(Window& w, const StateVec& s) { w.q = s.q; w.p = s.p; w.bias = s.bias; w.rho = s.rho; w.ld = s.ld; }

// Evaluator::Plus: x (+) delta with the SO(3) right-multiplicative update
// (ceres_local_param.h:137-145) and box projection (parameter_block.h Plus).
void plusState(const Window& w, const StateVec& x, const std::vector<double>& dc, const std::vector<double>& dl,
               double alpha, StateVec& out) {
  out = x;
  const int nK = w.nK(), nB = w.nB(), nL = w.nL();
  for (int k = 0; k < nK; ++k) {
    const double* d = &dc[w.idxKnot(k)];
    if (d[0] != 0.0 || d[1] != 0.0 || d[2] != 0.0) {
      Quat qn = so3_mul(Quat::fromPtr(&x.q[4 * k]), so3_exp(Vec3(alpha * d[0], alpha * d[1], alpha * d[2])));
      qn.toPtr(&out.q[4 * k]);
    }
    for (int c = 0; c < 3; ++c) out.p[3 * k + c] = x.p[3 * k + c] + alpha * d[3 + c];
  }
  for (int b = 0; b < nB; ++b)
    for (int c = 0; c < 6; ++c) out.bias[6 * b + c] = x.bias[6 * b + c] + alpha * dc[w.idxBias(b) + c];
  out.ld = x.ld + alpha * dc[w.idxLd()];
  if (!w.opt.fix_ld) out.ld = std::min(std::max(out.ld, w.opt.ld_lower), w.opt.ld_upper);
  for (int l = 0; l < nL; ++l) out.rho[l] = x.rho[l] + alpha * dl[l];
}

struct ActiveSet {
  std::vector<char> cam;  // np: non-constant && touched
  std::vector<char> lm;   // nL
};

double ambientNorm(const Window& w, const StateVec& a, const StateVec* b, const ActiveSet& act) {
  double s = 0;
  auto add = [&](double va, double vb) { const double d = va - vb; s += d * d; };
  for (int k = 0; k < w.nK(); ++k) {
    if (act.cam[w.idxKnot(k)])
      for (int c = 0; c < 4; ++c) add(a.q[4 * k + c], b ? b->q[4 * k + c] : 0.0);
    if (act.cam[w.idxKnot(k) + 3])
      for (int c = 0; c < 3; ++c) add(a.p[3 * k + c], b ? b->p[3 * k + c] : 0.0);
  }
  for (int n = 0; n < w.nB(); ++n)
    for (int c = 0; c < 6; ++c)
      if (act.cam[w.idxBias(n) + c]) add(a.bias[6 * n + c], b ? b->bias[6 * n + c] : 0.0);
  if (act.cam[w.idxLd()]) add(a.ld, b ? b->ld : 0.0);
  for (int l = 0; l < w.nL(); ++l)
    if (act.lm[l]) add(a.rho[l], b ? b->rho[l] : 0.0);
  return std::sqrt(s);
}

struct FunctionSample {
  double x = 0, value = 0, gradient = 0;
  bool value_is_valid = false, gradient_is_valid = false;
};

// polynomial.cc FindInterpolatingPolynomial / MinimizeInterpolatingPolynomial
double minimizeInterpolatingPolynomial(const std::vector<FunctionSample>& samples, double x_min, double x_max) {
  int nc = 0;
  for (const auto& s : samples) nc += int(s.value_is_valid) + int(s.gradient_is_valid);
  const int degree = nc - 1;
  std::vector<double> lhs(size_t(nc) * nc, 0.0), rhs(nc, 0.0);
  int row = 0;
  for (const auto& s : samples) {
    if (s.value_is_valid) {
      for (int j = 0; j <= degree; ++j) lhs[row * nc + j] = std::pow(s.x, degree - j);
      rhs[row++] = s.value;
    }
    if (s.gradient_is_valid) {
      for (int j = 0; j < degree; ++j) lhs[row * nc + j] = (degree - j) * std::pow(s.x, degree - j - 1);
      rhs[row++] = s.gradient;
    }
  }
  const std::vector<double> poly = SolveDenseFullPivot(lhs, rhs, nc);
  // MinimizePolynomial
  double opt_x = (x_min + x_max) / 2.0;
  double opt_v = EvaluatePolynomial(poly, opt_x);
  const double vmin = EvaluatePolynomial(poly, x_min);
  if (vmin < opt_v) { opt_v = vmin; opt_x = x_min; }
  const double vmax = EvaluatePolynomial(poly, x_max);
  if (vmax < opt_v) { opt_v = vmax; opt_x = x_max; }
  if (poly.size() > 2) {
    std::vector<double> roots;
    if (FindPolynomialRootsReal(DifferentiatePolynomial(poly), roots)) {
      for (double root : roots) {
        if (root < x_min || root > x_max) continue;
        const double v = EvaluatePolynomial(poly, root);
        if (v < opt_v) { opt_v = v; opt_x = root; }
      }
    }
  }
  for (const auto& s : samples) {
    if (s.x < x_min || s.x > x_max) continue;
    if (s.value_is_valid && s.value < opt_v) { opt_x = s.x; opt_v = s.value; }
  }
  return opt_x;
}

}  // namespace

Summary Window::solve(int max_iterations) {
  Summary sum;
  const double t_begin = now_s();
  // Ceres 1.14 Solver::Options defaults (solver.h)
  const double initial_radius = 1e4, max_radius = 1e16, min_radius = 1e-32;
  const double min_relative_decrease = 1e-3;
  const double min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
  const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  const int max_num_consecutive_invalid_steps = 5;
  // line search defaults used by the bounds-constrained path
  const double ls_sufficient_decrease = 1e-4, ls_max_step_contraction = 1e-3, ls_min_step_contraction = 0.6;
  const double ls_min_step_size = 1e-9;
  const int ls_max_iterations = 20;

  NormalEq ne, ne_trial;
  buildStructure(&ne);
  const int np_ = np(), nL_ = nL();
  const std::vector<char> cmask = constMask();
  const std::vector<char> touched = touchedMask();
  ActiveSet act;
  act.cam.assign(np_, 0);
  act.lm.assign(nL_, 0);
  for (int i = 0; i < np_; ++i) act.cam[i] = touched[i] && !cmask[i];
  for (int l = 0; l < nL_; ++l) act.lm[l] = touched[np_ + l];
  const bool is_constrained = !opt.fix_ld && touched[idxLd()];

  StateVec x = getState(*this);
  if (is_constrained) x.ld = std::min(std::max(x.ld, opt.ld_lower), opt.ld_upper);  // IterationZero: Plus(x, 0)
  setState(*this, x);
  double x_norm = ambientNorm(*this, x, nullptr, act);

  // iteration 0: residuals + Jacobian at x
  double t0 = now_s();
  double x_cost = assemble(kFull, &ne);
  sum.t_eval_s += now_s() - t0;
  sum.num_jacobian_evals++;
  sum.initial_cost = x_cost;

  // Jacobi scaling, computed once (trust_region_minimizer.cc EvaluateGradientAndJacobian)
  std::vector<double> sc(np_), sl(nL_);
  for (int i = 0; i < np_; ++i) sc[i] = 1.0 / (1.0 + std::sqrt(ne.Hcc[size_t(i) * np_ + i]));
  for (int l = 0; l < nL_; ++l) sl[l] = 1.0 / (1.0 + std::sqrt(ne.hl[l]));

  auto gradientMaxNorm = [&](const NormalEq& e, const StateVec& xs) {
    // |x - Plus(x, -g)|_inf; tangent-space norm for SO(3) blocks (documented simplification)
    double m = 0;
    for (int i = 0; i < np_; ++i) {
      if (!act.cam[i]) continue;
      double v = std::fabs(e.gc[i]);
      if (i == idxLd() && !opt.fix_ld) {
        const double proj = std::min(std::max(xs.ld - e.gc[i], opt.ld_lower), opt.ld_upper);
        v = std::fabs(xs.ld - proj);
      }
      m = std::max(m, v);
    }
    for (int l = 0; l < nL_; ++l)
      if (act.lm[l]) m = std::max(m, std::fabs(e.gl[l]));
    return m;
  };

  double radius = initial_radius, decrease_factor = 2.0;
  int num_consecutive_invalid = 0;
  bool last_step_successful = true;
  double gmax = gradientMaxNorm(ne, x);
  sum.num_successful_steps = 1;  // iteration 0 counts as successful in Ceres' summary

  std::vector<double> M(size_t(np_) * np_), rhs(np_), yc(np_), yl(nL_), dc(np_), dl(nL_), hh(nL_), tmp;
  int iter = 0;
  while (true) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (iter >= max_iterations) { sum.termination = kNoConvergence; break; }
    if (last_step_successful && gmax <= gradient_tolerance) { sum.termination = kConvergenceGradient; break; }
    if (radius < min_radius) { sum.termination = kMinRadius; break; }
    ++iter;

    // ---- ComputeTrustRegionStep: (Js'Js + D'D) y = Js' r, step = -y ----
    t0 = now_s();
    for (int i = 0; i < np_; ++i) {
      for (int j = i; j < np_; ++j) {
        const double v = sc[i] * ne.Hcc[size_t(i) * np_ + j] * sc[j];
        M[size_t(i) * np_ + j] = v;
        M[size_t(j) * np_ + i] = v;
      }
      const double diag = std::min(std::max(M[size_t(i) * np_ + i], min_lm_diagonal), max_lm_diagonal);
      M[size_t(i) * np_ + i] += diag / radius;
      rhs[i] = sc[i] * ne.gc[i];
    }
    const int ild = idxLd();
    for (int l = 0; l < nL_; ++l) {
      const double hs = sl[l] * sl[l] * ne.hl[l];
      const double diag = std::min(std::max(hs, min_lm_diagonal), max_lm_diagonal);
      hh[l] = hs + diag / radius;
      const double inv = 1.0 / hh[l];
      const int lo = ne.lo[l], hi = ne.hi[l];
      const double* Wl = &ne.W[ne.woff[l]];
      const double gls = sl[l] * ne.gl[l];
      // Ws = sl * W o sc over [lo,hi) plus the line-delay entry
      tmp.resize(size_t(hi - lo) + 1);
      for (int a = lo; a < hi; ++a) tmp[a - lo] = sl[l] * Wl[a - lo] * sc[a];
      const double wl = sl[l] * ne.wld[l] * sc[ild];
      tmp[hi - lo] = wl;
      for (int a = lo; a < hi; ++a) {
        const double wa = tmp[a - lo] * inv;
        if (wa == 0.0) continue;
        double* Ma = &M[size_t(a) * np_];
        for (int b = lo; b < hi; ++b) Ma[b] -= wa * tmp[b - lo];
        Ma[ild] -= wa * wl;
        M[size_t(ild) * np_ + a] -= wa * wl;
        rhs[a] -= wa * gls;
      }
      M[size_t(ild) * np_ + ild] -= wl * inv * wl;
      rhs[ild] -= wl * inv * gls;
    }
    for (int i = 0; i < np_; ++i)
      if (cmask[i]) {
        for (int j = 0; j < np_; ++j) M[size_t(i) * np_ + j] = M[size_t(j) * np_ + i] = 0.0;
        M[size_t(i) * np_ + i] = 1.0;
        rhs[i] = 0.0;
      }
    sum.t_schur_s += now_s() - t0;
    t0 = now_s();
    bool solved = cholesky_lower(M.data(), np_);
    if (solved) {
      yc = rhs;
      cholesky_solve(M.data(), np_, yc.data());
      for (int i = 0; i < np_ && solved; ++i) solved = std::isfinite(yc[i]);
    }
    sum.t_solve_s += now_s() - t0;
    sum.num_linear_solves++;

    bool step_is_valid = false;
    double model_cost_change = 0;
    if (solved) {
      for (int l = 0; l < nL_; ++l) {
        const int lo = ne.lo[l], hi = ne.hi[l];
        const double* Wl = &ne.W[ne.woff[l]];
        double s = sl[l] * ne.gl[l];
        for (int a = lo; a < hi; ++a) s -= sl[l] * Wl[a - lo] * sc[a] * yc[a];
        s -= sl[l] * ne.wld[l] * sc[ild] * yc[ild];
        yl[l] = s / hh[l];
      }
      for (int i = 0; i < np_; ++i) dc[i] = -sc[i] * yc[i];
      for (int l = 0; l < nL_; ++l) dl[l] = -sl[l] * yl[l];
      // model_cost_change = -(J d)'(r + J d / 2) = -g'd - d'Hd/2
      double gd = 0, dHd = 0;
      for (int i = 0; i < np_; ++i) gd += ne.gc[i] * dc[i];
      for (int l = 0; l < nL_; ++l) gd += ne.gl[l] * dl[l];
      for (int i = 0; i < np_; ++i) {
        if (dc[i] == 0.0) continue;
        double s = 0.5 * ne.Hcc[size_t(i) * np_ + i] * dc[i];
        for (int j = i + 1; j < np_; ++j) s += ne.Hcc[size_t(i) * np_ + j] * dc[j];
        dHd += 2.0 * dc[i] * s;
      }
      for (int l = 0; l < nL_; ++l) {
        const int lo = ne.lo[l], hi = ne.hi[l];
        const double* Wl = &ne.W[ne.woff[l]];
        double wd = ne.wld[l] * dc[ild];
        for (int a = lo; a < hi; ++a) wd += Wl[a - lo] * dc[a];
        dHd += 2.0 * dl[l] * wd + ne.hl[l] * dl[l] * dl[l];
      }
      model_cost_change = -gd - 0.5 * dHd;
      step_is_valid = model_cost_change > 0.0;
    }
    if (!step_is_valid) {
      // HandleInvalidStep -> LevenbergMarquardtStrategy::StepIsInvalid == StepRejected(0)
      ++sum.num_unsuccessful_steps;
      last_step_successful = false;
      if (++num_consecutive_invalid >= max_num_consecutive_invalid_steps) { sum.termination = kFailure; break; }
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      continue;
    }
    num_consecutive_invalid = 0;

    // ---- DoLineSearch (bounds constrained problems only) ----
    double step_scale = 1.0;
    if (is_constrained) {
      double initial_gradient = 0;
      for (int i = 0; i < np_; ++i) initial_gradient += ne.gc[i] * dc[i];
      for (int l = 0; l < nL_; ++l) initial_gradient += ne.gl[l] * dl[l];
      double dir_max = 0;
      for (int i = 0; i < np_; ++i) dir_max = std::max(dir_max, std::fabs(dc[i]));
      for (int l = 0; l < nL_; ++l) dir_max = std::max(dir_max, std::fabs(dl[l]));
      if (ne_trial.np == 0) buildStructure(&ne_trial);
      auto evalSample = [&](double a, FunctionSample& out) {
        out = FunctionSample();
        out.x = a;
        StateVec xt;
        plusState(*this, x, dc, dl, a, xt);
        setState(*this, xt);
        double t1 = now_s();
        out.value = assemble(kGradient, &ne_trial);
        sum.t_eval_s += now_s() - t1;
        sum.num_jacobian_evals++;
        setState(*this, x);
        if (!std::isfinite(out.value)) return;
        out.value_is_valid = true;
        double gsum = 0;
        for (int i = 0; i < np_; ++i) gsum += ne_trial.gc[i] * dc[i];
        for (int l = 0; l < nL_; ++l) gsum += ne_trial.gl[l] * dl[l];
        out.gradient = gsum;
        out.gradient_is_valid = std::isfinite(gsum);
      };
      FunctionSample initial;
      initial.x = 0; initial.value = x_cost; initial.gradient = initial_gradient;
      initial.value_is_valid = initial.gradient_is_valid = true;
      FunctionSample previous, current;
      evalSample(1.0, current);
      bool success = true;
      int ls_iters = 0;
      while (!current.value_is_valid ||
             current.value > x_cost + ls_sufficient_decrease * initial_gradient * current.x) {
        ++ls_iters;
        if (ls_iters >= ls_max_iterations) { success = false; break; }
        double step_size;
        const double smin = ls_max_step_contraction * current.x, smax = ls_min_step_contraction * current.x;
        if (!current.value_is_valid) {
          step_size = std::min(std::max(current.x * 0.5, smin), smax);
        } else {
          std::vector<FunctionSample> samples;
          samples.push_back(initial);
          samples.push_back(current);
          if (previous.value_is_valid) samples.push_back(previous);
          step_size = minimizeInterpolatingPolynomial(samples, smin, smax);
        }
        if (step_size * dir_max < ls_min_step_size) { success = false; break; }
        previous = current;
        evalSample(step_size, current);
      }
      sum.num_line_search_steps += ls_iters;
      if (success) step_scale = current.x;
    }

    // ---- ComputeCandidatePointAndEvaluateCost ----
    StateVec cand;
    plusState(*this, x, dc, dl, step_scale, cand);
    setState(*this, cand);
    t0 = now_s();
    const double candidate_cost = assemble(kCost, &ne);
    sum.t_eval_s += now_s() - t0;
    sum.num_cost_evals++;

    // ParameterToleranceReached
    const double step_norm = ambientNorm(*this, x, &cand, act);
    if (step_norm <= parameter_tolerance * (x_norm + parameter_tolerance)) {
      setState(*this, x);
      sum.termination = kConvergenceParameter;
      break;
    }
    // FunctionToleranceReached
    const double cost_change = x_cost - candidate_cost;
    if (std::fabs(cost_change) <= function_tolerance * x_cost) {
      setState(*this, x);
      sum.termination = kConvergenceFunction;
      break;
    }
    const double relative_decrease = cost_change / model_cost_change;
    if (relative_decrease > min_relative_decrease) {
      // HandleSuccessfulStep
      x = cand;
      x_norm = ambientNorm(*this, x, nullptr, act);
      t0 = now_s();
      x_cost = assemble(kFull, &ne);
      sum.t_eval_s += now_s() - t0;
      sum.num_jacobian_evals++;
      gmax = gradientMaxNorm(ne, x);
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3));
      radius = std::min(max_radius, radius);
      decrease_factor = 2.0;
      last_step_successful = true;
      ++sum.num_successful_steps;
    } else {
      // HandleUnsuccessfulStep
      setState(*this, x);
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      last_step_successful = false;
      ++sum.num_unsuccessful_steps;
    }
  }
  setState(*this, x);
  sum.iterations = iter;
  sum.final_cost = x_cost;
  sum.final_radius = radius;
  sum.t_total_s = now_s() - t_begin;
  return sum;
}

// ---------------------------------------------------------------------------
// marginalization (marginalization_factor.cpp:85-265)

namespace {
struct BlockKey {
  int type, index;
  bool operator<(const BlockKey& o) const {
    // deterministic order: knots (rot, pos per knot), bias (bg, ba per node), ld, rho
    auto rank = [](const BlockKey& k) {
      switch (k.type) {
        case kBlkRot: return std::make_pair(0, 2 * k.index);
        case kBlkPos: return std::make_pair(0, 2 * k.index + 1);
        case kBlkBg: return std::make_pair(1, 2 * k.index);
        case kBlkBa: return std::make_pair(1, 2 * k.index + 1);
        case kBlkLd: return std::make_pair(2, 0);
        default: return std::make_pair(3, k.index);
      }
    };
    return rank(*this) < rank(o);
  }
};
struct BlockInfo {
  bool dropped = false;
  int pos = -1;
};
struct RecordedFactor {
  int rows;
  std::vector<BlockKey> blocks;
  std::vector<double> J;  // rows x sum(local sizes), column-blocks in `blocks` order
  std::vector<double> r;
};
}  // namespace

bool Window::marginalize() {
  new_prior = Prior();
  if (!opt.is_marg_state) return false;
  const int later = opt.ctrl_to_be_opt_later, nowk = opt.ctrl_to_be_opt_now;
  const bool drop_knots = later > nowk;  // trajectory_estimator.cpp:161
  std::map<BlockKey, BlockInfo> blocks;
  std::vector<RecordedFactor> factors;
  auto touch = [&](const BlockKey& k, bool drop) {
    BlockInfo& b = blocks[k];
    b.dropped = b.dropped || drop;
  };
  const double* qd = q.data();
  const double* pd = p.data();

  // [1] old prior (trajectory_manager.cpp:166-203)
  if (prior.valid()) {
    bool any_drop = false;
    for (const PriorBlock& b : prior.blocks) {
      const bool isknot = (b.type == kBlkRot || b.type == kBlkPos);
      if ((isknot && b.index >= nowk && b.index < later) || ((b.type == kBlkBg || b.type == kBlkBa) && b.index == 0))
        any_drop = true;
    }
    if (any_drop) {
      RecordedFactor f;
      f.rows = prior.n;
      std::vector<double> dx;
      priorResidual(*this, prior, dx, f.r);
      // columns re-packed in block order
      int ncols = 0;
      for (const PriorBlock& b : prior.blocks) ncols += blockLocalSize(b.type);
      f.J.assign(size_t(prior.n) * ncols, 0.0);
      int c0 = 0;
      for (const PriorBlock& b : prior.blocks) {
        const bool isknot = (b.type == kBlkRot || b.type == kBlkPos);
        const bool drop = (isknot && b.index >= nowk && b.index < later) ||
                          ((b.type == kBlkBg || b.type == kBlkBa) && b.index == 0);
        f.blocks.push_back({b.type, b.index});
        touch({b.type, b.index}, drop);
        for (int d = 0; d < blockLocalSize(b.type); ++d)
          for (int i = 0; i < prior.n; ++i) f.J[size_t(i) * ncols + c0 + d] = prior.J[size_t(i) * prior.n + b.col + d];
        c0 += blockLocalSize(b.type);
      }
      factors.push_back(std::move(f));
    }
  }
  // [2] image factors flagged marg (trajectory_manager.cpp:206-236, estimator.cpp:325-331)
  for (const ImageObs& o : img) {
    if (!o.marg) continue;
    ImageEval e;
    EvaluateImage(grid, cal, qd, pd, rho[o.lm], ld, o, true, e);
    if (!e.ok) continue;
    ApplyLossImage(opt.cauchy_marg, e, true);
    // parameter blocks: merged padded windows (rot then pos), inverse depth, line delay
    int f0, l0, f1, l1;
    knotWindow(o.ti, f0, l0);
    knotWindow(o.tj, f1, l1);
    std::vector<int> knots;
    for (int k = f0; k <= l0; ++k) knots.push_back(k);
    for (int k = f1; k <= l1; ++k)
      if (std::find(knots.begin(), knots.end(), k) == knots.end()) knots.push_back(k);
    RecordedFactor f;
    f.rows = 2;
    f.r = {e.r[0], e.r[1]};
    const int ncols = int(knots.size()) * 6 + 2;
    f.J.assign(size_t(2) * ncols, 0.0);
    int c0 = 0;
    for (int k : knots) {
      f.blocks.push_back({kBlkRot, k});
      touch({kBlkRot, k}, drop_knots && k < later);
      for (int side = 0; side < 2; ++side) {
        const int kk = k - int(e.s[side]);
        if (kk < 0 || kk > 3) continue;
        for (int c = 0; c < 3; ++c) {
          f.J[0 * ncols + c0 + c] += e.Jrot[side][kk][c];
          f.J[1 * ncols + c0 + c] += e.Jrot[side][kk][3 + c];
        }
      }
      c0 += 3;
    }
    for (int k : knots) {
      f.blocks.push_back({kBlkPos, k});
      touch({kBlkPos, k}, drop_knots && k < later);
      for (int side = 0; side < 2; ++side) {
        const int kk = k - int(e.s[side]);
        if (kk < 0 || kk > 3) continue;
        for (int c = 0; c < 3; ++c) {
          f.J[0 * ncols + c0 + c] += e.Jpos[side][kk][c];
          f.J[1 * ncols + c0 + c] += e.Jpos[side][kk][3 + c];
        }
      }
      c0 += 3;
    }
    f.blocks.push_back({kBlkRho, o.lm});
    touch({kBlkRho, o.lm}, true);
    f.J[0 * ncols + c0] = e.Jrho[0];
    f.J[1 * ncols + c0] = e.Jrho[1];
    ++c0;
    f.blocks.push_back({kBlkLd, 0});
    touch({kBlkLd, 0}, false);
    f.J[0 * ncols + c0] = e.Jld[0];
    f.J[1 * ncols + c0] = e.Jld[1];
    factors.push_back(std::move(f));
  }
  // [3] IMU factors flagged marg (trajectory_manager.cpp:239-253, estimator.cpp:249-257)
  for (const ImuObs& o : imu) {
    if (!o.marg) continue;
    ImuEval e;
    EvaluateImu(grid, cal, qd, pd, &bias[6 * o.bias_idx], &bias[6 * o.bias_idx + 3], o, true, e);
    if (!e.ok) continue;
    RecordedFactor f;
    f.rows = 6;
    f.r.assign(e.r, e.r + 6);
    const int ncols = 30;
    f.J.assign(size_t(6) * ncols, 0.0);
    int c0 = 0;
    for (int k = 0; k < 4; ++k) {
      const int gk = int(e.s) + k;
      f.blocks.push_back({kBlkRot, gk});
      touch({kBlkRot, gk}, drop_knots && gk < later);
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 3; ++c) f.J[size_t(r) * ncols + c0 + c] = e.Jrot[k][3 * r + c];
      c0 += 3;
    }
    for (int k = 0; k < 4; ++k) {
      const int gk = int(e.s) + k;
      f.blocks.push_back({kBlkPos, gk});
      touch({kBlkPos, gk}, drop_knots && gk < later);
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 3; ++c) f.J[size_t(r) * ncols + c0 + c] = e.Jpos[k][3 * r + c];
      c0 += 3;
    }
    f.blocks.push_back({kBlkBg, o.bias_idx});
    touch({kBlkBg, o.bias_idx}, true);
    for (int c = 0; c < 3; ++c) f.J[size_t(c) * ncols + c0 + c] = e.Jbg[c];
    c0 += 3;
    f.blocks.push_back({kBlkBa, o.bias_idx});
    touch({kBlkBa, o.bias_idx}, true);
    for (int c = 0; c < 3; ++c) f.J[size_t(3 + c) * ncols + c0 + c] = e.Jba[3 + c];
    factors.push_back(std::move(f));
  }
  // [4] bias factors flagged marg (trajectory_manager.cpp:256-263, estimator.cpp:280-285)
  for (const BiasObs& o : biasf) {
    if (!o.marg) continue;
    RecordedFactor f;
    f.rows = 6;
    f.r.resize(6);
    EvaluateBias(&bias[6 * o.i], &bias[6 * o.j], o, f.r.data());
    const int ncols = 12;
    f.J.assign(size_t(6) * ncols, 0.0);
    // block order: bg_i, bg_j, ba_i, ba_j ; drop {0, 2}
    f.blocks = {{kBlkBg, o.i}, {kBlkBg, o.j}, {kBlkBa, o.i}, {kBlkBa, o.j}};
    touch({kBlkBg, o.i}, true);
    touch({kBlkBg, o.j}, false);
    touch({kBlkBa, o.i}, true);
    touch({kBlkBa, o.j}, false);
    for (int c = 0; c < 3; ++c) {
      f.J[size_t(c) * ncols + 0 + c] = -o.sqrt_info[c];
      f.J[size_t(c) * ncols + 3 + c] = o.sqrt_info[c];
      f.J[size_t(3 + c) * ncols + 6 + c] = -o.sqrt_info[3 + c];
      f.J[size_t(3 + c) * ncols + 9 + c] = o.sqrt_info[3 + c];
    }
    factors.push_back(std::move(f));
  }
  if (factors.empty()) return false;

  // index dropped blocks first, then kept (marginalize():180-195)
  int pos = 0;
  for (auto& kv : blocks)
    if (kv.second.dropped) { kv.second.pos = pos; pos += blockLocalSize(kv.first.type); }
  const int m = pos;
  for (auto& kv : blocks)
    if (!kv.second.dropped) { kv.second.pos = pos; pos += blockLocalSize(kv.first.type); }
  const int n = pos - m;
  if (n <= 0) return false;

  std::vector<double> A(size_t(pos) * pos, 0.0), b(pos, 0.0);
  for (const RecordedFactor& f : factors) {
    int ncols = 0;
    std::vector<int> gcol;
    for (const BlockKey& k : f.blocks) {
      const int base = blocks[k].pos;
      for (int d = 0; d < blockLocalSize(k.type); ++d) gcol.push_back(base + d);
      ncols += blockLocalSize(k.type);
    }
    for (int a = 0; a < ncols; ++a) {
      double g = 0;
      for (int r = 0; r < f.rows; ++r) g += f.J[size_t(r) * ncols + a] * f.r[r];
      b[gcol[a]] += g;
      for (int c = 0; c < ncols; ++c) {
        double h = 0;
        for (int r = 0; r < f.rows; ++r) h += f.J[size_t(r) * ncols + a] * f.J[size_t(r) * ncols + c];
        A[size_t(gcol[a]) * pos + gcol[c]] += h;
      }
    }
  }
  // Amm pseudo-inverse through the eigen-decomposition (:240-244), eps = 1e-30
  const double eps = 1e-30;
  std::vector<double> Amm(size_t(m) * m), ev, V;
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) Amm[size_t(i) * m + j] = 0.5 * (A[size_t(i) * pos + j] + A[size_t(j) * pos + i]);
  std::vector<double> Amm_inv(size_t(m) * m, 0.0);
  if (m > 0) {
    jacobi_eigh(Amm, m, ev, V);
    for (int k = 0; k < m; ++k) {
      if (!(ev[k] > eps)) continue;
      const double inv = 1.0 / ev[k];
      for (int i = 0; i < m; ++i) {
        const double vi = V[size_t(i) * m + k] * inv;
        for (int j = 0; j < m; ++j) Amm_inv[size_t(i) * m + j] += vi * V[size_t(j) * m + k];
      }
    }
  }
  // A' = Arr - Arm Amm^-1 Amr ; b' = brr - Arm Amm^-1 bmm (:246-252)
  std::vector<double> T(size_t(n) * m, 0.0);  // Arm * Amm_inv
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < m; ++k) {
      const double a = A[size_t(m + i) * pos + k];
      if (a == 0.0) continue;
      for (int j = 0; j < m; ++j) T[size_t(i) * m + j] += a * Amm_inv[size_t(k) * m + j];
    }
  std::vector<double> Ap(size_t(n) * n), bp(n);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) {
      double s = A[size_t(m + i) * pos + (m + j)];
      for (int k = 0; k < m; ++k) s -= T[size_t(i) * m + k] * A[size_t(k) * pos + (m + j)];
      Ap[size_t(i) * n + j] = s;
    }
    double s = b[m + i];
    for (int k = 0; k < m; ++k) s -= T[size_t(i) * m + k] * b[k];
    bp[i] = s;
  }
  // Eigen's SelfAdjointEigenSolver reads the lower triangle only (:254)
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) Ap[size_t(i) * n + j] = Ap[size_t(j) * n + i];
  std::vector<double> ev2, V2;
  jacobi_eigh(Ap, n, ev2, V2);
  new_prior.n = n;
  new_prior.J.assign(size_t(n) * n, 0.0);
  new_prior.r.assign(n, 0.0);
  for (int k = 0; k < n; ++k) {
    const double S = ev2[k] > eps ? ev2[k] : 0.0;
    const double Sinv = ev2[k] > eps ? 1.0 / ev2[k] : 0.0;
    const double ss = std::sqrt(S), sis = std::sqrt(Sinv);
    double vb = 0;
    for (int i = 0; i < n; ++i) {
      new_prior.J[size_t(k) * n + i] = ss * V2[size_t(i) * n + k];
      vb += V2[size_t(i) * n + k] * bp[i];
    }
    new_prior.r[k] = sis * vb;
  }
  for (const auto& kv : blocks) {
    if (kv.second.dropped) continue;
    PriorBlock pb;
    pb.type = kv.first.type;
    pb.index = kv.first.index;
    pb.col = kv.second.pos - m;
    const double* x = blockData(*this, pb.type, pb.index);
    for (int d = 0; d < 4; ++d) pb.x0[d] = d < blockGlobalSize(pb.type) ? x[d] : 0.0;
    new_prior.blocks.push_back(pb);
  }
  return true;
}

// ---------------------------------------------------------------------------
// trajectory_manager.cpp:485-516 + utils/eigen_utils.hpp:114-150

static void R2ypr(const Mat3& R, double ypr[3]) {
  const Vec3 n(R(0, 0), R(1, 0), R(2, 0)), o(R(0, 1), R(1, 1), R(2, 1)), a(R(0, 2), R(1, 2), R(2, 2));
  const double y = std::atan2(n.y, n.x);
  const double p = std::atan2(-n.z, n.x * std::cos(y) + n.y * std::sin(y));
  const double r = std::atan2(a.x * std::sin(y) - a.y * std::cos(y), -o.x * std::sin(y) + o.y * std::cos(y));
  ypr[0] = y / M_PI * 180.0; ypr[1] = p / M_PI * 180.0; ypr[2] = r / M_PI * 180.0;
}

// Eigen::Quaternion(Matrix3) (used by Sophus::SO3(Matrix) -> SE3d(rot_diff, tran_diff))
static Quat quatFromMatrix(const Mat3& m) {
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  double qv[4];  // x y z w
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    qv[3] = 0.5 * t;
    t = 0.5 / t;
    qv[0] = (m(2, 1) - m(1, 2)) * t;
    qv[1] = (m(0, 2) - m(2, 0)) * t;
    qv[2] = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    qv[i] = 0.5 * t;
    t = 0.5 / t;
    qv[3] = (m(k, j) - m(j, k)) * t;
    qv[j] = (m(j, i) + m(i, j)) * t;
    qv[k] = (m(k, i) + m(i, k)) * t;
  }
  return {qv[0], qv[1], qv[2], qv[3]};
}

void Window::gauge_realign(int min_idx, const double R0a[9], const double t0a[3]) {
  Mat3 R0;
  for (int i = 0; i < 9; ++i) R0.m[i] = R0a[i];
  const Vec3 t0(t0a[0], t0a[1], t0a[2]);
  const Mat3 R00 = so3_matrix(knotQ(q.data(), min_idx));
  const Vec3 t00 = knotP(p.data(), min_idx);
  double e0[3], e00[3];
  R2ypr(R0, e0);
  R2ypr(R00, e00);
  const double y_diff = e0[0] - e00[0];
  const double y = y_diff / 180.0 * M_PI;
  Mat3 rot_diff = Mat3::Identity();  // ypr2R(y_diff, 0, 0) == Rz(y)
  rot_diff(0, 0) = std::cos(y); rot_diff(0, 1) = -std::sin(y);
  rot_diff(1, 0) = std::sin(y); rot_diff(1, 1) = std::cos(y);
  if (std::fabs(std::fabs(e0[1]) - 90) < 1.0 || std::fabs(std::fabs(e00[1]) - 90) < 1.0)
    rot_diff = R0 * transpose(R00);
  const Vec3 tran_diff = t0 - rot_diff * t00;
  const Quat qd = quatFromMatrix(rot_diff);
  for (int i = min_idx; i < nK(); ++i) {
    const Quat qi = knotQ(q.data(), i);
    const Vec3 pi = knotP(p.data(), i);
    const Quat qn = so3_mul(qd, qi);                 // SE3 * SE3: so3 part
    const Vec3 pn = so3_rotate(qd, pi) + tran_diff;  // translation part
    qn.toPtr(&q[4 * i]);
    p[3 * i] = pn.x; p[3 * i + 1] = pn.y; p[3 * i + 2] = pn.z;
  }
}

}  // namespace ctvio_oracle
