"""The rotation schedule of the blocked Jacobi eigen-solver (ctrl-vio_b200/csrc/jacobi_blocked.cu: rr_pair, block_pair,
inner_pair), restated in numpy: every index pair is rotated exactly once per sweep, the rotations of an inner round are
disjoint, and the scheme converges like the classical cyclic method.  (The CUDA kernels themselves are checked against
LAPACK in tests/test_gpu_parity.py::test_eigen_solvers_match_lapack; this file pins the combinatorics they rely on and
runs without a GPU.)"""
import itertools

import numpy as np
import pytest


def rr_pair(m, rr, k):
    """pair k of round rr of the round-robin tournament on m (even) players, ascending (jacobi_blocked.cu: rr_pair)"""
    if k == 0:
        x, y = m - 1, rr
    else:
        x, y = (rr + k) % (m - 1), (rr + m - 1 - k) % (m - 1)
    return min(x, y), max(x, y)


def block_pair(nb, br, k):
    return (2 * k, 2 * k + 1) if br == 0 else rr_pair(nb, br - 1, k)


def inner_pair(within, t, i):
    if within:
        a, b = rr_pair(8, t, i & 3)
        o = 8 if (i & 4) else 0
        return a + o, b + o
    return i, 8 + ((i + t) & 7)


def sweep_rotations(nb):
    """[(block round, inner round, global p, global q)] of one sweep"""
    out = []
    for br in range(nb):
        within = br == 0
        for k in range(nb // 2):
            I, J = block_pair(nb, br, k)
            loc = list(range(8 * I, 8 * I + 8)) + list(range(8 * J, 8 * J + 8))
            for t in range(7 if within else 8):
                for i in range(8):
                    p, q = inner_pair(within, t, i)
                    out.append((br, t, loc[p], loc[q]))
    return out


@pytest.mark.parametrize("nb", [2, 4, 6, 12, 14])
def test_every_index_pair_is_rotated_exactly_once_per_sweep(nb):
    rots = sweep_rotations(nb)
    pairs = [tuple(sorted((p, q))) for _, _, p, q in rots]
    n = 8 * nb
    assert len(pairs) == n * (n - 1) // 2
    assert set(pairs) == set(itertools.combinations(range(n), 2))


@pytest.mark.parametrize("nb", [2, 6, 14])
def test_rotations_of_a_round_are_disjoint(nb):
    """the 8 rotations of an inner round of a pair problem, and the pair problems of a block round, touch disjoint indices:
    they commute, which is what lets one warp pair / one warp per block pair apply them side by side"""
    by_round = {}
    for br, t, p, q in sweep_rotations(nb):
        by_round.setdefault((br, t), []).extend((p, q))
    for idx in by_round.values():
        assert len(idx) == len(set(idx))
    for br in range(nb):
        blocks = [b for k in range(nb // 2) for b in block_pair(nb, br, k)]
        assert sorted(blocks) == list(range(nb))


def _rotation(app, aqq, apq):
    """jacobi_blocked.cu: rotation() - two reciprocal square roots, no division"""
    d, o = aqq - app, 2.0 * apq
    r2 = d * d + o * o
    if o == 0.0 or not (1e-280 < r2 < 1e280):
        return 1.0, 0.0
    ir = 1.0 / np.sqrt(r2)
    c2 = 0.5 + 0.5 * abs(d) * ir
    ic = 1.0 / np.sqrt(c2)
    return c2 * ic, np.copysign(0.5, d) * o * ir * ic


def test_rotation_formula_zeroes_the_pivot_and_is_orthonormal():
    rng = np.random.default_rng(7)
    for _ in range(2000):
        app, aqq = 10.0 ** rng.uniform(-8, 8, 2) * rng.choice([-1, 1], 2)
        apq = 10.0 ** rng.uniform(-12, 8) * rng.choice([-1, 1])
        c, s = _rotation(app, aqq, apq)
        assert abs(c * c + s * s - 1.0) <= 1e-15
        assert abs(s) <= c * (1 + 1e-15)  # |phi| <= pi / 4
        # column update x_p' = c x_p - s x_q, x_q' = s x_p + c x_q on both sides
        j = np.array([[c, s], [-s, c]])
        b = j.T @ np.array([[app, apq], [apq, aqq]]) @ j
        assert abs(b[0, 1]) <= 1e-15 * (abs(app) + abs(aqq) + abs(apq))


def test_blocked_schedule_converges_like_the_cyclic_method():
    """numpy model of the solver (one pass of rotations per pair problem, accumulated into Q, matrix updated as Q' A Q) on
    a prior-like matrix: reaches off(A)^2 <= 1e-30 diag(A)^2 in as many sweeps as the classical round-robin order"""
    rng = np.random.default_rng(3)
    n, nb = 40, 6
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = 10.0 ** rng.uniform(-4, 4, n)
    a0 = (q * ev) @ q.T
    a0 = 0.5 * (a0 + a0.T)

    def off2(a):
        o = a - np.diag(np.diag(a))
        return np.sum(o * o), np.sum(np.diag(a) ** 2)

    def blocked():
        N = 8 * nb
        a = np.zeros((N, N)); a[:n, :n] = a0
        for sweep in range(40):
            o, d = off2(a)
            if o <= 1e-30 * d:
                return sweep, a
            for br in range(nb):
                Q = np.eye(N)
                for k in range(nb // 2):
                    I, J = block_pair(nb, br, k)
                    loc = np.r_[8 * I:8 * I + 8, 8 * J:8 * J + 8]
                    s_ = a[np.ix_(loc, loc)].copy(); ql = np.eye(16)
                    for t in range(7 if br == 0 else 8):
                        jm = np.eye(16)
                        for i in range(8):
                            p, q_ = inner_pair(br == 0, t, i)
                            c, s = _rotation(s_[p, p], s_[q_, q_], s_[p, q_])
                            jm[p, p] = jm[q_, q_] = c; jm[p, q_] = s; jm[q_, p] = -s
                        s_ = jm.T @ s_ @ jm; ql = ql @ jm
                    Q[np.ix_(loc, loc)] = ql
                a = Q.T @ a @ Q
        return 40, a

    def classical():
        ne = n
        a = a0.copy()
        for sweep in range(40):
            o, d = off2(a)
            if o <= 1e-30 * d:
                return sweep
            for rr in range(ne - 1):
                jm = np.eye(ne)
                for k in range(ne // 2):
                    p, q_ = rr_pair(ne, rr, k)
                    c, s = _rotation(a[p, p], a[q_, q_], a[p, q_])
                    jm[p, p] = jm[q_, q_] = c; jm[p, q_] = s; jm[q_, p] = -s
                a = jm.T @ a @ jm
        return 40

    sb, a = blocked()
    sc = classical()
    assert sb <= sc + 2, (sb, sc)
    assert np.max(np.abs(np.sort(np.diag(a))[-n:] - np.linalg.eigvalsh(a0))) <= 1e-12 * np.linalg.norm(a0, 2)
