// ORACLE (test infrastructure, NOT product code).
// Small dense fp64 helpers standing in for the Eigen / CHOLMOD pieces the
// reference reaches through Ceres and marginalization_factor.cpp:
//   - dense Cholesky LL^T (stands in for Ceres SPARSE_NORMAL_CHOLESKY; the
//     factorisation is mathematically the same system, dense instead of sparse)
//   - cyclic Jacobi symmetric eigen-decomposition (stands in for
//     Eigen::SelfAdjointEigenSolver, marginalization_factor.cpp:241,254)
//   - polynomial root finding for the Armijo line search (Ceres polynomial.cc)
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

namespace ctvio_oracle {

// In-place lower Cholesky of row-major n x n (only lower triangle read/written).
// Returns false when a pivot is not strictly positive / not finite.
inline bool cholesky_lower(double* A, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    A[j * n + j] = d;
    const double inv = 1.0 / d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      const double* ai = A + i * n;
      const double* aj = A + j * n;
      for (int k = 0; k < j; ++k) s -= ai[k] * aj[k];
      A[i * n + j] = s * inv;
    }
  }
  return true;
}

inline void cholesky_solve(const double* L, int n, double* b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i * n + k] * b[k];
    b[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * b[k];
    b[i] = s / L[i * n + i];
  }
}

// Cyclic Jacobi eigen-decomposition of a symmetric row-major n x n matrix.
// On return eval[k] (ascending) and evec (row-major, column k = eigenvector k).
inline void jacobi_eigh(std::vector<double> A, int n, std::vector<double>& eval, std::vector<double>& evec) {
  evec.assign(size_t(n) * n, 0.0);
  for (int i = 0; i < n; ++i) evec[size_t(i) * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; ++i) {
      diag += A[size_t(i) * n + i] * A[size_t(i) * n + i];
      for (int j = i + 1; j < n; ++j) off += A[size_t(i) * n + j] * A[size_t(i) * n + j];
    }
    if (off <= 1e-60 || off <= 1e-32 * diag) break;
    for (int p = 0; p < n - 1; ++p) {
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[size_t(p) * n + q];
        if (apq == 0.0) continue;
        const double app = A[size_t(p) * n + p], aqq = A[size_t(q) * n + q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {  // columns p,q
          const double akp = A[size_t(k) * n + p], akq = A[size_t(k) * n + q];
          A[size_t(k) * n + p] = c * akp - s * akq;
          A[size_t(k) * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {  // rows p,q
          const double apk = A[size_t(p) * n + k], aqk = A[size_t(q) * n + k];
          A[size_t(p) * n + k] = c * apk - s * aqk;
          A[size_t(q) * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = evec[size_t(k) * n + p], vkq = evec[size_t(k) * n + q];
          evec[size_t(k) * n + p] = c * vkp - s * vkq;
          evec[size_t(k) * n + q] = s * vkp + c * vkq;
        }
      }
    }
  }
  eval.resize(n);
  for (int i = 0; i < n; ++i) eval[i] = A[size_t(i) * n + i];
  // sort ascending (Eigen convention)
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return eval[a] < eval[b]; });
  std::vector<double> ev2(n), V2(size_t(n) * n);
  for (int k = 0; k < n; ++k) {
    ev2[k] = eval[idx[k]];
    for (int i = 0; i < n; ++i) V2[size_t(i) * n + k] = evec[size_t(i) * n + idx[k]];
  }
  eval.swap(ev2);
  evec.swap(V2);
}

// ---- polynomial helpers (Ceres 1.14 internal/ceres/polynomial.cc) ----------
// Coefficients highest degree first.
inline double EvaluatePolynomial(const std::vector<double>& poly, double x) {
  double v = 0.0;
  for (double c : poly) v = v * x + c;
  return v;
}
inline std::vector<double> DifferentiatePolynomial(const std::vector<double>& poly) {
  const int degree = int(poly.size()) - 1;
  if (degree == 0) return {0.0};
  std::vector<double> d(degree);
  for (int i = 0; i < degree; ++i) d[i] = (degree - i) * poly[i];
  return d;
}
// Real parts of all roots (Ceres keeps the real parts of complex roots too).
inline bool FindPolynomialRootsReal(std::vector<double> poly, std::vector<double>& real) {
  real.clear();
  size_t lead = 0;
  while (lead + 1 < poly.size() && poly[lead] == 0.0) ++lead;  // RemoveLeadingZeros
  poly.erase(poly.begin(), poly.begin() + lead);
  const int degree = int(poly.size()) - 1;
  if (degree < 0) return false;
  if (degree == 0) return true;
  if (degree == 1) {
    real.push_back(-poly[1] / poly[0]);
    return true;
  }
  if (degree == 2) {
    const double a = poly[0], b = poly[1], c = poly[2];
    const double D = b * b - 4 * a * c;
    const double sqrt_D = std::sqrt(std::fabs(D));
    if (D >= 0) {
      if (b >= 0) {
        real.push_back((-b - sqrt_D) / (2.0 * a));
        real.push_back((2.0 * c) / (-b - sqrt_D));
      } else {
        real.push_back((2.0 * c) / (-b + sqrt_D));
        real.push_back((-b + sqrt_D) / (2.0 * a));
      }
    } else {
      real.push_back(-b / (2.0 * a));
      real.push_back(-b / (2.0 * a));
    }
    return true;
  }
  // degree >= 3: Durand-Kerner on the monic polynomial (Ceres uses the
  // eigenvalues of the balanced companion matrix; same roots).
  using cd = std::complex<double>;
  std::vector<cd> a(degree + 1);
  for (int i = 0; i <= degree; ++i) a[i] = poly[i] / poly[0];
  double radius = 0;
  for (int i = 1; i <= degree; ++i) radius = std::max(radius, std::abs(a[i]));
  radius = 1.0 + radius;
  std::vector<cd> z(degree);
  for (int i = 0; i < degree; ++i) z[i] = std::polar(radius * 0.5, 2.0 * M_PI * i / degree + 0.4);
  for (int it = 0; it < 500; ++it) {
    double change = 0;
    for (int i = 0; i < degree; ++i) {
      cd num = 0;
      for (int k = 0; k <= degree; ++k) num = num * z[i] + a[k];
      cd den = 1;
      for (int j = 0; j < degree; ++j)
        if (j != i) den *= (z[i] - z[j]);
      if (std::abs(den) == 0) den = 1e-300;
      const cd dz = num / den;
      z[i] -= dz;
      change = std::max(change, std::abs(dz));
    }
    if (change < 1e-15 * radius) break;
  }
  for (int i = 0; i < degree; ++i) real.push_back(z[i].real());
  return true;
}

// Solve a small dense system with full-pivot Gaussian elimination
// (FindInterpolatingPolynomial uses lhs.fullPivLu().solve(rhs)).
inline std::vector<double> SolveDenseFullPivot(std::vector<double> A, std::vector<double> b, int n) {
  std::vector<int> colperm(n);
  for (int i = 0; i < n; ++i) colperm[i] = i;
  for (int k = 0; k < n; ++k) {
    int pr = k, pc = k;
    double best = -1;
    for (int i = k; i < n; ++i)
      for (int j = k; j < n; ++j)
        if (std::fabs(A[i * n + j]) > best) {
          best = std::fabs(A[i * n + j]);
          pr = i; pc = j;
        }
    if (best <= 0) break;
    if (pr != k) {
      for (int j = 0; j < n; ++j) std::swap(A[k * n + j], A[pr * n + j]);
      std::swap(b[k], b[pr]);
    }
    if (pc != k) {
      for (int i = 0; i < n; ++i) std::swap(A[i * n + k], A[i * n + pc]);
      std::swap(colperm[k], colperm[pc]);
    }
    for (int i = k + 1; i < n; ++i) {
      const double f = A[i * n + k] / A[k * n + k];
      if (f == 0) continue;
      for (int j = k; j < n; ++j) A[i * n + j] -= f * A[k * n + j];
      b[i] -= f * b[k];
    }
  }
  std::vector<double> y(n, 0.0);
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < n; ++j) s -= A[i * n + j] * y[j];
    y[i] = (A[i * n + i] != 0) ? s / A[i * n + i] : 0.0;
  }
  std::vector<double> x(n);
  for (int i = 0; i < n; ++i) x[colperm[i]] = y[i];
  return x;
}

}  // namespace ctvio_oracle
