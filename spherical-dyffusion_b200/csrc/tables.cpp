#include "tables.h"

#include <algorithm>
#include <cmath>

namespace sfno {

static const double kPi = 3.14159265358979323846264338327950288;

void legendre_gauss(int n, std::vector<double>& nodes, std::vector<double>& weights) {
  nodes.assign(n, 0.0);
  weights.assign(n, 0.0);
  for (int i = 0; i < (n + 1) / 2; ++i) {
    // root i counted from +1; Newton on P_n with the Chebyshev-like initial guess
    double x = std::cos(kPi * (i + 0.75) / (n + 0.5));
    double dp = 1.0;
    for (int it = 0; it < 100; ++it) {
      double p0 = 1.0, p1 = x;
      for (int k = 1; k < n; ++k) {
        double p2 = ((2.0 * k + 1.0) * x * p1 - k * p0) / (k + 1.0);
        p0 = p1;
        p1 = p2;
      }
      if (n == 0) p1 = 1.0;
      dp = n * (x * p1 - p0) / (x * x - 1.0);
      double dx = p1 / dp;
      x -= dx;
      if (std::fabs(dx) < 1e-16) break;
    }
    // re-evaluate derivative at the converged root
    double p0 = 1.0, p1 = x;
    for (int k = 1; k < n; ++k) {
      double p2 = ((2.0 * k + 1.0) * x * p1 - k * p0) / (k + 1.0);
      p0 = p1;
      p1 = p2;
    }
    dp = n * (x * p1 - p0) / (x * x - 1.0);
    double w = 2.0 / ((1.0 - x * x) * dp * dp);
    nodes[n - 1 - i] = x;
    nodes[i] = -x;
    weights[n - 1 - i] = w;
    weights[i] = w;
  }
  if (n % 2 == 1) nodes[n / 2] = 0.0;
}

void clenshaw_curtis(int n, std::vector<double>& nodes, std::vector<double>& weights) {
  nodes.assign(n, 0.0);
  weights.assign(n, 0.0);
  const int N = n - 1;
  for (int k = 0; k < n; ++k) nodes[k] = std::cos(kPi - kPi * k / N);  // cos(linspace(pi, 0, n))
  if (n == 2) {
    weights[0] = weights[1] = 1.0;
    return;
  }
  // closed form equal to the FFT construction: w_k = c_k/N * (1 - sum_j b_j/(4j^2-1) cos(2 j theta_k))
  for (int k = 0; k < n; ++k) {
    double theta = kPi * k / N;
    double s = 0.0;
    for (int j = 1; j <= N / 2; ++j) {
      double b = (2 * j == N) ? 1.0 : 2.0;
      s += b / (4.0 * j * j - 1.0) * std::cos(2.0 * j * theta);
    }
    double c = (k == 0 || k == N) ? 1.0 : 2.0;
    weights[k] = c / N * (1.0 - s);
  }
}

bool build_sht_tables(int nlat, int nlon, int lmax, int mmax, int grid, ShtTables& t) {
  if (nlat < 2 || nlon < 2 || lmax < 1 || mmax < 1) return false;
  std::vector<double> nodes, w;
  if (grid == 0) legendre_gauss(nlat, nodes, w);
  else if (grid == 1) clenshaw_curtis(nlat, nodes, w);
  else return false;
  t.nlat = nlat; t.nlon = nlon; t.lmax = lmax; t.mmax = mmax; t.grid = grid;
  t.quad_w = w;
  // colatitudes north -> south: flip(arccos(nodes)); x = cos(colat)
  t.cost.resize(nlat);
  for (int k = 0; k < nlat; ++k) t.cost[k] = std::cos(std::acos(nodes[nlat - 1 - k]));

  const int n = std::max(lmax, mmax);
  // p[m][l][k] on the working size n, then truncated
  std::vector<double> p((size_t)n * n * nlat, 0.0);
  auto P = [&](int m, int l) { return p.data() + ((size_t)m * n + l) * nlat; };
  for (int k = 0; k < nlat; ++k) P(0, 0)[k] = 1.0 / std::sqrt(4.0 * kPi);
  for (int l = 1; l < n; ++l) {
    const double a = std::sqrt(2.0 * l + 1.0);
    for (int k = 0; k < nlat; ++k) {
      const double x = t.cost[k];
      P(l - 1, l)[k] = a * x * P(l - 1, l - 1)[k];
      P(l, l)[k] = std::sqrt((2.0 * l + 1.0) * (1.0 + x) * (1.0 - x) / 2.0 / l) * P(l - 1, l - 1)[k];
    }
  }
  for (int l = 2; l < n; ++l) {
    for (int m = 0; m < l - 1; ++m) {
      const double a = std::sqrt((2.0 * l - 1.0) / (l - m) * (2.0 * l + 1.0) / (l + m));
      const double b = std::sqrt((double)(l + m - 1) / (l - m) * (2.0 * l + 1.0) / (2.0 * l - 3.0) * (l - m - 1) / (l + m));
      double* out = P(m, l);
      const double* p1 = P(m, l - 1);
      const double* p2 = P(m, l - 2);
      for (int k = 0; k < nlat; ++k) out[k] = t.cost[k] * a * p1[k] - b * p2[k];
    }
  }
  t.pct.assign((size_t)mmax * lmax * nlat, 0.0);
  t.weights.assign((size_t)mmax * lmax * nlat, 0.0);
  for (int m = 0; m < mmax; ++m) {
    const double sign = (m & 1) ? -1.0 : 1.0;  // Condon-Shortley phase
    for (int l = 0; l < lmax; ++l) {
      const double* src = P(m, l);
      double* d0 = t.pct.data() + ((size_t)m * lmax + l) * nlat;
      double* d1 = t.weights.data() + ((size_t)m * lmax + l) * nlat;
      for (int k = 0; k < nlat; ++k) {
        d0[k] = sign * src[k];
        d1[k] = sign * src[k] * w[k];
      }
    }
  }
  return true;
}

static inline void cs(int m, int j, int nlon, double& c, double& s) {
  long long r = ((long long)m * j) % nlon;  // exact angle reduction
  double ang = 2.0 * kPi * (double)r / (double)nlon;
  c = std::cos(ang);
  s = std::sin(ang);
  if (r == 0) { c = 1.0; s = 0.0; }
  if (2 * r == nlon) { c = -1.0; s = 0.0; }
}

void build_dft_forward(int nlon, int mmax, std::vector<double>& e) {
  e.assign((size_t)2 * mmax * nlon, 0.0);
  const double scale = 2.0 * kPi / nlon;
  for (int m = 0; m < mmax; ++m)
    for (int j = 0; j < nlon; ++j) {
      double c, s;
      cs(m, j, nlon, c, s);
      e[(size_t)(2 * m) * nlon + j] = scale * c;
      e[(size_t)(2 * m + 1) * nlon + j] = -scale * s;
    }
}

void build_dft_inverse(int nlon, int mmax, std::vector<double>& e) {
  e.assign((size_t)nlon * 2 * mmax, 0.0);
  for (int j = 0; j < nlon; ++j)
    for (int m = 0; m < mmax; ++m) {
      double c, s;
      cs(m, j, nlon, c, s);
      const bool self_conj = (m == 0) || (2 * m == nlon);  // DC / Nyquist: weight 1, imaginary part ignored
      const double cm = self_conj ? 1.0 : 2.0;
      e[(size_t)j * 2 * mmax + 2 * m] = cm * c;
      e[(size_t)j * 2 * mmax + 2 * m + 1] = self_conj ? 0.0 : -cm * s;
    }
}

}  // namespace sfno
