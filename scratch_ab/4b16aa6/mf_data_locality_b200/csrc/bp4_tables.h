// bp4_tables.h -- host-side construction of the 1-D quadrature and basis tables.
// Stand-in for the deal.II pieces the reference pulls them from: QGauss<1>
// (poisson_operator.h:107), QGaussLobatto<1> / FE_Q support points (benchmark.h:91,129),
// ShapeInfo::shape_values / shape_gradients_collocation (poisson_operator.h:461,553).
#pragma once
#include <cmath>
#include <vector>

#include "bp4_cell.cuh"

namespace bp4
{
  // Legendre polynomial P_n and derivative at x in [-1,1]
  inline void legendre(int n, long double x, long double &p, long double &dp)
  {
    long double p0 = 1.0L, p1 = x;
    if (n == 0)
      {
        p  = 1.0L;
        dp = 0.0L;
        return;
      }
    for (int k = 2; k <= n; ++k)
      {
        const long double pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
        p0                   = p1;
        p1                   = pk;
      }
    p  = p1;
    dp = n * (x * p1 - p0) / (x * x - 1.0L);
  }

  // Gauss-Legendre on [0,1]
  inline void gauss_01(int n, std::vector<double> &x, std::vector<double> &w)
  {
    x.resize(n);
    w.resize(n);
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int i = 0; i < n; ++i)
      {
        long double z = -std::cos(pi * (i + 0.75L) / (n + 0.5L));
        long double p, dp;
        for (int it = 0; it < 100; ++it)
          {
            legendre(n, z, p, dp);
            const long double dz = p / dp;
            z -= dz;
            if (std::fabs((double)dz) < 1e-19)
              break;
          }
        legendre(n, z, p, dp);
        x[i] = (double)(0.5L * (z + 1.0L));
        w[i] = (double)(1.0L / ((1.0L - z * z) * dp * dp));
      }
    for (int i = 0; i < n / 2; ++i) // enforce symmetry
      {
        const double xs = 0.5 * (x[i] + (1.0 - x[n - 1 - i]));
        x[i]            = xs;
        x[n - 1 - i]    = 1.0 - xs;
        const double ws = 0.5 * (w[i] + w[n - 1 - i]);
        w[i] = w[n - 1 - i] = ws;
      }
    if (n % 2)
      x[n / 2] = 0.5;
  }

  // Gauss-Lobatto on [0,1]: end points plus the roots of P'_{n-1}
  inline void gauss_lobatto_01(int n, std::vector<double> &x, std::vector<double> &w)
  {
    x.resize(n);
    w.resize(n);
    const int         m  = n - 1;
    const long double pi = 3.141592653589793238462643383279502884L;
    std::vector<long double> z(n);
    z[0] = -1.0L;
    z[m] = 1.0L;
    for (int i = 1; i < m; ++i)
      {
        long double zi = -std::cos(pi * i / m);
        for (int it = 0; it < 100; ++it)
          {
            // Newton on q(z) = P'_m(z); q' from the Legendre ODE
            long double p, dp;
            legendre(m, zi, p, dp);
            const long double ddp = (2 * zi * dp - m * (m + 1) * p) / (1.0L - zi * zi);
            const long double dz  = dp / ddp;
            zi -= dz;
            if (std::fabs((double)dz) < 1e-19)
              break;
          }
        z[i] = zi;
      }
    for (int i = 0; i < n; ++i)
      {
        long double p, dp;
        if (i == 0 || i == m)
          p = (i == 0 && (m % 2)) ? -1.0L : 1.0L;
        else
          legendre(m, z[i], p, dp);
        x[i] = (double)(0.5L * (z[i] + 1.0L));
        w[i] = (double)(1.0L / (m * (m + 1) * p * p));
      }
    for (int i = 0; i < n / 2; ++i)
      {
        const double xs = 0.5 * (x[i] + (1.0 - x[n - 1 - i]));
        x[i]            = xs;
        x[n - 1 - i]    = 1.0 - xs;
        const double ws = 0.5 * (w[i] + w[n - 1 - i]);
        w[i] = w[n - 1 - i] = ws;
      }
    if (n % 2)
      x[n / 2] = 0.5;
  }

  // l_i(pt) for the Lagrange basis on `nodes`
  inline double lagrange_value(const std::vector<double> &nodes, int i, double pt)
  {
    long double v = 1.0L;
    for (size_t j = 0; j < nodes.size(); ++j)
      if ((int)j != i)
        v *= ((long double)pt - nodes[j]) / ((long double)nodes[i] - nodes[j]);
    return (double)v;
  }

  // l_i'(pt)
  inline double lagrange_deriv(const std::vector<double> &nodes, int i, double pt)
  {
    long double s = 0.0L;
    for (size_t k = 0; k < nodes.size(); ++k)
      {
        if ((int)k == i)
          continue;
        long double term = 1.0L / ((long double)nodes[i] - nodes[k]);
        for (size_t j = 0; j < nodes.size(); ++j)
          if ((int)j != i && j != k)
            term *= ((long double)pt - nodes[j]) / ((long double)nodes[i] - nodes[j]);
        s += term;
      }
    return (double)s;
  }

  template <int P>
  inline void fill_tab(Tab<P> &tb)
  {
    constexpr int       N = P + 1, Q = P + 2;
    std::vector<double> xn, wn, xq, wq;
    gauss_lobatto_01(N, xn, wn);
    gauss_01(Q, xq, wq);
    for (int q = 0; q < Q; ++q)
      {
        tb.xq[q] = xq[q];
        tb.wq[q] = wq[q];
        for (int i = 0; i < N; ++i)
          {
            tb.S[i][q]  = lagrange_value(xn, i, xq[q]);
            tb.Dn[i][q] = lagrange_deriv(xn, i, xq[q]);
          }
        for (int i = 0; i < Q; ++i)
          tb.D[i][q] = lagrange_deriv(xq, i, xq[q]);
      }
    // even-odd halves, formed from the rounded entries
    auto halves = [](auto &M, auto &fp, auto &fm, auto &sp, auto &sm, const int ni, const int no) {
      for (int i = 0; i < ni; ++i)
        for (int q = 0; q < no; ++q)
          {
            fp[i][q] = 0.5 * (M[i][q] + M[ni - 1 - i][q]);
            fm[i][q] = 0.5 * (M[i][q] - M[ni - 1 - i][q]);
            sp[i][q] = 0.5 * (M[i][q] + M[i][no - 1 - q]);
            sm[i][q] = 0.5 * (M[i][q] - M[i][no - 1 - q]);
          }
    };
    halves(tb.S, tb.Sfp, tb.Sfm, tb.Ssp, tb.Ssm, N, Q);
    halves(tb.Dn, tb.Dnfp, tb.Dnfm, tb.Dnsp, tb.Dnsm, N, Q);
    halves(tb.D, tb.Dfp, tb.Dfm, tb.Dsp, tb.Dsm, Q, Q);
  }
} // namespace bp4
