// Quadrature tables handed to the kernels by value.  Selection follows Dune::QuadratureRules semantics
// (smallest rule exact for the requested order; only direct call site in the reference:
// estimators/block-swipdg.hh:59-60, all other uses are inside dune-gdt): Gauss-Legendre on lines, its tensor
// product on squares, the classic 1/3/4/6/7-point rules on triangles up to order 5 and a conical product beyond.
#pragma once
#include <cmath>

#include "common.hpp"

namespace hdd {

constexpr int kMaxLinePts = 8;
constexpr int kMaxElemPts = 36;

struct LineRule {  // on [0,1]
  int n;
  double x[kMaxLinePts], w[kMaxLinePts];
};
struct ElemRule {  // reference triangle (weights sum to 1/2) or unit square
  int n;
  double x[kMaxElemPts], y[kMaxElemPts], w[kMaxElemPts];
};

inline LineRule gauss_legendre_01(int n) {
  LineRule r{};
  r.n = n;
  double z[kMaxLinePts], w[kMaxLinePts];  // on [-1,1]
  switch (n) {
    case 1: z[0] = 0.0; w[0] = 2.0; break;
    case 2: z[0] = -1.0 / std::sqrt(3.0); z[1] = -z[0]; w[0] = w[1] = 1.0; break;
    case 3:
      z[0] = -std::sqrt(0.6); z[1] = 0.0; z[2] = std::sqrt(0.6);
      w[0] = w[2] = 5.0 / 9.0; w[1] = 8.0 / 9.0;
      break;
    case 4: {
      const double a = std::sqrt(3.0 / 7.0 - 2.0 / 7.0 * std::sqrt(1.2));
      const double b = std::sqrt(3.0 / 7.0 + 2.0 / 7.0 * std::sqrt(1.2));
      z[0] = -b; z[1] = -a; z[2] = a; z[3] = b;
      w[1] = w[2] = (18.0 + std::sqrt(30.0)) / 36.0;
      w[0] = w[3] = (18.0 - std::sqrt(30.0)) / 36.0;
      break;
    }
    case 5: {
      const double a = std::sqrt(5.0 - 2.0 * std::sqrt(10.0 / 7.0)) / 3.0;
      const double b = std::sqrt(5.0 + 2.0 * std::sqrt(10.0 / 7.0)) / 3.0;
      z[0] = -b; z[1] = -a; z[2] = 0.0; z[3] = a; z[4] = b;
      w[2] = 128.0 / 225.0;
      w[1] = w[3] = (322.0 + 13.0 * std::sqrt(70.0)) / 900.0;
      w[0] = w[4] = (322.0 - 13.0 * std::sqrt(70.0)) / 900.0;
      break;
    }
    default: {
      if (n < 1 || n > kMaxLinePts) HDD_THROW(HDD_ERR_NOT_IMPLEMENTED, "Gauss-Legendre rule with " << n << " points");
      // eigenvalue-free fallback: Newton on P_n from Chebyshev-like starting values, refined in long double
      for (int i = 0; i < n; ++i) {
        long double t = std::cos(3.14159265358979323846L * (n - i - 0.25L) / (n + 0.5L));
        long double dp = 1.0L;
        for (int it = 0; it < 60; ++it) {
          long double p0 = 1.0L, p1 = t;
          for (int k = 2; k <= n; ++k) {
            const long double p2 = ((2 * k - 1) * t * p1 - (k - 1) * p0) / k;
            p0 = p1;
            p1 = p2;
          }
          dp = n * (t * p1 - p0) / (t * t - 1.0L);
          const long double dt = p1 / dp;
          t -= dt;
          if (std::fabs(double(dt)) < 1e-19) break;
        }
        z[i] = double(t);
        w[i] = double(2.0L / ((1.0L - t * t) * dp * dp));
      }
    }
  }
  for (int i = 0; i < n; ++i) {
    r.x[i] = 0.5 * (z[i] + 1.0);
    r.w[i] = 0.5 * w[i];
  }
  return r;
}

inline LineRule line_rule(int order) { return gauss_legendre_01(order / 2 + 1); }

inline ElemRule square_rule(int order) {
  const LineRule g = line_rule(order);
  if (g.n * g.n > kMaxElemPts) HDD_THROW(HDD_ERR_NOT_IMPLEMENTED, "square rule of order " << order);
  ElemRule r{};
  r.n = 0;
  for (int j = 0; j < g.n; ++j)
    for (int i = 0; i < g.n; ++i) {
      r.x[r.n] = g.x[i];
      r.y[r.n] = g.x[j];
      r.w[r.n] = g.w[i] * g.w[j];
      ++r.n;
    }
  return r;
}

inline ElemRule triangle_rule(int order) {
  ElemRule r{};
  r.n = 0;
  auto put = [&r](double x, double y, double w) {
    r.x[r.n] = x; r.y[r.n] = y; r.w[r.n] = w; ++r.n;
  };
  auto orbit3 = [&put](double a, double w) {  // (a,a), (1-2a,a), (a,1-2a)
    put(a, a, w); put(1.0 - 2.0 * a, a, w); put(a, 1.0 - 2.0 * a, w);
  };
  if (order <= 1) {
    put(1.0 / 3.0, 1.0 / 3.0, 0.5);
  } else if (order == 2) {
    put(2.0 / 3.0, 1.0 / 6.0, 1.0 / 6.0);
    put(1.0 / 6.0, 2.0 / 3.0, 1.0 / 6.0);
    put(1.0 / 6.0, 1.0 / 6.0, 1.0 / 6.0);
  } else if (order == 3) {
    put(1.0 / 3.0, 1.0 / 3.0, -9.0 / 32.0);
    put(3.0 / 5.0, 1.0 / 5.0, 25.0 / 96.0);
    put(1.0 / 5.0, 3.0 / 5.0, 25.0 / 96.0);
    put(1.0 / 5.0, 1.0 / 5.0, 25.0 / 96.0);
  } else if (order == 4) {
    const double root = std::sqrt(38.0 - 44.0 * std::sqrt(2.0 / 5.0));
    const double wroot = std::sqrt(213125.0 - 53320.0 * std::sqrt(10.0));
    orbit3((8.0 - std::sqrt(10.0) + root) / 18.0, (620.0 + wroot) / 7440.0);
    orbit3((8.0 - std::sqrt(10.0) - root) / 18.0, (620.0 - wroot) / 7440.0);
  } else if (order == 5) {
    const double s = std::sqrt(15.0);
    put(1.0 / 3.0, 1.0 / 3.0, 9.0 / 80.0);
    orbit3((6.0 - s) / 21.0, (155.0 - s) / 2400.0);
    orbit3((6.0 + s) / 21.0, (155.0 + s) / 2400.0);
  } else {
    const LineRule gu = gauss_legendre_01((order + 1) / 2 + 1);
    const LineRule gt = gauss_legendre_01(order / 2 + 1);
    if (gu.n * gt.n > kMaxElemPts) HDD_THROW(HDD_ERR_NOT_IMPLEMENTED, "triangle rule of order " << order);
    for (int i = 0; i < gu.n; ++i)
      for (int j = 0; j < gt.n; ++j) put(gu.x[i], (1.0 - gu.x[i]) * gt.x[j], gu.w[i] * gt.w[j] * (1.0 - gu.x[i]));
  }
  return r;
}

inline ElemRule element_rule(int kind, int order) {
  return kind == HDD_SIMPLEX2D ? triangle_rule(order) : square_rule(order);
}

}  // namespace hdd
