// Stand-in for the linear-interpolation subset of GSL the reference uses
// (gsl_spline_{alloc,init,eval} with gsl_interp_linear). Semantics follow GSL's linear_eval:
// bsearch for x[i] <= xv < x[i+1] clamped to [0, n-2], y_i + (xv-x_i)/(x_{i+1}-x_i)*(y_{i+1}-y_i).
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <vector>
struct gsl_interp_type {};
static const gsl_interp_type gsl_interp_linear_obj{};
static const gsl_interp_type* gsl_interp_linear = &gsl_interp_linear_obj;
struct gsl_interp_accel { size_t cache = 0; };
struct gsl_spline { std::vector<double> x, y; };
inline gsl_spline* gsl_spline_alloc(const gsl_interp_type*, size_t n) { auto* s = new gsl_spline; s->x.resize(n); s->y.resize(n); return s; }
inline int gsl_spline_init(gsl_spline* s, const double* x, const double* y, size_t n) { s->x.assign(x, x + n); s->y.assign(y, y + n); return 0; }
inline gsl_interp_accel* gsl_interp_accel_alloc() { return new gsl_interp_accel; }
inline double gsl_spline_eval(const gsl_spline* s, double xv, gsl_interp_accel*) {
  const auto& x = s->x; size_t n = x.size(); size_t lo = 0, hi = n - 1;
  while (hi > lo + 1) { size_t m = (hi + lo) / 2; if (x[m] > xv) hi = m; else lo = m; }
  double dx = x[lo + 1] - x[lo];
  return s->y[lo] + (xv - x[lo]) / dx * (s->y[lo + 1] - s->y[lo]);
}
