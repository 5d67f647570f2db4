/* C restatement of the arithmetic the reference takes from OUTSIDE its own tree for the normal
 * scores -- TEST INFRASTRUCTURE ONLY (checker of csrc/as241.cuh; never linked into the product).
 *
 *   z = statistics.NormalDist().inv_cdf(p)         diagnostics.py:124,131
 *     = Wichura AS241 / PPND16 in the operation order of CPython's
 *       Modules/_statisticsmodule.c (_normal_dist_inv_cdf), whose tail branches call the C library's log().
 *
 * glibc >= 2.28's log() is ARM Optimized Routines' table-driven algorithm (math/log.c, MIT licence); on an
 * x86-64 CPU with FMA the loader selects the build whose multiply-adds are fused.  `oracle_log` restates that
 * build operation by operation (the fused operations are explicit fma() calls here; everything else is
 * compiled with -ffp-contract=off), with the constants of oracle/log_table.h (tools/gen_log_table.py).
 * Pinned by tests/test_oracle_golden.py: oracle_log == math.log and oracle_inv_cdf ==
 * statistics.NormalDist().inv_cdf, bit for bit, over millions of points of the domain the path uses
 * (p = (k/2)/M, tails p <= 0.075).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "log_table.h"

static const double LOG_TAB[256] = {B200REF_LOG_TABLE};

static inline uint64_t as_u64(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double as_f64(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

/* log(x) for normal x outside [1 - 2^-4, 1 + 0x1.09p-4) -- the only range the AS241 tails reach (x <= 0.075).
 * Inside that range (and for zero / subnormal / non-finite x) the C library is called. */
double oracle_log(double x) {
  const uint64_t ix = as_u64(x);
  const uint64_t LO = 0x3fee000000000000ull, HI = 0x3ff1090000000000ull;
  const uint32_t top = (uint32_t)(ix >> 48);
  if (ix - LO < HI - LO || top - 0x0010u >= 0x7ff0u - 0x0010u) return log(x);
  const uint64_t OFF = 0x3fe6000000000000ull;
  const uint64_t tmp = ix - OFF;
  const int i = (int)((tmp >> 45) & 127u);
  const int64_t k = (int64_t)tmp >> 52;
  const uint64_t iz = ix - (tmp & (0xfffull << 52));
  const double invc = LOG_TAB[2 * i], logc = LOG_TAB[2 * i + 1];
  const double z = as_f64(iz);
  const double r = fma(z, invc, -1.0);
  const double kd = (double)k;
  const double w = fma(kd, B200REF_LOG_LN2HI, logc);
  const double hi = w + r;
  const double lo = fma(kd, B200REF_LOG_LN2LO, (w - hi) + r);
  const double r2 = r * r;
  const double r3 = r * r2;
  const double t1 = fma(r, B200REF_LOG_A2, B200REF_LOG_A1);
  const double t2 = fma(r, B200REF_LOG_A4, B200REF_LOG_A3);
  const double t3 = fma(r2, B200REF_LOG_A0, lo);
  const double t4 = fma(t2, r2, t1);
  const double t5 = fma(r3, t4, t3);
  return t5 + hi;
}

#define H8(c0, c1, c2, c3, c4, c5, c6, c7, r) \
  ((((((((c0) * (r) + (c1)) * (r) + (c2)) * (r) + (c3)) * (r) + (c4)) * (r) + (c5)) * (r) + (c6)) * (r) + (c7))

double oracle_inv_cdf(double p) {
  const double q = p - 0.5;
  double num, den, r, x;
  if (fabs(q) <= 0.425) {
    r = 0.180625 - q * q;
    num = H8(2.5090809287301226727e+3, 3.3430575583588128105e+4, 6.7265770927008700853e+4, 4.5921953931549871457e+4,
             1.3731693765509461125e+4, 1.9715909503065514427e+3, 1.3314166789178437745e+2, 3.3871328727963666080e+0, r) * q;
    den = H8(5.2264952788528545610e+3, 2.8729085735721942674e+4, 3.9307895800092710610e+4, 2.1213794301586595867e+4,
             5.3941960214247511077e+3, 6.8718700749205790830e+2, 4.2313330701600911252e+1, 1.0, r);
    x = num / den;
    return 0.0 + x;
  }
  r = (q <= 0.0) ? p : 1.0 - p;
  r = sqrt(-oracle_log(r));
  if (r <= 5.0) {
    r = r - 1.6;
    num = H8(7.74545014278341407640e-4, 2.27238449892691845833e-2, 2.41780725177450611770e-1, 1.27045825245236838258e+0,
             3.64784832476320460504e+0, 5.76949722146069140550e+0, 4.63033784615654529590e+0, 1.42343711074968357734e+0, r);
    den = H8(1.05075007164441684324e-9, 5.47593808499534494600e-4, 1.51986665636164571966e-2, 1.48103976427480074590e-1,
             6.89767334985100004550e-1, 1.67638483018380384940e+0, 2.05319162663775882187e+0, 1.0, r);
  } else {
    r = r - 5.0;
    num = H8(2.01033439929228813265e-7, 2.71155556874348757815e-5, 1.24266094738807843860e-3, 2.65321895265761230930e-2,
             2.96560571828504891230e-1, 1.78482653991729133580e+0, 5.46378491116411436990e+0, 6.65790464350110377720e+0, r);
    den = H8(2.04426310338993978564e-15, 1.42151175831644588870e-7, 1.84631831751005468180e-5, 7.86869131145613259100e-4,
             1.48753612908506148525e-2, 1.36929880922735805310e-1, 5.99832206555887937690e-1, 1.0, r);
  }
  x = num / den;
  if (q < 0.0) x = -x;
  return 0.0 + x;
}

void oracle_log_array(const double* x, double* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = oracle_log(x[i]);
}

/* T[k] = inv_cdf((k / 2) / M), k = 1 .. 2M-1; T[0] = 0 (unused) -- the table k_ztable builds on the device */
void oracle_ztable(int64_t m, double* out) {
  out[0] = 0.0;
  for (int64_t k = 1; k < 2 * m; ++k) out[k] = oracle_inv_cdf(((double)k * 0.5) / (double)m);
}
