// Normal-score table: z = Phi^-1(p) by Wichura's AS241 (PPND16), evaluated in the
// operation order of CPython's statistics._normal_dist_inv_cdf, which is what the
// reference calls (diagnostics.py:124,131).  Every multiply and add is an explicit
// round-to-nearest intrinsic so nvcc cannot contract a*b+c into an FMA: the
// central branch (|p - 0.5| <= 0.425, 85 % of draws) is pure + * / and therefore
// bit-identical to the CPU.  The tail branches go through log(): the stdlib takes it from the C
// library, and glibc's log (ARM Optimized Routines' table-driven algorithm, fused multiply-adds in the
// build x86-64 CPUs with FMA run) is restated here operation by operation (`glibc_log`, constants in
// log_table.cuh), so the tails are bit-identical to the stdlib too (tests: 0 ulp over the whole table).
#pragma once
#include "common.cuh"
#include "log_table.cuh"

namespace b200ref {

__device__ __forceinline__ double rmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double radd(double a, double b) { return __dadd_rn(a, b); }

#define B200REF_HORNER8(c0, c1, c2, c3, c4, c5, c6, c7, r)                                      \
  radd(rmul(radd(rmul(radd(rmul(radd(rmul(radd(rmul(radd(rmul(radd(rmul((c0), (r)), (c1)), (r)), \
       (c2)), (r)), (c3)), (r)), (c4)), (r)), (c5)), (r)), (c6)), (r)), (c7))

__device__ const double kLogTab[256] = {B200REF_LOG_TABLE};

// log(x) for normal x outside [1 - 2^-4, 1 + 0x1.09p-4): all the AS241 tails need (x <= 0.075).
// x = 2^k z with z in [0x1.6p-1, 0x1.6p0); c_i = centre of z's subinterval; r = z / c_i - 1 (one fma on the
// tabulated 1 / c_i); log x = k ln2 + log c_i + log1p(r), log1p by a degree-5 polynomial.  The fused and the
// separately rounded operations are exactly those of glibc's FMA build (checked against its machine code and,
// through the C restatement the tests load, against math.log on millions of points).
__device__ __forceinline__ double glibc_log(double x) {
  const uint64_t ix = (uint64_t)__double_as_longlong(x);
  const uint64_t tmp = ix - 0x3fe6000000000000ull;
  const int i = (int)((tmp >> 45) & 127u);
  const long long k = (long long)tmp >> 52;
  const double z = __longlong_as_double((long long)(ix - (tmp & (0xfffull << 52))));
  const double invc = kLogTab[2 * i], logc = kLogTab[2 * i + 1];
  const double r = __fma_rn(z, invc, -1.0);
  const double kd = (double)k;
  const double w = __fma_rn(kd, B200REF_LOG_LN2HI, logc);
  const double hi = __dadd_rn(w, r);
  const double lo = __fma_rn(kd, B200REF_LOG_LN2LO, __dadd_rn(__dsub_rn(w, hi), r));
  const double r2 = __dmul_rn(r, r);
  const double r3 = __dmul_rn(r, r2);
  const double t1 = __fma_rn(r, B200REF_LOG_A2, B200REF_LOG_A1);
  const double t2 = __fma_rn(r, B200REF_LOG_A4, B200REF_LOG_A3);
  const double t3 = __fma_rn(r2, B200REF_LOG_A0, lo);
  const double t4 = __fma_rn(t2, r2, t1);
  const double t5 = __fma_rn(r3, t4, t3);
  return __dadd_rn(t5, hi);
}

__device__ __forceinline__ double as241_inv_cdf(double p) {
  const double q = __dsub_rn(p, 0.5);
  if (fabs(q) <= 0.425) {
    const double r = __dsub_rn(0.180625, rmul(q, q));
    const double num = rmul(
        B200REF_HORNER8(2.5090809287301226727e+3, 3.3430575583588128105e+4, 6.7265770927008700853e+4,
                        4.5921953931549871457e+4, 1.3731693765509461125e+4, 1.9715909503065514427e+3,
                        1.3314166789178437745e+2, 3.3871328727963666080e+0, r),
        q);
    const double den =
        B200REF_HORNER8(5.2264952788528545610e+3, 2.8729085735721942674e+4, 3.9307895800092710610e+4,
                        2.1213794301586595867e+4, 5.3941960214247511077e+3, 6.8718700749205790830e+2,
                        4.2313330701600911252e+1, 1.0, r);
    // stdlib returns mu + x * sigma with mu = 0, sigma = 1: maps -0.0 to +0.0
    return radd(0.0, __ddiv_rn(num, den));
  }
  double r = (q <= 0.0) ? p : __dsub_rn(1.0, p);
  r = __dsqrt_rn(-glibc_log(r));  // r <= 0.075: inside glibc_log's domain
  double num, den;
  if (r <= 5.0) {
    r = __dsub_rn(r, 1.6);
    num = B200REF_HORNER8(7.74545014278341407640e-4, 2.27238449892691845833e-2, 2.41780725177450611770e-1,
                          1.27045825245236838258e+0, 3.64784832476320460504e+0, 5.76949722146069140550e+0,
                          4.63033784615654529590e+0, 1.42343711074968357734e+0, r);
    den = B200REF_HORNER8(1.05075007164441684324e-9, 5.47593808499534494600e-4, 1.51986665636164571966e-2,
                          1.48103976427480074590e-1, 6.89767334985100004550e-1, 1.67638483018380384940e+0,
                          2.05319162663775882187e+0, 1.0, r);
  } else {
    r = __dsub_rn(r, 5.0);
    num = B200REF_HORNER8(2.01033439929228813265e-7, 2.71155556874348757815e-5, 1.24266094738807843860e-3,
                          2.65321895265761230930e-2, 2.96560571828504891230e-1, 1.78482653991729133580e+0,
                          5.46378491116411436990e+0, 6.65790464350110377720e+0, r);
    den = B200REF_HORNER8(2.04426310338993978564e-15, 1.42151175831644588870e-7, 1.84631831751005468180e-5,
                          7.86869131145613259100e-4, 1.48753612908506148525e-2, 1.36929880922735805310e-1,
                          5.99832206555887937690e-1, 1.0, r);
  }
  double x = __ddiv_rn(num, den);
  if (q < 0.0) x = -x;
  return radd(0.0, x);
}

// T[k] = inv_cdf((k / 2) / M), k = 1 .. 2M-1.  With i / j the 0-based inclusive /
// exclusive bounds of a tie run in the pooled ascending order, the reference's
// avg_rank = (i + 1 + j) / 2.0 (diagnostics.py:118) and p = (avg_rank - 0.5) / n
// (:130) give p = ((i + j) / 2) / n with a single rounding (the divide), so
// k = i + j indexes every value z can take.  2M entries, L2-resident.
// odd[t] = T[2t + 1] is the score of an untied draw at sorted position t: a dense copy that the
// fold kernel reads coalesced (ties, the rare case, go to T itself).
__global__ void k_ztable(double* __restrict__ tab, double* __restrict__ odd, uint32_t m) {
  const uint32_t n = 2u * m;
  const double dm = (double)m;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    double z = 0.0;
    if (k > 0) z = as241_inv_cdf(__ddiv_rn(rmul((double)k, 0.5), dm));
    tab[k] = z;
    if (k & 1u) odd[k >> 1] = z;
  }
}

}  // namespace b200ref
