/*
 * abides_dd_math.h -- deterministic, (almost always) correctly rounded log / exp / pow.
 *
 * Why this exists: the reference draws its random variates and evaluates its
 * estimators through libm (numpy legacy `log` in gauss/exponential,
 * `math.exp` in util/oracle/SparseMeanRevertingOracle.py:105-106, float `**`
 * in agent/ZeroIntelligenceAgent.py:198-231 and model/LatencyModel.py:138) and
 * then rounds to integers.  CUDA's libdevice and glibc differ in the last ulp
 * on a few percent of arguments, which would flip an `int(round(.))` every few
 * thousand draws and desynchronise a run.  Both the CUDA engine and the CPU
 * oracle therefore evaluate these three functions with the SAME double-double
 * algorithm below, built only from IEEE add/mul/fma (bit-identical on x86 and
 * sm_100a as long as no implicit contraction happens: gcc -ffp-contract=off,
 * nvcc -fmad=false).  The relative error before the final rounding is < 2^-85,
 * i.e. the result equals the correctly rounded one except when the exact value
 * lies within 2^-32 ulp of a rounding boundary; glibc (< 1 ulp, not correctly
 * rounded) agrees with it on all but ~1e-3 of arguments, and a disagreement only
 * matters when it also straddles an integer rounding boundary.
 *
 * Domain: finite x; log needs x > 0 and normal; exp/pow results must be normal.
 */
#ifndef ABIDES_DD_MATH_H
#define ABIDES_DD_MATH_H

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define ABM_HD __host__ __device__ __forceinline__
#ifdef ABM_INLINE_BIG
#define ABM_HD_BIG __host__ __device__ __forceinline__   /* lane engine: a fetched instruction serves 32 instances */
#else
#define ABM_HD_BIG __host__ __device__ __noinline__      /* keep the big bodies out of every call site */
#endif
#define ABM_FMA(a, b, c) fma((a), (b), (c))
#else
#define ABM_HD static inline
#define ABM_HD_BIG static
#define ABM_FMA(a, b, c) __builtin_fma((a), (b), (c))
#endif

typedef struct { double hi, lo; } abm_dd;

ABM_HD abm_dd abm_mk(double hi, double lo) { abm_dd r; r.hi = hi; r.lo = lo; return r; }

ABM_HD abm_dd abm_two_sum(double a, double b) {
  double s = a + b; double bb = s - a; double e = (a - (s - bb)) + (b - bb);
  return abm_mk(s, e);
}
ABM_HD abm_dd abm_fast_two_sum(double a, double b) {      /* |a| >= |b| */
  double s = a + b; double e = b - (s - a);
  return abm_mk(s, e);
}
ABM_HD abm_dd abm_two_prod(double a, double b) {
  double p = a * b; double e = ABM_FMA(a, b, -p);
  return abm_mk(p, e);
}
ABM_HD abm_dd abm_dd_add(abm_dd a, abm_dd b) {
  abm_dd s = abm_two_sum(a.hi, b.hi);
  abm_dd t = abm_two_sum(a.lo, b.lo);
  s.lo = s.lo + t.hi; s = abm_fast_two_sum(s.hi, s.lo);
  s.lo = s.lo + t.lo; s = abm_fast_two_sum(s.hi, s.lo);
  return s;
}
ABM_HD abm_dd abm_dd_add_d(abm_dd a, double b) {
  abm_dd s = abm_two_sum(a.hi, b);
  s.lo = s.lo + a.lo;
  return abm_fast_two_sum(s.hi, s.lo);
}
ABM_HD abm_dd abm_dd_neg(abm_dd a) { return abm_mk(-a.hi, -a.lo); }
ABM_HD abm_dd abm_dd_mul(abm_dd a, abm_dd b) {
  abm_dd p = abm_two_prod(a.hi, b.hi);
  p.lo = p.lo + (a.hi * b.lo + a.lo * b.hi);
  return abm_fast_two_sum(p.hi, p.lo);
}
ABM_HD abm_dd abm_dd_mul_d(abm_dd a, double b) {
  abm_dd p = abm_two_prod(a.hi, b);
  p.lo = p.lo + a.lo * b;
  return abm_fast_two_sum(p.hi, p.lo);
}
ABM_HD abm_dd abm_dd_scale2(abm_dd a, double pow2) {      /* exact: pow2 is a power of two */
  return abm_mk(a.hi * pow2, a.lo * pow2);
}
ABM_HD_BIG abm_dd abm_dd_div(abm_dd a, abm_dd b) {
  double q1 = a.hi / b.hi;
  abm_dd r = abm_dd_add(a, abm_dd_neg(abm_dd_mul_d(b, q1)));
  double q2 = r.hi / b.hi;
  r = abm_dd_add(r, abm_dd_neg(abm_dd_mul_d(b, q2)));
  double q3 = r.hi / b.hi;
  abm_dd q = abm_fast_two_sum(q1, q2);
  return abm_dd_add_d(q, q3);
}

ABM_HD double abm_bits_to_double(uint64_t u) { double d; memcpy(&d, &u, sizeof d); return d; }
ABM_HD uint64_t abm_double_to_bits(double d) { uint64_t u; memcpy(&u, &d, sizeof u); return u; }
/* 2^k for -1022 <= k <= 1023 */
ABM_HD double abm_pow2i(int k) { return abm_bits_to_double((uint64_t)(k + 1023) << 52); }

/* e^a - 1 for |a| <= 2^-10, Taylor to a^10/10! (truncation < 2^-115 relative). */
ABM_HD_BIG abm_dd abm_slow_expm1_tiny(abm_dd a) {
  abm_dd p = abm_mk(0x1.27e4fb7789f5cp-22, 0x1.cbbc05b4fa99ap-76);                       /* 1/10! */
  p = abm_dd_add(abm_dd_mul(p, a), abm_mk(0x1.71de3a556c734p-19, -0x1.c154f8ddc6c00p-73)); /* 1/9! */
  p = abm_dd_add(abm_dd_mul(p, a), abm_mk(0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-76));  /* 1/8! */
  p = abm_dd_add(abm_dd_mul(p, a), abm_mk(0x1.a01a01a01a01ap-13, 0x1.a01a01a01a01ap-73));  /* 1/7! */
  p = abm_dd_add(abm_dd_mul(p, a), abm_mk(0x1.6c16c16c16c17p-10, -0x1.f49f49f49f49fp-65)); /* 1/6! */
  p = abm_dd_add(abm_dd_mul(p, a), abm_mk(0x1.1111111111111p-7, 0x1.1111111111111p-63));   /* 1/5! */
  p = abm_dd_add(abm_dd_mul(p, a), abm_mk(0x1.5555555555555p-5, 0x1.5555555555555p-59));   /* 1/4! */
  p = abm_dd_add(abm_dd_mul(p, a), abm_mk(0x1.5555555555555p-3, 0x1.5555555555555p-57));   /* 1/3! */
  p = abm_dd_add_d(abm_dd_mul(p, a), 0.5);
  p = abm_dd_add_d(abm_dd_mul(p, a), 1.0);
  return abm_dd_mul(p, a);
}

/* exp of a double-double argument, |a| < 700, as an UNSCALED double-double m and an
 * exponent k with result = m * 2^k, m in [~0.7, ~1.42]. */
ABM_HD_BIG abm_dd abm_slow_exp_dd_k(abm_dd a, int* kout) {
  const abm_dd LN2 = abm_mk(0x1.62e42fefa39efp-1, 0x1.abc9e3b39803fp-56);
  double kf = rint(a.hi * 0x1.71547652b82fep+0);
  abm_dd r = abm_dd_add(a, abm_dd_neg(abm_dd_mul_d(LN2, kf)));
  r = abm_dd_scale2(r, 0x1p-9);                               /* |r| <= ln2/2 / 512 < 2^-10 */
  abm_dd p = abm_slow_expm1_tiny(r);
  for (int i = 0; i < 9; ++i)                                 /* (1+p)^2 - 1 = 2p + p^2 */
    p = abm_dd_add(abm_dd_scale2(p, 2.0), abm_dd_mul(p, p));
  *kout = (int)kf;
  return abm_dd_add_d(p, 1.0);
}

ABM_HD double abm_scale_by_pow2(double m, int k) {
  /* m * 2^k without double rounding in the normal range; splits large |k|. */
  if (k > 1000)  { m *= abm_pow2i(1000);  k -= 1000; }
  if (k < -1000) { m *= abm_pow2i(-1000); k += 1000; }
  if (k > 1023) k = 1023;
  if (k < -1022) k = -1022;
  return m * abm_pow2i(k);
}

ABM_HD_BIG double abm_slow_exp(double x) {
  if (x == 0.0) return 1.0;
  if (x > 709.0) return INFINITY;
  if (x < -708.0) return 0.0;
  int k; abm_dd m = abm_slow_exp_dd_k(abm_mk(x, 0.0), &k);
  return abm_scale_by_pow2(m.hi, k);
}

/* log(x) as a double-double; x > 0, normal. */
ABM_HD_BIG abm_dd abm_slow_log_dd(double x) {
  const abm_dd LN2 = abm_mk(0x1.62e42fefa39efp-1, 0x1.abc9e3b39803fp-56);
  uint64_t bits = abm_double_to_bits(x);
  int e = (int)((bits >> 52) & 0x7ff) - 1023;
  double m = abm_bits_to_double((bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL); /* [1,2) */
  if (m > 0x1.6a09e667f3bcdp+0) { m *= 0.5; e += 1; }                                 /* [sqrt.5, sqrt2) */
  /* s = (m-1)/(m+1) in double-double; m-1 is exact. */
  abm_dd num = abm_mk(m - 1.0, 0.0);
  abm_dd den = abm_two_sum(m, 1.0);
  abm_dd s = abm_dd_div(num, den);
  abm_dd lm;
  if (fabs(m - 1.0) < 0x1p-9) {
    /* near 1: 2*atanh(s) directly, |s| < 2^-10, s^2 < 2^-20: 6 terms reach 2^-120. */
    abm_dd s2 = abm_dd_mul(s, s);
    abm_dd p = abm_mk(0x1.745d1745d1746p-4, -0x1.745d1745d1746p-59);                       /* 1/11 */
    p = abm_dd_add(abm_dd_mul(p, s2), abm_mk(0x1.c71c71c71c71cp-4, 0x1.c71c71c71c71cp-58)); /* 1/9 */
    p = abm_dd_add(abm_dd_mul(p, s2), abm_mk(0x1.2492492492492p-3, 0x1.2492492492492p-57)); /* 1/7 */
    p = abm_dd_add(abm_dd_mul(p, s2), abm_mk(0x1.999999999999ap-3, -0x1.999999999999ap-57));/* 1/5 */
    p = abm_dd_add(abm_dd_mul(p, s2), abm_mk(0x1.5555555555555p-2, 0x1.5555555555555p-56)); /* 1/3 */
    p = abm_dd_add_d(abm_dd_mul(p, s2), 1.0);
    lm = abm_dd_scale2(abm_dd_mul(p, s), 2.0);
  } else {
    /* rough double estimate y0 of log(m) (atanh series, ~2^-50), then one Newton step
     * through the double-double exp: log m = y0 + log(m e^-y0) = y0 + t - t^2/2, t ~ 2^-50. */
    double sh = s.hi, z = sh * sh;
    double q = 1.0 / 23.0;
    q = q * z + 1.0 / 21.0; q = q * z + 1.0 / 19.0; q = q * z + 1.0 / 17.0; q = q * z + 1.0 / 15.0;
    q = q * z + 1.0 / 13.0; q = q * z + 1.0 / 11.0; q = q * z + 1.0 / 9.0;  q = q * z + 1.0 / 7.0;
    q = q * z + 1.0 / 5.0;  q = q * z + 1.0 / 3.0;  q = q * z + 1.0;
    double y0 = 2.0 * sh * q;
    int k; abm_dd em = abm_slow_exp_dd_k(abm_mk(-y0, 0.0), &k);   /* |y0| <= 0.35 => k == 0 or -1/+1 */
    em = abm_dd_scale2(em, abm_pow2i(k));
    abm_dd t = abm_dd_add_d(abm_dd_mul_d(em, m), -1.0);
    abm_dd t2h = abm_dd_scale2(abm_dd_mul(t, t), 0.5);
    lm = abm_dd_add(abm_dd_add_d(t, y0), abm_dd_neg(t2h));
  }
  if (e == 0) return lm;
  return abm_dd_add(abm_dd_mul_d(LN2, (double)e), lm);
}

ABM_HD double abm_slow_log(double x) { return abm_slow_log_dd(x).hi; }

/* x ** y for x > 0 (C pow semantics for the finite, positive-base cases the reference hits). */
ABM_HD_BIG double abm_slow_pow(double x, double y) {
  if (y == 0.0 || x == 1.0) return 1.0;
  if (y == 1.0) return x;
  abm_dd l = abm_slow_log_dd(x);
  abm_dd p = abm_dd_mul_d(l, y);
  if (p.hi > 709.0) return INFINITY;
  if (p.hi < -708.0) return 0.0;
  int k; abm_dd m = abm_slow_exp_dd_k(p, &k);
  return abm_scale_by_pow2(m.hi, k);
}

/* x ** y given a precomputed l = abm_slow_log_dd(x): bit-identical to abm_slow_pow(x, y). */
ABM_HD_BIG double abm_slow_pow_with_log(abm_dd l, double x, double y) {
  if (y == 0.0 || x == 1.0) return 1.0;
  if (y == 1.0) return x;
  abm_dd p = abm_dd_mul_d(l, y);
  if (p.hi > 709.0) return INFINITY;
  if (p.hi < -708.0) return 0.0;
  int k; abm_dd m = abm_slow_exp_dd_k(p, &k);
  return abm_scale_by_pow2(m.hi, k);
}


/* ------------------------------------------------------------------------------------------------
 * Production versions: table-driven, ~40 flops each, relative error before the final rounding
 * < 2^-66 (log) / 2^-69 (exp), i.e. the correctly rounded result except when the exact value lies
 * within ~2^-13 ulp of a rounding boundary.  The abm_slow_* functions above (error < 2^-85) are kept
 * as the cross-check (tests/test_math.py).  Tables: include/abides_math_tables.h (generated).
 * ---------------------------------------------------------------------------------------------- */
#include "abides_math_tables.h"

static const double abm_exp_tab_h[2 * ABM_EXP_N] = { ABM_EXP_TAB_VALUES };
static const double abm_log_tab_h[3 * ABM_LOG_N] = { ABM_LOG_TAB_VALUES };
#if defined(__CUDACC__)
static __device__ __constant__ double abm_exp_tab_d[2 * ABM_EXP_N] = { ABM_EXP_TAB_VALUES };
static __device__ __constant__ double abm_log_tab_d[3 * ABM_LOG_N] = { ABM_LOG_TAB_VALUES };
#endif
#if defined(__CUDA_ARCH__)
#define ABM_EXP_TAB abm_exp_tab_d
#define ABM_LOG_TAB abm_log_tab_d
#else
#define ABM_EXP_TAB abm_exp_tab_h
#define ABM_LOG_TAB abm_log_tab_h
#endif

/* log(x) as a double-double; x > 0, normal. */
ABM_HD_BIG abm_dd abm_log_dd(double x) {
  uint64_t bits = abm_double_to_bits(x);
  int e = (int)((bits >> 52) & 0x7ff) - 1023;
  double z = abm_bits_to_double((bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);  /* [1,2) */
  if (z >= 1.5) { z *= 0.5; e += 1; }                                                   /* [0.75,1.5) */
  int i = z < 1.0 ? (int)((z - 0.75) * 256.0) : 64 + (int)((z - 1.0) * 128.0);
  double invc = ABM_LOG_TAB[3 * i], lch = ABM_LOG_TAB[3 * i + 1], lcl = ABM_LOG_TAB[3 * i + 2];
  /* r = z * invc - 1 exactly as rh + rl (|r| <= 2^-7; r = z - 1 exactly in the two cells touching 1) */
  double ph = z * invc, pl = ABM_FMA(z, invc, -ph);
  abm_dd r = abm_fast_two_sum(ph - 1.0, pl);
  double h = r.hi;
  /* log1p(r) = r - r^2/2 + h^3 (1/3 - h/4 + ... + h^8/11) */
  double sqh = h * h, sql = ABM_FMA(h, h, -sqh);
  double t2h = -0.5 * sqh, t2l = -0.5 * sql - h * r.lo;
  double q = 1.0 / 11.0;
  q = ABM_FMA(q, h, -1.0 / 10.0); q = ABM_FMA(q, h, 1.0 / 9.0); q = ABM_FMA(q, h, -1.0 / 8.0);
  q = ABM_FMA(q, h, 1.0 / 7.0);   q = ABM_FMA(q, h, -1.0 / 6.0); q = ABM_FMA(q, h, 1.0 / 5.0);
  q = ABM_FMA(q, h, -1.0 / 4.0);  q = ABM_FMA(q, h, 1.0 / 3.0);
  q = q * (sqh * h);
  abm_dd s1 = abm_two_sum(lch, h);
  abm_dd s2 = abm_two_sum(s1.hi, t2h);
  double lo = (s1.lo + s2.lo) + (((lcl + r.lo) + t2l) + q);
  abm_dd res = abm_fast_two_sum(s2.hi, lo);
  if (e != 0) {
    double eh = (double)e * ABM_LN2_HI42, el = (double)e * ABM_LN2_LO42;      /* eh exact */
    abm_dd u = abm_two_sum(eh, res.hi);
    res = abm_fast_two_sum(u.hi, u.lo + (res.lo + el));
  }
  return res;
}

ABM_HD double abm_log(double x) { return abm_log_dd(x).hi; }

/* exp(ah + al), |ah| < 709, |al| <~ ulp(ah) */
ABM_HD_BIG double abm_exp_dd(double ah, double al) {
  double kf = rint(ah * ABM_128_LN2);
  abm_dd rr = abm_two_sum(ABM_FMA(-kf, ABM_LN2_128_HI, ah) /* exact */, ABM_FMA(-kf, ABM_LN2_128_LO, al));
  double r = rr.hi, rl = rr.lo;                                /* |r| <= ln2/256, |rl| <= ulp(r)/2 */
  int k = (int)kf, j = k & (ABM_EXP_N - 1), e = k >> 7;
  double th = ABM_EXP_TAB[2 * j], tl = ABM_EXP_TAB[2 * j + 1];
  double p = ABM_FMA(r, 1.0 / 720.0, 1.0 / 120.0);
  p = ABM_FMA(r, p, 1.0 / 24.0); p = ABM_FMA(r, p, 1.0 / 6.0); p = ABM_FMA(r, p, 0.5);
  double q = ABM_FMA(r * r, p, rl);                            /* expm1(r + rl) - r */
  double sh = th * r, sl = ABM_FMA(th, r, -sh);
  double lo = sl + ABM_FMA(th, q, tl);
  abm_dd u = abm_fast_two_sum(th, sh);
  double m = u.hi + (u.lo + lo);
  return abm_scale_by_pow2(m, e);
}

ABM_HD double abm_exp(double x) {
  if (x > 709.0) return INFINITY;
  if (x < -708.0) return 0.0;
  return abm_exp_dd(x, 0.0);
}

/* x ** y given l = abm_log_dd(x), x > 0 (C pow semantics for the finite cases the reference hits). */
ABM_HD_BIG double abm_pow_with_log(abm_dd l, double x, double y) {
  if (y == 0.0 || x == 1.0) return 1.0;
  if (y == 1.0) return x;
  double ph = l.hi * y, pl = ABM_FMA(l.hi, y, -ph) + l.lo * y;
  if (ph > 709.0) return INFINITY;
  if (ph < -708.0) return 0.0;
  return abm_exp_dd(ph, pl);
}
ABM_HD double abm_pow(double x, double y) {
  if (y == 0.0 || x == 1.0) return 1.0;
  if (y == 1.0) return x;
  return abm_pow_with_log(abm_log_dd(x), x, y);
}

/* x ** 3 (model/LatencyModel.py:138): the exact cube in double-double, rounded once. */
ABM_HD double abm_cube(double x) {
  abm_dd sq = abm_two_prod(x, x);
  abm_dd c = abm_dd_mul_d(sq, x);
  return c.hi;
}

#endif /* ABIDES_DD_MATH_H */
