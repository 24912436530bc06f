// Device building blocks of the per-star path: Philox4x32-10, table search and
// interpolation with NumPy/SciPy semantics, HEALPix ang2pix.  sm_100a only.
//
// Arithmetic that feeds an integer index or a comparison (pixel numbers, CDF
// index, efficiency compare) is written with explicit round-to-nearest
// intrinsics so nvcc cannot contract a*b+c into an FMA: the reference runs
// un-fused IEEE double arithmetic (NumPy / healpix_cxx on x86-64).
#pragma once
#include <stdint.h>
#include "../../include/streamsim_b200.h"

// rare per-star paths: out of line (__noinline__) or inline in a cold branch
#ifndef SS_COLD
#define SS_COLD __forceinline__
#endif

namespace ss {

__device__ __forceinline__ double nan_d() { return __longlong_as_double(0x7ff8000000000000LL); }

// ---------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// The same ten rounds with the key schedule precomputed by the host (rk[2r], rk[2r+1] =
// key + r * Weyl constants): as launch constants the round keys are constant-bank operands
// of the XORs, so a round is two wide multiplies and two three-input XORs.
struct PhiloxKeys {
  uint32_t rk[20];
  // fold[b] = (rk[0] ^ hi(M1 * b), rk[2] ^ lo(M1 * b)), M1 = 0xCD9E8D57: philox_block_c
  uint32_t fold[2][2];
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKeys& K) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ K.rk[2 * r], lo1, hi0 ^ c.w ^ K.rk[2 * r + 1], lo0);
  }
  return c;
}

__device__ __forceinline__ uint4 philox_block(unsigned long long star, uint32_t block,
                                              const PhiloxKeys& K) {
  return philox4x32_10(make_uint4((uint32_t)star, (uint32_t)(star >> 32), block, 0u), K);
}

// The same block for a compile-time block number, with that constant folded into the keys by
// the host (PhiloxKeys.fold).  In rounds 1 and 2 the block number reaches a key XOR as a
// compile-time constant -- (constant ^ register ^ key) does not fit one LOP3 with the key as a
// constant-bank operand, so the compiler fetches the key with LDC into a register; behind a
// batch of table loads that LDC shares their scoreboard and the round waits for the loads
// (6.7 % of the production kernel's stall samples).  With the constant folded every XOR is
// (register ^ register ^ uniform).
template <uint32_t BLOCK>
__device__ __forceinline__ uint4 philox_block_c(unsigned long long star, const PhiloxKeys& K) {
  const uint32_t s_lo = (uint32_t)star, s_hi = (uint32_t)(star >> 32);
  // round 1: counter (s_lo, s_hi, BLOCK, 0) -> (hi(M1 BLOCK) ^ s_hi ^ k0, M1 BLOCK, hi(M0 s_lo) ^ k1, M0 s_lo)
  uint32_t x = s_hi ^ K.fold[BLOCK][0];
  uint32_t z = __umulhi(0xD2511F53u, s_lo) ^ K.rk[1], w = 0xD2511F53u * s_lo;
  // round 2: y = M1 BLOCK is folded into fold[BLOCK][1]
  uint4 c;
  {
    const uint32_t hi0 = __umulhi(0xD2511F53u, x), lo0 = 0xD2511F53u * x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, z), lo1 = 0xCD9E8D57u * z;
    c = make_uint4(hi1 ^ K.fold[BLOCK][1], lo1, hi0 ^ w ^ K.rk[3], lo0);
  }
#pragma unroll
  for (int r = 2; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ K.rk[2 * r], lo1, hi0 ^ c.w ^ K.rk[2 * r + 1], lo0);
  }
  return c;
}

// two 32-bit words -> double in [0,1) with 52 random bits: the bits become the
// mantissa of a double in [1,2), minus 1 (two integer ops and one DADD; no I2F)
__device__ __forceinline__ double u52(uint32_t a, uint32_t b) {
  return __hiloint2double((int)(0x3ff00000u | (a >> 12)), (int)b) - 1.0;
}

__device__ __forceinline__ uint4 philox_block(unsigned long long star, uint32_t block,
                                              unsigned long long seed) {
  const uint4 c = make_uint4((uint32_t)star, (uint32_t)(star >> 32), block, 0u);
  return philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

// one 32-bit word -> the double w (exact): the word becomes the low mantissa bits of 2^52
// (one DADD, no I2F)
__device__ __forceinline__ double word_to_double(uint32_t w) {
  return __hiloint2double(0x43300000, (int)w) - 4503599627370496.0;
}

// one 32-bit word -> double in [0,1), exact: w * 2^-32
__device__ __forceinline__ double u32(uint32_t w) { return word_to_double(w) * (1.0 / 4294967296.0); }

// Box-Muller from two 32-bit words, evaluated in fp32 (like cuRAND's curand_normal): the
// draw is a random number, not a quantity with a parity requirement, and everything
// computed *from* it downstream stays in double.  Radius from y = (~wa + 0.5) / 2^32 in
// (0, 1): small y (the tail) keeps full relative precision in fp32; |z| <= 6.7.
__device__ __forceinline__ void bm32f(uint32_t wa, uint32_t wb, float& z0, float& z1) {
  const float y = ((float)(~wa) + 0.5f) * 2.3283064365386963e-10f;
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * __log2f(y)));
  // angle = 2 pi wb / 2^32, taken modulo 2 pi into [-pi, pi) (wb read as a signed word) where
  // the MUFU sine / cosine are good to 5e-7 absolute
  const float ang = (float)(int)wb * 4.6566128730773926e-10f * 3.14159265358979f;
  z0 = r * __cosf(ang);
  z1 = r * __sinf(ang);
}

__device__ __forceinline__ void bm32(uint32_t wa, uint32_t wb, double& z0, double& z1) {
  float a, b;
  bm32f(wa, wb, a, b);
  z0 = (double)a;
  z1 = (double)b;
}

// ---------------------------------------------------------------- elementary functions
// Own double-precision log10 / 10^x / atan2 / small-angle sincos for the hot loop.  The
// CUDA library versions are accurate but spell every polynomial coefficient as two
// 32-bit immediates (two UMOV per DFMA: 12 % of the kernel's issued instructions in the
// round-1 ncu captures); here the coefficients sit in the constant bank and arrive two
// per LDCU.128.  Designed for <= 2 ulp on their stated domains; tests/test_gpu_math.py holds
// them to <= 4 ulp of NumPy (sqrt: <= 1 ulp) through ss_math_probe.
__constant__ __align__(16) double kSinC[8] = {
    1.0, -1.0 / 6, 1.0 / 120, -1.0 / 5040, 1.0 / 362880, -1.0 / 39916800, 1.0 / 6227020800.0,
    -1.0 / 1307674368000.0};
__constant__ __align__(16) double kCosC[8] = {
    -0.5, 1.0 / 24, -1.0 / 720, 1.0 / 40320, -1.0 / 3628800, 1.0 / 479001600,
    -1.0 / 87178291200.0, 1.0 / 20922789888000.0};
// (2/ln 10)/(2k+1), k = 0..10: log10(m) = s * sum_k kLogC[k] s^(2k), s = (m-1)/(m+1)
__constant__ __align__(16) double kLogC[12] = {
    0.8685889638065036, 0.2895296546021679, 0.17371779276130073, 0.12408413768664338,
    0.09650988486738929, 0.07896263307331851, 0.06681453567742336, 0.05790593092043358,
    0.0510934684592061, 0.04571520862139493, 0.041361379228881126, 0.0};
// (ln 2)^k / k!, k = 1..14
__constant__ __align__(16) double kExp2C[14] = {
    0.6931471805599453, 0.24022650695910072, 0.05550410866482158, 0.009618129107628477,
    0.0013333558146428443, 0.0001540353039338161, 1.5252733804059841e-05,
    1.321548679014431e-06, 1.01780860092397e-07, 7.054911620801123e-09,
    4.4455382718708116e-10, 2.5678435993488206e-11, 1.3691488853904128e-12,
    6.778726354822545e-14};
// (-1)^k/(2k+1), k = 1..7: atan(r) = r + r^3 * (...)
__constant__ __align__(16) double kAtanC[8] = {
    -1.0 / 3, 1.0 / 5, -1.0 / 7, 1.0 / 9, -1.0 / 11, 1.0 / 13, -1.0 / 15, 0.0};
// atan(i/8) as (hi, lo), i = 0..8
__constant__ __align__(16) double kAtanT[32] = {
    0.0, 0.0,
    0.12435499454676144, -3.1253241424539383e-18,
    0.24497866312686414, 1.0698755618734451e-17,
    0.35877067027057225, -2.4623815582638635e-17,
    0.4636476090008061, 2.2698777452961687e-17,
    0.5585993153435624, -5.4556305485916264e-18,
    0.6435011087932844, 1.5834785051444286e-17,
    0.7188299996216245, -2.1478388444456983e-17,
    0.7853981633974483, 3.061616997868383e-17};
// scalars: [0] log2(10) hi, [1] lo, [2] log10(2), [3] pi/2 hi, [4] lo, [5] pi hi, [6] lo,
// [7] 180/pi, [8] pi/180, [9] 2/pi, [10] 1.5*2^52, [11] 1/1.0857362
__constant__ __align__(16) double kNum[12] = {
    3.321928094887362, 1.661617516973592e-16, 0.3010299956639812, 1.5707963267948966,
    6.123233995736766e-17, 3.141592653589793, 1.2246467991473532e-16, 57.29577951308232,
    0.017453292519943295, 0.6366197723675814, 6755399441055744.0, 1.0 / 1.0857362};

// The *_nb functions are branch-free (no fallback, no slow path) so that several of them
// placed in one basic block interleave their dependency chains: the kernel is bound by
// the latency of dependent double-precision instructions at 8 warps per scheduler, not
// by instruction count.  Their domains are the callers' responsibility.

// 1/d for a normal positive double: MUFU.RCP64H seed (20 bits) + one Newton step (40 bits);
// div_pos finishes with a residual correction of the quotient
__device__ __forceinline__ double rcp40_pos(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  return fma(fma(-d, r, 1.0), r, r);
}

// a/d, d normal positive (<= 1 ulp)
__device__ __forceinline__ double div_pos(double a, double d) {
  const double r = rcp40_pos(d);
  const double q = a * r;
  return fma(fma(-q, d, a), r, q);
}

// sqrt(a), a normal positive or 0 (NaN propagates); correctly rounded up to rare ties
__device__ __forceinline__ double sqrt_nb(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double g = a * y, h = 0.5 * y;
  double r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  g = fma(fma(-g, g, a), h, g);
  return (a == 0.0) ? 0.0 : g;
}

// log10(t) for normal t > 0 (anything else: garbage, the caller selects)
__device__ __forceinline__ double log10_nb(double t) {
  int hi = __double2hiint(t);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  const bool up = hi >= 0x3ff6a09f;                        // m in [sqrt(1/2), sqrt 2)
  hi -= up ? 0x00100000 : 0;
  e += up ? 1 : 0;
  const double m = __hiloint2double(hi, __double2loint(t));
  const double s = div_pos(m - 1.0, m + 1.0);
  const double s2 = s * s, s4 = s2 * s2;
  // sum_{k=1..10} kLogC[k] s2^(k-1), even / odd halves evaluated side by side
  double pe = kLogC[9], po = kLogC[10];
  pe = fma(pe, s4, kLogC[7]); po = fma(po, s4, kLogC[8]);
  pe = fma(pe, s4, kLogC[5]); po = fma(po, s4, kLogC[6]);
  pe = fma(pe, s4, kLogC[3]); po = fma(po, s4, kLogC[4]);
  pe = fma(pe, s4, kLogC[1]); po = fma(po, s4, kLogC[2]);
  const double p = fma(po, s2, pe);
  const double lm = fma(s * s2, p, s * kLogC[0]);          // log10(m)
  return fma((double)e, kNum[2], lm);
}

// 10^x for |x| < 300 (NaN propagates)
__device__ __forceinline__ double exp10_nb(double x) {
  const double t = fma(x, kNum[0], kNum[10]);               // round(x log2 10) in the low word
  const int n = __double2loint(t);
  const double nd = t - kNum[10];
  const double f = fma(x, kNum[1], fma(x, kNum[0], -nd));   // |f| <= 1/2
  const double f2 = f * f;
  // 2^f = 1 + f * sum_{k=0..13} kExp2C[k] f^k, even / odd halves side by side
  double pe = kExp2C[12], po = kExp2C[13];
#pragma unroll
  for (int k = 10; k >= 0; k -= 2) {
    pe = fma(pe, f2, kExp2C[k]);
    po = fma(po, f2, kExp2C[k + 1]);
  }
  const double p = fma(fma(po, f, pe), f, 1.0);
  return p * __hiloint2double((n + 1023) << 20, 0);
}

__device__ __forceinline__ double exp10_fast(double x) {
  return (fabs(x) < 300.0 || x != x) ? exp10_nb(x) : exp10(x);
}

// atan2(y, x) for finite arguments whose larger magnitude is a normal number (0 for
// x = y = 0).  q = min/max in [0,1] is reduced against the nearest i/8:
// atan(q) = atan(i/8) + atan((q - c)/(1 + q c)), |r| <= 1/16, one division.
__device__ __forceinline__ double atan2_nb(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  const bool steep = ay > ax;
  const double mx = steep ? ay : ax, mn = steep ? ax : ay;
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(mx));
  const double t = fma(mn * r0, 8.0, kNum[10]);
  const int i = __double2loint(t) & 15;                     // 0..8
  const double c = (t - kNum[10]) * 0.125;
  const double r = div_pos(fma(-c, mx, mn), fma(c, mn, mx));
  const double r2 = r * r, r4 = r2 * r2;
  // atan(r) = r + r^3 sum_k kAtanC[k] r2^k (k = 0..6), even / odd halves side by side
  double pe = fma(kAtanC[6], r4, kAtanC[4]), po = kAtanC[5];
  pe = fma(pe, r4, kAtanC[2]); po = fma(po, r4, kAtanC[3]);
  pe = fma(pe, r4, kAtanC[0]); po = fma(po, r4, kAtanC[1]);
  const double p = fma(po, r2, pe);
  double a = kAtanT[2 * i] + (fma(r * r2, p, r) + kAtanT[2 * i + 1]);
  a = steep ? (kNum[3] - a) + kNum[4] : a;
  a = (x < 0.0) ? (kNum[5] - a) + kNum[6] : a;
  a = (mx == 0.0) ? 0.0 : a;
  return (y < 0.0) ? -a : a;
}

// Two-wide forms: the same operations, element by element, as the scalar functions above (so
// the results are bit-identical), written out side by side so that every polynomial
// coefficient is fetched from the constant bank once for both elements -- constant loads were
// 13 % of the production kernel's issued instructions -- and the two chains interleave.
__device__ __forceinline__ void div_pos2(double a0, double d0, double a1, double d1, double& q0,
                                         double& q1) {
  double r0, r1;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d0));
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r1) : "d"(d1));
  r0 = fma(fma(-d0, r0, 1.0), r0, r0);
  r1 = fma(fma(-d1, r1, 1.0), r1, r1);
  q0 = a0 * r0;
  q1 = a1 * r1;
  q0 = fma(fma(-q0, d0, a0), r0, q0);
  q1 = fma(fma(-q1, d1, a1), r1, q1);
}

__device__ __forceinline__ void log10_nb2(double t0, double t1, double& o0, double& o1) {
  int h0 = __double2hiint(t0), h1 = __double2hiint(t1);
  int e0 = (h0 >> 20) - 1023, e1 = (h1 >> 20) - 1023;
  h0 = (h0 & 0x000fffff) | 0x3ff00000;
  h1 = (h1 & 0x000fffff) | 0x3ff00000;
  const bool up0 = h0 >= 0x3ff6a09f, up1 = h1 >= 0x3ff6a09f;
  h0 -= up0 ? 0x00100000 : 0;
  h1 -= up1 ? 0x00100000 : 0;
  e0 += up0 ? 1 : 0;
  e1 += up1 ? 1 : 0;
  const double m0 = __hiloint2double(h0, __double2loint(t0)), m1 = __hiloint2double(h1, __double2loint(t1));
  double s0, s1;
  div_pos2(m0 - 1.0, m0 + 1.0, m1 - 1.0, m1 + 1.0, s0, s1);
  const double a0 = s0 * s0, a1 = s1 * s1, b0 = a0 * a0, b1 = a1 * a1;
  double pe0 = kLogC[9], po0 = kLogC[10], pe1 = kLogC[9], po1 = kLogC[10];
#pragma unroll
  for (int k = 7; k >= 1; k -= 2) {
    const double ce = kLogC[k], co = kLogC[k + 1];
    pe0 = fma(pe0, b0, ce); po0 = fma(po0, b0, co);
    pe1 = fma(pe1, b1, ce); po1 = fma(po1, b1, co);
  }
  const double p0 = fma(po0, a0, pe0), p1 = fma(po1, a1, pe1);
  const double c0 = kLogC[0], l2 = kNum[2];
  const double lm0 = fma(s0 * a0, p0, s0 * c0), lm1 = fma(s1 * a1, p1, s1 * c0);
  o0 = fma((double)e0, l2, lm0);
  o1 = fma((double)e1, l2, lm1);
}

__device__ __forceinline__ void exp10_nb2(double x0, double x1, double& o0, double& o1) {
  const double lh = kNum[0], ll = kNum[1], magic = kNum[10];
  const double t0 = fma(x0, lh, magic), t1 = fma(x1, lh, magic);
  const int n0 = __double2loint(t0), n1 = __double2loint(t1);
  const double nd0 = t0 - magic, nd1 = t1 - magic;
  const double f0 = fma(x0, ll, fma(x0, lh, -nd0)), f1 = fma(x1, ll, fma(x1, lh, -nd1));
  const double g0 = f0 * f0, g1 = f1 * f1;
  double pe0 = kExp2C[12], po0 = kExp2C[13], pe1 = kExp2C[12], po1 = kExp2C[13];
#pragma unroll
  for (int k = 10; k >= 0; k -= 2) {
    const double ce = kExp2C[k], co = kExp2C[k + 1];
    pe0 = fma(pe0, g0, ce); po0 = fma(po0, g0, co);
    pe1 = fma(pe1, g1, ce); po1 = fma(po1, g1, co);
  }
  const double p0 = fma(fma(po0, f0, pe0), f0, 1.0), p1 = fma(fma(po1, f1, pe1), f1, 1.0);
  o0 = p0 * __hiloint2double((n0 + 1023) << 20, 0);
  o1 = p1 * __hiloint2double((n1 + 1023) << 20, 0);
}

// (lon, lat) of a unit vector: atan2(y0, x0) for any signs and atan2(y1, x1) with x1 >= 0
// (a hypotenuse), max(|x|, |y|) > 0 for both (|w| = 1)
__device__ __forceinline__ void atan2_nb2(double y0, double x0, double y1, double x1, double& o0,
                                          double& o1) {
  const double ax0 = fabs(x0), ay0 = fabs(y0), ay1 = fabs(y1);
  const bool st0 = ay0 > ax0, st1 = ay1 > x1;
  const double mx0 = st0 ? ay0 : ax0, mn0 = st0 ? ax0 : ay0;
  const double mx1 = st1 ? ay1 : x1, mn1 = st1 ? x1 : ay1;
  double r0, r1;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(mx0));
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r1) : "d"(mx1));
  const double magic = kNum[10];
  const double t0 = fma(mn0 * r0, 8.0, magic), t1 = fma(mn1 * r1, 8.0, magic);
  const int i0 = __double2loint(t0) & 15, i1 = __double2loint(t1) & 15;
  const double c0 = (t0 - magic) * 0.125, c1 = (t1 - magic) * 0.125;
  double q0, q1;
  div_pos2(fma(-c0, mx0, mn0), fma(c0, mn0, mx0), fma(-c1, mx1, mn1), fma(c1, mn1, mx1), q0, q1);
  const double a0 = q0 * q0, a1 = q1 * q1, b0 = a0 * a0, b1 = a1 * a1;
  const double k6 = kAtanC[6], k4 = kAtanC[4], k5 = kAtanC[5];
  double pe0 = fma(k6, b0, k4), po0 = k5, pe1 = fma(k6, b1, k4), po1 = k5;
#pragma unroll
  for (int k = 2; k >= 0; k -= 2) {
    const double ce = kAtanC[k], co = kAtanC[k + 1];
    pe0 = fma(pe0, b0, ce); po0 = fma(po0, b0, co);
    pe1 = fma(pe1, b1, ce); po1 = fma(po1, b1, co);
  }
  const double p0 = fma(po0, a0, pe0), p1 = fma(po1, a1, pe1);
  double v0 = kAtanT[2 * i0] + (fma(q0 * a0, p0, q0) + kAtanT[2 * i0 + 1]);
  double v1 = kAtanT[2 * i1] + (fma(q1 * a1, p1, q1) + kAtanT[2 * i1 + 1]);
  const double hp = kNum[3], hpl = kNum[4];
  v0 = st0 ? (hp - v0) + hpl : v0;
  v1 = st1 ? (hp - v1) + hpl : v1;
  v0 = (x0 < 0.0) ? (kNum[5] - v0) + kNum[6] : v0;
  o0 = (y0 < 0.0) ? -v0 : v0;
  o1 = (y1 < 0.0) ? -v1 : v1;
}

__device__ SS_COLD double2 sincos_cold(double b) {
  double s, c;
  sincos(b, &s, &c);
  return make_double2(s, c);
}

// sin and cos of a small angle (|b| <= 0.26 rad ~ 15 deg: the transverse stream
// coordinate) by Taylor series to full double precision; larger angles take the
// library path.
__device__ __forceinline__ void sincos_small(double b, double& s, double& c) {
  if (fabs(b) <= 0.26) {
    const double x = b * b;
    double ps = kSinC[7], pc = kCosC[7];
#pragma unroll
    for (int k = 6; k >= 0; --k) {
      ps = fma(ps, x, kSinC[k]);
      pc = fma(pc, x, kCosC[k]);
    }
    s = b * ps;
    c = fma(pc, x, 1.0);
  } else {
    const double2 sc = sincos_cold(b);
    s = sc.x;
    c = sc.y;
  }
}

// Box-Muller in double from the same two words (SSParams.normals_f64): y = (~wa + 0.5) / 2^32,
// radius sqrt(-2 ln y), angle pi * (int32)wb / 2^31 -- bm32f's formulas, every step in double
__device__ __forceinline__ void bm64(uint32_t wa, uint32_t wb, double& z0, double& z1) {
  const double y = (word_to_double(~wa) + 0.5) * (1.0 / 4294967296.0);
  const double r = sqrt_nb(-4.605170185988092 * log10_nb(y));      // -2 ln 10 * log10(y)
  double s, c;
  sincospi((double)(int)wb * 4.656612873077393e-10, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

__device__ __forceinline__ void normal_pair(bool f64, uint32_t wa, uint32_t wb, double& z0, double& z1) {
  if (f64) bm64(wa, wb, z0, z1);
  else bm32(wa, wb, z0, z1);
}

// ---------------------------------------------------------------- table search
// last j in [lo, hi) with xp[j] <= x.  Requires xp[lo] <= x and (hi == n or xp[hi] > x).
__device__ __forceinline__ int search_le(const double* __restrict__ xp, int lo, int hi, double x) {
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (xp[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

// Uniform-bucket guide over a sorted table: bucket g = floor((x - x0) * inv) maps to
// guide[g] = last knot <= the bucket's left edge, so the knot interval of x is found
// with a short forward scan instead of a binary search (the backward step only
// absorbs rounding of the bucket index at bucket edges).  The guide is int32 data
// stored inside the double blob; g_off is in doubles.
struct Guide {
  const int32_t* __restrict__ g;
  int n;            // buckets (0 = no guide: binary search)
  double x0, inv;
};

__device__ __forceinline__ Guide guide_of(const SSFunc& f, const double* __restrict__ tab) {
  Guide G;
  G.g = reinterpret_cast<const int32_t*>(tab + f.g_off);
  G.n = f.g_n;
  G.x0 = f.g_x0;
  G.inv = f.g_inv;
  return G;
}

// Guide entries carry two flag bits (engine.BlobPacker.add_bucket_guide): the index is the
// last knot at or below every x that can land in the bucket; SS_GUIDE_ONE says exactly one
// more knot can lie at or below such an x, SS_GUIDE_MANY says several can.
#define SS_GUIDE_ONE 0x80000000u
#define SS_GUIDE_MANY 0x40000000u
#define SS_GUIDE_INDEX 0x3fffffffu

// last j in [0, n) with xp[j] <= x; requires xp[0] <= x <= xp[n-1]
__device__ __forceinline__ int locate(const double* __restrict__ xp, int n, const Guide& G, double x) {
  if (G.n <= 0) return search_le(xp, 0, n, x);
  int b = (int)((x - G.x0) * G.inv);
  b = max(0, min(b, G.n - 1));
  int j = (int)((unsigned)G.g[b] & SS_GUIDE_INDEX);
  while (j > 0 && xp[j] > x) --j;
  while (j + 1 < n && xp[j + 1] <= x) ++j;
  return j;
}

// np.interp(x, xp, fp, left=fill_lo, right=fill_hi) with host-precomputed slopes
// (numpy/_core/src/multiarray/compiled_base.c arr_interp): NaN x -> NaN, exact knot
// hit -> fp[j], otherwise slope*(x - xp[j]) + fp[j] in un-fused double arithmetic.
// lin_at evaluates at an already located interval (tables sharing their abscissae,
// e.g. the Pal 5 track centre and width, are located once).
__device__ __forceinline__ double lin_at(const double* __restrict__ xp, const double* __restrict__ fp,
                                         int n, int j, double x) {
  const double xj = xp[j], fj = fp[j];
  if (j == n - 1 || xj == x) return fj;
  return __dadd_rn(__dmul_rn(fp[n + j], __dsub_rn(x, xj)), fj);
}

// The guided look-up of a LINEAR table in two halves, so that curves tabulated on the same
// abscissae (the survey's photo-error and efficiency curves) are located once per star.
__device__ __forceinline__ int lin_locate(const SSFunc& f, const double* __restrict__ tab, double x) {
  const double* __restrict__ xp = tab + f.x_off;
  const int n = f.n;
  if (f.g_n <= 0) {                                        // tiny table without a guide
    if (!(x >= xp[0])) return 0;                           // below, or NaN: lin_value fills
    return (x >= xp[n - 1]) ? n - 1 : search_le(xp, 0, n, x);
  }
  // bucket -> knot with at most one compare; no search loop, no loads of xp[0] / xp[n-1]
  int b = (int)((x - f.g_x0) * f.g_inv);                  // NaN -> 0
  b = max(0, min(b, f.g_n - 1));
  const unsigned g = (unsigned)reinterpret_cast<const int32_t*>(tab + f.g_off)[b];
  int j = (int)(g & SS_GUIDE_INDEX);
  if (g & SS_GUIDE_MANY) {
    while (j > 0 && xp[j] > x) --j;
    while (j + 1 < n && xp[j + 1] <= x) ++j;
  } else {
    j += ((g & SS_GUIDE_ONE) && x >= xp[min(j + 1, n - 1)]) ? 1 : 0;
  }
  return j;
}

// value at an already located interval; the range and NaN cases are selects against the
// table ends kept in the descriptor (p[2], p[3])
__device__ __forceinline__ double lin_value(const SSFunc& f, const double* __restrict__ tab, int j,
                                            double x, double fill_lo, double fill_hi) {
  const double* __restrict__ xp = tab + f.x_off;
  const double* __restrict__ fp = tab + f.c_off;
  const int n = f.n;
  const double xj = xp[j], fj = fp[j];
  double r = (j == n - 1 || xj == x) ? fj : __dadd_rn(__dmul_rn(fp[n + j], __dsub_rn(x, xj)), fj);
  r = (x < f.p[2]) ? fill_lo : r;
  r = (x > f.p[3]) ? fill_hi : r;
  return (x != x) ? x : r;
}

__device__ __forceinline__ double lin_eval(const SSFunc& f, const double* __restrict__ tab,
                                           double x, double fill_lo, double fill_hi) {
  const double* __restrict__ xp = tab + f.x_off;
  const double* __restrict__ fp = tab + f.c_off;
  const int n = f.n;
  if (f.g_n > 0) return lin_value(f, tab, lin_locate(f, tab, x), x, fill_lo, fill_hi);
  if (x != x) return x;
  if (x < xp[0]) return fill_lo;
  if (x > xp[n - 1]) return fill_hi;
  return lin_at(xp, fp, n, search_le(xp, 0, n, x), x);
}

// scipy CubicSpline(bc_type='natural', extrapolate=False)(x): PPoly in powers of
// (x - x_j), coefficients highest power first, NaN outside [x_0, x_{n-1}].
__device__ __forceinline__ double spline_eval(const double* __restrict__ xk,
                                              const double* __restrict__ c, int n, double x) {
  if (!(x >= xk[0] && x <= xk[n - 1])) return nan_d();
  int j = search_le(xk, 0, n, x);
  if (j > n - 2) j = n - 2;
  const double s = x - xk[j];
  const double* cj = c + 4 * j;
  return ((cj[0] * s + cj[1]) * s + cj[2]) * s + cj[3];
}

// BoundedFunction.__call__ (reference functions.py:64-76)
__device__ __forceinline__ double fn_eval(const SSFunc& f, const double* __restrict__ tab, double x) {
  double y;
  switch (f.kind) {
    case SS_FN_CONSTANT:
      y = f.p[0];
      break;
    case SS_FN_LINE:
      y = __dadd_rn(__dmul_rn(f.p[0], x), f.p[1]);
      break;
    case SS_FN_SINUSOID: {
      const double norm = f.p[0] / 2.0;
      const double arg = __ddiv_rn(__dmul_rn(6.283185307179586, __dsub_rn(x, f.p[2])), f.p[1]);
      y = __dadd_rn(__dadd_rn(__dmul_rn(norm, cos(arg)), norm), f.p[3]);
      break;
    }
    case SS_FN_LINEAR:
      y = lin_eval(f, tab, x, nan_d(), nan_d());
      break;
    case SS_FN_SPLINE:
      y = spline_eval(tab + f.x_off, tab + f.c_off, f.n, x);
      break;
    case SS_FN_SPLINE_PRODUCT:
      // LinearDensityCubicSplineInterpolation has no bounds mask of its own (functions.py:349-359)
      return 2.5066282746310002 * spline_eval(tab + f.x_off, tab + f.c_off, f.n, x) *
             spline_eval(tab + f.x2_off, tab + f.c2_off, f.n2, x);
    default:
      y = nan_d();
  }
  return (x >= f.xmin && x <= f.xmax) ? y : nan_d();
}

// ---------------------------------------------------------------- HEALPix
// healpy.ang2pix(nside, lon, lat, lonlat=True) = healpix_cxx T_Healpix_Base<int64>::
// ang2pix -> loc2pix.  The angle preparation is shared by every nside.
struct HpLoc {
  double z, tt, sth;
  bool have_sth, valid;
};

__device__ __forceinline__ HpLoc hp_locate(double lon_deg, double lat_deg) {
  const double D2R = 0.017453292519943295;      // NPY_PI/180 (np.radians)
  const double HALFPI = 1.5707963267948966;     // np.pi/2
  const double INV_HALFPI = 0.6366197723675813430755350534900574;
  HpLoc L;
  const double theta = __dsub_rn(HALFPI, __dmul_rn(lat_deg, D2R));
  const double phi = __dmul_rn(lon_deg, D2R);
  // healpy check_theta_valid: 0 <= theta <= pi + 1e-5; NaN fails both
  L.valid = (theta >= 0.0) && (theta <= 3.141592653589793 + 1e-5) && (phi == phi) &&
            (fabs(phi) < 1e300);
  L.have_sth = (theta < 0.01) || (theta > 3.14159 - 0.01);
  L.z = cos(theta);
  L.sth = L.have_sth ? sin(theta) : 0.0;
  const double v = __dmul_rn(phi, INV_HALFPI);
  double tt;
  if (v >= 0.0) {
    tt = (v < 4.0) ? v : fmod(v, 4.0);
  } else {
    tt = fmod(v, 4.0) + 4.0;
    if (tt == 4.0) tt = 0.0;
  }
  L.tt = tt;
  return L;
}

// The fused path: z = cos(colatitude) = w_z and tt = longitude / (pi/2) straight from the
// rotated unit vector (the reference goes through degrees and back; equal to ~1e-16).  The
// callers keep the reference's own chain within 0.8 deg of a pole.
__device__ __forceinline__ HpLoc hp_from_vector(double wz, double lon_rad) {
  HpLoc L;
  L.z = wz;
  const double v = __dmul_rn(lon_rad, kNum[9]);
  L.tt = (v < 0.0) ? v + 4.0 : v;
  if (L.tt >= 4.0) L.tt = 0.0;
  L.sth = 0.0;
  L.have_sth = false;
  L.valid = true;
  return L;
}

__device__ __forceinline__ double hp_polar_tmp(const HpLoc& L, double za, double ns) {
  return (za < 0.99 || !L.have_sth) ? __dmul_rn(ns, sqrt(__dmul_rn(3.0, __dsub_rn(1.0, za))))
                                    : __ddiv_rn(__dmul_rn(ns, L.sth), sqrt(__ddiv_rn(__dadd_rn(1.0, za), 3.0)));
}

__device__ __forceinline__ long long hp_ring(const HpLoc& L, long long nside) {
  const double ns = (double)nside;
  const double za = fabs(L.z);
  if (za <= 2.0 / 3.0) {
    const long long nl4 = 4 * nside;
    const double t1 = __dmul_rn(ns, __dadd_rn(0.5, L.tt));
    const double t2 = __dmul_rn(__dmul_rn(ns, L.z), 0.75);
    const long long jp = (long long)__dsub_rn(t1, t2);
    const long long jm = (long long)__dadd_rn(t1, t2);
    const long long ir = nside + 1 + jp - jm;
    const long long kshift = 1 - (ir & 1);
    const long long t = jp + jm - nside + kshift + 1 + nl4 + nl4;
    const long long ip = ((nside & (nside - 1)) == 0 && nside > 1) ? ((t >> 1) & (nl4 - 1))
                                                                   : ((t >> 1) % nl4);
    return 2 * nside * (nside - 1) + (ir - 1) * nl4 + ip;
  }
  const double tp = __dsub_rn(L.tt, (double)(long long)L.tt);
  const double tmp = hp_polar_tmp(L, za, ns);
  const long long jp = (long long)__dmul_rn(tp, tmp);
  const long long jm = (long long)__dmul_rn(__dsub_rn(1.0, tp), tmp);
  const long long ir = jp + jm + 1;
  const long long ip = (long long)__dmul_rn(L.tt, (double)ir);
  return (L.z > 0) ? 2 * ir * (ir - 1) + ip : 12 * nside * nside - 2 * ir * (ir + 1) + ip;
}

__device__ __forceinline__ unsigned long long hp_spread_bits(unsigned long long v) {
  v = (v | (v << 16)) & 0x0000FFFF0000FFFFULL;
  v = (v | (v << 8)) & 0x00FF00FF00FF00FFULL;
  v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0FULL;
  v = (v | (v << 2)) & 0x3333333333333333ULL;
  v = (v | (v << 1)) & 0x5555555555555555ULL;
  return v;
}

__device__ __forceinline__ long long hp_nest(const HpLoc& L, long long nside) {
  const double ns = (double)nside;
  const double za = fabs(L.z);
  const int order = 63 - __clzll(nside);
  long long ix, iy, face;
  if (za <= 2.0 / 3.0) {
    const double t1 = __dmul_rn(ns, __dadd_rn(0.5, L.tt));
    const double t2 = __dmul_rn(__dmul_rn(ns, L.z), 0.75);
    const long long jp = (long long)__dsub_rn(t1, t2);
    const long long jm = (long long)__dadd_rn(t1, t2);
    const long long ifp = jp >> order, ifm = jm >> order;
    face = (ifp == ifm) ? (ifp | 4) : ((ifp < ifm) ? ifp : (ifm + 8));
    ix = jm & (nside - 1);
    iy = nside - (jp & (nside - 1)) - 1;
  } else {
    long long ntt = (long long)L.tt;
    if (ntt > 3) ntt = 3;
    const double tp = __dsub_rn(L.tt, (double)ntt);
    const double tmp = hp_polar_tmp(L, za, ns);
    long long jp = (long long)__dmul_rn(tp, tmp);
    long long jm = (long long)__dmul_rn(__dsub_rn(1.0, tp), tmp);
    if (jp > nside - 1) jp = nside - 1;
    if (jm > nside - 1) jm = nside - 1;
    if (L.z >= 0) { ix = nside - jm - 1; iy = nside - jp - 1; face = ntt; }
    else { ix = jp; iy = jm; face = ntt + 8; }
  }
  return (face << (2 * order)) + (long long)(hp_spread_bits((unsigned long long)ix) |
                                             (hp_spread_bits((unsigned long long)iy) << 1));
}

__device__ __forceinline__ long long hp_pix(const HpLoc& L, long long nside, int nest) {
  return nest ? hp_nest(L, nside) : hp_ring(L, nside);
}

// ---------------------------------------------------------------- survey getters
// log10 of the statistical photometric error, Survey.get_photo_error (reference
// surveys.py:270-286); `bright` is set for saturated stars, whose error is the
// host-evaluated constant 10**f(dsat-1).
__device__ __forceinline__ double photo_error_log(const SSParams& p, const double* __restrict__ tab,
                                                  int band, double mag, double maglim, bool& bright) {
  const double delta = __dsub_rn(mag, maglim);
  const double sat = p.saturation[band];
  bright = mag < sat;
  if (bright) return 0.0;
  return (delta <= p.delta_saturation && mag >= sat) ? p.log_err_at_dsat
                                                     : lin_eval(p.photo_error, tab, delta, p.photo_error.p[0], p.photo_error.p[0]);
}

// statistical (+) systematic error in quadrature (reference surveys.py:288-290)
__device__ __forceinline__ double photo_error_total(const SSParams& p, int band, double loge,
                                                    bool bright) {
  const double stat = bright ? p.err_bright : exp10_fast(loge);
  const double sys = p.sys_error[band];
  return sqrt(__dadd_rn(__dmul_rn(stat, stat), __dmul_rn(sys, sys)));
}

__device__ __forceinline__ double photo_error(const SSParams& p, const double* __restrict__ tab,
                                              int band, double mag, double maglim) {
  bool bright;
  const double loge = photo_error_log(p, tab, band, mag, maglim, bright);
  return photo_error_total(p, band, loge, bright);
}

// Survey.get_efficiency (reference surveys.py:473-493)
__device__ __forceinline__ double efficiency(const SSParams& p, const SSFunc& f,
                                             const double* __restrict__ tab, int band,
                                             double mag, double maglim) {
  const double sat = p.saturation[band];
  if (maglim < sat || maglim != maglim) return 0.0;
  if (mag < sat) return 0.0;
  const double delta = __dsub_rn(mag, maglim);
  if (mag > sat && delta <= p.delta_saturation) return 1.0;
  return lin_eval(f, tab, delta, f.p[0], f.p[0]);
}

// the same two getters at an already located table interval, branch-free (the fused
// observe step; guided tables only -- validate() insists on a guide for the survey curves)
__device__ __forceinline__ double photo_error_log_at(const SSParams& p, const double* __restrict__ tab,
                                                     double sat, double mag, double delta, int j,
                                                     bool& bright) {
  bright = mag < sat;
  const double v = lin_value(p.photo_error, tab, j, delta, p.photo_error.p[0], p.photo_error.p[0]);
  const double w = (delta <= p.delta_saturation && mag >= sat) ? p.log_err_at_dsat : v;
  return bright ? 0.0 : w;
}

__device__ __forceinline__ double efficiency_at(const SSParams& p, const SSFunc& f,
                                                const double* __restrict__ tab, double sat,
                                                double mag, double maglim, double delta, int j) {
  double r = lin_value(f, tab, j, delta, f.p[0], f.p[0]);
  r = (mag > sat && delta <= p.delta_saturation) ? 1.0 : r;
  r = (mag < sat) ? 0.0 : r;
  return (maglim < sat || maglim != maglim) ? 0.0 : r;
}

// the same two getters on the direct look-up form of the curves (SSFunc.q_*): the framed
// segment records answer every range / NaN case by themselves
__device__ __forceinline__ int q_locate(const SSFunc& f, const double* __restrict__ tab, double x) {
  int b = __double2int_rd(__dmul_rn(__dsub_rn(x, f.q_x0), f.q_inv)) + 1;     // NaN -> 1
  b = max(0, min(b, f.q_n + 1));
  const double2* __restrict__ rec = reinterpret_cast<const double2*>(tab + f.q_off) + 2 * b;
  const double2 r0 = rec[0];
  const int j = __double2loint(rec[1].x);
  return j + (x >= r0.x ? 1 : 0) + (x >= r0.y ? 1 : 0);
}

__device__ __forceinline__ double q_value(const SSFunc& f, const double* __restrict__ tab, int j,
                                          double x) {
  const double2* __restrict__ seg = reinterpret_cast<const double2*>(tab + f.q_seg_off) + 2 * j;
  const double2 s0 = seg[0];
  const double slope = seg[1].x;
  return __dadd_rn(__dmul_rn(slope, __dsub_rn(x, s0.x)), s0.y);
}

__device__ __forceinline__ double photo_error_log_q(const SSParams& p, const double* __restrict__ tab,
                                                    double sat, double mag, double delta, int j,
                                                    bool& bright) {
  bright = mag < sat;
  const double v = q_value(p.photo_error, tab, j, delta);
  const double w = (delta <= p.delta_saturation && mag >= sat) ? p.log_err_at_dsat : v;
  return bright ? 0.0 : w;
}

__device__ __forceinline__ double efficiency_q(const SSParams& p, const SSFunc& f,
                                               const double* __restrict__ tab, double sat,
                                               double mag, double maglim, double delta, int j) {
  double r = q_value(f, tab, j, delta);
  r = (mag > sat && delta <= p.delta_saturation) ? 1.0 : r;
  r = (mag < sat) ? 0.0 : r;
  return (maglim < sat || maglim != maglim) ? 0.0 : r;
}

// ---------------------------------------------------------------- per-star observe arithmetic
// Everything of StreamInjector.inject's per-star body that follows the map gathers
// (reference observed.py:163-291), shared by the generic kernel and the production
// kernel so that both produce the same bytes.
struct ObsOut {
  double mobs[2], err[2];
  bool observed, perfect, unseen, bad[2];
};

// sample_measured_magnitudes in the short form
//   F = 3631*10^(-0.4 m), F_obs = F + (F*err/1.0857362)*z, m_obs = -2.5 log10(F_obs/3631)
//   <=>  F_obs > 0 <=> t > 0,  m_obs = m - 2.5 log10(t),  t = 1 + (err/1.0857362)*z
// (one log10 instead of exp10 + log10 + two divisions; equal to rounding while the flux
// neither under- nor overflows), all in double.  A bad magnitude (F_obs <= 0 or NaN: the
// reference's "BAD_MAG") is returned as NaN -- a good one is never NaN.
__device__ __forceinline__ void noisy_mag_f64(const SSParams& p, int b, double mt, double loge,
                                              bool bright, double z, double& err, double& mobs) {
  const double stat = bright ? p.err_bright : exp10_nb(loge);
  const double sys = p.sys_error[b];                                // surveys.py:288-290
  err = sqrt_nb(__dadd_rn(__dmul_rn(stat, stat), __dmul_rn(sys, sys)));
  const double t = 1.0 + err * kNum[11] * z;
  mobs = (t > 0.0) ? mt - 2.5 * log10_nb(t) : nan_d();
}

// both bands at once (bit-identical to two noisy_mag_f64 calls)
__device__ __forceinline__ void noisy_mag_f64x2(const SSParams& p, const double* mt, const double* loge,
                                                const bool* bright, const double* z, double* err,
                                                double* mobs) {
  double s0, s1;
  exp10_nb2(loge[0], loge[1], s0, s1);
  const double eb = p.err_bright;
  s0 = bright[0] ? eb : s0;
  s1 = bright[1] ? eb : s1;
  const double y0 = p.sys_error[0], y1 = p.sys_error[1];
  err[0] = sqrt_nb(__dadd_rn(__dmul_rn(s0, s0), __dmul_rn(y0, y0)));
  err[1] = sqrt_nb(__dadd_rn(__dmul_rn(s1, s1), __dmul_rn(y1, y1)));
  const double kf = kNum[11];
  const double t0 = 1.0 + err[0] * kf * z[0], t1 = 1.0 + err[1] * kf * z[1];
  double l0, l1;
  log10_nb2(t0, t1, l0, l1);
  mobs[0] = (t0 > 0.0) ? mt[0] - 2.5 * l0 : nan_d();
  mobs[1] = (t1 > 0.0) ? mt[1] - 2.5 * l1 : nan_d();
}

// exact flux-space noise for magnitudes whose flux under-/overflows the short form, or
// log-errors outside exp10_nb's domain (observed.py:863-904, :984-1035); rare, out of line.
// Returns (err, mobs).
__device__ SS_COLD double2 noisy_mag_exact(const SSParams& p, int b, double mt, double loge,
                                                bool bright, double z) {
  const double err = photo_error_total(p, b, loge, bright);
  double mobs = nan_d();
  if (fabs(mt) < 200.0) {
    const double t = 1.0 + err * kNum[11] * z;
    if (t > 0.0) mobs = mt - 2.5 * log10(t);
  } else {
    const double flux = __dmul_rn(3631.0, exp10(__dmul_rn(-0.4, mt)));
    const double sig = __ddiv_rn(__dmul_rn(flux, err), 1.0857362);
    const double fobs = __dadd_rn(flux, __dmul_rn(sig, z));
    if (fobs > 0.0) mobs = __dmul_rn(-2.5, log10(__ddiv_rn(fobs, 3631.0)));
  }
  return make_double2(err, mobs);
}

// the double short form, out of line: SS_MATH_FAST's redo inside its guard band
__device__ __noinline__ double2 noisy_mag_redo(const SSParams& p, int b, double mt, double loge,
                                               bool bright, double z) {
  double err, mobs;
  noisy_mag_f64(p, b, mt, loge, bright, z, err, mobs);
  return make_double2(err, mobs);
}

// share: bit 0 = the completeness curve sits on the photo-error curve's abscissae, bit 1 = so
// does the detection-only curve (same blob offsets: BlobPacker de-duplicates): located once
__device__ __forceinline__ int curve_share(const SSParams& p) {
  const SSFunc& e = p.photo_error;
  const SSFunc& c = p.completeness;
  const SSFunc& d = p.detection;
  const int sc = (c.n == e.n && c.x_off == e.x_off && c.g_n == e.g_n && c.g_off == e.g_off &&
                  c.q_n == e.q_n && c.q_off == e.q_off) ? 1 : 0;
  const int sd = (d.n == e.n && d.x_off == e.x_off && d.g_n == e.g_n && d.g_off == e.g_off &&
                  d.q_n == e.q_n && d.q_off == e.q_off) ? 2 : 0;
  return sc | sd;
}

// QCURVE: the survey curves are read through their direct look-up form (the production
// kernel; the launcher checks that every curve has one), else through the guided search.
template <int MATH, bool QCURVE = false>
__device__ __forceinline__ void observe_star(const SSParams& p, const double* __restrict__ tab,
                                             int share, bool band_g, bool band_r, double mag_g,
                                             double mag_r, double ebv, double ml_g, double ml_r,
                                             double z_g, double z_r, double u_det, double u_det2,
                                             ObsOut& o) {
  // Stage 1: per band, the table look-ups.  Stage 2: the error / noise arithmetic of BOTH
  // bands as one straight-line block, so that the two dependency chains interleave.
  double ext2[2], mt2[2], loge2[2], zf2[2], dl2[2], ml2[2];
  int j2[2];
  bool bright2[2], have2[2], good2[2];
  bool plain = true;
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    have2[b] = (b == SS_BAND_G) ? band_g : band_r;
    const double mag = b == SS_BAND_G ? mag_g : mag_r;
    ml2[b] = b == SS_BAND_G ? ml_g : ml_r;
    zf2[b] = b == SS_BAND_G ? z_g : z_r;
    ext2[b] = __dmul_rn(p.coeff_extinc[b], ebv);                   // surveys.py:228
    mt2[b] = __dadd_rn(mag, ext2[b]);                              // observed.py:167
    dl2[b] = __dsub_rn(mt2[b], ml2[b]);
    bright2[b] = false;
    good2[b] = false;
    loge2[b] = nan_d();
    j2[b] = 0;
    o.err[b] = o.mobs[b] = nan_d();
    if (!have2[b]) continue;
    if (QCURVE) {
      j2[b] = q_locate(p.photo_error, tab, dl2[b]);
      loge2[b] = photo_error_log_q(p, tab, p.saturation[b], mt2[b], dl2[b], j2[b], bright2[b]);
    } else {
      j2[b] = lin_locate(p.photo_error, tab, dl2[b]);
      loge2[b] = photo_error_log_at(p, tab, p.saturation[b], mt2[b], dl2[b], j2[b], bright2[b]);  // surveys.py:234-286
    }
    plain = plain && fabs(mt2[b]) < 200.0 && !(fabs(loge2[b]) >= 300.0);
  }
  // completeness / detection Bernoulli trials in the survey's completeness band (observed.py:224-243)
  const bool cr = p.completeness_band == SS_BAND_R;
  const double mt_c = cr ? mt2[1] : mt2[0], ml_c = cr ? ml2[1] : ml2[0], dl_c = cr ? dl2[1] : dl2[0];
  const double sat_c = cr ? p.saturation[1] : p.saturation[0];
  const int j_p = cr ? j2[1] : j2[0];
  bool flag_c, flag_d = false;
  if (QCURVE) {
    const int j_c = (share & 1) ? j_p : q_locate(p.completeness, tab, dl_c);
    flag_c = u_det <= efficiency_q(p, p.completeness, tab, sat_c, mt_c, ml_c, dl_c, j_c);
    if (p.perfect_galstarsep) {
      const int j_d = (share & 2) ? j_p : q_locate(p.detection, tab, dl_c);
      flag_d = u_det2 <= efficiency_q(p, p.detection, tab, sat_c, mt_c, ml_c, dl_c, j_d);
    }
  } else {
    const int j_c = (share & 1) ? j_p : lin_locate(p.completeness, tab, dl_c);
    flag_c = u_det <= efficiency_at(p, p.completeness, tab, sat_c, mt_c, ml_c, dl_c, j_c);
    if (p.perfect_galstarsep) {
      const int j_d = (share & 2) ? j_p : lin_locate(p.detection, tab, dl_c);
      flag_d = u_det2 <= efficiency_at(p, p.detection, tab, sat_c, mt_c, ml_c, dl_c, j_d);
    }
  }
  o.unseen = ml_c != ml_c;
  if (plain && MATH == SS_MATH_F64 && band_g && band_r) {
    noisy_mag_f64x2(p, mt2, loge2, bright2, zf2, o.err, o.mobs);
    good2[0] = o.mobs[0] == o.mobs[0];
    good2[1] = o.mobs[1] == o.mobs[1];
  } else if (plain) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      if (MATH == SS_MATH_FAST) {
        // fp32 values on the MUFU units.  stat carries ~1e-6 relative error, so does
        // x = err*z/1.0857362; the log amplifies the error of t = 1 + x by 1/t: inside the
        // guard band the star is redone in double (it also owns the sign decision there).
        const float lf = (float)loge2[b];
        const float stat = bright2[b] ? (float)p.err_bright : exp2f(lf * 3.3219281f);
        const float sys = (float)p.sys_error[b];
        const float ef = sqrtf(fmaf(stat, stat, sys * sys));
        const float x = ef * 0.92103404f * (float)zf2[b];
        const float t = 1.0f + x;
        if (fabsf(t) < 0.012f * (1.0f + fabsf(x)) || fabsf(lf) >= 30.0f) {   // NaN: not taken
          const double2 r = noisy_mag_redo(p, b, mt2[b], loge2[b], bright2[b], zf2[b]);
          o.err[b] = r.x;
          o.mobs[b] = r.y;
        } else {
          o.err[b] = (double)ef;
          o.mobs[b] = (t > 0.0f) ? mt2[b] - (double)(0.75257499f * __log2f(t)) : nan_d();
        }
      } else {
        noisy_mag_f64(p, b, mt2[b], loge2[b], bright2[b], zf2[b], o.err[b], o.mobs[b]);
      }
      good2[b] = o.mobs[b] == o.mobs[b];
    }
  } else {
#pragma unroll
    for (int b = 0; b < 2; ++b)
      if (have2[b]) {
        const double2 r = noisy_mag_exact(p, b, mt2[b], loge2[b], bright2[b], zf2[b]);
        o.err[b] = r.x;
        o.mobs[b] = r.y;
        good2[b] = r.y == r.y;
      }
  }
  bool valid = true, snr_ok = true, snr_r_ok = true;
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    o.bad[b] = false;
    if (!have2[b]) {
      o.err[b] = o.mobs[b] = nan_d();
      continue;
    }
    if (good2[b] && p.dust_correction) o.mobs[b] = __dsub_rn(o.mobs[b], ext2[b]);  // observed.py:194-208
    valid = valid && good2[b];
    o.bad[b] = !good2[b];
    // 1/magerr >= 5 (observed.py:266-281) decided on the monotone pre-image log10(err_stat)
    const bool ok = bright2[b] ? (p.snr_bright_ok[b] != 0) : (loge2[b] <= p.snr_loge_max[b]);
    if (p.snr_cut[b]) snr_ok = snr_ok && ok;
    if (b == SS_BAND_R) snr_r_ok = ok;                             // observed.py:283-286
  }
  o.observed = valid && flag_c && snr_ok;
  o.perfect = p.perfect_galstarsep && valid && flag_d && snr_ok && snr_r_ok;
}

// histogram bin in fp32 (SSParams.hist_*): t = ((float)v - lo) * inv, two roundings; -1 when
// outside or NaN
__device__ __forceinline__ int hist_bin(float v, float lo, float inv) {
  const float t = __fmul_rn(__fsub_rn(v, lo), inv);
  return (t >= 0.0f && t < (float)SS_HIST_BINS) ? (int)t : -1;
}

// ---------------------------------------------------------------- HEALPix, 32-bit variant
// ang2pix RING from (z, tt) for nside <= 8192 (npix < 2^31): the same floating-point
// expressions as hp_ring, integer part in 32 bits.
__device__ __forceinline__ int hp_ring32(double z, double tt, int nside) {
  const double ns = (double)nside;
  const double za = fabs(z);
  if (za <= 2.0 / 3.0) {
    const int nl4 = 4 * nside;
    const double t1 = __dmul_rn(ns, __dadd_rn(0.5, tt));
    const double t2 = __dmul_rn(__dmul_rn(ns, z), 0.75);
    const int jp = __double2int_rz(__dsub_rn(t1, t2));
    const int jm = __double2int_rz(__dadd_rn(t1, t2));
    const int ir = nside + 1 + jp - jm;
    const int kshift = 1 - (ir & 1);
    const int t = jp + jm - nside + kshift + 1 + nl4 + nl4;
    const int ip = ((nside & (nside - 1)) == 0 && nside > 1) ? ((t >> 1) & (nl4 - 1))
                                                             : ((t >> 1) % nl4);
    return 2 * nside * (nside - 1) + (ir - 1) * nl4 + ip;
  }
  const double tp = __dsub_rn(tt, (double)__double2int_rz(tt));
  const double tmp = __dmul_rn(ns, sqrt(__dmul_rn(3.0, __dsub_rn(1.0, za))));
  const int jp = __double2int_rz(__dmul_rn(tp, tmp));
  const int jm = __double2int_rz(__dmul_rn(__dsub_rn(1.0, tp), tmp));
  const int ir = jp + jm + 1;
  const int ip = __double2int_rz(__dmul_rn(tt, (double)ir));
  return (z > 0) ? 2 * ir * (ir - 1) + ip : 12 * nside * nside - 2 * ir * (ir + 1) + ip;
}

}  // namespace ss
