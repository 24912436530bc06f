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

// one 32-bit word -> double in [0,1), exact
__device__ __forceinline__ double u32(uint32_t w) { return (double)w * (1.0 / 4294967296.0); }

// Box-Muller from two 32-bit words, evaluated in fp32 (like cuRAND's curand_normal): the
// draw is a random number, not a quantity with a parity requirement, and everything
// computed *from* it downstream stays in double.  y = (wa+0.5)/2^32 in (0,1) feeds the
// radius; for y close to 1 the complement (~wa+0.5)/2^32 = 1-y is exact and a short
// series replaces log(y), which would lose relative accuracy there.  |z| <= 6.7.
__device__ __forceinline__ void bm32(uint32_t wa, uint32_t wb, double& z0, double& z1) {
  const float y = ((float)wa + 0.5f) * 2.3283064365386963e-10f;
  const float c1 = ((float)(~wa) + 0.5f) * 2.3283064365386963e-10f;
  const float l_series = -c1 * (1.0f + c1 * (0.5f + c1 * (0.33333334f + c1 * 0.25f)));
  const float l = (c1 < 0.03f) ? l_series : __logf(y);
  const float r = sqrtf(-2.0f * l);
  float s, c;
  sincospif((float)wb * 4.6566128730773926e-10f, &s, &c);   // angle/pi = 2*wb/2^32
  z0 = (double)(r * c);
  z1 = (double)(r * s);
}

// sin and cos of a small angle (|b| <= 0.26 rad ~ 15 deg: the transverse stream
// coordinate) by Taylor series to full double precision -- ~20 FMAs instead of a
// general-range sincos; larger angles take the library path.
__device__ __forceinline__ void sincos_small(double b, double& s, double& c) {
  if (fabs(b) <= 0.26) {
    const double x = b * b;
    s = b * (1.0 + x * (-1.0 / 6 + x * (1.0 / 120 + x * (-1.0 / 5040 + x * (1.0 / 362880 +
        x * (-1.0 / 39916800 + x * (1.0 / 6227020800.0 + x * (-1.0 / 1307674368000.0))))))));
    c = 1.0 + x * (-0.5 + x * (1.0 / 24 + x * (-1.0 / 720 + x * (1.0 / 40320 +
        x * (-1.0 / 3628800 + x * (1.0 / 479001600 + x * (-1.0 / 87178291200.0 +
        x * (1.0 / 20922789888000.0))))))));
  } else {
    sincos(b, &s, &c);
  }
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

// last j in [0, n) with xp[j] <= x; requires xp[0] <= x <= xp[n-1]
__device__ __forceinline__ int locate(const double* __restrict__ xp, int n, const Guide& G, double x) {
  if (G.n <= 0) return search_le(xp, 0, n, x);
  int b = (int)((x - G.x0) * G.inv);
  b = max(0, min(b, G.n - 1));
  int j = G.g[b];
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

__device__ __forceinline__ double lin_eval(const SSFunc& f, const double* __restrict__ tab,
                                           double x, double fill_lo, double fill_hi) {
  const double* __restrict__ xp = tab + f.x_off;
  const int n = f.n;
  if (x != x) return x;
  if (x < xp[0]) return fill_lo;
  if (x > xp[n - 1]) return fill_hi;
  return lin_at(xp, tab + f.c_off, n, locate(xp, n, guide_of(f, tab), x), x);
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
                                                     : lin_eval(p.photo_error, tab, delta, 1.0, 1.0);
}

// statistical (+) systematic error in quadrature (reference surveys.py:288-290)
__device__ __forceinline__ double photo_error_total(const SSParams& p, int band, double loge,
                                                    bool bright) {
  const double stat = bright ? p.err_bright : exp10(loge);
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
  return lin_eval(f, tab, delta, 0.0, 0.0);
}

}  // namespace ss
