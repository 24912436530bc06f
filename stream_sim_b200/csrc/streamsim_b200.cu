// libstreamsim_b200.so -- fused per-star generate -> rotate -> observe kernel for
// NVIDIA B200 (sm_100a) and its C ABI (include/streamsim_b200.h).
//
// One persistent CTA per SM.  The small lookup tables (track / distance-modulus
// knots, isochrone, IMF CDF, efficiency and photo-error curves, search guides)
// are packed by the host into one blob that each CTA pulls into shared memory
// with a single TMA bulk copy (cp.async.bulk + mbarrier); the 100000-point
// phi1 CDF and the HEALPix maps stay in HBM/L2 and are read through the
// read-only path.  Every star is independent: thread t handles stars
// t, t+T, t+2T, ... so that each output column is written as fully coalesced
// 128-byte lines.  Randomness is counter-based (Philox4x32-10 keyed by the
// global star index) so results do not depend on the launch geometry or on how
// the index range is sharded across GPUs.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <utility>

#include "device_math.cuh"

#ifndef SS_BLOCK
#define SS_BLOCK 1024
#endif
// geometry of the production kernel k_fused: threads per CTA and resident CTAs per SM
// (the register budget follows: 65536 / (block * ctas) per thread)
#ifndef SS_FUSED_BLOCK
#define SS_FUSED_BLOCK 1024
#endif
#ifndef SS_FUSED_MIN_CTAS
#define SS_FUSED_MIN_CTAS 1
#endif
// 1: 32-bit integer arithmetic for the RING index of maps up to nside 8192
#ifndef SS_HP32
#define SS_HP32 1
#endif

// 1: k_fused writes pairs of fp64 columns with 128-bit stores (neighbouring lanes exchange one
// value each).  A measured build option: STG.E.128 in the SASS (profiles/r02_vecstores_sass_
// excerpt.txt), parity subset green, 2 % slower than the scalar stores in the round-2 interim
// kernel and 8 % in the final one (profiles/r02_build_variants.json) -- the bytes per L1 pass
// are the same and the exchange costs issue slots.  Off.
#ifndef SS_VEC_STORES
#define SS_VEC_STORES 0
#endif
// (cache hints on the column stores -- st.global.cs / .wt / .cg -- change nothing: measured)

// what-if builds (development only; results are wrong on purpose): leave one component of
// k_fused out to measure what it costs.  1 = no map gather, 2 = no column stores, 4 = no
// histogram / counter updates, 8 = no observe arithmetic, 16 = no rotation, 32 = no row loads,
// 64 = no bucket entries (steps straight from the words) and half a grid row
#ifndef SS_WHATIF
#define SS_WHATIF 0
#endif

namespace {

// k_fused's dynamic shared memory behind the blob: three 16-byte staging slots per thread
constexpr int kFusedStage = 3 * 16 * SS_FUSED_BLOCK;

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return code;
}

#define SS_CUDA(call)                                                        \
  do {                                                                       \
    cudaError_t e_ = (call);                                                 \
    if (e_ != cudaSuccess) return fail(SS_ECUDA, #call ": %s", cudaGetErrorString(e_)); \
  } while (0)

struct KArgs {
  SSParams p;
  SSTables t;
  SSMaps m;
  SSCatalog in;
  SSDraws d;
  SSOut o;
  unsigned long long n, offset, seed;
  int blob_in_smem;
  uint32_t blob_bytes;
  // derived launch constants (filled by ss_run): the Philox round keys of `seed` and the
  // histogram origins / inverse bin widths rounded to fp32
  ss::PhiloxKeys keys;
  float hist_lo[SS_N_HIST], hist_inv[SS_N_HIST];
  int vec_ok;   // every floating-point column is 16-byte aligned (SS_VEC_STORES)
  int grid_wshift, iso_wshift;   // bucket of a word table = w >> shift
};

// ---------------------------------------------------------------- smem staging
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// Pull `bytes` (multiple of 16) from global into shared memory with the TMA bulk
// engine; every thread of the CTA returns once the bytes have landed.
__device__ __forceinline__ void stage_blob(void* dst, const void* src, uint32_t bytes,
                                           unsigned long long* bar) {
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
    // the bulk engine takes at most 2^20-1 bytes of tx count per barrier phase; the
    // blob is < 227 KB.  Issue in <= 64 KB pieces.
    uint32_t done = 0;
    while (done < bytes) {
      const uint32_t piece = min(bytes - done, 65536u);
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
          ::"r"(smem_u32((const char*)dst + done)), "l"((const char*)src + done), "r"(piece),
          "r"(smem_u32(bar))
          : "memory");
      done += piece;
    }
  }
  // bounded wait: a lost transaction traps instead of hanging the GPU
  uint32_t ok = 0;
  for (int spin = 0; spin < (1 << 22) && !ok; ++spin) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
        " selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar))
        : "memory");
  }
  if (!ok) __trap();
}

// ---------------------------------------------------------------- helpers
template <typename T>
__device__ __forceinline__ void put(void* col, unsigned long long i, double v) {
  if (col) reinterpret_cast<T*>(col)[i] = (T)v;
}

template <typename T, bool SURE>
__device__ __forceinline__ void putc(void* col, unsigned long long i, double v) {
#if SS_WHATIF & 2
  if (v == 1.2345e-300) reinterpret_cast<T*>(col)[i] = (T)v;
#else
  if (SURE || col) reinterpret_cast<T*>(col)[i] = (T)v;
#endif
}

// Two columns, one 16-byte store per lane: neighbouring lanes swap one value each (even lanes
// end up with column A of both stars, odd lanes with column B), so a warp writes its 2 x 256
// bytes as 32 STG.E.128 instead of 64 STG.E.64.  `full` = the whole warp is inside the range
// (warp-uniform); the ragged last warp and fp32 columns take the scalar stores.
template <typename OutT>
__device__ __forceinline__ void put2(void* ca, void* cb, unsigned long long i, double va, double vb,
                                     bool full) {
  if (sizeof(OutT) == 8 && full && ca && cb) {
    const bool odd = (threadIdx.x & 1) != 0;
    const double recv = __shfl_xor_sync(0xffffffffu, odd ? va : vb, 1);
    const double2 v = odd ? make_double2(recv, vb) : make_double2(va, recv);
    double* const col = reinterpret_cast<double*>(odd ? cb : ca);
    *reinterpret_cast<double2*>(col + (i & ~1ULL)) = v;
  } else {
    put<OutT>(ca, i, va);
    put<OutT>(cb, i, vb);
  }
}

// centre and spread of a TrackModel at x (reference model.py:517-518)
__device__ __forceinline__ void track_eval(const SSFunc& fc, const SSFunc& fs, const double* tab,
                                           double x, double& c, double& s) {
  if (fc.kind == SS_FN_LINEAR && fs.kind == SS_FN_LINEAR && fc.x_off == fs.x_off &&
      fc.n == fs.n) {
    // centre and width tabulated on the same abscissae (e.g. Pal 5): locate once
    const double* __restrict__ xp = tab + fc.x_off;
    const int n = fc.n;
    if (x >= xp[0] && x <= xp[n - 1]) {
      const int j = ss::locate(xp, n, ss::guide_of(fc, tab), x);
      c = ss::lin_at(xp, tab + fc.c_off, n, j, x);
      s = ss::lin_at(xp, tab + fs.c_off, n, j, x);
    } else {
      c = s = ss::nan_d();
    }
    if (!(x >= fc.xmin && x <= fc.xmax)) c = ss::nan_d();
    if (!(x >= fs.xmin && x <= fs.xmax)) s = ss::nan_d();
  } else {
    c = ss::fn_eval(fc, tab, x);
    s = ss::fn_eval(fs, tab, x);
  }
}

// TrackModel.sample (reference model.py:513-544): value = draw*scale + loc
__device__ __forceinline__ double track_draw(double c, double s, int sampler, double draw,
                                             unsigned& domain_err) {
  if (sampler == SS_SAMPLER_GAUSSIAN) {
    if (!(s >= 0.0)) domain_err++;
    return __dadd_rn(__dmul_rn(draw, s), c);
  }
  const double lo = __dsub_rn(c, s);
  const double hi = __dadd_rn(lo, __dmul_rn(2.0, s));
  const double scale = __dsub_rn(hi, lo);
  if (!(scale >= 0.0)) domain_err++;
  return __dadd_rn(__dmul_rn(draw, scale), lo);
}

// One row per point of the phi1 grid (SS_DENSITY_GRID): everything the generate and
// rotate stages need that depends on phi1 only, evaluated once on the host with the
// reference's own NumPy/SciPy calls.  64 bytes = two L2 sectors per star.
struct GridRow {
  double thr;      // smallest u whose sample index exceeds this row's index
  double x;        // phi1 grid value
  double c, s;     // track centre, spread at x
  double dc, ds;   // distance-modulus centre, spread at x
  double sinx, cosx;  // sin/cos of radians(x)
};

// 32 bytes per lane in ONE load instruction (LDG.E.ENL2.256, read-only path): the scattered
// table records cost the L1 data pipe one pass per instruction and lane, whatever the width --
// the kernel's busiest unit -- so a 64-byte row is fetched as two of these instead of four
// 16-byte loads.  `p` must be 32-byte aligned.
struct __align__(32) Double4 {
  double a, b, c, d;
};

__device__ __forceinline__ Double4 ldg256(const void* p) {
  Double4 r;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
      : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d)
      : "l"(p));
  return r;
}

// thresholds thr[lo .. lo+k-1] lie in one bucket of the u -> step table: guess the step by
// position inside the bucket, fetch BOTH neighbouring thresholds of the guess at once (one
// round trip when the guess is right), then walk if it was not
__device__ __forceinline__ int step_scan(const double* __restrict__ thr, int lo, int k, double u,
                                         uint32_t lut_n) {
  const double fb = u * (double)lut_n;
  int j = lo + (int)((fb - (double)(int)fb) * (double)k);   // candidate: first idx with u < thr[idx]
  const int hi = lo + k;
  const double below = __ldg(&thr[max(j - 1, lo)]), above = __ldg(&thr[min(j, hi - 1)]);
  if (j > lo && u < below) {
    --j;
    while (j > lo && u < __ldg(&thr[j - 1])) --j;
  } else if (j < hi && u >= above) {
    ++j;
    while (j < hi && u >= __ldg(&thr[j])) ++j;
  }
  return j;
}

// One 16-byte entry per bucket of u (SSTables.grid_lut / iso_lut): the step at the bucket's
// left edge, the number k of thresholds inside the bucket and the first of them.  For k <= 1
// (82 % of the stars at 2^17 buckets of the phi1 table) the step costs ONE round trip and no
// loop; fuller buckets finish with the scan.  (Measured: 2^17 buckets beat 2^15 and 2^20;
// 32-byte entries holding three thresholds cost more in registers than they save in scans.)
__device__ __forceinline__ int step_from_entry(const double2& e, const double* __restrict__ thr,
                                               double u, uint32_t lut_n) {
  const long long packed = __double_as_longlong(e.x);
  const int lo = (int)(packed & 0xffffffffLL), k = (int)(packed >> 32);
  int idx = lo + ((k >= 1 && u >= e.y) ? 1 : 0);
  if (k > 1 && u >= e.y) idx = step_scan(thr, lo, k, u, lut_n);
  return idx;
}

// The same look-up for a 32-bit draw, in integers (SSTables.grid_wlut / iso_wlut): the entry
// holds the step at the bucket's left edge and the first three thresholds inside it, as words;
// fuller buckets (0.1 % of the draws at 2^18 buckets of the Pal 5 table) finish with a binary
// search on the word thresholds.
__device__ __forceinline__ int step_from_words(const uint4& e, const uint32_t* __restrict__ wthr,
                                               uint32_t w, int nthr) {
  const int lo = (int)(e.x & 0xFFFFFu);
  int idx = lo + (w > e.y ? 1 : 0) + (w > e.z ? 1 : 0) + (w > e.w ? 1 : 0);
  const int k = (int)(e.x >> 20);
  if (k > 3 && w > e.w) {
    int a = lo + 3, b = (k == 4095) ? nthr : lo + k;
    while (a < b) {
      const int mid = (a + b) >> 1;
      if (w > __ldg(&wthr[mid])) a = mid + 1; else b = mid;
    }
    idx = a;
  }
  return idx;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src)
               : "memory");
}

__device__ __forceinline__ void load_grid_row(const double* __restrict__ grid_rows, long long idx,
                                              GridRow& row) {
  const Double4 r0 = ldg256(grid_rows + 8 * idx), r1 = ldg256(grid_rows + 8 * idx + 4);
  row.thr = r0.a; row.x = r0.b; row.c = r0.c; row.s = r0.d;
  row.dc = r1.a; row.ds = r1.b; row.sinx = r1.c; row.cosx = r1.d;
}

// inverse_transform_sample (reference samplers.py:55-76) as a direct lookup: the
// sample index rint(interp(u)) is a step function of u, so the step a draw falls in is
// the number of host-computed exact thresholds <= u, and each step has its own row (the
// index may step down as well as up: non-monotone CDFs).  A uniform table over u gives
// the step at the bucket's edges; the few thresholds inside the bucket settle it.
__device__ __forceinline__ bool grid_sample(const KArgs& a, double u, GridRow& row) {
  const SSParams& p = a.p;
  long long idx;
  if (u != u) return false;
  if (u < 0.0 || u > p.cdf_last) {
    idx = p.grid_nthr + 1;                    // interp1d fill_value = 0.0: the fill row
  } else if (u >= 1.0) {                      // off the bucket table (never a [0,1) draw)
    int lo = 0, hi = p.grid_nthr;             // number of thresholds <= u
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(&a.t.grid_thr[mid]) <= u) lo = mid + 1; else hi = mid;
    }
    idx = lo;
  } else {
    const int b = (int)(u * (double)p.grid_lut_n);
    idx = step_from_entry(__ldg(reinterpret_cast<const double2*>(a.t.grid_lut) + b), a.t.grid_thr, u,
                          (uint32_t)p.grid_lut_n);
  }
  load_grid_row(a.t.grid_rows, idx, row);
  return true;
}

// inverse_transform_sample (reference samplers.py:55-76) on the host-built tables
__device__ __forceinline__ double invcdf_phi1(const KArgs& a, const int32_t* __restrict__ itab,
                                              double u) {
  const SSParams& p = a.p;
  const int n = p.cdf_n;
  double y;
  if (!(u >= p.cdf_first && u <= p.cdf_last)) {
    y = (u != u) ? u : 0.0;  // fill_value=0.0 outside; NaN stays NaN
  } else {
    const double2* __restrict__ cp = reinterpret_cast<const double2*>(a.t.cdf_pairs);
    int lo = 0, hi = n;
    if (p.cdf_guide_n > 0 && u >= 0.0 && u < 1.0) {
      const int32_t* guide = itab + 2 * p.cdf_guide_off;
      const int g = (int)(u * (double)p.cdf_guide_n);
      lo = guide[g];
      hi = guide[g + 1] + 1;
      if (hi > n) hi = n;
    }
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(&cp[mid].x) <= u) lo = mid; else hi = mid;
    }
    const double2 e = __ldg(&cp[lo]);
    const double base = p.cdf_identity ? (double)lo : (double)__ldg(&a.t.cdf_perm[lo]);
    y = (lo == n - 1 || e.x == u) ? base : __dadd_rn(__dmul_rn(e.y, __dsub_rn(u, e.x)), base);
  }
  if (y != y) return y;
  const long long idx = __double2ll_rn(y);  // np.rint: half to even
  if (a.t.grid) return __ldg(&a.t.grid[idx]);
  return (idx == n - 1) ? p.density_hi
                        : __dadd_rn(__dmul_rn((double)idx, p.grid_step), p.density_lo);
}

// The isochrone magnitudes are a function of ONE uniform: mass = F_IMF^-1(u) is piecewise
// linear in u on the 10000 IMF-CDF knots, mag_b(mass) piecewise linear on the isochrone
// rows, so mag_b(u) is piecewise linear on the union of both sets of breakpoints.  The host
// tabulates that composite (values at the breakpoints from the reference's own two-stage
// arithmetic) and the device finds the segment with the same one-round-trip bucket table
// as the phi1 grid: no search loops in shared memory, no division, 100 KB less shared
// memory.  Rows: 32 bytes = (a_g, dg/du, a_r, dr/du) with mag_b(u) = fma(slope_b, u, a_b) inside
// segment k, a_b = mag_b(U_k) - slope_b * U_k folded on the host in extended precision: one
// 256-bit load per star.  thr[k] = U_{k+1}.
__device__ __forceinline__ void iso_row_eval(const double* __restrict__ iso_rows, long long idx,
                                             double u, double& m1, double& m2) {
  const Double4 r = ldg256(iso_rows + 4 * idx);
  m1 = fma(r.b, u, r.a);
  m2 = fma(r.d, u, r.c);
}

__device__ __forceinline__ void iso_sample_direct(const KArgs& a, double u, double& m1, double& m2,
                                                  unsigned& domain_err) {
  const SSParams& p = a.p;
  if (!(u >= 0.0 && u <= 1.0)) {             // interp1d(bounds_error=True) raises
    domain_err++;
    m1 = m2 = ss::nan_d();
    return;
  }
  long long idx;
  if (u >= 1.0) {
    idx = p.iso_n - 1;
  } else {
    const int b = (int)(u * (double)p.iso_lut_n);
    idx = step_from_entry(__ldg(reinterpret_cast<const double2*>(a.t.iso_lut) + b), a.t.iso_thr, u,
                          (uint32_t)p.iso_lut_n);
  }
  iso_row_eval(a.t.iso_rows, idx, u, m1, m2);
}

// stream frame -> ICRS (reference observed.py:478-575; astropy: lon = atan2(y, x),
// lat = atan2(z, hypot(x, y))).  sl, cl = sin / cos of phi1; the transverse angle phi2 is
// small (Taylor series).  |w| = 1, so the branch-free atan2 / sqrt are inside their domains
// and the two chains interleave.
__device__ __forceinline__ void rotate_star(const SSParams& p, double sl, double cl, double phi2,
                                            double& wx, double& wy, double& wz, double& lon,
                                            double& ra, double& dec) {
  double sb, cb;
  ss::sincos_small(phi2 * ss::kNum[8], sb, cb);
  const double v0 = cb * cl, v1 = cb * sl, v2 = sb;
  wx = p.R[0] * v0 + p.R[3] * v1 + p.R[6] * v2;
  wy = p.R[1] * v0 + p.R[4] * v1 + p.R[7] * v2;
  wz = p.R[2] * v0 + p.R[5] * v1 + p.R[8] * v2;
  const double R2D = ss::kNum[7];
  ss::atan2_nb2(wy, wx, wz, ss::sqrt_nb(wx * wx + wy * wy), lon, dec);
  dec *= R2D;
  ra = lon * R2D;
  ra = (ra < 0.0) ? ra + 360.0 : ra;
}

// ugali imf.sample + isochrone interp1d (reference model.py:590-602)
__device__ __forceinline__ void iso_sample(const SSParams& p, const double* __restrict__ tab,
                                           const int32_t* __restrict__ itab, double u,
                                           double& m1, double& m2, unsigned& domain_err) {
  const double* __restrict__ cdf = tab + p.imf_cdf_off;
  const int n = p.imf_n;
  if (!(u >= cdf[0] && u <= cdf[n - 1])) {  // interp1d(bounds_error=True) raises
    domain_err++;
    m1 = m2 = ss::nan_d();
    return;
  }
  int j;
  if (p.imf_guide_n > 0 && u >= 0.0 && u < 1.0) {
    // guide[g] = last cdf entry <= g/G <= u: short forward scan
    const int32_t* guide = itab + 2 * p.imf_guide_off;
    const double fb = u * (double)p.imf_guide_n;
    const int b = (int)fb;
    const int lo = guide[b], hi = guide[b + 1];
    // the CDF is smooth: start from the position-proportional guess inside the bucket so
    // that the (warp-divergent) fix-up loops run once or twice instead of hi-lo times
    j = lo + (int)((fb - (double)b) * (double)(hi - lo));
    while (j > lo && cdf[j] > u) --j;
    while (j + 1 < n && cdf[j + 1] <= u) ++j;
  } else {
    j = ss::search_le(cdf, 0, n, u);
  }
  const double mj = (j == n - 1) ? p.imf_mass_hi
                                 : __dadd_rn(__dmul_rn((double)j, p.imf_mass_step), p.imf_mass_lo);
  double mass;
  const double cj = cdf[j];
  if (j == n - 1 || cj == u) {
    mass = mj;
  } else {
    const double mj1 = (j + 1 == n - 1)
                           ? p.imf_mass_hi
                           : __dadd_rn(__dmul_rn((double)(j + 1), p.imf_mass_step), p.imf_mass_lo);
    const double slope = __ddiv_rn(__dsub_rn(mj1, mj), __dsub_rn(cdf[j + 1], cj));
    mass = __dadd_rn(__dmul_rn(slope, __dsub_rn(u, cj)), mj);
  }
  const double* __restrict__ mi = tab + p.iso_mass_off;
  const int ni = p.iso_n;
  if (!(mass >= mi[0] && mass <= mi[ni - 1])) {
    domain_err++;
    m1 = m2 = ss::nan_d();
    return;
  }
  ss::Guide G;
  G.g = itab + 2 * p.iso_guide_off;
  G.n = p.iso_guide_n;
  G.x0 = p.iso_guide_x0;
  G.inv = p.iso_guide_inv;
  const int k = ss::locate(mi, ni, G, mass);
  m1 = ss::lin_at(mi, tab + p.iso_mag_off[0], ni, k, mass);
  m2 = ss::lin_at(mi, tab + p.iso_mag_off[1], ni, k, mass);
}

// ---------------------------------------------------------------- the star kernels
// Random draws (include/streamsim_b200.h): two Philox blocks per star.
//   POS: w0 -> u_phi1, w1 -> u_mass, (w2, w3) -> z_phi2, z_dist (or two uniforms)
//   OBS: (w0, w1) -> z_r, z_g, w2 -> u_det, w3 -> u_det2
//
// k_stars: the generic kernel -- any stage mask, injected draws, partially filled
// catalogs, every density kind, NESTED maps, draw export.
// MATH = SS_MATH_F64: every value in double.  MATH = SS_MATH_FAST: identical decisions
// (pixels, flags, counters), magerr / mag_obs VALUES from fp32 MUFU arithmetic with a
// per-star double redo inside a guard band (ss::observe_star).
template <uint32_t ST, typename OutT, int MATH>
__global__ void __launch_bounds__(SS_BLOCK, 1) k_stars(const __grid_constant__ KArgs a) {
  extern __shared__ __align__(16) unsigned char smem_blob[];
  __shared__ unsigned long long s_bar;
  __shared__ unsigned int s_cnt[SS_N_COUNTERS];
  __shared__ unsigned int s_hist[SS_N_HIST * SS_HIST_BINS];

  const SSParams& p = a.p;
  const double* const d_u_phi1 = a.d.u_phi1;
  const double* const d_u_mass = a.d.u_mass;
  const double* const d_n_phi2 = a.d.n_phi2;
  const double* const d_n_dist = a.d.n_dist;
  const double* const d_z_g = a.d.z_flux[SS_BAND_G];
  const double* const d_z_r = a.d.z_flux[SS_BAND_R];
  const double* const d_u_det = a.d.u_det;
  const double* const d_u_det2 = a.d.u_det2;
  const double* const in_phi1 = a.in.phi1;
  const double* const in_phi2 = a.in.phi2;
  const double* const in_dist = a.in.dist;
  double* const draws_out = a.o.draws_out;
  const int density_kind = p.density_kind;
  const bool has_dist = p.has_dist != 0;
  const bool has_iso = p.has_iso != 0;
  const bool band_g = p.band_present[SS_BAND_G] != 0;
  const bool band_r = p.band_present[SS_BAND_R] != 0;
  const int nest = p.nest;
  const bool want_cnt = a.o.counters != nullptr;
  const bool want_hist = want_cnt && p.hist_enable;
  const int share = ss::curve_share(p);

  for (int i = threadIdx.x; i < SS_N_COUNTERS; i += blockDim.x) s_cnt[i] = 0;
  if (want_hist)
    for (int i = threadIdx.x; i < SS_N_HIST * SS_HIST_BINS; i += blockDim.x) s_hist[i] = 0;

  // `tab`: the hot part of the blob (first blob_hot_bytes), staged into shared memory.
  // `stab`: base for the stream-model tables (track / distance-modulus knots, inverse-CDF
  // guide) -- the same shared copy, or, when the phi1 grid table makes them cold, the
  // global blob (shared memory not spent on them is L1 cache for the gathers).
  const double* tab;
  if (a.blob_in_smem) {
    stage_blob(smem_blob, a.t.blob, a.blob_bytes, &s_bar);
    tab = reinterpret_cast<const double*>(smem_blob);
  } else {
    tab = reinterpret_cast<const double*>(a.t.blob);
  }
  const double* stab = (a.t.blob_hot_bytes < a.t.blob_bytes)
                           ? reinterpret_cast<const double*>(a.t.blob) : tab;
  const int32_t* itab = reinterpret_cast<const int32_t*>(tab);
  const int32_t* sitab = reinterpret_cast<const int32_t*>(stab);
  __syncthreads();

  // Summary counters go straight to shared memory (uniform-address atomicAdd: REDUX +
  // one ATOMS per warp) instead of living in registers across the whole loop body.
#define SS_COUNT(slot, cond) \
  do { if (want_cnt) atomicAdd(&s_cnt[slot], (cond) ? 1u : 0u); } while (0)

  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
       i += stride) {
    const unsigned long long gi = a.offset + i;
    unsigned c_dom = 0;
    double phi1 = 0, phi2 = 0, dist = ss::nan_d(), mag_g = ss::nan_d(), mag_r = ss::nan_d();
    double ra = 0, dec = 0, wx = 0, wy = 0, wz = 0, lon = 0;
    bool on_grid = false;   // phi1 came from the grid table: row holds f(phi1) values
    GridRow row;
    row.c = row.s = row.dc = row.ds = row.sinx = row.cosx = row.x = 0.0;

    // ------------------------------------------------ generate (model.py:106-157)
    if (ST & SS_STAGE_GENERATE) {
      const bool z_track = p.track_sampler == SS_SAMPLER_GAUSSIAN;
      const bool z_dist = p.dist_sampler == SS_SAMPLER_GAUSSIAN;
      const bool all_given = d_u_phi1 && (d_u_mass || !has_iso) && d_n_phi2 && (d_n_dist || !has_dist);
      uint4 w = make_uint4(0u, 0u, 0u, 0u);
      if (!all_given) w = ss::philox_block(gi, SS_DRAW_BLOCK_POS, a.keys);
      double u_phi1 = d_u_phi1 ? d_u_phi1[i] : ss::u32(w.x);
      const double u_mass = d_u_mass ? d_u_mass[i] : ss::u32(w.y);
      double z0 = 0, z1 = 0;
      if (!all_given && (z_track || (has_dist && z_dist))) ss::normal_pair(p.normals_f64 != 0, w.z, w.w, z0, z1);
      const double n_phi2 = d_n_phi2 ? d_n_phi2[i] : (z_track ? z0 : ss::u32(w.z));
      const double n_dist = d_n_dist ? d_n_dist[i] : (z_dist ? z1 : ss::u32(w.w));

      const double phi1_given = in_phi1 ? in_phi1[i] : ss::nan_d();
      if (phi1_given == phi1_given) {
        phi1 = phi1_given;
      } else if (density_kind == SS_DENSITY_GRID) {
        on_grid = grid_sample(a, u_phi1, row);
        phi1 = on_grid ? row.x : ss::nan_d();
      } else if (density_kind == SS_DENSITY_UNIFORM) {
        phi1 = __dadd_rn(__dmul_rn(u_phi1, __dsub_rn(p.density_hi, p.density_lo)), p.density_lo);
      } else if (density_kind == SS_DENSITY_INVCDF) {
        phi1 = invcdf_phi1(a, sitab, u_phi1);
      } else if (density_kind == SS_DENSITY_GAUSSIAN) {  // the u_phi1 slot carries a normal draw
        if (!d_u_phi1) {
          double zz;
          const uint4 wz4 = ss::philox_block(gi, SS_DRAW_BLOCK_DENSITY_Z, a.keys);
          ss::normal_pair(p.normals_f64 != 0, wz4.x, wz4.y, u_phi1, zz);
        }
        phi1 = __dadd_rn(__dmul_rn(u_phi1, p.density_hi), p.density_lo);
      } else {
        phi1 = ss::nan_d();
      }
      if (draws_out) {
        draws_out[SS_DRAW_U_PHI1 * a.n + i] = u_phi1;
        draws_out[SS_DRAW_U_MASS * a.n + i] = u_mass;
        draws_out[SS_DRAW_N_PHI2 * a.n + i] = n_phi2;
        draws_out[SS_DRAW_N_DIST * a.n + i] = n_dist;
      }
      const double phi2_given = in_phi2 ? in_phi2[i] : ss::nan_d();
      if (phi2_given == phi2_given) {
        phi2 = phi2_given;
      } else {
        double c = row.c, s = row.s;
        if (!on_grid) track_eval(p.track_center, p.track_spread, stab, phi1, c, s);
        phi2 = track_draw(c, s, p.track_sampler, n_phi2, c_dom);
      }
      const double dist_given = in_dist ? in_dist[i] : ss::nan_d();
      if (dist_given == dist_given) dist = dist_given;
      if (has_dist || has_iso) {
        if (!(dist_given == dist_given) && has_dist) {
          double c = row.dc, s = row.ds;
          if (!on_grid) track_eval(p.dist_center, p.dist_spread, stab, phi1, c, s);
          dist = track_draw(c, s, p.dist_sampler, n_dist, c_dom);
        }
        if (has_iso) {
          double m1, m2;
          if (p.iso_direct) iso_sample_direct(a, u_mass, m1, m2, c_dom);
          else iso_sample(p, tab, itab, u_mass, m1, m2, c_dom);
          mag_g = __dadd_rn(m1, dist);
          mag_r = __dadd_rn(m2, dist);
        }
      }
      if (c_dom && want_cnt) atomicAdd(&s_cnt[SS_CNT_DOMAIN], c_dom);
      put<OutT>(a.o.phi1, i, phi1);
      put<OutT>(a.o.phi2, i, phi2);
      put<OutT>(a.o.dist, i, dist);
      put<OutT>(a.o.mag[SS_BAND_G], i, mag_g);
      put<OutT>(a.o.mag[SS_BAND_R], i, mag_r);
    }

    // ------------------------------------------------ rotate (observed.py:573)
    if (ST & SS_STAGE_ROTATE) {
      if (!(ST & SS_STAGE_GENERATE)) {
        phi1 = in_phi1[i];
        phi2 = in_phi2[i];
      }
      double sl = row.sinx, cl = row.cosx;
      if (!on_grid) sincospi(phi1 * (1.0 / 180.0), &sl, &cl);
      rotate_star(p, sl, cl, phi2, wx, wy, wz, lon, ra, dec);
      put<OutT>(a.o.ra, i, ra);
      put<OutT>(a.o.dec, i, dec);
    }

    // ------------------------------------------------ observe (observed.py:146-291)
    if (ST & SS_STAGE_OBSERVE) {
      if (!(ST & SS_STAGE_ROTATE)) {
        ra = a.in.ra[i];
        dec = a.in.dec[i];
      }
      if (!(ST & SS_STAGE_GENERATE)) {
        mag_g = band_g ? a.in.mag[SS_BAND_G][i] : ss::nan_d();
        mag_r = band_r ? a.in.mag[SS_BAND_R][i] : ss::nan_d();
      }
      ss::HpLoc L;
      if ((ST & SS_STAGE_ROTATE) && fabs(wz) < 0.9999 && lon == lon) {
        L = ss::hp_from_vector(wz, lon);
      } else {
        L = ss::hp_locate(ra, dec);
      }
      double ebv = ss::nan_d(), ml_g = ss::nan_d(), ml_r = ss::nan_d();
      if (L.valid) {
        const long long pe = ss::hp_pix(L, p.nside_ebv, nest);
        ebv = __ldg(&a.m.ebv[pe]);
        long long pg = -1;
        if (band_g) {
          pg = (p.nside_maglim[SS_BAND_G] == p.nside_ebv) ? pe
                                                          : ss::hp_pix(L, p.nside_maglim[SS_BAND_G], nest);
          ml_g = __ldg(&a.m.maglim[SS_BAND_G][pg]);
        }
        if (band_r) {
          const long long pr =
              (p.nside_maglim[SS_BAND_R] == p.nside_ebv) ? pe
              : (pg >= 0 && p.nside_maglim[SS_BAND_R] == p.nside_maglim[SS_BAND_G])
                  ? pg
                  : ss::hp_pix(L, p.nside_maglim[SS_BAND_R], nest);
          ml_r = __ldg(&a.m.maglim[SS_BAND_R][pr]);
        }
        if (a.o.pix) a.o.pix[i] = ss::hp_pix(L, p.nside_pix, nest);
      } else {
        if (want_cnt) atomicAdd(&s_cnt[SS_CNT_BADCOORD], 1u);
        if (a.o.pix) a.o.pix[i] = -1;
      }

      const bool need_d2 = p.perfect_galstarsep != 0;
      const bool all_given = d_z_g && d_z_r && d_u_det && (d_u_det2 || !need_d2);
      uint4 w = make_uint4(0u, 0u, 0u, 0u);
      double z_r = 0, z_g = 0;
      if (!all_given) {
        w = ss::philox_block(gi, SS_DRAW_BLOCK_OBS, a.keys);
        ss::normal_pair(p.normals_f64 != 0, w.x, w.y, z_r, z_g);
      }
      if (d_z_g) z_g = d_z_g[i];
      if (d_z_r) z_r = d_z_r[i];
      const double u_det = d_u_det ? d_u_det[i] : ss::u32(w.z);
      const double u_det2 = d_u_det2 ? d_u_det2[i] : ss::u32(w.w);
      if (draws_out) {
        draws_out[SS_DRAW_Z_G * a.n + i] = z_g;
        draws_out[SS_DRAW_Z_R * a.n + i] = z_r;
        draws_out[SS_DRAW_U_DET * a.n + i] = u_det;
        draws_out[SS_DRAW_U_DET2 * a.n + i] = u_det2;
      }
      ss::ObsOut o;
      ss::observe_star<MATH>(p, tab, share, band_g, band_r, mag_g, mag_r, ebv, ml_g, ml_r, z_g,
                             z_r, u_det, u_det2, o);
      SS_COUNT(SS_CNT_UNSEEN, o.unseen);
      if (band_g) {
        SS_COUNT(SS_CNT_BADMAG_G, o.bad[SS_BAND_G]);
        put<OutT>(a.o.mag_obs[SS_BAND_G], i, o.mobs[SS_BAND_G]);
        put<OutT>(a.o.magerr[SS_BAND_G], i, o.err[SS_BAND_G]);
      }
      if (band_r) {
        SS_COUNT(SS_CNT_BADMAG_R, o.bad[SS_BAND_R]);
        put<OutT>(a.o.mag_obs[SS_BAND_R], i, o.mobs[SS_BAND_R]);
        put<OutT>(a.o.magerr[SS_BAND_R], i, o.err[SS_BAND_R]);
      }
      if (a.o.flag_observed) a.o.flag_observed[i] = o.observed ? 1 : 0;
      if (a.o.flag_perfect) a.o.flag_perfect[i] = o.perfect ? 1 : 0;
      SS_COUNT(SS_CNT_OBSERVED, o.observed);
      if (p.perfect_galstarsep) SS_COUNT(SS_CNT_PERFECT, o.perfect);
    }

    if (want_hist) {
      // only quantities this launch produced (or was handed) are binned
      const bool have_pos = (ST & SS_STAGE_GENERATE) || (ST & SS_STAGE_ROTATE);
      const bool have_dist = (ST & SS_STAGE_GENERATE) != 0;
      const bool have_mag = (ST & SS_STAGE_GENERATE) || (ST & SS_STAGE_OBSERVE);
      const double hv[SS_N_HIST] = {phi1, phi2, dist, mag_g, mag_g - mag_r};
      const bool hb[SS_N_HIST] = {have_pos, have_pos, have_dist, have_mag, have_mag};
#pragma unroll
      for (int h = 0; h < SS_N_HIST; ++h) {
        if (!hb[h]) continue;
        const int bin = ss::hist_bin((float)hv[h], a.hist_lo[h], a.hist_inv[h]);
        if (bin >= 0) atomicAdd(&s_hist[h * SS_HIST_BINS + bin], 1u);
      }
    }
  }

  if (want_cnt) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&a.o.counters[SS_CNT_TOTAL], a.n);
    __syncthreads();
    for (int k = threadIdx.x; k < SS_N_COUNTERS; k += blockDim.x)
      if (s_cnt[k]) atomicAdd(&a.o.counters[k], (unsigned long long)s_cnt[k]);
    if (want_hist)
      for (int k = threadIdx.x; k < SS_N_HIST * SS_HIST_BINS; k += blockDim.x)
        if (s_hist[k]) atomicAdd(&a.o.counters[SS_N_COUNTERS + k], (unsigned long long)s_hist[k]);
  }
}

// k_fused: the production kernel -- free-running generate -> rotate -> observe with the
// phi1 grid table, the composite isochrone table, both bands, RING maps and the survey
// curves in shared memory (the launcher checks all of that; anything else runs k_stars).
// Same arithmetic as k_stars through the same device functions, so both produce the same
// bytes (tests/test_gpu_parity.py::test_fused_kernel_equals_generic); what differs is the
// schedule:
//   * the star index is 32-bit inside a launch (ss_run splits longer ranges);
//   * the draws are 32-bit words, so the two step functions (phi1 grid, isochrone) are
//     decided in integers on the word tables (SSTables.grid_wlut / iso_wlut): one 16-byte
//     bucket entry with three thresholds settles 99.9 % of the draws without a scan;
//   * the POS Philox block of the NEXT star is computed right behind the map gather and its
//     two bucket entries are requested with cp.async into per-thread shared-memory slots (no
//     registers in flight): their L2 round trip is over when the next iteration starts;
//   * the three row loads of an iteration are issued back to back and the star's OBS draw
//     block (Philox + fp32 Box-Muller) runs behind them;
//   * the phi1 histogram bin comes with the grid row; the other four are binned in fp32
//     right after the generate stage, before the rotation raises the register pressure;
//   * PACKED: the three survey maps interleaved as one 32-byte record per pixel -- one
//     sector per star instead of three.
//   * ALLCOLS: every floating-point column and flag_observed are requested (the launcher
//     checks) -- no per-store NULL tests.
template <typename OutT, int MATH, bool HIST, bool PACKED, bool ALLCOLS>
__global__ void __launch_bounds__(SS_FUSED_BLOCK, SS_FUSED_MIN_CTAS)
k_fused(const __grid_constant__ KArgs a) {
  extern __shared__ __align__(16) unsigned char smem_blob[];
  __shared__ unsigned long long s_bar;
  __shared__ unsigned int s_cnt[SS_N_COUNTERS];
  // one extra slot at the end collects the out-of-range values: no branch around the atomics
  __shared__ unsigned int s_hist[HIST ? SS_N_HIST * SS_HIST_BINS + 1 : 1];
  constexpr int kDump = SS_N_HIST * SS_HIST_BINS;

  const SSParams& p = a.p;
  const bool want_cnt = a.o.counters != nullptr;
  const int share = ss::curve_share(p);
  for (int i = threadIdx.x; i < SS_N_COUNTERS; i += blockDim.x) s_cnt[i] = 0;
  if (HIST)
    for (int i = threadIdx.x; i <= kDump; i += blockDim.x) s_hist[i] = 0;
  // The five per-star summary counts live in three registers as 16-bit fields (a thread
  // sees at most 2^31 / (148 * 1024) < 2^16 stars per launch) and are reduced once, after
  // the loop: one add per count and star instead of a warp reduction and a shared atomic.
  uint32_t c_unseen_badg = 0, c_badr_obs = 0, c_perfect = 0;
  stage_blob(smem_blob, a.t.blob, a.blob_bytes, &s_bar);
  const double* const tab = reinterpret_cast<const double*>(smem_blob);
  __syncthreads();

  const uint32_t n = (uint32_t)a.n;
  const uint32_t stride = gridDim.x * blockDim.x;
  const bool z_track = p.track_sampler == SS_SAMPLER_GAUSSIAN;
  const bool z_dist = p.dist_sampler == SS_SAMPLER_GAUSSIAN;
  const bool nf64 = p.normals_f64 != 0;

  const uint4* __restrict__ const glut = reinterpret_cast<const uint4*>(a.t.grid_wlut);
  const uint4* __restrict__ const ilut = reinterpret_cast<const uint4*>(a.t.iso_wlut);
  const int gshift = a.grid_wshift, ishift = a.iso_wshift;
  uint4 wp = make_uint4(0u, 0u, 0u, 0u), eg = wp, ei = wp;
  // this thread's three 16-byte staging slots behind the blob (conflict-free: slot k of
  // thread t at 16 * (k * blockDim + t))
  uint4* const stg = reinterpret_cast<uint4*>(smem_blob + a.blob_bytes) + threadIdx.x;
  auto prepare = [&](uint32_t star) {
    const uint4 w = ss::philox_block_c<SS_DRAW_BLOCK_POS>(a.offset + star, a.keys);
    stg[0] = w;
#if !(SS_WHATIF & 64)
    cp_async16(stg + SS_FUSED_BLOCK, glut + (w.x >> gshift));
    cp_async16(stg + 2 * SS_FUSED_BLOCK, ilut + (w.y >> ishift));
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
  };
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) prepare(i);
  for (; i < n; i += stride) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    wp = stg[0];
    eg = stg[SS_FUSED_BLOCK];
    ei = stg[2 * SS_FUSED_BLOCK];
    // ------------------------------------------------ generate (model.py:106-157)
    const double u_mass = ss::u32(wp.y);
    // both steps first, then the three row loads back to back, then work that needs none of
    // them; the empty asm keeps every loaded register alive up to that point -- otherwise the
    // register allocator hands the unused high word of row.thr to the code in between, whose
    // first write then waits for the load (a write-after-write hazard on a load in flight)
#if SS_WHATIF & 64
    const long long sg = ((unsigned long long)wp.x * (unsigned)p.grid_nthr) >> 32;
    const long long si = ((unsigned long long)wp.y * (unsigned)(p.iso_n - 1)) >> 32;
#else
    const long long sg = step_from_words(eg, a.t.grid_wthr, wp.x, p.grid_nthr);
    const long long si = step_from_words(ei, a.t.iso_wthr, wp.y, p.iso_n - 1);
#endif
#if SS_WHATIF & 32
    Double4 g0, g1, ri;
    g0.a = 0.0; g0.b = 1e-4 * (double)sg - 5.0; g0.c = 0.1; g0.d = 0.2; g1.a = 16.5; g1.b = 0.05;
    g1.c = 1e-5 * (double)sg; g1.d = 1.0 - g1.c; ri.a = 5.0; ri.b = 1e-3 * (double)si; ri.c = 4.5; ri.d = ri.b;
#else
#if SS_WHATIF & 64
    const Double4 g0 = ldg256(a.t.grid_rows + 8 * sg);
    Double4 g1;
    g1.a = 16.5 + 0.1 * g0.b; g1.b = 1e-4; g1.c = sin(g0.b * 0.017453292519943295); g1.d = cos(g0.b * 0.017453292519943295);
#else
    const Double4 g0 = ldg256(a.t.grid_rows + 8 * sg), g1 = ldg256(a.t.grid_rows + 8 * sg + 4);
#endif
    const Double4 ri = ldg256(a.t.iso_rows + 4 * si);
#endif
    const uint4 wo = ss::philox_block_c<SS_DRAW_BLOCK_OBS>(a.offset + i, a.keys);
    // the flux-noise normals travel to the observe stage as two fp32 values -- or, with
    // double-precision normals, as the two words they are made of there
    uint32_t c_r = wo.x, c_g = wo.y;
    double nz0 = 0.0, nz1 = 0.0;
    if (nf64) {
      if (z_track || z_dist) ss::bm64(wp.z, wp.w, nz0, nz1);
    } else {
      float zf_r, zf_g, zf0 = 0.0f, zf1 = 0.0f;
      ss::bm32f(wo.x, wo.y, zf_r, zf_g);
      if (z_track || z_dist) ss::bm32f(wp.z, wp.w, zf0, zf1);
      c_r = __float_as_uint(zf_r);
      c_g = __float_as_uint(zf_g);
      nz0 = (double)zf0;
      nz1 = (double)zf1;
    }
    asm volatile("" ::"d"(g0.a), "d"(g0.b), "d"(g0.c), "d"(g0.d), "d"(g1.a), "d"(g1.b), "d"(g1.c),
                 "d"(g1.d), "d"(ri.a), "d"(ri.b), "d"(ri.c), "d"(ri.d));
    GridRow row;
    row.thr = g0.a; row.x = g0.b; row.c = g0.c; row.s = g0.d;
    row.dc = g1.a; row.ds = g1.b; row.sinx = g1.c; row.cosx = g1.d;
    const double m1 = fma(ri.b, u_mass, ri.a), m2 = fma(ri.d, u_mass, ri.c);
    // the two track draws: normals (fp32 Box-Muller) or uniforms, per the samplers
    const double n_phi2 = z_track ? nz0 : ss::u32(wp.z);
    const double n_dist = z_dist ? nz1 : ss::u32(wp.w);
    unsigned c_dom = 0;
    const double phi2 = track_draw(row.c, row.s, p.track_sampler, n_phi2, c_dom);
    const double dist = track_draw(row.dc, row.ds, p.dist_sampler, n_dist, c_dom);
    if (c_dom && want_cnt) atomicAdd(&s_cnt[SS_CNT_DOMAIN], c_dom);
    const double mag_g = __dadd_rn(m1, dist), mag_r = __dadd_rn(m2, dist);
#if SS_VEC_STORES
    const bool full = a.vec_ok && (i | 31u) < n;                   // warp-uniform
    put2<OutT>(a.o.phi1, a.o.phi2, i, row.x, phi2, full);
    put2<OutT>(a.o.dist, a.o.mag[SS_BAND_G], i, dist, mag_g, full);
#else
    putc<OutT, ALLCOLS>(a.o.phi1, i, row.x);
    putc<OutT, ALLCOLS>(a.o.phi2, i, phi2);
    putc<OutT, ALLCOLS>(a.o.dist, i, dist);
    putc<OutT, ALLCOLS>(a.o.mag[SS_BAND_G], i, mag_g);
#endif
    putc<OutT, ALLCOLS>(a.o.mag[SS_BAND_R], i, mag_r);
    if (HIST && !(SS_WHATIF & 4)) {
      const int b0 = __double2loint(row.thr);                        // host-binned phi1
      atomicAdd(&s_hist[b0 >= 0 ? b0 : kDump], 1u);
      const float hv[4] = {(float)phi2, (float)dist, (float)mag_g, (float)(mag_g - mag_r)};
#pragma unroll
      for (int h = 1; h < SS_N_HIST; ++h) {
        const int bin = ss::hist_bin(hv[h - 1], a.hist_lo[h], a.hist_inv[h]);
        atomicAdd(&s_hist[bin >= 0 ? h * SS_HIST_BINS + bin : kDump], 1u);
      }
    }

    // ------------------------------------------------ rotate (observed.py:573)
    double wx, wy, wz, lon, ra, dec;
#if SS_WHATIF & 16
    wx = row.cosx; wy = row.sinx; wz = phi2 * 0.01; lon = 4.0 + 0.1 * row.sinx; ra = lon * 57.0; dec = wz * 57.0;
#else
    rotate_star(p, row.sinx, row.cosx, phi2, wx, wy, wz, lon, ra, dec);
#endif
#if SS_VEC_STORES
    put2<OutT>(a.o.ra, a.o.dec, i, ra, dec, full);
#else
    putc<OutT, ALLCOLS>(a.o.ra, i, ra);
    putc<OutT, ALLCOLS>(a.o.dec, i, dec);
#endif

    // ------------------------------------------------ observe (observed.py:146-291)
    ss::HpLoc L;
    if (fabs(wz) < 0.9999 && lon == lon) L = ss::hp_from_vector(wz, lon);
    else L = ss::hp_locate(ra, dec);
    double ebv = ss::nan_d(), ml_g = ss::nan_d(), ml_r = ss::nan_d(), map_pad = 0.0;
    if (L.valid) {
#if SS_HP32
      const long long pe = (L.have_sth || p.nside_ebv > 8192) ? ss::hp_ring(L, p.nside_ebv)
                                                             : (long long)ss::hp_ring32(L.z, L.tt, p.nside_ebv);
#else
      const long long pe = ss::hp_ring(L, p.nside_ebv);
#endif
      if (SS_WHATIF & 1) {
        ebv = 0.05 + 1e-9 * (double)(pe & 1023); ml_g = 26.0; ml_r = 25.7;
      } else if (PACKED) {
        if (a.m.packed32) {                                          // fp32 records (option)
          const float4 r = __ldg(reinterpret_cast<const float4*>(a.m.packed32) + pe);
          ebv = (double)r.x; ml_g = (double)r.y; ml_r = (double)r.z; map_pad = (double)r.w;
        } else {
          const Double4 r = ldg256(a.m.packed + 4 * pe);
          ebv = r.a; ml_g = r.b; ml_r = r.c; map_pad = r.d;
        }
      } else {
        ebv = __ldg(&a.m.ebv[pe]);
        const long long pg = (p.nside_maglim[SS_BAND_G] == p.nside_ebv)
                                 ? pe : ss::hp_ring(L, p.nside_maglim[SS_BAND_G]);
        ml_g = __ldg(&a.m.maglim[SS_BAND_G][pg]);
        const long long pr = (p.nside_maglim[SS_BAND_R] == p.nside_ebv) ? pe
                             : (p.nside_maglim[SS_BAND_R] == p.nside_maglim[SS_BAND_G])
                                 ? pg : ss::hp_ring(L, p.nside_maglim[SS_BAND_R]);
        ml_r = __ldg(&a.m.maglim[SS_BAND_R][pr]);
      }
      if (a.o.pix) a.o.pix[i] = (p.nside_pix == p.nside_ebv) ? pe : ss::hp_ring(L, p.nside_pix);
    } else {
      if (want_cnt) atomicAdd(&s_cnt[SS_CNT_BADCOORD], 1u);
      if (a.o.pix) a.o.pix[i] = -1;
    }
    if (i + stride < n) prepare(i + stride);                       // n <= 2^31: no wrap-around
    // (the gathered record stays whole until here: see the row loads above)
    asm volatile("" ::"d"(ebv), "d"(ml_g), "d"(ml_r), "d"(map_pad));
    double z_r = (double)__uint_as_float(c_r), z_g = (double)__uint_as_float(c_g);
    if (nf64) ss::bm64(c_r, c_g, z_r, z_g);
    ss::ObsOut o;
#if SS_WHATIF & 8
    o.mobs[0] = mag_g + ebv; o.mobs[1] = mag_r + ml_g; o.err[0] = ml_r; o.err[1] = z_r + z_g;
    o.observed = wo.z > wo.w; o.perfect = false; o.unseen = false; o.bad[0] = o.bad[1] = false;
#else
    ss::observe_star<MATH, true>(p, tab, share, true, true, mag_g, mag_r, ebv, ml_g, ml_r, z_g, z_r,
                                 ss::u32(wo.z), ss::u32(wo.w), o);
#endif
#if SS_VEC_STORES
    put2<OutT>(a.o.mag_obs[SS_BAND_G], a.o.magerr[SS_BAND_G], i, o.mobs[SS_BAND_G], o.err[SS_BAND_G], full);
    put2<OutT>(a.o.mag_obs[SS_BAND_R], a.o.magerr[SS_BAND_R], i, o.mobs[SS_BAND_R], o.err[SS_BAND_R], full);
#else
    putc<OutT, ALLCOLS>(a.o.mag_obs[SS_BAND_G], i, o.mobs[SS_BAND_G]);
    putc<OutT, ALLCOLS>(a.o.magerr[SS_BAND_G], i, o.err[SS_BAND_G]);
    putc<OutT, ALLCOLS>(a.o.mag_obs[SS_BAND_R], i, o.mobs[SS_BAND_R]);
    putc<OutT, ALLCOLS>(a.o.magerr[SS_BAND_R], i, o.err[SS_BAND_R]);
#endif
    if (ALLCOLS || a.o.flag_observed) a.o.flag_observed[i] = o.observed ? 1 : 0;
    if (a.o.flag_perfect) a.o.flag_perfect[i] = o.perfect ? 1 : 0;
    c_unseen_badg += (o.unseen ? 1u : 0u) + (o.bad[SS_BAND_G] ? 0x10000u : 0u);
    c_badr_obs += (o.bad[SS_BAND_R] ? 1u : 0u) + (o.observed ? 0x10000u : 0u);
    c_perfect += o.perfect ? 1u : 0u;

  }

  if (want_cnt) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&a.o.counters[SS_CNT_TOTAL], a.n);
    atomicAdd(&s_cnt[SS_CNT_UNSEEN], c_unseen_badg & 0xffffu);
    atomicAdd(&s_cnt[SS_CNT_BADMAG_G], c_unseen_badg >> 16);
    atomicAdd(&s_cnt[SS_CNT_BADMAG_R], c_badr_obs & 0xffffu);
    atomicAdd(&s_cnt[SS_CNT_OBSERVED], c_badr_obs >> 16);
    atomicAdd(&s_cnt[SS_CNT_PERFECT], c_perfect);
    __syncthreads();
    for (int k = threadIdx.x; k < SS_N_COUNTERS; k += blockDim.x)
      if (s_cnt[k]) atomicAdd(&a.o.counters[k], (unsigned long long)s_cnt[k]);
    if (HIST)
      for (int k = threadIdx.x; k < SS_N_HIST * SS_HIST_BINS; k += blockDim.x)
        if (s_hist[k]) atomicAdd(&a.o.counters[SS_N_COUNTERS + k], (unsigned long long)s_hist[k]);
  }
}

// ---------------------------------------------------------------- small kernels
__global__ void k_frame_fraction(SSParams p, SSMaps m, SSCatalog in, unsigned long long* counters,
                                 unsigned long long n) {
  unsigned inside = 0;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    double sl, cl, sb, cb;
    sincospi(in.phi1[i] * (1.0 / 180.0), &sl, &cl);
    sincospi(in.phi2[i] * (1.0 / 180.0), &sb, &cb);
    const double v0 = cb * cl, v1 = cb * sl, v2 = sb;
    const double wx = p.R[0] * v0 + p.R[3] * v1 + p.R[6] * v2;
    const double wy = p.R[1] * v0 + p.R[4] * v1 + p.R[7] * v2;
    const double wz = p.R[2] * v0 + p.R[5] * v1 + p.R[8] * v2;
    double ra = atan2(wy, wx) * 57.29577951308232;
    if (ra < 0.0) ra += 360.0;
    const double dec = atan2(wz, sqrt(wx * wx + wy * wy)) * 57.29577951308232;
    const ss::HpLoc L = ss::hp_locate(ra, dec);
    if (L.valid) inside += m.mask[ss::hp_pix(L, m.nside_mask, p.nest)] != 0;
  }
  const unsigned s = __reduce_add_sync(0xffffffffu, inside);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(&counters[SS_CNT_INSIDE], (unsigned long long)s);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&counters[SS_CNT_TOTAL], n);
}

__global__ void k_photo_error(SSParams p, const double* tab, int band, const double* mag,
                              const double* maglim, double* out, unsigned long long n) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += stride)
    out[i] = ss::photo_error(p, tab, band, mag[i], maglim[i]);
}

__global__ void k_efficiency(SSParams p, const double* tab, int band, int detection_only,
                             const double* mag, const double* maglim, double* out,
                             unsigned long long n) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += stride)
    out[i] = ss::efficiency(p, detection_only ? p.detection : p.completeness, tab, band, mag[i],
                            maglim[i]);
}

__global__ void k_eval_function(SSFunc f, const double* tab, const double* x, double* out,
                                unsigned long long n) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += stride)
    out[i] = ss::fn_eval(f, tab, x[i]);
}

__global__ void k_ang2pix(int nside, int nest, const double* lon, const double* lat,
                          long long* pix, unsigned long long n) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    const ss::HpLoc L = ss::hp_locate(lon[i], lat[i]);
    pix[i] = L.valid ? ss::hp_pix(L, nside, nest) : -1;
  }
}

__global__ void k_math_probe(int fn, const double* x, const double* y2, double* out,
                             unsigned long long n) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    const double v = x[i];
    double r;
    switch (fn) {
      case SS_MATH_FN_LOG10: r = ss::log10_nb(v); break;
      case SS_MATH_FN_EXP10: r = ss::exp10_nb(v); break;
      case SS_MATH_FN_ATAN2: r = ss::atan2_nb(v, y2[i]); break;
      default: r = ss::sqrt_nb(v); break;
    }
    out[i] = r;
  }
}

// rv.rvs(loc, scale): draw * scale + loc per element (reference samplers.py:94-138); the
// draw is injected or the star's z_phi2 / u_phi2 (POS block words 2, 3)
__global__ void k_loc_scale(int normal, const double* loc, const double* scale, double loc0,
                            double scale0, const double* draws, double* out, unsigned long long n,
                            unsigned long long offset, const ss::PhiloxKeys keys) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    double d;
    if (draws) {
      d = draws[i];
    } else {
      const uint4 w = ss::philox_block(offset + i, SS_DRAW_BLOCK_POS, keys);
      double z0, z1;
      ss::bm32(w.z, w.w, z0, z1);
      d = normal ? z0 : ss::u32(w.z);
    }
    out[i] = __dadd_rn(__dmul_rn(d, scale ? scale[i] : scale0), loc ? loc[i] : loc0);
  }
}

// StreamInjector.sample_measured_magnitudes (reference observed.py:863-904, :984-1035) as
// written -- flux-space Gaussian noise, NaN where the noisy flux is not positive ("BAD_MAG");
// z injected or the star's z_r (OBS block words 0, 1)
__global__ void k_noisy_mag(const double* mag, const double* err, const double* z_in, double* out,
                            unsigned long long n, unsigned long long offset,
                            const ss::PhiloxKeys keys) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    double z;
    if (z_in) {
      z = z_in[i];
    } else {
      const uint4 w = ss::philox_block(offset + i, SS_DRAW_BLOCK_OBS, keys);
      double z1;
      ss::bm32(w.x, w.y, z, z1);
    }
    const double flux = __dmul_rn(3631.0, exp10(__dmul_rn(-0.4, mag[i])));
    const double sig = __ddiv_rn(__dmul_rn(flux, err[i]), 1.0857362);
    const double fobs = __dadd_rn(flux, __dmul_rn(sig, z));
    out[i] = (fobs > 0.0) ? __dmul_rn(-2.5, log10(__ddiv_rn(fobs, 3631.0))) : ss::nan_d();
  }
}

// Bernoulli trial u <= prob (reference observed.py:906-959); u injected or the star's u_det
// (OBS block word 2)
__global__ void k_bernoulli(const double* prob, const double* u_in, uint8_t* out,
                            unsigned long long n, unsigned long long offset,
                            const ss::PhiloxKeys keys) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    const double u = u_in ? u_in[i]
                          : ss::u32(ss::philox_block(offset + i, SS_DRAW_BLOCK_OBS, keys).z);
    out[i] = (u <= prob[i]) ? 1 : 0;
  }
}

// ---------------------------------------------------------------- host side
// Philox4x32 key schedule of a 64-bit seed: round r uses key + r * (0x9E3779B9, 0xBB67AE85)
ss::PhiloxKeys philox_keys(unsigned long long seed) {
  ss::PhiloxKeys K;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    K.rk[2 * r] = k0;
    K.rk[2 * r + 1] = k1;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  for (uint32_t b = 0; b < 2; ++b) {
    const unsigned long long m = 0xCD9E8D57ULL * b;
    K.fold[b][0] = K.rk[0] ^ (uint32_t)(m >> 32);
    K.fold[b][1] = K.rk[2] ^ (uint32_t)m;
  }
  return K;
}

struct DevInfo {
  int device = -1, sms = 0, smem_optin = 0;
};

int device_info(DevInfo& d) {
  SS_CUDA(cudaGetDevice(&d.device));
  static thread_local DevInfo cache;
  if (cache.device == d.device) { d = cache; return SS_OK; }
  int major = 0;
  SS_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d.device));
  if (major != 10) return fail(SS_ENODEV, "device is not sm_100 (Blackwell B200)%s");
  SS_CUDA(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, d.device));
  SS_CUDA(cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, d.device));
  cache = d;
  return SS_OK;
}

// static shared memory of the star kernels (barrier + counters + histograms), rounded up
constexpr int kStaticSmem = 8 + 4 * SS_N_COUNTERS + 4 * SS_N_HIST * SS_HIST_BINS + 256;

int plan_launch(const DevInfo& d, int64_t blob_bytes, uint64_t n, int& smem, int& grid) {
  smem = (blob_bytes > 0 && blob_bytes + kStaticSmem <= d.smem_optin) ? (int)blob_bytes : 0;
  const uint64_t blocks_needed = (n + SS_BLOCK - 1) / SS_BLOCK;
  // persistent CTAs, one per SM: a 1024-thread CTA at 64 registers per thread owns the
  // whole register file
  const uint64_t cap = (uint64_t)d.sms;
  grid = (int)(blocks_needed < cap ? (blocks_needed ? blocks_needed : 1) : cap);
  return SS_OK;
}

// the opt-in dynamic shared-memory limit is a per-device, per-kernel attribute: remember
// the largest size configured for each (kernel, device)
template <typename K>
int launch_kernel(K kern, const KArgs& a, int smem, int grid, int block, cudaStream_t stream) {
  static thread_local std::map<std::pair<const void*, int>, int> configured;
  int dev = 0;
  SS_CUDA(cudaGetDevice(&dev));
  int& have = configured.emplace(std::make_pair((const void*)kern, dev), -1).first->second;
  if (smem > have) {
    SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // experiment knob: share of the unified L1 / shared memory to reserve as shared memory
    // (percent; unset = the driver's choice)
    if (const char* c = getenv("SS_SMEM_CARVEOUT"))
      SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(c)));
    have = smem;
  }
  kern<<<grid, block, smem, stream>>>(a);
  SS_CUDA(cudaGetLastError());
  return SS_OK;
}

template <uint32_t ST, typename OutT, int MATH = SS_MATH_F64>
int launch_t(const KArgs& a, int smem, int grid, cudaStream_t stream) {
  return launch_kernel(k_stars<ST, OutT, MATH>, a, smem, grid, SS_BLOCK, stream);
}

// the production kernel's preconditions (see k_fused)
bool fused_eligible(const KArgs& a, int smem) {
  const SSParams& p = a.p;
  const SSDraws& d = a.d;
  return p.fused_kernel && smem > 0 && p.density_kind == SS_DENSITY_GRID && !p.nest && p.has_dist &&
         p.has_iso && p.iso_direct && p.band_present[0] && p.band_present[1] &&
         !d.u_phi1 && !d.u_mass && !d.n_phi2 && !d.n_dist && !d.z_flux[0] && !d.z_flux[1] &&
         !d.u_det && !d.u_det2 && !a.in.phi1 && !a.in.phi2 && !a.in.dist && !a.o.draws_out &&
         p.photo_error.q_n > 0 && p.completeness.q_n > 0 &&
         (!p.perfect_galstarsep || p.detection.q_n > 0) && a.n <= (1ULL << 31) &&
         p.grid_wlut_n > 0 && p.iso_wlut_n > 0 && a.t.grid_wlut && a.t.grid_wthr && a.t.iso_wlut &&
         a.t.iso_wthr;
}

// dynamic shared memory of k_fused: the hot blob, three 16-byte staging slots per thread (kFusedStage)
int fused_smem(const KArgs& a, int blob_smem) {
  return blob_smem + kFusedStage;
}

template <typename OutT, int MATH, bool HIST>
int launch_fused_p(const KArgs& a, int smem, int grid, cudaStream_t s) {
  const SSParams& p = a.p;
  const bool packed = (a.m.packed || a.m.packed32) && a.m.nside_packed == p.nside_ebv &&
                      p.nside_maglim[0] == p.nside_ebv && p.nside_maglim[1] == p.nside_ebv;
  const int total = fused_smem(a, smem);
  const SSOut& o = a.o;
  const bool all = o.phi1 && o.phi2 && o.dist && o.mag[0] && o.mag[1] && o.ra && o.dec && o.mag_obs[0] &&
                   o.mag_obs[1] && o.magerr[0] && o.magerr[1] && o.flag_observed;
  if (packed && all)
    return launch_kernel(k_fused<OutT, MATH, HIST, true, true>, a, total, grid, SS_FUSED_BLOCK, s);
  return packed ? launch_kernel(k_fused<OutT, MATH, HIST, true, false>, a, total, grid, SS_FUSED_BLOCK, s)
                : launch_kernel(k_fused<OutT, MATH, HIST, false, false>, a, total, grid, SS_FUSED_BLOCK, s);
}

template <typename OutT>
int launch_fused(const KArgs& a, const DevInfo& d, int smem, cudaStream_t s) {
  const uint64_t blocks_needed = (a.n + SS_FUSED_BLOCK - 1) / SS_FUSED_BLOCK;
  const uint64_t cap = (uint64_t)d.sms * SS_FUSED_MIN_CTAS;
  const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
  const bool hist = a.o.counters && a.p.hist_enable;
  if (a.p.math_mode == SS_MATH_FAST)
    return hist ? launch_fused_p<OutT, SS_MATH_FAST, true>(a, smem, grid, s)
                : launch_fused_p<OutT, SS_MATH_FAST, false>(a, smem, grid, s);
  return hist ? launch_fused_p<OutT, SS_MATH_F64, true>(a, smem, grid, s)
              : launch_fused_p<OutT, SS_MATH_F64, false>(a, smem, grid, s);
}

template <typename OutT>
int launch_stage(uint32_t st, const KArgs& a, const DevInfo& d, int smem, int grid, cudaStream_t s) {
  const bool fast = a.p.math_mode == SS_MATH_FAST;
#ifdef SS_DEV_BUILD
  if (st != 7) return fail(SS_EINVAL, "development build: fused stage mask only%s");
  {
#else
  switch (st) {
    case 1: return launch_t<1, OutT>(a, smem, grid, s);
    case 2: return launch_t<2, OutT>(a, smem, grid, s);
    case 3: return launch_t<3, OutT>(a, smem, grid, s);
    case 4: return fast ? launch_t<4, OutT, SS_MATH_FAST>(a, smem, grid, s) : launch_t<4, OutT>(a, smem, grid, s);
    case 6: return fast ? launch_t<6, OutT, SS_MATH_FAST>(a, smem, grid, s) : launch_t<6, OutT>(a, smem, grid, s);
    case 7:
#endif
      if (fused_eligible(a, smem) &&
          SS_FUSED_MIN_CTAS * (fused_smem(a, smem) + kStaticSmem + 1024) <= d.smem_optin)
        return launch_fused<OutT>(a, d, smem, s);
      return fast ? launch_t<7, OutT, SS_MATH_FAST>(a, smem, grid, s) : launch_t<7, OutT>(a, smem, grid, s);
  }
  return fail(SS_EINVAL, "unsupported stage mask%s");
}

int check_func(const SSFunc& f, int64_t blob_doubles, const char* name) {
  if (f.kind == SS_FN_LINEAR || f.kind == SS_FN_SPLINE || f.kind == SS_FN_SPLINE_PRODUCT) {
    const int64_t need_c = f.kind == SS_FN_LINEAR ? 2LL * f.n : 4LL * (f.n - 1);
    if (f.n < 2 || f.x_off < 0 || f.c_off < 0 || f.x_off + (int64_t)f.n > blob_doubles ||
        f.c_off + need_c > blob_doubles)
      return fail(SS_EINVAL, "function table '%s' lies outside the blob", name);
    if (f.q_n != 0 &&
        (f.q_n < 1 || f.q_nseg < 3 || f.q_off < 0 || f.q_seg_off < 0 || (f.q_off & 1) || (f.q_seg_off & 1) ||
         f.q_off + 4LL * (f.q_n + 2) > blob_doubles || f.q_seg_off + 4LL * f.q_nseg > blob_doubles))
      return fail(SS_EINVAL, "direct look-up table of '%s' lies outside the blob", name);
  } else if (f.kind < SS_FN_NONE || f.kind > SS_FN_SPLINE_PRODUCT) {
    return fail(SS_EINVAL, "unknown function kind in '%s'", name);
  }
  return SS_OK;
}

int validate(uint32_t st, const SSParams* p, const SSTables* t, const SSMaps* m,
             const SSCatalog* in, const SSOut* out) {
  if (!p || !out) return fail(SS_EINVAL, "params and out must not be NULL%s");
  if (out->dtype != SS_DTYPE_F64 && out->dtype != SS_DTYPE_F32)
    return fail(SS_EINVAL, "SSOut.dtype must be SS_DTYPE_F64 or SS_DTYPE_F32%s");
  if (p->math_mode != SS_MATH_F64 && p->math_mode != SS_MATH_FAST)
    return fail(SS_EINVAL, "SSParams.math_mode must be SS_MATH_F64 or SS_MATH_FAST%s");
  const int64_t nd = t ? t->blob_bytes / 8 : 0;
  if (t && (t->blob_bytes % 16 != 0)) return fail(SS_EINVAL, "blob_bytes must be a multiple of 16%s");
  if (t && (t->blob_hot_bytes % 16 != 0 || t->blob_hot_bytes < 0 || t->blob_hot_bytes > t->blob_bytes))
    return fail(SS_EINVAL, "blob_hot_bytes must be a multiple of 16 within the blob%s");
  if (st & SS_STAGE_GENERATE) {
    if (!t) return fail(SS_EINVAL, "generate needs tables%s");
    int rc;
    if ((rc = check_func(p->track_center, nd, "track.center"))) return rc;
    if ((rc = check_func(p->track_spread, nd, "track.spread"))) return rc;
    if (p->track_center.kind == SS_FN_NONE || p->track_spread.kind == SS_FN_NONE)
      return fail(SS_EINVAL, "track center/spread are required%s");
    if (p->has_dist) {
      if ((rc = check_func(p->dist_center, nd, "distance_modulus.center"))) return rc;
      if ((rc = check_func(p->dist_spread, nd, "distance_modulus.spread"))) return rc;
    }
    if (p->density_kind == SS_DENSITY_GRID) {
      if (!t->grid_rows || !t->grid_lut || !t->grid_thr || p->grid_nthr < 0 || p->grid_lut_n < 1)
        return fail(SS_EINVAL, "grid density needs grid_rows, grid_lut and grid_thr%s");
      if ((uintptr_t)t->grid_rows & 31) return fail(SS_EINVAL, "grid_rows must be 32-byte aligned%s");
      // u * buckets must be exact for bucket = floor(u * buckets) to agree with the host's edges
      if (p->grid_lut_n & (p->grid_lut_n - 1)) return fail(SS_EINVAL, "grid_lut_n must be a power of two%s");
      if (p->grid_wlut_n < 0 || (p->grid_wlut_n & (p->grid_wlut_n - 1)) || p->grid_wlut_n == 1 ||
          (p->grid_wlut_n && (!t->grid_wlut || !t->grid_wthr || ((uintptr_t)t->grid_wlut & 15))))
        return fail(SS_EINVAL, "grid_wlut_n must be 0 or a power of two >= 2 with 16-byte aligned grid_wlut, grid_wthr%s");
    }
    if (p->density_kind == SS_DENSITY_INVCDF) {
      if (!t->cdf_pairs || p->cdf_n < 2) return fail(SS_EINVAL, "inverse-CDF density needs cdf_pairs%s");
      if (!p->cdf_identity && !t->cdf_perm) return fail(SS_EINVAL, "non-monotone CDF needs cdf_perm%s");
      if (p->cdf_guide_n & (p->cdf_guide_n - 1)) return fail(SS_EINVAL, "cdf_guide_n must be a power of two%s");
    }
    if (p->has_iso) {
      if (!p->has_dist && !(in && in->dist))
        return fail(SS_EINVAL, "isochrone magnitudes need a distance modulus%s");
      if (p->iso_direct) {
        if (!t->iso_rows || !t->iso_lut || !t->iso_thr || p->iso_n < 2 || p->iso_lut_n < 1)
          return fail(SS_EINVAL, "direct isochrone lookup needs iso_rows, iso_lut, iso_thr%s");
        if ((uintptr_t)t->iso_rows & 31) return fail(SS_EINVAL, "iso_rows must be 32-byte aligned%s");
        if (p->iso_lut_n & (p->iso_lut_n - 1)) return fail(SS_EINVAL, "iso_lut_n must be a power of two%s");
        if (p->iso_wlut_n < 0 || (p->iso_wlut_n & (p->iso_wlut_n - 1)) || p->iso_wlut_n == 1 ||
            (p->iso_wlut_n && (!t->iso_wlut || !t->iso_wthr || ((uintptr_t)t->iso_wlut & 15))))
          return fail(SS_EINVAL, "iso_wlut_n must be 0 or a power of two >= 2 with 16-byte aligned iso_wlut, iso_wthr%s");
      } else {
        if (p->iso_n < 2 || p->imf_n < 2) return fail(SS_EINVAL, "isochrone/IMF tables too short%s");
        if (p->imf_guide_n & (p->imf_guide_n - 1))
          return fail(SS_EINVAL, "imf_guide_n must be a power of two%s");
      }
    }
  }
  if ((st & SS_STAGE_ROTATE) && !(st & SS_STAGE_GENERATE)) {
    if (!in || !in->phi1 || !in->phi2) return fail(SS_EINVAL, "rotate needs catalog phi1/phi2%s");
  }
  if (st & SS_STAGE_OBSERVE) {
    if (!t || !m) return fail(SS_EINVAL, "observe needs tables and maps%s");
    if (!p->band_present[SS_BAND_R])
      return fail(SS_EINVAL, "band 'r' is required (reference observed.py:256)%s");
    if (!m->ebv) return fail(SS_EINVAL, "E(B-V) map not loaded%s");
    if (m->packed && ((uintptr_t)m->packed & 31))
      return fail(SS_EINVAL, "SSMaps.packed must be 32-byte aligned%s");
    if (m->packed32 && ((uintptr_t)m->packed32 & 15))
      return fail(SS_EINVAL, "SSMaps.packed32 must be 16-byte aligned%s");
    for (int b = 0; b < 2; ++b)
      if (p->band_present[b] && !m->maglim[b]) return fail(SS_EINVAL, "maglim map missing for a requested band%s");
    if (!p->band_present[p->completeness_band & 1])
      return fail(SS_EINVAL, "the completeness band must be in bands (reference observed.py:246-252)%s");
    int rc;
    if ((rc = check_func(p->photo_error, nd, "photo_error"))) return rc;
    if ((rc = check_func(p->completeness, nd, "completeness"))) return rc;
    if (p->photo_error.kind != SS_FN_LINEAR || p->completeness.kind != SS_FN_LINEAR)
      return fail(SS_EINVAL, "photo_error/completeness must be SS_FN_LINEAR tables%s");
    if (p->perfect_galstarsep) {
      if ((rc = check_func(p->detection, nd, "detection"))) return rc;
      if (p->detection.kind != SS_FN_LINEAR) return fail(SS_EINVAL, "detection table missing%s");
    }
    if (p->nest) {
      const int ns[3] = {p->nside_ebv, p->nside_maglim[0], p->nside_maglim[1]};
      for (int k = 0; k < 3; ++k)
        if (ns[k] & (ns[k] - 1)) return fail(SS_EINVAL, "NESTED maps need power-of-two nside%s");
    }
    if (!(st & SS_STAGE_ROTATE) && (!in || !in->ra || !in->dec))
      return fail(SS_EINVAL, "observe needs catalog ra/dec%s");
    if (!(st & SS_STAGE_GENERATE)) {
      if (!in) return fail(SS_EINVAL, "observe needs catalog magnitudes%s");
      for (int b = 0; b < 2; ++b)
        if (p->band_present[b] && !in->mag[b]) return fail(SS_EINVAL, "catalog magnitude column missing%s");
    } else if (!p->has_iso) {
      return fail(SS_EINVAL, "generate+observe needs an isochrone%s");
    }
  }
  return SS_OK;
}

}  // namespace

// ================================================================ C ABI
extern "C" {

int ss_abi_version(void) { return SS_ABI_VERSION; }
const char* ss_last_error(void) { return g_err; }

int ss_struct_sizes(int32_t* s) {
  if (!s) return fail(SS_EINVAL, "sizes7 is NULL%s");
  s[0] = sizeof(SSFunc); s[1] = sizeof(SSParams); s[2] = sizeof(SSTables); s[3] = sizeof(SSMaps);
  s[4] = sizeof(SSCatalog); s[5] = sizeof(SSDraws); s[6] = sizeof(SSOut);
  return SS_OK;
}

int ss_launch_info(int64_t blob_bytes, uint64_t n, int32_t* smem_bytes, int32_t* grid, int32_t* block) {
  DevInfo d;
  int rc = device_info(d);
  if (rc) return rc;
  int smem, g;
  plan_launch(d, blob_bytes, n, smem, g);
  if (smem_bytes) *smem_bytes = smem;
  if (grid) *grid = g;
  if (block) *block = SS_BLOCK;
  return SS_OK;
}

static size_t elem_size(int32_t dtype);

int ss_run(uint32_t stages, const SSParams* p, const SSTables* t, const SSMaps* m,
           const SSCatalog* in, const SSDraws* d, const SSOut* out, uint64_t n,
           uint64_t star_offset, uint64_t seed, void* stream) {
  if (n == 0) return SS_OK;   // nothing to do (empty arrays may legitimately be NULL)
  int rc = validate(stages, p, t, m, in, out);
  if (rc) return rc;
  // The production kernel indexes stars with 32 bits inside a launch.  A longer free-running
  // range is cut into 2^30-star launches: draws are keyed by the global index, so the result
  // is the same bytes (columns advance with the range).
  const uint64_t kSplit = 1ULL << 30;
  static const SSDraws no_draws = {};
  static const SSCatalog no_cat = {};
  if (stages == (SS_STAGE_GENERATE | SS_STAGE_ROTATE | SS_STAGE_OBSERVE) && n > (1ULL << 31) &&
      p->fused_kernel && (!d || !memcmp(d, &no_draws, sizeof(SSDraws))) &&
      (!in || !memcmp(in, &no_cat, sizeof(SSCatalog))) && !out->draws_out) {
    const size_t es = elem_size(out->dtype);
    for (uint64_t lo = 0; lo < n; lo += kSplit) {
      SSOut o = *out;
      void** cols[11] = {&o.phi1, &o.phi2, &o.dist, &o.mag[0], &o.mag[1], &o.ra, &o.dec,
                         &o.mag_obs[0], &o.mag_obs[1], &o.magerr[0], &o.magerr[1]};
      for (void** c : cols)
        if (*c) *c = (char*)*c + lo * es;
      if (o.flag_observed) o.flag_observed += lo;
      if (o.flag_perfect) o.flag_perfect += lo;
      if (o.pix) o.pix += lo;
      const uint64_t cn = (n - lo < kSplit) ? n - lo : kSplit;
      if ((rc = ss_run(stages, p, t, m, nullptr, nullptr, &o, cn, star_offset + lo, seed, stream)))
        return rc;
    }
    return SS_OK;
  }
  DevInfo dev;
  if ((rc = device_info(dev))) return rc;
  KArgs a;
  memset(&a, 0, sizeof(a));
  a.p = *p;
  if (t) a.t = *t;
  if (m) a.m = *m;
  if (in) a.in = *in;
  if (d) a.d = *d;
  a.o = *out;
  a.n = n;
  a.offset = star_offset;
  a.seed = seed;
  a.keys = philox_keys(seed);
  for (int h = 0; h < SS_N_HIST; ++h) {
    a.hist_lo[h] = (float)p->hist_lo[h];
    a.hist_inv[h] = (float)p->hist_inv_width[h];
  }
  {
    const void* cols[11] = {out->phi1, out->phi2, out->dist, out->mag[0], out->mag[1], out->ra, out->dec,
                            out->mag_obs[0], out->mag_obs[1], out->magerr[0], out->magerr[1]};
    uintptr_t bits = 0;
    for (const void* c : cols) bits |= (uintptr_t)c;
    a.vec_ok = (bits & 15) == 0;
  }
  for (int b = p->grid_wlut_n; b > 1; b >>= 1) a.grid_wshift++;    // log2 buckets ...
  for (int b = p->iso_wlut_n; b > 1; b >>= 1) a.iso_wshift++;
  a.grid_wshift = 32 - a.grid_wshift;                              // ... -> bucket = w >> shift
  a.iso_wshift = 32 - a.iso_wshift;
  int smem, grid;
  if (a.t.blob_hot_bytes <= 0 || a.t.blob_hot_bytes > a.t.blob_bytes)
    a.t.blob_hot_bytes = a.t.blob_bytes;
  plan_launch(dev, a.t.blob_hot_bytes, n, smem, grid);
  a.blob_in_smem = smem > 0;
  a.blob_bytes = (uint32_t)a.t.blob_hot_bytes;                 // what stage_blob copies
  if (!a.blob_in_smem) a.t.blob_hot_bytes = a.t.blob_bytes;   // everything read from global
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
#ifdef SS_DEV_BUILD   // quick experiments: the double-output kernels only
  return launch_stage<double>(stages, a, dev, smem, grid, s);
#else
  return out->dtype == SS_DTYPE_F32 ? launch_stage<float>(stages, a, dev, smem, grid, s)
                                    : launch_stage<double>(stages, a, dev, smem, grid, s);
#endif
}

int ss_generate(const SSParams* p, const SSTables* t, const SSDraws* d, const SSOut* out,
                uint64_t n, uint64_t star_offset, uint64_t seed, void* stream) {
  return ss_run(SS_STAGE_GENERATE, p, t, nullptr, nullptr, d, out, n, star_offset, seed, stream);
}

int ss_rotate(const SSParams* p, const SSCatalog* in, const SSOut* out, uint64_t n, void* stream) {
  return ss_run(SS_STAGE_ROTATE, p, nullptr, nullptr, in, nullptr, out, n, 0, 0, stream);
}

int ss_observe(const SSParams* p, const SSTables* t, const SSMaps* m, const SSCatalog* in,
               const SSDraws* d, const SSOut* out, uint64_t n, uint64_t star_offset,
               uint64_t seed, void* stream) {
  const uint32_t st = (in && in->ra && in->dec) ? SS_STAGE_OBSERVE
                                                : (SS_STAGE_ROTATE | SS_STAGE_OBSERVE);
  return ss_run(st, p, t, m, in, d, out, n, star_offset, seed, stream);
}

int ss_generate_observe(const SSParams* p, const SSTables* t, const SSMaps* m, const SSDraws* d,
                        const SSOut* out, uint64_t n, uint64_t star_offset, uint64_t seed,
                        void* stream) {
  return ss_run(SS_STAGE_GENERATE | SS_STAGE_ROTATE | SS_STAGE_OBSERVE, p, t, m, nullptr, d, out,
                n, star_offset, seed, stream);
}

static int small_grid(uint64_t n, int* grid) {
  DevInfo d;
  int rc = device_info(d);
  if (rc) return rc;
  const uint64_t need = (n + 255) / 256;
  const uint64_t cap = (uint64_t)d.sms * 8;
  *grid = (int)(need < cap ? (need ? need : 1) : cap);
  return SS_OK;
}

int ss_frame_fraction(const SSParams* p, const SSMaps* m, const SSCatalog* in,
                      unsigned long long* counters, uint64_t n, void* stream) {
  if (!p || !m || !in || !counters || !m->mask || !in->phi1 || !in->phi2 || m->nside_mask < 1)
    return fail(SS_EINVAL, "ss_frame_fraction: params, mask, phi1, phi2 and counters are required%s");
  if (n == 0) return SS_OK;
  int grid, rc;
  if ((rc = small_grid(n, &grid))) return rc;
  k_frame_fraction<<<grid, 256, 0, (cudaStream_t)stream>>>(*p, *m, *in, counters, n);
  SS_CUDA(cudaGetLastError());
  return SS_OK;
}

int ss_photo_error(const SSParams* p, const SSTables* t, int32_t band, const double* mag,
                   const double* maglim, double* out, uint64_t n, void* stream) {
  if (!p || !t || !t->blob || !mag || !maglim || !out || band < 0 || band > 1)
    return fail(SS_EINVAL, "ss_photo_error: bad argument%s");
  int rc = check_func(p->photo_error, t->blob_bytes / 8, "photo_error");
  if (rc) return rc;
  if (p->photo_error.kind != SS_FN_LINEAR) return fail(SS_EINVAL, "Photo error model not loaded%s");
  if (n == 0) return SS_OK;
  int grid;
  if ((rc = small_grid(n, &grid))) return rc;
  k_photo_error<<<grid, 256, 0, (cudaStream_t)stream>>>(*p, (const double*)t->blob, band, mag,
                                                        maglim, out, n);
  SS_CUDA(cudaGetLastError());
  return SS_OK;
}

int ss_efficiency(const SSParams* p, const SSTables* t, int32_t band, int32_t detection_only,
                  const double* mag, const double* maglim, double* out, uint64_t n, void* stream) {
  if (!p || !t || !t->blob || !mag || !maglim || !out || band < 0 || band > 1)
    return fail(SS_EINVAL, "ss_efficiency: bad argument%s");
  const SSFunc& f = detection_only ? p->detection : p->completeness;
  int rc = check_func(f, t->blob_bytes / 8, "efficiency");
  if (rc) return rc;
  if (f.kind != SS_FN_LINEAR) return fail(SS_EINVAL, "Efficiency function not loaded%s");
  if (n == 0) return SS_OK;
  int grid;
  if ((rc = small_grid(n, &grid))) return rc;
  k_efficiency<<<grid, 256, 0, (cudaStream_t)stream>>>(*p, (const double*)t->blob, band,
                                                       detection_only, mag, maglim, out, n);
  SS_CUDA(cudaGetLastError());
  return SS_OK;
}

int ss_eval_function(const SSFunc* f, const SSTables* t, const double* x, double* out, uint64_t n,
                     void* stream) {
  if (!f || !x || !out) return fail(SS_EINVAL, "ss_eval_function: bad argument%s");
  const bool tabular = f->kind >= SS_FN_LINEAR;
  if (tabular && (!t || !t->blob)) return fail(SS_EINVAL, "ss_eval_function: table blob required%s");
  int rc = check_func(*f, t ? t->blob_bytes / 8 : 0, "function");
  if (rc) return rc;
  if (n == 0) return SS_OK;
  int grid;
  if ((rc = small_grid(n, &grid))) return rc;
  k_eval_function<<<grid, 256, 0, (cudaStream_t)stream>>>(*f, t ? (const double*)t->blob : nullptr,
                                                          x, out, n);
  SS_CUDA(cudaGetLastError());
  return SS_OK;
}

int ss_ang2pix(int32_t nside, int32_t nest, const double* lon, const double* lat, int64_t* pix,
               uint64_t n, void* stream) {
  if (nside < 1 || !lon || !lat || !pix) return fail(SS_EINVAL, "ss_ang2pix: bad argument%s");
  if (nest && (nside & (nside - 1))) return fail(SS_EINVAL, "NESTED needs power-of-two nside%s");
  if (n == 0) return SS_OK;
  int grid, rc;
  if ((rc = small_grid(n, &grid))) return rc;
  k_ang2pix<<<grid, 256, 0, (cudaStream_t)stream>>>(nside, nest, lon, lat, (long long*)pix, n);
  SS_CUDA(cudaGetLastError());
  return SS_OK;
}

int ss_math_probe(int32_t fn, const double* x, const double* y2, double* out, uint64_t n,
                  void* stream) {
  if (fn < SS_MATH_FN_LOG10 || fn > SS_MATH_FN_SQRT || !x || !out || (fn == SS_MATH_FN_ATAN2 && !y2))
    return fail(SS_EINVAL, "ss_math_probe: bad argument%s");
  if (n == 0) return SS_OK;
  int grid, rc;
  if ((rc = small_grid(n, &grid))) return rc;
  k_math_probe<<<grid, 256, 0, (cudaStream_t)stream>>>(fn, x, y2, out, n);
  SS_CUDA(cudaGetLastError());
  return SS_OK;
}

int ss_loc_scale(int32_t normal, const double* loc, const double* scale, double loc0,
                 double scale0, const double* draws, double* out, uint64_t n, uint64_t star_offset,
                 uint64_t seed, void* stream) {
  if (!out) return fail(SS_EINVAL, "ss_loc_scale: out is NULL%s");
  if (n == 0) return SS_OK;
  int grid, rc;
  if ((rc = small_grid(n, &grid))) return rc;
  k_loc_scale<<<grid, 256, 0, (cudaStream_t)stream>>>(normal, loc, scale, loc0, scale0, draws, out,
                                                      n, star_offset, philox_keys(seed));
  SS_CUDA(cudaGetLastError());
  return SS_OK;
}

int ss_noisy_magnitudes(const double* mag, const double* err, const double* z, double* out,
                        uint64_t n, uint64_t star_offset, uint64_t seed, void* stream) {
  if (!mag || !err || !out) return fail(SS_EINVAL, "ss_noisy_magnitudes: bad argument%s");
  if (n == 0) return SS_OK;
  int grid, rc;
  if ((rc = small_grid(n, &grid))) return rc;
  k_noisy_mag<<<grid, 256, 0, (cudaStream_t)stream>>>(mag, err, z, out, n, star_offset,
                                                      philox_keys(seed));
  SS_CUDA(cudaGetLastError());
  return SS_OK;
}

int ss_bernoulli(const double* prob, const double* u, uint8_t* out, uint64_t n,
                 uint64_t star_offset, uint64_t seed, void* stream) {
  if (!prob || !out) return fail(SS_EINVAL, "ss_bernoulli: bad argument%s");
  if (n == 0) return SS_OK;
  int grid, rc;
  if ((rc = small_grid(n, &grid))) return rc;
  k_bernoulli<<<grid, 256, 0, (cudaStream_t)stream>>>(prob, u, out, n, star_offset,
                                                      philox_keys(seed));
  SS_CUDA(cudaGetLastError());
  return SS_OK;
}

// ---------------------------------------------------------------- host-buffer variants
namespace {
struct ColSpec { size_t off_out; int is_float; };  // float-typed column (dtype) or fixed type
}

static size_t elem_size(int32_t dtype) { return dtype == SS_DTYPE_F32 ? 4 : 8; }
static uint64_t align256(uint64_t v) { return (v + 255) & ~255ULL; }

// workspace layout per buffer (two buffers): 11 float columns, 2 flag columns, pix
int64_t ss_host_workspace_bytes(uint64_t chunk, int32_t dtype) {
  const uint64_t per = 11 * align256(chunk * elem_size(dtype)) + 2 * align256(chunk) +
                       align256(chunk * 8) + 4 * align256(chunk * 8);
  return (int64_t)(2 * per + align256(SS_COUNTERS_LEN * 8));
}

// two private streams (kernel / copies) and their events, created once per host thread and
// device and kept for the life of the thread
struct HostPipe {
  int device = -1;
  cudaStream_t compute = nullptr, copy = nullptr;
  cudaEvent_t done[2] = {nullptr, nullptr}, drained[2] = {nullptr, nullptr}, loaded[2] = {nullptr, nullptr};
  int init(int dev) {
    if (device == dev) return SS_OK;
    release();
    SS_CUDA(cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking));
    SS_CUDA(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      SS_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
      SS_CUDA(cudaEventCreateWithFlags(&drained[i], cudaEventDisableTiming));
      SS_CUDA(cudaEventCreateWithFlags(&loaded[i], cudaEventDisableTiming));
    }
    device = dev;
    return SS_OK;
  }
  void release() {
    for (int i = 0; i < 2; ++i) {
      if (done[i]) cudaEventDestroy(done[i]);
      if (drained[i]) cudaEventDestroy(drained[i]);
      if (loaded[i]) cudaEventDestroy(loaded[i]);
      done[i] = drained[i] = loaded[i] = nullptr;
    }
    if (compute) cudaStreamDestroy(compute);
    if (copy) cudaStreamDestroy(copy);
    compute = copy = nullptr;
    device = -1;
  }
  // no destructor: at thread / process exit the CUDA context may already be gone
};

static int host_pipeline_body(HostPipe& pipe, uint32_t stages, const SSParams* p, const SSTables* t,
                              const SSMaps* m, const SSCatalog* in_host, const SSOut* oh, uint64_t n,
                              uint64_t star_offset, uint64_t seed, void* workspace, uint64_t chunk) {
  const size_t es = elem_size(oh->dtype);
  int rc;
  // carve the workspace
  char* base = (char*)workspace;
  uint64_t cur = 0;
  auto take = [&](uint64_t bytes) { char* q = base + cur; cur += align256(bytes); return (void*)q; };
  SSOut dev[2];
  SSCatalog cat[2];
  void* const* host_cols[11] = {&oh->phi1, &oh->phi2, &oh->dist, &oh->mag[0], &oh->mag[1], &oh->ra,
                                &oh->dec, &oh->mag_obs[0], &oh->mag_obs[1], &oh->magerr[0],
                                &oh->magerr[1]};
  for (int b = 0; b < 2; ++b) {
    memset(&dev[b], 0, sizeof(SSOut));
    memset(&cat[b], 0, sizeof(SSCatalog));
    dev[b].dtype = oh->dtype;
    void** dev_cols[11] = {&dev[b].phi1, &dev[b].phi2, &dev[b].dist, &dev[b].mag[0], &dev[b].mag[1],
                           &dev[b].ra, &dev[b].dec, &dev[b].mag_obs[0], &dev[b].mag_obs[1],
                           &dev[b].magerr[0], &dev[b].magerr[1]};
    for (int c = 0; c < 11; ++c) {
      void* q = take(chunk * es);
      *dev_cols[c] = *host_cols[c] ? q : nullptr;
    }
    void* q = take(chunk); dev[b].flag_observed = oh->flag_observed ? (uint8_t*)q : nullptr;
    q = take(chunk);       dev[b].flag_perfect = oh->flag_perfect ? (uint8_t*)q : nullptr;
    q = take(chunk * 8);   dev[b].pix = oh->pix ? (int64_t*)q : nullptr;
    // catalog staging (observe-only): ra, dec, mag_g, mag_r
    cat[b].ra = (const double*)take(chunk * 8);
    cat[b].dec = (const double*)take(chunk * 8);
    cat[b].mag[0] = (const double*)take(chunk * 8);
    cat[b].mag[1] = (const double*)take(chunk * 8);
  }
  unsigned long long* dcnt = (unsigned long long*)take(SS_COUNTERS_LEN * 8);
  if (oh->counters) SS_CUDA(cudaMemsetAsync(dcnt, 0, SS_COUNTERS_LEN * 8, pipe.compute));

  const bool has_cat = in_host != nullptr;
  uint64_t k = 0;
  for (uint64_t lo = 0; lo < n; lo += chunk, ++k) {
    const int b = (int)(k & 1);
    const uint64_t cn = (n - lo < chunk) ? (n - lo) : chunk;
    // buffer b must have been drained by the copy stream two iterations ago
    if (k >= 2) SS_CUDA(cudaStreamWaitEvent(pipe.compute, pipe.drained[b], 0));
    SSCatalog cin;
    memset(&cin, 0, sizeof(cin));
    if (has_cat) {
      // H2D of this chunk's catalog columns on the copy stream, ahead of the kernel
      if (k >= 2) SS_CUDA(cudaStreamWaitEvent(pipe.copy, pipe.done[b], 0));
      const double* hsrc[4] = {in_host->ra, in_host->dec, in_host->mag[0], in_host->mag[1]};
      const double* ddst[4] = {cat[b].ra, cat[b].dec, cat[b].mag[0], cat[b].mag[1]};
      const double** slot[4] = {&cin.ra, &cin.dec, &cin.mag[0], &cin.mag[1]};
      for (int c = 0; c < 4; ++c) {
        if (!hsrc[c]) continue;
        SS_CUDA(cudaMemcpyAsync((void*)ddst[c], hsrc[c] + lo, cn * 8, cudaMemcpyHostToDevice, pipe.copy));
        *slot[c] = ddst[c];
      }
      SS_CUDA(cudaEventRecord(pipe.loaded[b], pipe.copy));
      SS_CUDA(cudaStreamWaitEvent(pipe.compute, pipe.loaded[b], 0));
    }
    SSOut o = dev[b];
    o.counters = oh->counters ? dcnt : nullptr;
    rc = ss_run(stages, p, t, m, has_cat ? &cin : nullptr, nullptr, &o, cn, star_offset + lo, seed,
                pipe.compute);
    if (rc) return rc;
    SS_CUDA(cudaEventRecord(pipe.done[b], pipe.compute));
    SS_CUDA(cudaStreamWaitEvent(pipe.copy, pipe.done[b], 0));
    void* dcols[11] = {o.phi1, o.phi2, o.dist, o.mag[0], o.mag[1], o.ra, o.dec, o.mag_obs[0],
                       o.mag_obs[1], o.magerr[0], o.magerr[1]};
    for (int c = 0; c < 11; ++c)
      if (dcols[c])
        SS_CUDA(cudaMemcpyAsync((char*)*host_cols[c] + lo * es, dcols[c], cn * es,
                                cudaMemcpyDeviceToHost, pipe.copy));
    if (o.flag_observed)
      SS_CUDA(cudaMemcpyAsync(oh->flag_observed + lo, o.flag_observed, cn, cudaMemcpyDeviceToHost, pipe.copy));
    if (o.flag_perfect)
      SS_CUDA(cudaMemcpyAsync(oh->flag_perfect + lo, o.flag_perfect, cn, cudaMemcpyDeviceToHost, pipe.copy));
    if (o.pix)
      SS_CUDA(cudaMemcpyAsync(oh->pix + lo, o.pix, cn * 8, cudaMemcpyDeviceToHost, pipe.copy));
    SS_CUDA(cudaEventRecord(pipe.drained[b], pipe.copy));
  }
  SS_CUDA(cudaStreamSynchronize(pipe.compute));
  if (oh->counters) {
    unsigned long long tmp[SS_COUNTERS_LEN];
    SS_CUDA(cudaMemcpyAsync(tmp, dcnt, sizeof(tmp), cudaMemcpyDeviceToHost, pipe.copy));
    SS_CUDA(cudaStreamSynchronize(pipe.copy));
    for (int i = 0; i < SS_COUNTERS_LEN; ++i) oh->counters[i] += tmp[i];
  }
  SS_CUDA(cudaStreamSynchronize(pipe.copy));
  return SS_OK;
}

static int host_pipeline(uint32_t stages, const SSParams* p, const SSTables* t, const SSMaps* m,
                         const SSCatalog* in_host, const SSOut* oh, uint64_t n,
                         uint64_t star_offset, uint64_t seed, void* workspace,
                         int64_t workspace_bytes, uint64_t chunk) {
  if (!oh || !workspace) return fail(SS_EINVAL, "host variant: out and workspace are required%s");
  if (chunk == 0) return fail(SS_EINVAL, "chunk must be > 0%s");
  if (workspace_bytes < ss_host_workspace_bytes(chunk, oh->dtype))
    return fail(SS_EINVAL, "workspace too small (see ss_host_workspace_bytes)%s");
  static thread_local HostPipe pipe;
  int dev = 0;
  SS_CUDA(cudaGetDevice(&dev));
  int rc = pipe.init(dev);
  if (rc) return rc;
  // The call is synchronous and works on private streams: everything the caller queued
  // before it (table / map uploads, the allocation that produced the workspace, earlier
  // readers of the host buffers) must have finished first.
  SS_CUDA(cudaDeviceSynchronize());
  rc = host_pipeline_body(pipe, stages, p, t, m, in_host, oh, n, star_offset, seed, workspace, chunk);
  if (rc) {
    // copies into the caller's buffers may still be in flight: let them land (or fail)
    // before the caller can free anything; keep the first error's message
    char keep[sizeof(g_err)];
    memcpy(keep, g_err, sizeof(keep));
    cudaStreamSynchronize(pipe.compute);
    cudaStreamSynchronize(pipe.copy);
    memcpy(g_err, keep, sizeof(keep));
  }
  return rc;
}

int ss_generate_observe_host(const SSParams* p, const SSTables* t, const SSMaps* m,
                             const SSOut* out_host, uint64_t n, uint64_t star_offset,
                             uint64_t seed, void* workspace, int64_t workspace_bytes,
                             uint64_t chunk) {
  return host_pipeline(SS_STAGE_GENERATE | SS_STAGE_ROTATE | SS_STAGE_OBSERVE, p, t, m, nullptr,
                       out_host, n, star_offset, seed, workspace, workspace_bytes, chunk);
}

int ss_observe_host(const SSParams* p, const SSTables* t, const SSMaps* m,
                    const SSCatalog* in_host, const SSOut* out_host, uint64_t n,
                    uint64_t star_offset, uint64_t seed, void* workspace,
                    int64_t workspace_bytes, uint64_t chunk) {
  if (!in_host || !in_host->ra || !in_host->dec)
    return fail(SS_EINVAL, "ss_observe_host needs host ra/dec columns%s");
  return host_pipeline(SS_STAGE_OBSERVE, p, t, m, in_host, out_host, n, star_offset, seed,
                       workspace, workspace_bytes, chunk);
}

}  // extern "C"
