/* streamsim_b200.h -- C ABI of libstreamsim_b200.so
 *
 * B200 (sm_100a) implementation of the per-star generate-and-observe path of
 * LSSTDESC/stream_sim.  The reference has no FFI layer: its boundary is the
 * Python API itself.  These entry points are what the
 * bodies of the following reference methods bind to (see INTEGRATION.md for the
 * ctypes stubs a maintainer would add):
 *
 *   ss_generate          <- StreamModel.sample            stream_sim/model.py:106-157
 *   ss_rotate            <- StreamInjector.phi_to_radec   stream_sim/observed.py:478-575
 *   ss_observe           <- StreamInjector.inject (body)  stream_sim/observed.py:146-291
 *   ss_generate_observe  <- the three above fused, one pass, nothing materialised
 *   ss_frame_fraction    <- StreamInjector._find_gc_frame trial body
 *                           stream_sim/observed.py:661-668,700-718
 *   ss_photo_error       <- Survey.get_photo_error        stream_sim/surveys.py:234-292
 *   ss_efficiency        <- Survey.get_efficiency         stream_sim/surveys.py:399-493
 *   ss_eval_function     <- BoundedFunction.__call__      stream_sim/functions.py:64-76
 *
 * Conventions
 *   - plain C, POD structs, plain pointers and sizes; no torch types.
 *   - every array pointer is a DEVICE pointer unless the name ends in _host.
 *   - calls are stream-ordered on `stream` (a cudaStream_t passed as void*),
 *     never allocate, never synchronise (the *_host variants are the exception:
 *     they own the copies and return when the host buffers are complete).
 *   - return 0 on success, a negative SS_E* code otherwise; ss_last_error()
 *     gives a thread-local message.
 *   - a NULL output column is simply not written; a NULL draw array means
 *     "derive that draw from Philox4x32-10" (counter-based, see SS_DRAW_BLOCK_*).
 */
#ifndef STREAMSIM_B200_H
#define STREAMSIM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SS_ABI_VERSION 15

/* error codes */
#define SS_OK 0
#define SS_EINVAL (-1)   /* bad argument (message says which) */
#define SS_ECUDA (-2)    /* CUDA runtime error (message carries cudaGetErrorString) */
#define SS_ENODEV (-3)   /* no sm_100 device */

/* pipeline stages (bit mask) */
#define SS_STAGE_GENERATE 1u /* phi1, phi2, dist, mag_g, mag_r        (model.py:106-157) */
#define SS_STAGE_ROTATE 2u   /* (phi1,phi2) -> (ra,dec)                (observed.py:573)  */
#define SS_STAGE_OBSERVE 4u  /* pix, extinction, errors, noise, flags  (observed.py:146-291) */

/* function kinds  (stream_sim/functions.py) */
#define SS_FN_NONE 0
#define SS_FN_CONSTANT 1 /* p[0]=value                                  :99-102  */
#define SS_FN_LINE 2     /* p[0]=slope p[1]=intercept                   :130-133 */
#define SS_FN_SINUSOID 3 /* p[0]=amplitude p[1]=period p[2]=phase p[3]=offset :164-169 */
#define SS_FN_LINEAR 4   /* interp1d linear table: x[n], y[n], slope[n] :193-199 */
#define SS_FN_SPLINE 5   /* natural cubic PPoly: x[n], c[4*(n-1)]       :255-274 */
#define SS_FN_SPLINE_PRODUCT 6 /* sqrt(2pi)*I(x)*S(x): two SS_FN_SPLINE blocks :356-359 */

/* samplers  (stream_sim/model.py:513-527) */
#define SS_SAMPLER_GAUSSIAN 0
#define SS_SAMPLER_UNIFORM 1

/* density kinds (stream_sim/samplers.py) */
#define SS_DENSITY_UNIFORM 0  /* :105-120 */
#define SS_DENSITY_INVCDF 1   /* :55-76,158-161 */
#define SS_DENSITY_GAUSSIAN 2 /* :123-138 */
#define SS_DENSITY_NONE 3     /* phi1 always comes from SSCatalog.phi1 */
#define SS_DENSITY_GRID 4     /* INVCDF with a monotone CDF as a direct lookup: exact
                                 index thresholds + per-grid-point rows (SSTables.grid_*) */

/* band indices used in every [2] array below */
#define SS_BAND_G 0
#define SS_BAND_R 1

/* arithmetic of the transcendental-heavy steps (SSParams.math_mode) */
#define SS_MATH_F64 0   /* every value in double (<= 1e-9 of the oracle)               */
#define SS_MATH_FAST 1  /* same decisions, same positions and true magnitudes; the VALUES of
                           magerr_b / mag_b_obs come from fp32 10^x / sqrt / log2 (MUFU), with
                           a per-star double redo when the noisy flux is within a guard band
                           of zero.  pix and every flag are identical to SS_MATH_F64 (all
                           comparisons stay in double, the SNR cut is decided on log10(err));
                           magerr, mag_obs agree to ~2e-6 relative (north star: 1e-5).     */

/* output element type of the floating-point columns */
#define SS_DTYPE_F64 0
#define SS_DTYPE_F32 1

/* Philox4x32-10 draw blocks: counter = (star_lo, star_hi, block, 0),
 * key = (seed_lo, seed_hi), output words (w0,w1,w2,w3).  Two blocks per star:
 *   U32(a)   = a * 2^-32                       (uniform in [0,1), exact in double)
 *   BM32(a,b): fp32 Box-Muller, radius sqrt(-2 ln y), y = (~a + 0.5)/2^32, angle 2*pi*b/2^32
 *   BM64(a,b): the same in double (SSParams.normals_f64): angle = pi * (int32)b / 2^31
 *   block 0 (POS): u_phi1 = U32(w0), u_mass = U32(w1),
 *                  (z_phi2, z_dist) = BM32(w2,w3)   [Uniform samplers: U32(w2), U32(w3)]
 *   block 1 (OBS): (z_r, z_g) = BM32(w0,w1) flux noise,
 *                  u_det = U32(w2) (completeness), u_det2 = U32(w3) (detection-only)
 *   block 3:       z_phi1 = BM32(w0,w1).first      (Gaussian density only)
 * Every uniform decides a table step or a Bernoulli trial; 32 bits resolve both far
 * below what any test can see (an inverse-CDF step of the 10^5-point grid spans ~4*10^4
 * values), the bucket of a power-of-two look-up table is a shift of the word, and one
 * Philox block serves four draws.
 * The normals are random draws, generated in fp32 like cuRAND's curand_normal;
 * everything computed from them is double.  A star's draws depend only on
 * (seed, global star index): any sharding of the index range over launches or
 * GPUs gives identical per-star results.  SSOut.draws_out exports the draws used. */
#define SS_DRAW_BLOCK_POS 0
#define SS_DRAW_BLOCK_OBS 1
#define SS_DRAW_BLOCK_DENSITY_Z 3
/* rows of SSOut.draws_out */
#define SS_DRAW_U_PHI1 0
#define SS_DRAW_N_PHI2 1
#define SS_DRAW_N_DIST 2
#define SS_DRAW_U_MASS 3
#define SS_DRAW_Z_G 4
#define SS_DRAW_Z_R 5
#define SS_DRAW_U_DET 6
#define SS_DRAW_U_DET2 7
#define SS_N_DRAWS 8

/* counters written (atomically accumulated) into SSOut.counters */
#define SS_CNT_TOTAL 0
#define SS_CNT_OBSERVED 1     /* flag_observed                                  */
#define SS_CNT_PERFECT 2      /* flag_perfect_galstarsep                        */
#define SS_CNT_BADMAG_G 3     /* flux_obs <= 0 or NaN in g ("BAD_MAG")          */
#define SS_CNT_BADMAG_R 4
#define SS_CNT_UNSEEN 5       /* maglim NaN in the completeness band            */
#define SS_CNT_DOMAIN 6       /* !(scale >= 0) in a sampler -> scipy ValueError */
#define SS_CNT_BADCOORD 7     /* ra/dec NaN or |dec| > 90 -> healpy ValueError  */
#define SS_CNT_INSIDE 8       /* ss_frame_fraction: stars inside the mask       */
#define SS_N_COUNTERS 16
#define SS_HIST_BINS 256
#define SS_N_HIST 5           /* phi1, phi2, dist, mag_g, g-r                   */
/* SSOut.counters layout: [SS_N_COUNTERS] then SS_N_HIST x [SS_HIST_BINS], uint64 */
#define SS_COUNTERS_LEN (SS_N_COUNTERS + SS_N_HIST * SS_HIST_BINS)

typedef struct SSFunc {
  int32_t kind;
  int32_t n;      /* knots */
  int32_t x_off;  /* offsets into SSTables.blob, in doubles */
  int32_t c_off;  /* LINEAR: y[n] then slope[n]; SPLINE: c[4*(n-1)], highest power first */
  double xmin, xmax; /* result is NaN outside [xmin,xmax] */
  double p[4];
  int32_t n2, x2_off, c2_off; /* second spline of SS_FN_SPLINE_PRODUCT */
  /* optional uniform-bucket search guide over x[] (LINEAR tables): int32 guide[g_n]
   * stored inside the double blob at g_off (in doubles); bucket = (x-g_x0)*g_inv */
  int32_t g_n;
  int32_t g_off, pad_;
  double g_x0, g_inv;
  /* optional direct look-up form of a LINEAR table (the survey curves in the production
   * kernel): no search, no range / NaN selects.
   *   segments  seg[q_nseg][4 doubles] = (x_j, f_j, slope_j, 0) at q_seg_off: the table's own
   *     intervals framed by constant ones for x below the first knot (fill), at the last
   *     knot and above it (fill), so that  value = slope * (x - x_j) + f_j  (two roundings)
   *     is interp1d's answer for every finite or NaN x;
   *   buckets   rec[q_n + 2][4 doubles] = (x_a, x_b, j as int32 in the low word, 0) at q_off:
   *     bucket = clamp(floor((x - q_x0) * q_inv) + 1, 0, q_n + 1), interval = j + (x >= x_a)
   *     + (x >= x_b)  (+inf = no such knot).  q_n = 0: absent. */
  int32_t q_n, q_off, q_seg_off, q_nseg;
  double q_x0, q_inv;
} SSFunc;

typedef struct SSParams {
  /* ---- generate: density (samplers.py) */
  int32_t density_kind;
  int32_t cdf_n;        /* grid points (reference: 100000)                       */
  int32_t cdf_guide_n;  /* power of two; guide has cdf_guide_n+1 int32 entries   */
  int32_t cdf_identity; /* 1: interp1d's mergesort permutation is the identity   */
  double density_lo, density_hi; /* uniform: xmin,xmax; invcdf: grid start,stop; gaussian: mu,sigma */
  double grid_step;     /* np.linspace step                                      */
  double cdf_first, cdf_last; /* sorted cdf[0], cdf[n-1]; outside -> index 0     */
  /* ---- generate: track and distance modulus (model.py:500-544) */
  SSFunc track_center, track_spread, dist_center, dist_spread;
  int32_t track_sampler, dist_sampler;
  int32_t has_dist, has_iso;
  /* ---- generate: isochrone (model.py:575-604; ugali imf.sample + interp1d) */
  int32_t iso_n;        /* rows of (mass_init, mag_1, mag_2), sorted by mass      */
  int32_t iso_mass_off, iso_mag_off[2]; /* blob offsets (doubles); mag[n] then slope[n] */
  int32_t iso_guide_n, iso_guide_off;   /* uniform-bucket guide over mass_init     */
  int32_t imf_n;        /* IMF CDF steps (ugali: 10000)                           */
  int32_t imf_cdf_off;  /* blob offset of cdf[imf_n]                              */
  int32_t imf_guide_n;  /* power of two                                           */
  int32_t imf_guide_off;/* blob offset (in doubles) of int32 guide[imf_guide_n+1] */
  int32_t cdf_guide_off;/* blob offset (in doubles) of int32 guide[cdf_guide_n+1] */
  int32_t iso_direct;   /* 1: mags(u) tabulated on the composite breakpoints (SSTables.iso_*);
                           iso_n is then the number of composite rows, iso_lut_n the buckets */
  double imf_mass_lo, imf_mass_hi, imf_mass_step; /* np.linspace(lo,hi,imf_n)     */
  double iso_guide_x0, iso_guide_inv;
  /* ---- rotate: R rows = (origin, pole x origin, pole), ICRS -> stream frame */
  double R[9];
  /* ---- observe */
  int32_t nside_pix;        /* inject(nside=...) ; only used for SSOut.pix        */
  int32_t nside_ebv;
  int32_t nside_maglim[2];
  int32_t band_present[2];  /* band listed in inject(bands=...); 'r' is mandatory  */
  int32_t completeness_band;
  int32_t dust_correction, perfect_galstarsep;
  int32_t snr_cut[2];       /* band in detection_mag_cut (and in bands)           */
  int32_t nest;             /* 0 RING (the reference), 1 NESTED maps              */
  double coeff_extinc[2], saturation[2], sys_error[2];
  double delta_saturation;
  double snr_err_max;       /* largest e with fl(1/e) >= SNR_min (observed.py:267-279) */
  /* the same cut decided on x = log10(statistical error): magerr(x) = sqrt((10**x)**2 + sys**2)
   * is monotone, so 1/magerr >= SNR_min  <=>  x <= snr_loge_max[b], the largest double for
   * which the REFERENCE arithmetic (NumPy on the host) passes; saturated stars carry the
   * constant error err_bright, whose verdict is snr_bright_ok[b] */
  double snr_loge_max[2];
  int32_t snr_bright_ok[2];
  double log_err_at_dsat;   /* f(dsat)          surveys.py:278                    */
  double err_bright;        /* 10**f(dsat-1)    surveys.py:284                    */
  SSFunc photo_error;       /* SS_FN_LINEAR, fill 1.0 both sides (surveys.py:1412-1417) */
  SSFunc completeness;      /* SS_FN_LINEAR, fill 0.0           (surveys.py:1361-1366) */
  SSFunc detection;         /* detection-only table, used when perfect_galstarsep */
  /* ---- histograms, binned in fp32: t = ((float)v - (float)lo) * (float)inv_width (two
   * roundings, no FMA), bin = (int)t if 0 <= t < SS_HIST_BINS, dropped otherwise */
  int32_t hist_enable;
  int32_t grid_lut_n;       /* SS_DENSITY_GRID: buckets of the u -> index table (power of two) */
  double hist_lo[SS_N_HIST], hist_inv_width[SS_N_HIST];
  int32_t math_mode;        /* SS_MATH_*                                          */
  int32_t iso_lut_n;        /* buckets of SSTables.iso_lut (iso_direct; power of two) */
  int32_t grid_nthr;        /* SS_DENSITY_GRID: thresholds (= steps - 1) of the u -> index map */
  int32_t fused_kernel;     /* 1 (default): the production launch may take the pipelined kernel
                               k_fused; 0: always the generic k_stars (tests compare the two) */
  int32_t grid_wlut_n;      /* buckets of SSTables.grid_wlut (power of two; 0 = absent)  */
  int32_t iso_wlut_n;       /* buckets of SSTables.iso_wlut  (power of two; 0 = absent)  */
  int32_t normals_f64;      /* 0 (default): the four normal draws of a star come from an fp32
                               Box-Muller on the MUFU units (BM32 below); 1: the same two words
                               per pair through a double-precision Box-Muller (BM64) -- the star
                               kernels only; costs 12.5 % */
} SSParams;

typedef struct SSTables {
  const void* blob;          /* packed small tables (doubles; int32 guides embedded)    */
  int64_t blob_bytes;        /* multiple of 16                                    */
  int64_t blob_hot_bytes;    /* leading part staged into shared memory (0 = all); the
                                stream-model tables may lie beyond it when a phi1 grid
                                table is present (they are then read from global memory,
                                and only on the generic path)                       */
  const double* cdf_pairs;   /* [cdf_n][2] = (sorted cdf_j, slope_j)              */
  const int32_t* cdf_perm;   /* [cdf_n] original index of sorted entry j; NULL if identity */
  const double* grid;        /* [cdf_n] explicit phi1 grid, or NULL = start + j*step */
  /* SS_DENSITY_GRID: the sample index of inverse_transform_sample (samplers.py:71-75) is a
   * step function of u; step s = #{k : thr[k] <= u}, k < grid_nthr.  One row per step plus one
   * for interp1d's fill value: rows[grid_nthr + 2][8] = (thr[s], x, centre, spread, dm_centre,
   * dm_spread, sin(radians x), cos(radians x)) with x the phi1 grid value the step maps to (the
   * index may also step DOWN: non-monotone CDFs of spline densities); rows[grid_nthr + 1] is the
   * fill row (grid index 0).  Column 0 is free for the host: its low 32 bits hold the phi1
   * histogram bin of x (int32, -1 = outside), so the production kernel never bins phi1.
   * grid_thr[grid_nthr + 1] = the thresholds, then +inf.
   * grid_lut[grid_lut_n][2 doubles]: bucket b = [b, b+1)/grid_lut_n of u ->
   * (int32 lo | int32 k << 32 as the bits of the first double, thr[lo] as the second):
   * lo = step at the bucket's left edge, k = thresholds inside the bucket. */
  const double* grid_rows;
  const double* grid_lut;
  const double* grid_thr;
  /* iso_direct: rows[iso_n][4] = (a_g, dmag_g/du, a_r, dmag_r/du) on the union of the
   * IMF-CDF knots and the isochrone knots mapped to u: mag_b(u) = fma(slope_b, u, a_b) inside
   * segment k (a_b = mag_b(U_k) - slope_b * U_k, folded on the host in extended precision);
   * 32-byte aligned.  thr[k] = U_{k+1} (last = +inf); lut[iso_lut_n][2] packed exactly like
   * grid_lut.  grid_rows and SSMaps.packed must be 32-byte aligned as well. */
  const double* iso_rows;
  const double* iso_lut;
  const double* iso_thr;
  /* The same two step functions for a 32-BIT draw u = w * 2^-32 (the Philox path), in
   * integer form: thr[k] <= u  <=>  w >= ceil(thr[k] * 2^32)  <=>  w > wthr[k] with
   * wthr[k] = ceil(thr[k] * 2^32) - 1 (0xFFFFFFFF = never; thresholds <= 0 are folded into the
   * first bucket's step), so a step is decided by integer compares and three thresholds fit
   * in one 16-byte bucket entry:
   *   wlut[wlut_n][4 x uint32] = (lo | min(k, 4095) << 20, t0, t1, t2): lo = step at the bucket's
   *   left edge (< 2^20), k = thresholds inside the bucket, t0..t2 = the first three of them
   *   (0xFFFFFFFF where there is none): step = lo + (w > t0) + (w > t1) + (w > t2), complete for
   *   k <= 3; fuller buckets finish on wthr[lo + 3 .. lo + k) (all of wthr[lo + 3 ..) when k
   *   saturates at 4095).  bucket = w >> (32 - log2 wlut_n).
   * grid_wthr[grid_nthr], iso_wthr[iso_n - 1].  Optional (SSParams.grid_wlut_n / iso_wlut_n
   * = 0: absent); the pipelined production kernel requires both. */
  const uint32_t* grid_wlut;
  const uint32_t* grid_wthr;
  const uint32_t* iso_wlut;
  const uint32_t* iso_wthr;
} SSTables;

typedef struct SSMaps {
  const double* ebv;         /* [12*nside_ebv^2]                                  */
  const double* maglim[2];   /* [12*nside_maglim[b]^2], NaN = unseen              */
  const uint8_t* mask;       /* ss_frame_fraction only: [12*nside_mask^2]         */
  int32_t nside_mask;
  /* optional: the three maps interleaved, [12*nside_packed^2][4] = (ebv, maglim_g,
   * maglim_r, 0) -- one 32-byte sector per star instead of three gathers.  Only valid
   * when the three maps share one nside (0 = absent). */
  int32_t nside_packed;
  const double* packed;
  /* optional, instead of `packed`: the same records in fp32, [12*nside_packed^2][4 floats] --
   * one 16-byte gather, half the footprint.  NOT the reference's arithmetic: the survey maps
   * are rounded to fp32 first (the host then also hands the separate maps over rounded, so
   * that every kernel sees the same values); a labelled option, never the default. */
  const float* packed32;
} SSMaps;

typedef struct SSCatalog {   /* per-star inputs of the stages that are not run */
  /* ROTATE without GENERATE reads phi1/phi2.  GENERATE treats non-NULL phi1 / phi2 /
   * dist as a partially filled catalog: finite entries are kept, NaN entries are
   * sampled (StreamModel.complete_catalog, reference model.py:261-298). */
  const double* phi1;
  const double* phi2;
  const double* dist;
  const double* ra;          /* OBSERVE without ROTATE                           */
  const double* dec;
  const double* mag[2];      /* OBSERVE without GENERATE (true apparent mags)    */
} SSCatalog;

typedef struct SSDraws {     /* injected draws; NULL -> Philox */
  const double* u_phi1;      /* z for SS_DENSITY_GAUSSIAN                        */
  const double* n_phi2;      /* z (Gaussian sampler) or u (Uniform sampler)      */
  const double* n_dist;
  const double* u_mass;
  const double* z_flux[2];   /* per band                                         */
  const double* u_det;
  const double* u_det2;
} SSDraws;

typedef struct SSOut {
  int32_t dtype, pad_;       /* SS_DTYPE_*: element type of the columns below    */
  void* phi1;  void* phi2;  void* dist;
  void* mag[2];              /* true apparent magnitudes (no extinction)         */
  void* ra;  void* dec;
  void* mag_obs[2];          /* NaN where the reference writes "BAD_MAG"         */
  void* magerr[2];
  uint8_t* flag_observed;
  uint8_t* flag_perfect;
  int64_t* pix;              /* ang2pix at nside_pix                             */
  unsigned long long* counters; /* [SS_COUNTERS_LEN], accumulated, or NULL       */
  double* draws_out;         /* [SS_N_DRAWS][n] the draws actually used, or NULL  */
} SSOut;

int ss_abi_version(void);
const char* ss_last_error(void);
/* sizeof() of SSFunc, SSParams, SSTables, SSMaps, SSCatalog, SSDraws, SSOut, in that
 * order: lets a binding verify its struct mirrors. */
int ss_struct_sizes(int32_t* sizes7);
/* dynamic shared memory the star kernel will request for this blob (0 = blob
 * stays in global memory) and the grid it will launch for n stars. */
int ss_launch_info(int64_t blob_bytes, uint64_t n, int32_t* smem_bytes, int32_t* grid, int32_t* block);

/* Generic launcher: run `stages` for stars [star_offset, star_offset+n). */
int ss_run(uint32_t stages, const SSParams* p, const SSTables* t, const SSMaps* m,
           const SSCatalog* in, const SSDraws* d, const SSOut* out,
           uint64_t n, uint64_t star_offset, uint64_t seed, void* stream);

/* Named entry points = ss_run with a fixed stage mask. */
int ss_generate(const SSParams* p, const SSTables* t, const SSDraws* d, const SSOut* out,
                uint64_t n, uint64_t star_offset, uint64_t seed, void* stream);
int ss_rotate(const SSParams* p, const SSCatalog* in, const SSOut* out, uint64_t n, void* stream);
int ss_observe(const SSParams* p, const SSTables* t, const SSMaps* m, const SSCatalog* in,
               const SSDraws* d, const SSOut* out,
               uint64_t n, uint64_t star_offset, uint64_t seed, void* stream);
int ss_generate_observe(const SSParams* p, const SSTables* t, const SSMaps* m,
                        const SSDraws* d, const SSOut* out,
                        uint64_t n, uint64_t star_offset, uint64_t seed, void* stream);

/* One trial of the great-circle frame search: rotate (phi1,phi2) with p->R,
 * ang2pix at m->nside_mask, count stars whose mask byte is non-zero into
 * counters[SS_CNT_INSIDE] (and n into counters[SS_CNT_TOTAL]). */
int ss_frame_fraction(const SSParams* p, const SSMaps* m, const SSCatalog* in,
                      unsigned long long* counters, uint64_t n, void* stream);

/* Element-wise survey getters on device arrays (fp64 in, fp64 out). */
int ss_photo_error(const SSParams* p, const SSTables* t, int32_t band,
                   const double* mag, const double* maglim, double* out, uint64_t n, void* stream);
int ss_efficiency(const SSParams* p, const SSTables* t, int32_t band, int32_t detection_only,
                  const double* mag, const double* maglim, double* out, uint64_t n, void* stream);
int ss_eval_function(const SSFunc* f, const SSTables* t, const double* x, double* out,
                     uint64_t n, void* stream);
int ss_ang2pix(int32_t nside, int32_t nest, const double* lon_deg, const double* lat_deg,
               int64_t* pix, uint64_t n, void* stream);

/* The kernel's own double-precision elementary functions, element-wise on device arrays,
 * so that tests can hold them against libm/NumPy: fn = SS_MATH_FN_LOG10 (x > 0),
 * _EXP10 (|x| < 300), _ATAN2 (y = x[], x = y2[]), _SQRT (x >= 0).  `y2` is only read by
 * _ATAN2.  Replaces nothing in the reference (NumPy ufuncs np.log10, 10**x, np.arctan2,
 * np.sqrt at observed.py:984-1035, surveys.py:288-290 and astropy's lon/lat). */
#define SS_MATH_FN_LOG10 0
#define SS_MATH_FN_EXP10 1
#define SS_MATH_FN_ATAN2 2
#define SS_MATH_FN_SQRT 3
int ss_math_probe(int32_t fn, const double* x, const double* y2, double* out, uint64_t n,
                  void* stream);

/* Stand-alone pieces of the path, element-wise on device arrays (the reference exposes them
 * as public methods; inside ss_generate_observe the same arithmetic is fused).
 *  ss_loc_scale: out = draw * scale + loc, scipy's rv.rvs(loc, scale) (samplers.py:94-138).
 *    loc / scale: per-element arrays or NULL = the scalars loc0 / scale0; draws: injected
 *    standard draws or NULL = the star's own z_phi2 (normal != 0) / u_phi2 (Philox POS block).
 *  ss_noisy_magnitudes: StreamInjector.sample_measured_magnitudes (observed.py:863-904) as
 *    written (flux-space noise); NaN where the reference writes "BAD_MAG"; z injected or the
 *    star's z_r.
 *  ss_bernoulli: out = (u <= prob), StreamInjector.detect_flag's trial (observed.py:906-959);
 *    u injected or the star's u_det. */
int ss_loc_scale(int32_t normal, const double* loc, const double* scale, double loc0,
                 double scale0, const double* draws, double* out, uint64_t n, uint64_t star_offset,
                 uint64_t seed, void* stream);
int ss_noisy_magnitudes(const double* mag, const double* err, const double* z, double* out,
                        uint64_t n, uint64_t star_offset, uint64_t seed, void* stream);
int ss_bernoulli(const double* prob, const double* u, uint8_t* out, uint64_t n,
                 uint64_t star_offset, uint64_t seed, void* stream);

/* Host-buffer variant of ss_generate_observe: every SSOut column pointer (and
 * counters) is a HOST pointer (pinned for full speed).  Stars are produced in
 * chunks into the caller-provided device workspace and copied back on a second
 * stream while the next chunk computes.  Returns when `out` is complete.
 *   workspace: device buffer of workspace_bytes >= ss_host_workspace_bytes(chunk, dtype)
 */
int64_t ss_host_workspace_bytes(uint64_t chunk, int32_t dtype);
int ss_generate_observe_host(const SSParams* p, const SSTables* t, const SSMaps* m,
                             const SSOut* out_host, uint64_t n, uint64_t star_offset,
                             uint64_t seed, void* workspace, int64_t workspace_bytes,
                             uint64_t chunk);
/* Host-buffer variant of ss_observe: catalog columns (ra, dec, mag[2]) and all
 * outputs are HOST pointers; same chunked two-stream pipeline. */
int ss_observe_host(const SSParams* p, const SSTables* t, const SSMaps* m,
                    const SSCatalog* in_host, const SSOut* out_host, uint64_t n,
                    uint64_t star_offset, uint64_t seed, void* workspace,
                    int64_t workspace_bytes, uint64_t chunk);

#ifdef __cplusplus
}
#endif
#endif /* STREAMSIM_B200_H */
