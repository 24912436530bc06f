"""CPU oracle for the per-star generate-and-observe path of LSSTDESC/stream_sim.

TEST INFRASTRUCTURE ONLY.  Nothing under ``stream_sim_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and there only as the checker
(or as the CPU arm being timed), never as the shipped path.

It is a draw-injected NumPy restatement of the reference algorithm: every
random number the reference would pull from MT19937 / PCG64 is an explicit
input array here, so the CUDA path can be compared star by star.

Pinning status (see tests/test_oracle_golden.py and DESIGN.md section 3):
  * generate path (phi1, phi2, dist)         -- PINNED bit-equal against the real
    reference ``StreamModel.sample`` (tests/golden/generate_*.npz, made by
    scripts/make_golden.py importing /root/reference).
  * photo-error / flux noise / completeness  -- PINNED bit-equal against the real
    reference ``StreamInjector.inject`` + ``Survey`` getters (imported through
    stub modules for the absent healpy/astropy/gala; tests/golden/inject_*.npz).
  * function evaluators                      -- PINNED by the reference's own
    known-answer vectors (reference tests/test_functions.py:8-35).
  * stream-frame -> ICRS rotation            -- pinned only by the 5 printed rows of
    notebooks/tutorial_inject_stream.ipynb cell 15 (5e-7 deg).
  * HEALPix ang2pix (healpy/healpix_cxx)     -- PARITY UNPINNED against healpy
    itself (not installed, not vendored); restated from the published
    algorithm, checked against healpy's docstring examples (from memory) and
    geometric identities.
  * isochrone / IMF sampling (ugali)         -- PARITY UNPINNED (ugali and the
    PARSEC tables are absent); restated from the published ugali algorithm.
"""
from __future__ import annotations

import numpy as np
import scipy.interpolate

# --------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al. 2011, Random123).  The device derives all of its
# free-running draws from this generator; restating it here lets the tests feed
# the *same* draws to the NumPy chain.
# --------------------------------------------------------------------------
_PH_M0 = np.uint64(0xD2511F53)
_PH_M1 = np.uint64(0xCD9E8D57)
_PH_W0 = np.uint32(0x9E3779B9)
_PH_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds.  All args broadcastable uint32."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32).copy() for c in
                      np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PH_M0 * c0.astype(np.uint64)
            p1 = _PH_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_PH_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_PH_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def philox_block(seed, index, block):
    """The four 32-bit words of draw-block ``block`` for global star ``index``.

    counter = (index_lo, index_hi, block, 0), key = (seed_lo, seed_hi) --
    the layout documented in include/streamsim_b200.h (SS_DRAW_BLOCK_*).
    """
    index = np.asarray(index, dtype=np.uint64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return philox4x32_10((index & _MASK32).astype(np.uint32),
                         (index >> np.uint64(32)).astype(np.uint32),
                         np.uint32(block), np.uint32(0),
                         seed & 0xFFFFFFFF, seed >> 32)


def _u32(a):
    return a.astype(np.float64) * (1.0 / 4294967296.0)


def bm32(wa, wb):
    """The device's fp32 Box-Muller from two 32-bit words, emulated with NumPy float32:
    radius from y = (~wa + 0.5) / 2^32, angle 2 pi wb / 2^32.  NOT bit-identical to the CUDA
    intrinsics (MUFU log2 / sqrt / sin / cos: agrees to ~1e-6, less near z = 0): tests that
    need the exact normals read them back from the kernel (SSOut.draws_out) instead."""
    f = np.float32
    y = ((~wa).astype(f) + f(0.5)) * f(2.3283064365386963e-10)
    r = np.sqrt(f(-1.3862943611198906) * np.log2(y).astype(f))
    ang = wb.astype(np.float64) * 4.6566128730773926e-10
    return ((r * np.cos(np.pi * ang).astype(f)).astype(np.float64),
            (r * np.sin(np.pi * ang).astype(f)).astype(np.float64))


def bm64(wa, wb):
    """The device's double-precision Box-Muller (SSParams.normals_f64, csrc ``bm64``): the
    same two words, y = (~wa + 0.5) / 2^32, radius sqrt(-2 ln y), angle pi * int32(wb) / 2^31.
    Agrees with the kernel to a few ulp (its own log10 / sqrt, CUDA's sincospi)."""
    y = ((~wa).astype(np.float64) + 0.5) * (1.0 / 4294967296.0)
    r = np.sqrt(-2.0 * np.log(y))
    ang = np.pi * (wb.astype(np.int32).astype(np.float64) * 4.656612873077393e-10)
    return r * np.cos(ang), r * np.sin(ang)


def device_draws(seed, start, n, normals="f32"):
    """All per-star draws the kernel derives for stars [start, start+n): two Philox blocks
    per star, layout of include/streamsim_b200.h (SS_DRAW_BLOCK_POS / _OBS).

    Returns the dict accepted by :func:`generate` / :func:`observe`.  The uniforms
    (u_phi1, u_mass, u_phi2, u_dist, u_det, u_det2; one 32-bit word each) are exactly the
    device's; the normals (z_phi2, z_dist, z_r, z_g) are the float32 emulation of
    :func:`bm32` (``normals="f64"``: of :func:`bm64`).
    """
    bm = bm64 if normals == "f64" else bm32
    idx = np.arange(start, start + n, dtype=np.uint64)
    w = philox_block(seed, idx, 0)
    u_phi1, u_mass = _u32(w[0]), _u32(w[1])
    z_phi2, z_dist = bm(w[2], w[3])
    u_phi2, u_dist = _u32(w[2]), _u32(w[3])
    w = philox_block(seed, idx, 1)
    z_r, z_g = bm(w[0], w[1])
    u_det, u_det2 = _u32(w[2]), _u32(w[3])
    return dict(u_phi1=u_phi1, u_mass=u_mass, u_phi2=u_phi2, u_dist=u_dist,
                z_phi2=z_phi2, z_dist=z_dist, z_r=z_r, z_g=z_g,
                u_det=u_det, u_det2=u_det2)


# --------------------------------------------------------------------------
# Functions  (reference stream_sim/functions.py)
# --------------------------------------------------------------------------
def _interp1d_linear(xp, fp, x, fill):
    """scipy interp1d(kind='linear', bounds_error=False, fill_value=fill) on
    1-D float data: stable sort of xp, np.interp, then out-of-range -> fill
    (scipy/interpolate/_interpolate.py)."""
    xp = np.asarray(xp, dtype=float)
    fp = np.asarray(fp, dtype=float)
    order = np.argsort(xp, kind="mergesort")
    xs, fs = xp[order], fp[order]
    x = np.asarray(x, dtype=float)
    y = np.interp(x, xs, fs)
    y = np.where((x < xs[0]) | (x > xs[-1]), fill, y)
    return y


class OFunc:
    """One phi1 -> y evaluator, built from a config dict ``{'type': ..., ...}``.

    Follows functions.py:64-76 (bounds -> NaN), :99-102 (Constant), :130-133
    (Line), :164-169 (Sinusoid), :193-199 (Interpolation), :255-274
    (CubicSplineInterpolation), :356-359 (LinearDensityCubicSplineInterpolation).
    """

    def __init__(self, cfg, data_root="."):
        import os
        import pandas as pd
        cfg = dict(cfg)
        self.kind = cfg.pop("type").lower()
        self.xmin, self.xmax = -np.inf, np.inf
        k = self.kind

        def _path(p):
            return p if os.path.isabs(p) else os.path.join(data_root, p)

        if k == "constant":
            self.value = cfg.get("value", 1.0)
        elif k == "line":
            self.slope = cfg.get("slope", 0.0)
            self.intercept = cfg.get("intercept", 0.0)
        elif k == "sinusoid":
            self.amplitude = cfg.get("amplitude", 1.0)
            self.period = cfg.get("period", 1.0)
            self.phase = cfg.get("phase", 0.0)
            self.offset = cfg.get("offset", 0.0)
        elif k in ("interpolation", "file", "fileinterpolation"):
            if k == "interpolation":
                xv, yv = np.atleast_1d(cfg["xvals"]), np.atleast_1d(cfg["yvals"])
            else:
                df = pd.read_csv(_path(cfg["filename"]))
                cols = cfg.get("columns") or list(df.columns[[0, 1]])
                xv, yv = df[cols[0]].values, df[cols[1]].values
            self.xv, self.yv = xv, yv
            self.xmin, self.xmax = np.min(xv), np.max(xv)
        elif k in ("cubicsplineinterpolation", "filecubicsplineinterpolation"):
            if k.startswith("file"):
                df = pd.read_csv(_path(cfg["filename"]))
                if cfg.get("spline_type"):
                    df = df.loc[df["type"] == cfg["spline_type"], :]
                if cfg.get("stream_name"):
                    df = df.loc[df["stream"] == cfg["stream_name"], :]
                nodes = df[cfg.get("nodes_name", "phi1")].values
                vals = df[cfg.get("node_vals_name", "mean")].values
            else:
                nodes, vals = np.atleast_1d(cfg["nodes"]), np.atleast_1d(cfg["node_values"])
            self.nodes, self.vals = nodes, vals
            self.xmin, self.xmax = np.min(nodes), np.max(nodes)
        elif k in ("lineardensitycubicsplineinterpolation",
                   "filelineardensitycubicsplineinterpolation"):
            if k.startswith("file"):
                df = pd.read_csv(_path(cfg["filename"]))
                if cfg.get("stream_name"):
                    df = df.loc[df["stream"] == cfg["stream_name"]]
                si = df["type"] == "peak_intensity"
                ss = df["type"] == "spread"
                inodes, ivals = df.loc[si, "phi1"].values, df.loc[si, "mean"].values
                snodes, svals = df.loc[ss, "phi1"].values, df.loc[ss, "mean"].values
            else:
                inodes, ivals = np.asarray(cfg["intensity_nodes"]), np.asarray(cfg["intensity_node_values"])
                snodes, svals = np.asarray(cfg["spread_nodes"]), np.asarray(cfg["spread_node_values"])
            self.peak = OFunc(dict(type="cubicsplineinterpolation", nodes=inodes, node_values=ivals))
            self.spread = OFunc(dict(type="cubicsplineinterpolation", nodes=snodes, node_values=svals))
            self.xmin = np.max([inodes.min(), snodes.min()])
            self.xmax = np.min([inodes.max(), snodes.max()])
        else:
            raise Exception(f"Unrecognized function: {k}")
        if "xmin" in cfg:
            self.xmin = cfg["xmin"]
        if "xmax" in cfg:
            self.xmax = cfg["xmax"]

    def _evaluate(self, x):
        k = self.kind
        if k == "constant":
            return self.value * np.ones(len(x))
        if k == "line":
            return self.slope * x + self.intercept
        if k == "sinusoid":
            norm = self.amplitude / 2.0
            return norm * np.cos(2 * np.pi * (x - self.phase) / self.period) + norm + self.offset
        if k in ("interpolation", "file", "fileinterpolation"):
            return _interp1d_linear(self.xv, self.yv, x, np.nan)
        if k in ("cubicsplineinterpolation", "filecubicsplineinterpolation"):
            return scipy.interpolate.CubicSpline(self.nodes, self.vals, bc_type="natural",
                                                 extrapolate=False)(x)
        return np.sqrt(2 * np.pi) * self.peak(x) * self.spread(x)

    def __call__(self, x):
        x = np.atleast_1d(np.asarray(x, dtype=float))
        y = self._evaluate(x)
        if self.kind.endswith("lineardensitycubicsplineinterpolation"):
            return y  # functions.py:349-354 -- no bounds mask on this class
        return np.where((x >= self.xmin) & (x <= self.xmax), y, np.nan)


# --------------------------------------------------------------------------
# Samplers / model  (reference samplers.py, model.py)
# --------------------------------------------------------------------------
NSTEPS = 100000


def inverse_cdf_tables(func, nsteps=NSTEPS):
    """Tables behind samplers.py:55-76,158-161: grid, pdf, cdf and interp1d's
    mergesorted (cdf, index) pair."""
    xg = np.linspace(func.xmin, func.xmax, int(nsteps))
    pdf = func(xg)
    cdf = np.cumsum(pdf)
    cdf /= cdf[-1]
    order = np.argsort(cdf, kind="mergesort")
    return xg, cdf[order], order.astype(float)


def sample_phi1(density_cfg, u, data_root="."):
    """phi1 from uniforms ``u``  (model.py:484-497; samplers.py:105-120,158-161)."""
    cfg = dict(density_cfg)
    kind = cfg["type"].lower()
    if kind == "uniform":
        lo, hi = cfg["xmin"], cfg["xmax"]
        return u * (hi - lo) + lo            # scipy.stats.uniform(loc, scale).rvs
    if kind == "gaussian":
        raise NotImplementedError("gaussian density needs normal draws; use z")
    func = OFunc(cfg, data_root)
    xg, cs, perm = inverse_cdf_tables(func)
    y = np.interp(u, cs, perm)
    y = np.where((u < cs[0]) | (u > cs[-1]), 0.0, y)
    return xg[np.rint(y).astype(int)]


def sample_track(track_cfg, phi1, z=None, u=None, data_root="."):
    """phi2 (or distance modulus) at phi1  (model.py:513-544)."""
    center = OFunc(track_cfg["center"], data_root)
    spread = OFunc(track_cfg["spread"], data_root)
    kind = track_cfg.get("sampler", "Gaussian").lower()
    c, s = center(phi1), spread(phi1)
    if kind == "gaussian":
        return z * s + c                     # norm(loc=c, scale=s).rvs
    if kind == "uniform":
        lo = c - s
        hi = lo + 2 * s
        return u * (hi - lo) + lo
    raise Exception(f"Unrecognized sampler: {kind}")


# ---- isochrone / IMF: restatement of ugali -- UNPINNED
def chabrier2003_pdf(mass, log_mode=True):
    log_mass = np.log10(mass)
    m_c, sigma, x = 0.079, 0.69, 1.3
    a, b = 1.31357499301, 0.279087531047
    dn = (log_mass <= 0) * a * np.exp(-(log_mass - np.log10(m_c)) ** 2 / (2 * sigma ** 2))
    dn = dn + (log_mass > 0) * a * b * mass ** (-x)
    return dn if log_mode else dn / (mass * np.log(10))


def imf_cdf(mass_min, mass_max, steps=10000):
    d_mass = (mass_max - mass_min) / float(steps)
    mass = np.linspace(mass_min, mass_max, steps)
    cdf = np.insert(np.cumsum(d_mass * chabrier2003_pdf(mass[1:], log_mode=False)), 0, 0.0)
    cdf = cdf / cdf[-1]
    return mass, cdf


def sample_mags(iso, u_mass, dist):
    """iso = dict(mass_init, mag_1, mag_2) -> (mag_g, mag_r)  (model.py:575-604)."""
    mi = np.asarray(iso["mass_init"], dtype=float)
    mgrid, cdf = imf_cdf(np.min(mi), np.max(mi))
    mass = np.interp(u_mass, cdf, mgrid)
    order = np.argsort(mi, kind="mergesort")
    m1 = np.interp(mass, mi[order], np.asarray(iso["mag_1"], dtype=float)[order]) + 0.0
    m2 = np.interp(mass, mi[order], np.asarray(iso["mag_2"], dtype=float)[order]) + 0.0
    return m1 + dist, m2 + dist


def generate(stream_cfg, draws, iso=None, data_root="."):
    """StreamModel.sample with injected draws (model.py:106-157).

    draws: u_phi1; z_phi2 or u_phi2; z_dist or u_dist; u_mass.
    """
    dens = stream_cfg.get("density") or stream_cfg.get("linear_density")
    phi1 = sample_phi1(dens, draws["u_phi1"], data_root)
    phi2 = sample_track(stream_cfg["track"], phi1, draws.get("z_phi2"), draws.get("u_phi2"), data_root)
    out = dict(phi1=phi1, phi2=phi2)
    if stream_cfg.get("distance_modulus"):
        out["dist"] = sample_track(stream_cfg["distance_modulus"], phi1,
                                   draws.get("z_dist"), draws.get("u_dist"), data_root)
        if iso is not None:
            out["mag_g"], out["mag_r"] = sample_mags(iso, draws["u_mass"], out["dist"])
    return out


# --------------------------------------------------------------------------
# Great-circle frame (gala/astropy restatement)
# --------------------------------------------------------------------------
def _unit(ra_deg, dec_deg):
    ra, dec = np.radians(ra_deg), np.radians(dec_deg)
    return np.array([np.cos(dec) * np.cos(ra), np.cos(dec) * np.sin(ra), np.sin(dec)])


def frame_from_endpoints(ra1, dec1, ra2, dec2):
    """Rotation matrix R (rows origin, pole x origin, pole): ICRS -> stream frame."""
    c1, c2 = _unit(ra1, dec1), _unit(ra2, dec2)
    pole = np.cross(c1, c2)
    pole /= np.linalg.norm(pole)
    origin = c1 + c2
    origin /= np.linalg.norm(origin)
    return np.stack([origin, np.cross(pole, origin), pole])


def random_endpoints(rng):
    """observed.py:695-698: z~U(-1,1), ra~U(0,360) -- one sky point, 2 draws."""
    z = rng.uniform(-1.0, 1.0)
    dec = np.degrees(np.arcsin(z))
    ra = rng.uniform(0.0, 360.0)
    return ra, dec


def find_frame(rng, phi1, phi2, mask, percentile_threshold=0.99, max_iter=1000):
    """StreamInjector._find_gc_frame (observed.py:651-679): first random frame whose
    fraction of stars inside the boolean RING mask reaches the threshold.
    Returns (R, endpoints, trials, fraction) or None."""
    nside = npix2nside(len(mask))
    for trial in range(1, max_iter + 1):
        e1, e2 = random_endpoints(rng), random_endpoints(rng)
        R = frame_from_endpoints(*e1, *e2)
        ra, dec = rotate(phi1, phi2, R)
        frac = np.count_nonzero(mask[ang2pix_ring(nside, ra, dec)]) / len(ra)
        if frac >= percentile_threshold:
            return R, (e1, e2), trial, frac
    return None


def rotate(phi1, phi2, R):
    """(phi1, phi2) deg in the stream frame -> (ra, dec) deg ICRS."""
    l, b = np.radians(phi1), np.radians(phi2)
    v = np.stack([np.cos(b) * np.cos(l), np.cos(b) * np.sin(l), np.sin(b)])
    w = R.T @ v
    ra = np.degrees(np.arctan2(w[1], w[0]))
    ra = np.where(ra < 0.0, ra + 360.0, ra)
    dec = np.degrees(np.arctan2(w[2], np.hypot(w[0], w[1])))
    return ra, dec


# --------------------------------------------------------------------------
# HEALPix ang2pix (healpix_cxx loc2pix restatement) -- UNPINNED
# --------------------------------------------------------------------------
_INV_HALFPI = 0.6366197723675813430755350534900574
_TWOTHIRD = 2.0 / 3.0


def _fmodulo4(v):
    pos = np.where(v < 4.0, v, np.fmod(v, 4.0))
    t = np.fmod(v, 4.0) + 4.0
    neg = np.where(t == 4.0, 0.0, t)
    return np.where(v >= 0, pos, neg)


def _zphi(lon, lat):
    theta = np.pi / 2.0 - np.radians(np.asarray(lat, dtype=float))
    phi = np.radians(np.asarray(lon, dtype=float))
    have_sth = (theta < 0.01) | (theta > 3.14159 - 0.01)
    z = np.cos(theta)
    sth = np.where(have_sth, np.sin(theta), 0.0)
    return z, phi, sth, have_sth


def ang2pix_ring(nside, lon, lat):
    """hp.ang2pix(nside, lon, lat, lonlat=True) in the RING scheme."""
    ns = int(nside)
    z, phi, sth, have_sth = _zphi(lon, lat)
    za = np.abs(z)
    tt = _fmodulo4(phi * _INV_HALFPI)
    nl4 = 4 * ns
    ncap = 2 * ns * (ns - 1)
    npix = 12 * ns * ns
    with np.errstate(invalid="ignore"):
        # equatorial
        t1 = ns * (0.5 + tt)
        t2 = ns * z * 0.75
        jp = (t1 - t2).astype(np.int64)
        jm = (t1 + t2).astype(np.int64)
        ir = ns + 1 + jp - jm
        kshift = 1 - (ir & 1)
        t = jp + jm - ns + kshift + 1 + nl4 + nl4
        ip = (t >> 1) & (nl4 - 1) if (ns & (ns - 1)) == 0 and ns > 1 else (t >> 1) % nl4
        pix_eq = ncap + (ir - 1) * nl4 + ip
        # polar caps
        tp = tt - tt.astype(np.int64)
        tmp = np.where((za < 0.99) | (~have_sth), ns * np.sqrt(3 * (1 - za)),
                       ns * sth / np.sqrt((1.0 + za) / 3.0))
        jp = (tp * tmp).astype(np.int64)
        jm = ((1.0 - tp) * tmp).astype(np.int64)
        ir = jp + jm + 1
        ip = (tt * ir).astype(np.int64)
        pix_pol = np.where(z > 0, 2 * ir * (ir - 1) + ip, npix - 2 * ir * (ir + 1) + ip)
    return np.where(za <= _TWOTHIRD, pix_eq, pix_pol)


def _spread_bits(v):
    v = v.astype(np.uint64)
    v = (v | (v << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
    v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
    v = (v | (v << np.uint64(2))) & np.uint64(0x3333333333333333)
    v = (v | (v << np.uint64(1))) & np.uint64(0x5555555555555555)
    return v


def ang2pix_nest(nside, lon, lat):
    """hp.ang2pix(nside, lon, lat, nest=True, lonlat=True)."""
    ns = int(nside)
    order = ns.bit_length() - 1
    z, phi, sth, have_sth = _zphi(lon, lat)
    za = np.abs(z)
    tt = _fmodulo4(phi * _INV_HALFPI)
    with np.errstate(invalid="ignore"):
        t1 = ns * (0.5 + tt)
        t2 = ns * z * 0.75
        jp = (t1 - t2).astype(np.int64)
        jm = (t1 + t2).astype(np.int64)
        ifp, ifm = jp >> order, jm >> order
        face_eq = np.where(ifp == ifm, ifp | 4, np.where(ifp < ifm, ifp, ifm + 8))
        ix_eq = jm & (ns - 1)
        iy_eq = ns - (jp & (ns - 1)) - 1
        ntt = np.minimum(3, tt.astype(np.int64))
        tp = tt - ntt
        tmp = np.where((za < 0.99) | (~have_sth), ns * np.sqrt(3 * (1 - za)),
                       ns * sth / np.sqrt((1.0 + za) / 3.0))
        jp = np.minimum((tp * tmp).astype(np.int64), ns - 1)
        jm = np.minimum(((1.0 - tp) * tmp).astype(np.int64), ns - 1)
        ix_p = np.where(z >= 0, ns - jm - 1, jp)
        iy_p = np.where(z >= 0, ns - jp - 1, jm)
        face_p = np.where(z >= 0, ntt, ntt + 8)
    eq = za <= _TWOTHIRD
    ix = np.where(eq, ix_eq, ix_p)
    iy = np.where(eq, iy_eq, iy_p)
    face = np.where(eq, face_eq, face_p)
    inter = _spread_bits(ix) | (_spread_bits(iy) << np.uint64(1))
    return (face.astype(np.int64) << (2 * order)) + inter.astype(np.int64)


def npix2nside(npix):
    ns = int(round(np.sqrt(npix / 12.0)))
    if 12 * ns * ns != npix:
        raise ValueError("Wrong pixel number (it is not 12*nside**2)")
    return ns


# --------------------------------------------------------------------------
# Survey tables and getters (reference surveys.py)
# --------------------------------------------------------------------------
def completeness_table(delta_mags, eff, delta_saturation=-10.4):
    """surveys.py:1344-1358: zero padding at the bright and faint ends."""
    dm = np.array(delta_mags, dtype=float)
    ef = np.array(eff, dtype=float)
    if dm.min() > delta_saturation:
        dm = np.insert(dm, 0, delta_saturation)
        ef = np.insert(ef, 0, 0.0)
    elif dm.min() == delta_saturation:
        dm = np.insert(dm, 0, delta_saturation - 1e-5)
        ef = np.insert(ef, 0, 0.0)
    else:
        ef[dm <= delta_saturation] = 0.0
    dm = np.append(dm, dm[-1] + 1)
    ef = np.append(ef, 0.0)
    return dm, ef


def photo_error_table(delta_mags, log_err, delta_saturation=-10.4):
    """surveys.py:1406-1409."""
    dm = np.array(delta_mags, dtype=float)
    le = np.array(log_err, dtype=float)
    if dm.min() > delta_saturation:
        dm = np.insert(dm, 0, delta_saturation)
        le = np.insert(le, 0, le[0])
    return dm, le


def get_photo_error(sv, band, mag, maglim):
    """surveys.py:234-292.  sv['photo_error'] = (x, y) from photo_error_table."""
    fx, fy = sv["photo_error"]
    f = lambda d: _interp1d_linear(fx, fy, d, 1.0)
    dsat = sv["delta_saturation"]
    sat = sv["saturation"][band]
    with np.errstate(invalid="ignore", over="ignore"):
        delta = mag - maglim
        stat = 10 ** np.where((delta <= dsat) & (mag >= sat), f(dsat), f(delta))
        stat = np.where(mag < sat, 10 ** f(dsat - 1), stat)
        sys_err = sv["sys_error"].get(band, 0.005)
        return np.sqrt(stat ** 2 + sys_err ** 2)


def get_efficiency(sv, band, mag, maglim, table="completeness"):
    """surveys.py:399-493."""
    fx, fy = sv[table]
    dsat = sv["delta_saturation"]
    sat = sv["saturation"][band]
    with np.errstate(invalid="ignore"):
        delta = mag - maglim
        c = np.where((mag > sat) & (delta <= dsat), 1.0, _interp1d_linear(fx, fy, delta, 0.0))
        c = np.where(mag < sat, 0.0, c)
        c = np.where((maglim < sat) | np.isnan(maglim), 0.0, c)
    return c


def observe(sv, ra, dec, mag_g, mag_r, draws, bands=("r", "g"), nside=4096,
            dust_correction=True, perfect_galstarsep=False, detection_mag_cut=("g",),
            snr_min=5.0):
    """StreamInjector.inject per-star body with injected draws (observed.py:146-291).

    sv: dict(ebv_map, maglim_maps{band}, coeff_extinc{band}, saturation{band},
    sys_error{band}, delta_saturation, completeness_band, photo_error=(x,y),
    completeness=(x,y)[, efficiency_detection=(x,y)]).
    draws: z_r, z_g (flux noise), u_det (completeness), u_det2 (detection-only).
    BAD_MAG is returned as NaN.
    """
    mags = dict(g=np.asarray(mag_g, dtype=float), r=np.asarray(mag_r, dtype=float))
    out = dict(pix=ang2pix_ring(nside, ra, dec))
    ns_ebv = npix2nside(len(sv["ebv_map"]))
    pix_ebv = out["pix"] if ns_ebv == nside else ang2pix_ring(ns_ebv, ra, dec)
    flag_c = flag_d = None
    for band in bands:
        ext = sv["coeff_extinc"][band] * sv["ebv_map"][pix_ebv]
        m_true = mags[band] + ext
        mlmap = sv["maglim_maps"][band]
        ns_ml = npix2nside(len(mlmap))
        pix_ml = out["pix"] if ns_ml == nside else ang2pix_ring(ns_ml, ra, dec)
        maglim = mlmap[pix_ml]
        err = get_photo_error(sv, band, m_true, maglim)
        with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
            flux = 3631.0 * 10 ** (-0.4 * m_true)
            sig = flux * err / 1.0857362
            fobs = flux + (0.0 + sig * draws["z_" + band])
            mobs = np.where(fobs > 0.0, -2.5 * np.log10(fobs / 3631.0), np.nan)
        if dust_correction:
            mobs = mobs - ext
        out["mag_%s_obs" % band] = mobs
        out["magerr_" + band] = err
        if band == sv["completeness_band"]:
            flag_c = draws["u_det"] <= get_efficiency(sv, band, m_true, maglim)
            if perfect_galstarsep:
                flag_d = draws["u_det2"] <= get_efficiency(sv, band, m_true, maglim,
                                                           "efficiency_detection")
    valid = ~np.isnan(out["mag_r_obs"])
    if "g" in bands:
        valid &= ~np.isnan(out["mag_g_obs"])
    flag = valid & flag_c
    perfect = valid & flag_d if perfect_galstarsep else None
    with np.errstate(divide="ignore", invalid="ignore"):
        for band in detection_mag_cut:
            if band not in bands:
                continue
            snr_ok = (1.0 / out["magerr_" + band]) >= snr_min
            flag &= snr_ok
            if perfect_galstarsep:
                perfect &= snr_ok
        if perfect_galstarsep:
            perfect &= (1.0 / out["magerr_r"]) >= snr_min
    out["flag_observed"] = flag
    if perfect_galstarsep:
        out["flag_perfect_galstarsep"] = perfect
    return out


def generate_observe(stream_cfg, iso, R, sv, draws, data_root=".", **kw):
    """The whole fused path: generate -> rotate -> observe."""
    out = generate(stream_cfg, draws, iso, data_root)
    out["ra"], out["dec"] = rotate(out["phi1"], out["phi2"], R)
    out.update(observe(sv, out["ra"], out["dec"], out["mag_g"], out["mag_r"], draws, **kw))
    return out
