"""Isochrone and initial-mass-function tables (host side).

The reference delegates magnitudes to ``ugali`` (``stream_sim/model.py:557-604``):
``iso.simulate`` draws initial masses from the IMF by inverse transform on a
10000-step linear-mass CDF and interpolates the isochrone's (mass_init, mag_1,
mag_2) rows linearly in mass.  ugali and its PARSEC
tables are not a dependency here: this module restates the published algorithm,
builds the two lookup tables once on the host, and the per-star work runs on the
GPU (csrc/streamsim_b200.cu ``iso_sample``).

An isochrone can come from
  * ugali, if it happens to be installed (``IsochroneTable.from_ugali``);
  * an explicit table: config keys ``mass_init``/``mag_1``/``mag_2`` or
    ``table: <csv with those columns>`` (``IsochroneTable.from_config``);
  * the built-in synthetic old metal-poor isochrone (``name: synthetic``), used
    by tests and benchmarks.
"""
from __future__ import annotations

import numpy as np

IMF_STEPS = 10000


def chabrier2003_pdf(mass, log_mode=True):
    """Chabrier (2003) IMF, dn/dlog10(m) (or dn/dm), ugali's constants."""
    log_mass = np.log10(mass)
    m_c, sigma, x = 0.079, 0.69, 1.3
    a, b = 1.31357499301, 0.279087531047     # normalisation; continuity at 1 Msun
    low = (log_mass <= 0) * a * np.exp(-(log_mass - np.log10(m_c)) ** 2 / (2 * sigma ** 2))
    high = (log_mass > 0) * a * b * mass ** (-x)
    dn_dlogm = low + high
    return dn_dlogm if log_mode else dn_dlogm / (mass * np.log(10))


_IMFS = {"chabrier2003": chabrier2003_pdf}


def imf_cdf_table(mass_min, mass_max, steps=IMF_STEPS, imf="chabrier2003"):
    """(mass grid, normalised CDF) exactly as ugali's ``IMF.sample`` builds them:
    left-rectangle cumulative sum of pdf on ``linspace(min, max, steps)``."""
    pdf = _IMFS[imf.lower()]
    d_mass = (mass_max - mass_min) / float(steps)
    mass = np.linspace(mass_min, mass_max, steps)
    cdf = np.insert(np.cumsum(d_mass * pdf(mass[1:], log_mode=False)), 0, 0.0)
    cdf = cdf / cdf[-1]
    return mass, cdf


class IsochroneTable:
    """(mass_init, mag_1, mag_2) rows sorted by mass, plus the IMF CDF."""

    def __init__(self, mass_init, mag_1, mag_2, imf="chabrier2003", band_1="g", band_2="r"):
        mass_init = np.asarray(mass_init, dtype=float)
        order = np.argsort(mass_init, kind="mergesort")     # interp1d's ordering
        self.mass_init = mass_init[order]
        self.mag_1 = np.asarray(mag_1, dtype=float)[order]
        self.mag_2 = np.asarray(mag_2, dtype=float)[order]
        if len(self.mass_init) < 2:
            raise ValueError("isochrone table needs at least two rows")
        self.band_1, self.band_2 = band_1, band_2
        self.imf = imf
        self.mass_grid, self.cdf = imf_cdf_table(self.mass_init[0], self.mass_init[-1], imf=imf)

    # g is always reported as mag_g, r as mag_r whatever band_1/band_2 order was given
    def mags_gr(self):
        if (self.band_1, self.band_2) == ("r", "g"):
            return self.mag_2, self.mag_1
        return self.mag_1, self.mag_2

    def composite(self):
        """mag_g(u), mag_r(u) as ONE piecewise-linear table over the uniform draw u.

        mass(u) (ugali's IMF inverse transform) is piecewise linear on the CDF knots and
        mag(mass) piecewise linear on the isochrone rows, so their composition is
        piecewise linear on the union of the CDF knots and of the u values at which the
        mass crosses an isochrone row.  Breakpoint values come from the two-stage
        arithmetic itself.  Returns (U, mag_g, dmag_g/du, mag_r, dmag_r/du), the slopes
        being 0 on the last row."""
        mg, mr = self.mags_gr()
        inner = self.mass_init[(self.mass_init > self.mass_grid[0])
                               & (self.mass_init < self.mass_grid[-1])]
        u_knots = np.interp(inner, self.mass_grid, self.cdf)
        U = np.unique(np.concatenate([self.cdf, u_knots]))
        U = U[(U >= 0.0) & (U <= 1.0)]
        mass = np.interp(U, self.cdf, self.mass_grid)
        G = np.interp(mass, self.mass_init, mg)
        R = np.interp(mass, self.mass_init, mr)
        dU = np.diff(U)
        sg = np.concatenate([np.diff(G) / dU, [0.0]])
        sr = np.concatenate([np.diff(R) / dU, [0.0]])
        return U, G, sg, R, sr

    @classmethod
    def synthetic(cls, n=256):
        from .synthetic import isochrone_table
        t = isochrone_table(n)
        return cls(t["mass_init"], t["mag_1"], t["mag_2"])

    @classmethod
    def from_ugali(cls, config):
        import ugali.isochrone           # ImportError propagates: caller decides
        cfg = {k: v for k, v in config.items() if k not in ("table", "mass_init", "mag_1", "mag_2")}
        iso = ugali.isochrone.factory(**cfg)
        return cls(iso.mass_init, iso.mag_1, iso.mag_2,
                   band_1=config.get("band_1", "g"), band_2=config.get("band_2", "r"))

    @classmethod
    def from_config(cls, config):
        """Build from an ``isochrone:`` config section (reference model.py:557-573)."""
        kw = dict(band_1=config.get("band_1", "g"), band_2=config.get("band_2", "r"))
        if "mass_init" in config:
            return cls(config["mass_init"], config["mag_1"], config["mag_2"], **kw)
        if "table" in config:
            import pandas as pd
            df = pd.read_csv(config["table"])
            return cls(df["mass_init"].values, df["mag_1"].values, df["mag_2"].values, **kw)
        if str(config.get("name", "")).lower() == "synthetic":
            return cls.synthetic()
        try:
            return cls.from_ugali(config)
        except ImportError as e:
            raise ImportError(
                "isochrone '%s' needs ugali and its isochrone library, which are not installed; "
                "pass an explicit table (mass_init/mag_1/mag_2 or table: file.csv) or "
                "name: synthetic" % config.get("name")) from e
