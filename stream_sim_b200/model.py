"""Stream models with the reference's configuration schema and class names
(``stream_sim/model.py``).  Objects are built on the host from the same YAML
sections; ``sample`` / ``complete_catalog`` run on the GPU through
:class:`stream_sim_b200.engine.StarEngine` and return pandas DataFrames with the
reference's columns.

Additive keywords (not in the reference): ``seed`` (counter-based Philox seed;
default: fresh entropy, i.e. irreproducible like the reference's global
MT19937), ``draws`` (dict of injected draws, for parity tests), ``device``.
"""
from __future__ import annotations

import copy
import warnings

import numpy as np
import pandas as pd

from .functions import function_factory
from .samplers import sampler_factory

_ALL_COLUMNS = ("phi1", "phi2", "dist", "mag_g", "mag_r", "mu1", "mu2", "rv")


class ConfigurableModel(object):
    """Common base of the model pieces: owns a private copy of its ``config`` section with the
    keyword overrides merged in, and builds itself from it (``None`` = an absent section:
    nothing is built).  Same contract as the reference's base class (model.py:25-41)."""

    def __init__(self, config, **kwargs):
        self._config = None if config is None else {**copy.deepcopy(config), **kwargs}
        if self._config is not None:
            self._create_model()

    def _create_model(self):
        """Build the sub-models from ``self._config`` (subclasses)."""

    def sample(self, size):
        """Draw ``size`` values (subclasses)."""


class _Parts:
    """Bag of sub-models handed to StarEngine when only a piece is sampled."""

    def __init__(self, density=None, track=None, distance_modulus=None, isochrone=None):
        self.density, self.track = density, track
        self.distance_modulus, self.isochrone = distance_modulus, isochrone
        if self.track is None or not hasattr(self.track, "center"):
            self.track = TrackModel(dict(center=dict(type="Constant", value=0.0),
                                         spread=dict(type="Constant", value=0.0)))


def _engine_for(parts, device=None):
    from .engine import StarEngine
    return StarEngine(stream=parts, device=device)


def _finish(engine, counters):
    summary = engine.summarize(counters)
    engine.raise_for_counters(summary)
    return summary


class StreamModel(ConfigurableModel):
    """High-level stream: density along phi1, track in phi2, distance modulus,
    isochrone photometry, (placeholder) kinematics."""

    def __init__(self, config, **kwargs):
        self._engine = None
        super().__init__(config, **kwargs)

    def _create_model(self):
        self.density = self._create_density()
        self.track = self._create_track()
        self.distance_modulus = self._create_distance_modulus()
        self.isochrone = self._create_isochrone()
        self.velocity = self._create_velocity()

    def _section(self, name, factory):
        cfg = self._config.get(name)
        return factory(cfg) if cfg else None

    def _create_density(self):
        return DensityModel(self._config.get("density"))

    def _create_linear_density(self):
        return DensityModel(self._config.get("linear_density"))

    def _create_track(self):
        return TrackModel(self._config.get("track"))

    def _create_distance_modulus(self):
        return self._section("distance_modulus", TrackModel)

    def _create_isochrone(self):
        cfg = self._config.get("isochrone")
        if not cfg:
            return None
        iso = IsochroneModel(cfg)
        iso.create_isochrone(cfg)
        return iso

    def _create_velocity(self):
        return self._section("velocity", VelocityModel)

    # ------------------------------------------------------------------ device
    def engine(self, device=None):
        """The device-resident tables of this model (built once, cached)."""
        if self._engine is None or (device is not None and str(self._engine.device) != str(device)):
            self._engine = _engine_for(self, device)
        return self._engine

    def sample(self, size, seed=None, draws=None, device=None):
        """Generate ``size`` stars -> DataFrame[phi1, phi2, dist, mu1, mu2, rv, mag_g, mag_r]
        (reference model.py:106-157; absent sub-models give ``None`` columns)."""
        from .engine import _fresh_seed
        n = int(size)
        eng = self.engine(device)
        counters = eng.new_counters()
        out = eng.generate(n, seed=_fresh_seed() if seed is None else seed, draws=draws,
                           counters=counters)
        _finish(eng, counters)
        host = {k: v.cpu().numpy() for k, v in out.items()}
        if self.velocity:
            mu1, mu2, rv = self.velocity.sample(host["phi1"])
        else:
            mu1, mu2, rv = None, None, None
        return pd.DataFrame({"phi1": host["phi1"], "phi2": host["phi2"], "dist": host.get("dist"),
                             "mu1": mu1, "mu2": mu2, "rv": rv,
                             "mag_g": host.get("mag_g"), "mag_r": host.get("mag_r")})

    # ------------------------------------------------------------------ partial fill
    def complete_catalog(self, catalog, columns_to_add=None, size=None, inplace=False,
                         save_path=None, verbose=True, seed=None, device=None):
        """Fill only the requested, missing stream columns of ``catalog``
        (reference model.py:159-351).  Existing finite phi1/phi2/dist values are
        kept; magnitudes (and velocities) are (re)sampled for every row."""
        from .engine import _fresh_seed
        targets = (list(_ALL_COLUMNS) if columns_to_add is None
                   else [c for c in columns_to_add if c in _ALL_COLUMNS])
        unknown = [] if columns_to_add is None else sorted(set(columns_to_add) - set(_ALL_COLUMNS))
        if unknown:
            warnings.warn(f"Ignoring unknown columns: {unknown}")

        def drop(cols, present, message):
            nonlocal targets
            if present is None and any(c in targets for c in cols):
                self._info(verbose, message)
            if present is None:
                targets = [c for c in targets if c not in cols]

        drop(("mag_g", "mag_r"), self.isochrone, "Isochrone model not defined; skipping magnitudes.")
        drop(("mu1", "mu2", "rv"), self.velocity, "Velocity model not defined; skipping velocities.")
        drop(("dist",), self.distance_modulus, "Distance modulus model not defined; skipping distances.")

        df, src_path = self._open_catalog(catalog, size=size, inplace=inplace)
        n_rows = len(df)
        want_mags = any(c in targets for c in ("mag_g", "mag_r"))
        mags_absent = any(c not in df.columns for c in ("mag_g", "mag_r"))
        need = {"phi1": "phi1" in targets, "phi2": "phi2" in targets,
                "dist": "dist" in targets or (want_mags and mags_absent)}
        missing = {c: self._missing_idx(df, c) for c in ("phi1", "phi2", "dist")}
        fill = {c: need[c] and len(missing[c]) > 0 for c in need}

        if fill["phi1"] and getattr(self.density, "density", None) is None:
            raise ValueError("Density model is required to sample phi1")
        phi1_ok_after = ("phi1" in df.columns and not df["phi1"].isna().any()) or \
                        ("phi1" in targets)
        if need["phi2"] and not phi1_ok_after:
            raise ValueError("phi1 required to sample phi2; include 'phi1' in columns_to_add "
                             "or provide it in catalog")
        if fill["dist"] and not phi1_ok_after:
            raise ValueError("phi1 required to sample dist; include 'phi1' in columns_to_add "
                             "or provide it in catalog")
        sample_mags = want_mags and mags_absent
        if want_mags and not mags_absent:
            self._info(verbose, "'mag_g' and 'mag_r' already exist; no sampling performed.")
        if sample_mags and "dist" not in df.columns and not fill["dist"]:
            raise ValueError("dist is required to sample apparent magnitudes; include 'dist' in "
                             "`columns_to_add` or provide it in catalog")

        if any(fill.values()) or sample_mags:
            # one launch: finite catalog entries are kept, NaN entries are sampled
            def col(name):
                if name in df.columns:
                    return df[name].to_numpy(dtype=float, na_value=np.nan)
                return np.full(n_rows, np.nan)
            parts = _Parts(density=self.density, track=self.track,
                           distance_modulus=self.distance_modulus if (fill["dist"]) else None,
                           isochrone=self.isochrone if sample_mags else None)
            eng = _engine_for(parts, device)
            counters = eng.new_counters()
            catalog_in = dict(phi1=col("phi1") if not fill["phi1"] or "phi1" in df.columns else None,
                              phi2=col("phi2") if not fill["phi2"] or "phi2" in df.columns else None,
                              dist=col("dist") if "dist" in df.columns else None)
            if not fill["phi2"]:
                # phi2 is not requested: keep the kernel from evaluating the track on NaN phi1
                catalog_in["phi2"] = np.zeros(n_rows)
            cols = ["phi1", "phi2", "dist"] + (["mag_g", "mag_r"] if sample_mags else [])
            out = eng.allocate(n_rows, cols)
            from . import _lib
            eng.run(_lib.STAGE_GENERATE, n_rows, out, catalog_in, None,
                    _fresh_seed() if seed is None else seed, 0, counters)
            _finish(eng, counters)
            host = {k: v.cpu().numpy() for k, v in out.items()}
            for c in ("phi1", "phi2", "dist"):
                if fill[c]:
                    idx = missing[c]
                    pos = df.index.get_indexer(idx)
                    df.loc[idx, c] = host[c][pos]
                    self._info(verbose, f"Filled {len(idx)} {c} values.")
            if sample_mags:
                if "mag_g" in df.columns or "mag_r" in df.columns:
                    self._info(verbose, "Overwriting existing mag_g and/or mag_r to keep colors "
                                        "consistent.")
                df["mag_g"] = host["mag_g"]
                df["mag_r"] = host["mag_r"]
                self._info(verbose, f"Filled magnitudes for {n_rows} rows.")

        if any(c in targets for c in ("mu1", "mu2", "rv")):
            if any(c in df.columns for c in ("mu1", "mu2", "rv")):
                self._info(verbose, "Velocity components already exist; no sampling performed.")
            else:
                if "phi1" not in df.columns or df["phi1"].isna().any():
                    raise ValueError("phi1 required to sample velocities")
                df["mu1"], df["mu2"], df["rv"] = self.velocity.sample(df["phi1"].to_numpy())
                self._info(verbose, f"Filled velocities for {n_rows} rows.")

        if save_path is not None:
            df.to_csv(save_path, index=False)
            self._info(verbose, f"Saved completed catalog to {save_path}.")
        elif isinstance(catalog, str) and inplace:
            df.to_csv(src_path, index=False)
            self._info(verbose, f"Overwrote original catalog at {src_path}.")
        return df

    def _missing_idx(self, df, col):
        """Index of rows where ``col`` is absent or NaN (reference model.py:353-371)."""
        if col not in df.columns:
            return df.index
        return df.index[df[col].isna()]

    def _info(self, verbose, msg):
        if verbose:
            print(msg)

    def _open_catalog(self, catalog, size=None, inplace=False):
        """Whatever the caller handed over -> (DataFrame with canonical column names, CSV path
        it came from or None).  Accepted, as in the reference (model.py:386-437): nothing
        (``size`` empty rows), a CSV path, a DataFrame (copied unless ``inplace``), a dict of
        columns.  A table without rows is replaced by ``size`` empty rows."""
        def blank():
            if size is None:
                raise ValueError("size must be provided when catalog is None" if catalog is None
                                 else "Empty catalog; provide size")
            return pd.DataFrame(index=np.arange(int(size)))

        origin = catalog if isinstance(catalog, str) else None
        if catalog is None:
            table = blank()
        elif origin is not None:
            table = pd.read_csv(origin)
        elif isinstance(catalog, dict):
            table = pd.DataFrame(catalog)
        elif isinstance(catalog, pd.DataFrame):
            table = catalog if inplace else catalog.copy()
        else:
            raise TypeError("catalog must be None, path, DataFrame, or dict")
        if not len(table):
            table = blank()
        return self._standardize_columns_name(table), origin

    _ALIASES = {
        "dist": ("dist", "distance", "distance_modulus"),
        "mag_g": ("mag_g", "g_mag", "g", "gmag", "magnitude_g"),
        "mag_r": ("mag_r", "r_mag", "r", "rmag", "magnitude_r"),
        "phi1": ("phi1", "phi_1", "Phi1", "Phi_1"),
        "phi2": ("phi2", "phi_2", "Phi2", "Phi_2"),
        "mu1": ("mu1", "mu_1"),
        "mu2": ("mu2", "mu_2"),
        "rv": ("rv", "radial_velocity", "v_radial"),
    }

    def _standardize_columns_name(self, catalog):
        """Rename known column-name variants (reference model.py:439-472; keys are
        lower-cased there, so capitalised variants only match if already lower-case)."""
        lookup = {alias.lower(): std for std, names in self._ALIASES.items() for alias in names}
        return catalog.rename(columns=lookup)


class DensityModel(ConfigurableModel):
    """Density along the stream; samples phi1."""

    def _create_model(self):
        kwargs = copy.deepcopy(self._config)
        type_ = kwargs.pop("type").lower()
        self.density = sampler_factory(type_, **kwargs)

    def sample(self, size, seed=None, draws=None, device=None):
        from .engine import _fresh_seed
        eng = _engine_for(_Parts(density=self), device)
        n = int(np.rint(size))
        out = eng.allocate(n, ["phi1"])
        from . import _lib
        d = None if draws is None else (draws if isinstance(draws, dict) else dict(u_phi1=draws, z_phi1=draws))
        eng.run(_lib.STAGE_GENERATE, n, out, None, d, _fresh_seed() if seed is None else seed)
        return out["phi1"].cpu().numpy()


class TrackModel(ConfigurableModel):
    """Transverse model: a centre and a spread function of phi1, sampled with a
    Gaussian (default) or Uniform law."""

    def _create_model(self):
        kwargs = copy.deepcopy(self._config["center"])
        self.center = function_factory(kwargs.pop("type").lower(), **kwargs)
        kwargs = copy.deepcopy(self._config["spread"])
        self.spread = function_factory(kwargs.pop("type").lower(), **kwargs)

    def _create_sampler(self, x):
        """Host-side frozen sampler at positions ``x`` (kept for API parity with
        reference model.py:513-527; ``sample`` does not go through it)."""
        type_ = self._config.get("sampler", "Gaussian").lower()
        if type_ == "gaussian":
            kwargs = dict(mu=self.center(x), sigma=self.spread(x))
        elif type_ == "uniform":
            xmin = self.center(x) - self.spread(x)
            kwargs = dict(xmin=xmin, xmax=xmin + 2 * self.spread(x))
        else:
            raise Exception(f"Unrecognized sampler: {type_}")
        self._sampler = sampler_factory(type_, **kwargs)

    def sample(self, x, seed=None, draws=None, device=None):
        """Sample the transverse coordinate at phi1 positions ``x`` on the GPU."""
        from .engine import _fresh_seed
        from . import _lib
        x = np.asarray(x, dtype=float)
        eng = _engine_for(_Parts(track=self), device)
        counters = eng.new_counters()
        out = eng.allocate(len(x), ["phi2"])
        d = None if draws is None else (draws if isinstance(draws, dict) else dict(z_phi2=draws, u_phi2=draws))
        eng.run(_lib.STAGE_GENERATE, len(x), out, dict(phi1=x), d,
                _fresh_seed() if seed is None else seed, 0, counters)
        _finish(eng, counters)
        return out["phi2"].cpu().numpy()


class DistanceModel(ConfigurableModel):
    pass


class IsochroneModel(ConfigurableModel):
    """Colour-magnitude sampling from an isochrone + IMF (see ``isochrone.py``)."""

    def __init__(self, config, **kwargs):
        self.table = None
        super().__init__(config, **kwargs)

    def create_isochrone(self, config):
        from .isochrone import IsochroneTable
        self.table = IsochroneTable.from_config(config)
        self.iso = self.table
        if "distance_modulus" in config:
            warnings.warn('Please use the "distance_modulus" section of the configuration file, '
                          "instead of the isochrone section, to define a distance module.")

    def sample(self, nstars, distance_modulus, seed=None, draws=None, device=None, **kwargs):
        """(mag_g, mag_r) of ``nstars`` stars at the given distance modulus
        (scalar or per star), reference model.py:575-604."""
        from .engine import _fresh_seed
        from . import _lib
        n = int(round(nstars))
        dm = np.broadcast_to(np.asarray(distance_modulus, dtype=float), (n,))
        eng = _engine_for(_Parts(isochrone=self), device)
        counters = eng.new_counters()
        out = eng.allocate(n, ["mag_g", "mag_r"])
        d = None if draws is None else (draws if isinstance(draws, dict) else dict(u_mass=draws))
        eng.run(_lib.STAGE_GENERATE, n, out, dict(phi1=np.zeros(n), phi2=np.zeros(n), dist=dm), d,
                _fresh_seed() if seed is None else seed, 0, counters)
        _finish(eng, counters)
        return out["mag_g"].cpu().numpy(), out["mag_r"].cpu().numpy()

    def _dist_to_modulus(self, distance):
        """Distance in pc -> distance modulus (unused by the pipeline, as upstream)."""
        if distance is None:
            return 0
        if np.all(distance == 0):
            warnings.warn("Distances are equal to 0, distance modulus has been set to 0.")
            return 0
        return 5 * np.log10(distance) - 5


class VelocityModel(ConfigurableModel):
    """Placeholder, as in the reference (model.py:621-633): NaN kinematics."""

    def sample(self, phi1):
        warnings.warn("VelocityModel not implemented!")
        if np.isscalar(phi1):
            return np.nan, np.nan, np.nan
        return tuple(np.nan * np.ones_like([phi1, phi1, phi1]))


class BackgroundModel(StreamModel):
    """Background model."""
    pass


class SplineStreamModel(StreamModel):
    """Spline-defined stream: ``linear_density`` + spline track, all keyed by
    ``stream_name``.  (Upstream's version cannot be instantiated -- it calls an
    undefined ``_create_distance``, model.py:670; here the distance-modulus
    section is honoured.)"""

    def __init__(self, config, **kwargs):
        stream_name = config["stream_name"] if config["stream_name"] else None
        config["linear_density"]["stream_name"] = stream_name
        config["track"]["center"]["stream_name"] = stream_name
        config["track"]["spread"]["stream_name"] = stream_name
        super().__init__(config, **kwargs)

    def _create_model(self):
        self.density = self._create_linear_density()
        self.track = self._create_track()
        self.distance_modulus = self._create_distance_modulus()
        self.isochrone = self._create_isochrone()
        self.velocity = self._create_velocity()
