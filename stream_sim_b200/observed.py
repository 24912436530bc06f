"""Survey injection with the reference's API (``stream_sim/observed.py``).

``StreamInjector.inject`` keeps the reference's signature, keywords, column
names, messages and errors; the per-star work -- stream-frame -> ICRS rotation,
HEALPix ``ang2pix``, map gathers, photometric-error model, flux-space noise,
completeness / detection Bernoulli draws, flag combination (reference
:146-291) -- is ONE fused CUDA kernel launch (plus the great-circle frame search
when the stream has to be placed on the sky, one small launch per trial).

Additive keywords of ``inject`` (all optional):
  ``rng_compat``  True  -> draw the flux-noise normals and detection uniforms on
                           the host from ``np.random.default_rng(seed)`` in the
                           reference's order and inject them: outputs are then
                           bit-compatible with the reference for the same seed.
                  False -> (default) counter-based Philox4x32-10 on the device.
  ``bad_mag``     "string" (default, reference behaviour: object column holding
                  floats and the sentinel "BAD_MAG") or "nan" (float column).
  ``device``      CUDA device.
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd

from . import _lib
from . import healpix as hp
from .frames import GreatCircleFrame, SkyPoint, SkyPositions
from .model import StreamModel
from .surveys import Survey


class StreamInjector:
    """Inject observational effects of a survey into stream catalogs."""

    mask_cache = {}

    def __init__(self, survey, **kwargs):
        """``survey``: a ready :class:`Survey`, or the name of one to load (keywords go to
        ``Survey.load``), as in the reference (observed.py:38-46)."""
        if not isinstance(survey, (str, Survey)):
            raise ValueError("survey must be a string or Survey instance.")
        self.survey = Survey.load(survey=survey, **kwargs) if isinstance(survey, str) else survey
        self._last_gc_frame = None       # what gc_frame="last" refers to
        self._engines = {}               # device engines, one per injection configuration

    # ------------------------------------------------------------------ engine cache
    def _engine(self, bands, nside, dust_correction, perfect_galstarsep, detection_mag_cut,
                stream=None, device=None):
        from .engine import StarEngine
        key = (tuple(bands), int(nside), bool(dust_correction), bool(perfect_galstarsep),
               tuple(detection_mag_cut), id(stream), str(device))
        if key not in self._engines:
            self._engines[key] = StarEngine(
                stream=stream, survey=self.survey, device=device, bands=tuple(bands), nside=nside,
                dust_correction=dust_correction, perfect_galstarsep=perfect_galstarsep,
                detection_mag_cut=tuple(detection_mag_cut))
        return self._engines[key]

    # ------------------------------------------------------------------ inject
    def inject(self, data, bands=["r", "g"], **kwargs):
        """Add observed quantities to ``data`` (reference observed.py:80-298).

        Returns the input columns plus ``ra, dec`` (if absent), sampled ``dist,
        mag_g, mag_r`` (if absent), ``mag_<band>_obs, magerr_<band>`` per band in
        ``bands`` order, ``flag_observed`` and, with ``perfect_galstarsep=True``,
        ``flag_perfect_galstarsep``.

        ``output`` (not in the reference): ``"pandas"`` (default) the reference's DataFrame;
        ``"table"`` a :class:`stream_sim_b200.table.StarTable` whose new columns stay on the
        GPU until used (no per-row Python objects; bad magnitudes are NaN / nulls);
        ``"arrow"`` a ``pyarrow.Table``; ``"parquet"`` writes ``path`` and returns it.
        """
        verbose = kwargs.get("verbose", True)
        perfect_galstarsep = kwargs.pop("perfect_galstarsep", False)
        rng_compat = kwargs.pop("rng_compat", False)
        bad_mag = kwargs.pop("bad_mag", "string")
        device = kwargs.pop("device", None)
        output = kwargs.pop("output", "pandas")
        path = kwargs.pop("path", None)
        if output not in ("pandas", "table", "arrow", "parquet"):
            raise ValueError("output must be 'pandas', 'table', 'arrow' or 'parquet'")
        if output == "parquet" and not path:
            raise ValueError("output='parquet' needs path=...")
        data = self._load_data(data)
        seed = kwargs.pop("seed", None)
        rng = np.random.default_rng(seed)

        for band in bands:
            if band not in ["r", "g"]:
                raise ValueError("Currently only 'r' and 'g' bands are supported.")
        if self.survey.completeness_band not in bands:
            raise ValueError(f"Detection flag requires '{self.survey.completeness_band}' band "
                             "to be in bands.")

        data = self.complete_data(data, rng=rng, seed=seed, bands=bands, device=device, **kwargs)
        nside = kwargs.pop("nside", 4096)
        dust_correction = kwargs.get("dust_correction", True)
        detection_mag_cut = kwargs.get("detection_mag_cut", ["g"])
        n = len(data)

        draws = None
        if rng_compat:
            # PCG64 stream in the reference's order (observed.py:154-243): per band in
            # `bands`: N normals; after the completeness band: N uniforms (+N with
            # perfect_galstarsep)
            draws = {}
            for band in bands:
                draws["z_" + band] = rng.standard_normal(n)
                if band == self.survey.completeness_band:
                    draws["u_det"] = rng.uniform(size=n)
                    if perfect_galstarsep:
                        draws["u_det2"] = rng.uniform(size=n)
        from .engine import _fresh_seed
        eng = self._engine(bands, nside, dust_correction, perfect_galstarsep, detection_mag_cut,
                           device=device)
        counters = eng.new_counters()
        catalog = dict(ra=data["ra"].to_numpy(dtype=float), dec=data["dec"].to_numpy(dtype=float))
        for band in bands:
            catalog["mag_" + band] = data["mag_" + band].to_numpy(dtype=float)
        out = eng.observe(catalog, seed=_fresh_seed() if seed is None else seed, draws=draws,
                          counters=counters)
        eng.raise_for_counters(eng.summarize(counters))
        if output != "pandas":
            return self._columnar_result(data, out, bands, perfect_galstarsep, output, path)
        host = {k: v.cpu().numpy() for k, v in out.items()}

        data = data.reset_index(drop=True)
        new_cols = {}
        for band in bands:
            if dust_correction and verbose:
                print(f"Applying dust correction for {band}-band on observed magnitudes.")
            mobs = host[f"mag_{band}_obs"]
            if bad_mag == "string":
                col = mobs.astype(object)
                col[np.isnan(mobs)] = "BAD_MAG"
                mobs = col
            new_cols[f"mag_{band}_obs"] = mobs
            new_cols[f"magerr_{band}"] = host[f"magerr_{band}"]
        data = pd.concat([data, pd.DataFrame(new_cols)], axis=1)
        for band in detection_mag_cut:
            if band not in bands:
                if verbose:
                    print(f"Warning: SNR cut requested for {band}-band but it's not in bands "
                          "list. Skipping.")
                continue
            if verbose:
                print(f"Applying detection cut on {band}-band with SNR >= 5.0")
        data["flag_observed"] = host["flag_observed"].astype(bool)
        if perfect_galstarsep:
            data["flag_perfect_galstarsep"] = host["flag_perfect_galstarsep"].astype(bool)
        if kwargs.get("save"):
            self._save_injected_data(data, kwargs.get("folder", None))
        return data

    @staticmethod
    def _columnar_result(data, out, bands, perfect_galstarsep, output, path):
        """The injection as a StarTable: the catalog's columns as they are, the kernel's
        columns still on the device; same column order as the DataFrame."""
        from .table import StarTable
        cols = {c: data[c].to_numpy() for c in data.columns}
        bad = []
        for band in bands:
            cols[f"mag_{band}_obs"] = out[f"mag_{band}_obs"]
            cols[f"magerr_{band}"] = out[f"magerr_{band}"]
            bad.append(f"mag_{band}_obs")
        cols["flag_observed"] = out["flag_observed"]
        if perfect_galstarsep:
            cols["flag_perfect_galstarsep"] = out["flag_perfect_galstarsep"]
        table = StarTable(cols, bad_mag_columns=bad)
        if output == "table":
            return table
        if output == "arrow":
            return table.to_arrow()
        return table.to_parquet(path)

    def _load_data(self, data):
        if isinstance(data, pd.DataFrame):
            return data
        if isinstance(data, str):
            extension = data.split(".")[-1]
            if extension == "csv":
                return pd.read_csv(data)
            if extension in ["xls", "xlsx"]:
                return pd.read_excel(data)
            raise ValueError(f"Unsupported file format: {extension}")
        raise ValueError("Unsupported file format")

    # ------------------------------------------------------------------ completion
    def complete_data(self, data, bands=["r", "g"], **kwargs):
        """Make sure ``ra, dec`` and ``mag_<band>`` exist (reference :332-476):
        rotate (phi1, phi2) onto the sky if needed, sample missing magnitudes from
        ``stream_config`` if given.  Returns a copy."""
        required = ["ra", "dec"] + [f"mag_{band}" for band in bands]
        rng = kwargs.pop("rng", None)
        seed = kwargs.pop("seed", None)
        device = kwargs.pop("device", None)
        data = data.copy()
        if not ("ra" in data.columns and "dec" in data.columns):
            if "phi1" not in data.columns or "phi2" not in data.columns:
                raise ValueError("Input data must contain either (ra, dec) or (phi1, phi2) columns.")
            if kwargs.get("gc_frame", None) == "last":
                kwargs["gc_frame"] = self._last_gc_frame
            sky = self.phi_to_radec(data["phi1"], data["phi2"], seed=seed, rng=rng, device=device,
                                    **kwargs)
            data.loc[:, "ra"] = sky.icrs.ra.deg
            data.loc[:, "dec"] = sky.icrs.dec.deg
        missing = [f"mag_{band}" for band in bands if f"mag_{band}" not in data.columns]
        if missing:
            stream_config = kwargs.get("stream_config", None)
            if stream_config is None:
                raise ValueError("`stream_config` must be provided to sample missing magnitudes.")
            data = StreamModel(stream_config).complete_catalog(
                data, inplace=True, columns_to_add=missing, seed=seed, device=device)
        for col in required:
            if col not in data.columns:
                raise ValueError(f"Input data must contain '{col}' column.")
        return data

    def phi_to_radec(self, phi1, phi2, gc_frame=None, seed=None, rng=None,
                     mask_type=["footprint"], device=None, **kwargs):
        """Stream coordinates (deg) -> ICRS sky positions (reference :478-575)."""
        from .engine import StarEngine
        phi1_arr = np.asarray(phi1, dtype=float)
        phi2_arr = np.asarray(phi2, dtype=float)
        if phi1_arr.size == 0 or phi2_arr.size == 0:
            raise ValueError("phi1 and phi2 cannot be empty arrays")
        if isinstance(gc_frame, str) and gc_frame == "last":
            gc_frame = self._last_gc_frame
        if gc_frame is None:
            gc_frame = self._find_gc_frame(rng=rng, seed=seed, mask_type=mask_type, phi1=phi1_arr,
                                           phi2=phi2_arr, device=device, **kwargs)
            if gc_frame is None:
                raise RuntimeError("No suitable great circle frame could be found.")
        gc_frame = GreatCircleFrame.coerce(gc_frame)
        self._last_gc_frame = gc_frame
        eng = StarEngine(frame=gc_frame, device=device)
        out = eng.rotate(phi1_arr.ravel(), phi2_arr.ravel())
        return SkyPositions(out["ra"].cpu().numpy(), out["dec"].cpu().numpy(), frame=gc_frame,
                            phi1=phi1_arr, phi2=phi2_arr)

    def _find_gc_frame(self, phi1=None, phi2=None, mask=None, mask_type=["footprint"],
                       percentile_threshold=0.99, max_iter=1000, rng=None, seed=None,
                       verbose=True, device=None, **kwargs):
        """Random great-circle frames until ``percentile_threshold`` of the stars
        fall inside the mask (reference :577-679).  Endpoints are drawn on the host
        from ``rng`` in the reference's order (4 scalars per trial), the fraction
        inside the mask is one ``ss_frame_fraction`` launch per trial."""
        import ctypes as C
        import torch
        from .engine import StarEngine, require_device
        say = print if verbose else (lambda *a, **k: None)
        rng = np.random.default_rng(seed) if rng is None else rng

        def random_frame():              # two endpoints = 4 scalars of the stream, in this order
            return GreatCircleFrame.from_endpoints(self._random_uniform_skycoord(rng),
                                                   self._random_uniform_skycoord(rng))

        if mask is not None:
            say("Using provided HEALPix mask for footprint checking.")
            healpix_mask = mask
        else:
            healpix_mask = self._create_mask(mask_type, verbose=verbose)
        if healpix_mask is None:         # nothing to test against: any frame will do
            say("No mask available, returning a random great circle frame.")
            return random_frame()
        if phi1 is None or phi2 is None:
            raise ValueError("phi1 and phi2 must be provided if no mask is given.")
        device = require_device(device)
        lib = _lib.load()
        eng = StarEngine(device=device)
        mask_t = torch.from_numpy(np.ascontiguousarray(healpix_mask, dtype=np.uint8)).to(device)
        maps = _lib.SSMaps()
        maps.mask, maps.nside_mask = mask_t.data_ptr(), hp.get_nside(healpix_mask)
        cat = _lib.SSCatalog()
        p1 = torch.from_numpy(np.ascontiguousarray(phi1, dtype=np.float64).ravel()).to(device)
        p2 = torch.from_numpy(np.ascontiguousarray(phi2, dtype=np.float64).ravel()).to(device)
        cat.phi1, cat.phi2 = p1.data_ptr(), p2.data_ptr()
        n = p1.numel()
        counters = eng.new_counters()
        trials = 0
        while trials < max_iter:
            trials += 1
            frame = random_frame()
            eng.set_frame(frame)
            counters.zero_()
            with torch.cuda.device(device):
                rc = lib.ss_frame_fraction(C.byref(eng.params), C.byref(maps), C.byref(cat),
                                           C.c_void_p(counters.data_ptr()), n,
                                           C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
            _lib.check(rc)
            frac = int(counters[_lib.CNT_INSIDE]) / n
            if frac >= percentile_threshold:
                if verbose:
                    print(f"Found suitable great circle frame after {trials} trials with "
                          f"{frac*100:.2f}% points inside the mask.")
                return frame
        if verbose:
            print(f"Could not find a suitable great circle frame after {max_iter} trials.")
        return None

    def _random_uniform_skycoord(self, rng):
        """One point uniform on the sphere: z~U(-1,1), ra~U(0,360) (reference :681-698)."""
        z = rng.uniform(-1.0, 1.0)
        dec = np.degrees(np.arcsin(z))
        ra = rng.uniform(0.0, 360.0)
        return SkyPoint(ra, dec)

    def _fraction_inside_mask(self, ra_deg, dec_deg, healpix_mask):
        pix = hp.ang2pix(hp.get_nside(healpix_mask), ra_deg, dec_deg, lonlat=True)
        return np.count_nonzero(np.asarray(healpix_mask)[pix]) / len(pix)

    def _create_mask(self, mask_type, verbose=True, ebv_threshold=0.2):
        """Boolean sky mask combining footprint / depth / dust cuts at the coarsest
        resolution involved, cached per (survey, types, threshold) (reference :720-861)."""
        say = print if verbose else (lambda *a, **k: None)
        if mask_type is None:
            say("⚠ No mask_type provided to build mask.")
            return None
        if not isinstance(mask_type, (str, list)):
            raise ValueError("mask_type must be a string, list of strings, or None.")
        # one canonical spelling per combination, so that ["ebv", "footprint"] and
        # ["footprint", "ebv"] share a cache entry; the dust threshold only matters with "ebv"
        mask_type = sorted([mask_type] if isinstance(mask_type, str) else mask_type)
        key = (getattr(self.survey, "name", "unknown"), tuple(mask_type),
               ebv_threshold if "ebv" in mask_type else None)
        cached = self.mask_cache.get(key)
        if cached is not None:
            say(f"✓ Using cached mask for {mask_type}")
            return cached
        say(f"Building new mask for {mask_type}...")

        def host(m):
            return m.cpu().numpy() if hasattr(m, "cpu") else np.asarray(m)

        maps = {}
        for m in mask_type:
            if "maglim" in m:
                band = m.split("_")[-1]
                if band not in self.survey.maglim_maps:
                    raise ValueError(f"Band '{band}' not found in survey magnitude limit maps. "
                                     f"Available: {list(self.survey.maglim_maps.keys())}")
                maps[m] = host(self.survey.maglim_maps[band])
            elif m in ["coverage", "footprint"]:
                if self.survey.coverage is None:
                    raise ValueError("Survey coverage map is not available.")
                maps[m] = host(self.survey.coverage)
            elif m == "ebv":
                if self.survey.ebv_map is None:
                    raise ValueError("Survey E(B-V) extinction map is not available.")
                maps[m] = host(self.survey.ebv_map)
            else:
                raise ValueError(f"Unknown mask type: '{m}'. Valid options: 'footprint', "
                                 "'coverage', 'maglim_<band>', 'ebv'")
        if not maps:
            raise ValueError(f"No valid maps found for mask_type: {mask_type}")
        nside_min = min(hp.get_nside(v) for v in maps.values())
        for m in maps:
            if hp.get_nside(maps[m]) != nside_min:
                if verbose:
                    print(f"  Resampling {m} from nside={hp.get_nside(maps[m])} to nside={nside_min}")
                maps[m] = hp.ud_grade(maps[m], nside_min)
        mask_map = np.ones(hp.nside2npix(nside_min), dtype=bool)
        with np.errstate(invalid="ignore"):
            for m in mask_type:
                if "maglim" in m:
                    band = m.split("_")[-1]
                    if band in self.survey.saturation:
                        mask_map &= maps[m] > self.survey.saturation[band]
                    else:
                        mask_map &= maps[m] > 0
                elif m in ["coverage", "footprint"]:
                    mask_map &= maps[m] > 0.5
                elif m == "ebv":
                    mask_map &= maps[m] < ebv_threshold
        self.mask_cache[key] = mask_map
        if verbose:
            print(f"✓ Mask created: valid pixels fraction = {np.sum(mask_map) / len(mask_map):.1f}")
            print(f"  Cached with key: {key}")
        return mask_map

    # ------------------------------------------------------------------ pieces of inject
    def sample_measured_magnitudes(self, mag_true, mag_err, **kwargs):
        """Noisy magnitudes from flux-space Gaussian noise (reference :863-904);
        "BAD_MAG" where the noisy flux is not positive.  One launch of
        ``ss_noisy_magnitudes``; the normals come from ``rng`` (NumPy Generator: the
        reference's stream), from ``z=`` or, by default, from the device's Philox stream
        (``seed=``)."""
        from .engine import noisy_magnitudes
        z = kwargs.pop("z", None)
        rng = kwargs.pop("rng", None)
        if z is None and rng is not None:
            z = rng.standard_normal(np.shape(mag_true))
        mobs = noisy_magnitudes(np.asarray(mag_true, dtype=float), np.asarray(mag_err, dtype=float),
                                z=z, seed=kwargs.pop("seed", None),
                                device=kwargs.pop("device", None))
        mobs = mobs.reshape(np.shape(mag_true))
        good = ~np.isnan(mobs)
        return np.where(good, mobs, "BAD_MAG")

    def detect_flag(self, pix, mag=None, band="r", **kwargs):
        """Bernoulli detection against the survey completeness at ``pix``
        (reference :906-959): the efficiency getter and the trial (``ss_bernoulli``)
        both run on the GPU; uniforms from ``rng`` / ``u=`` or the device's Philox
        stream (``seed=``)."""
        from .engine import bernoulli
        u = kwargs.pop("u", None)
        rng = kwargs.pop("rng", None)
        if u is None and rng is not None:
            u = rng.uniform(size=len(mag))
        maglim = self.survey.get_maglim(band, pixel=pix)
        if kwargs.get("perfect_galstarsep", False):
            compl = self.survey.get_detection_efficiency(band, mag, maglim)
        else:
            compl = self.survey.get_completeness(band, mag, maglim)
        return bernoulli(compl, u=u, seed=kwargs.pop("seed", None), device=kwargs.pop("device", None))

    def _save_injected_data(self, data, folder):
        if folder is None:
            here = os.path.dirname(os.path.abspath(__file__))
            folder = os.path.join(here, "..", "data/outputs/")
        if not os.path.exists(folder):
            os.makedirs(folder)
        file_name = os.path.join(folder, "data_injected.csv")
        print(f"Saving injected data to {file_name}")
        data.to_csv(file_name, index=False)

    @staticmethod
    def magToFlux(mag):
        """AB magnitude -> flux in Jy."""
        return 3631.0 * 10 ** (-0.4 * mag)

    @staticmethod
    def fluxToMag(flux):
        """Flux in Jy -> AB magnitude."""
        return -2.5 * np.log10(flux / 3631.0)

    @staticmethod
    def getFluxError(mag, mag_error):
        """Magnitude error -> flux error in Jy."""
        return StreamInjector.magToFlux(mag) * mag_error / 1.0857362

    def plot_stream_in_mask(self, data, mask_type, ebv_threshold=0.2, **kwargs):
        """Stars over the footprint / depth / dust mask (reference :1078-1134).  The mask is
        built exactly as for the frame search; drawing needs matplotlib, which is outside
        the GPU path (``plotting.py`` is out of scope): without it this
        raises ImportError after the same ValueError checks as the reference.  Returns
        ``(fig, ax)``; ``output_folder=`` saves ``stream_in_mask.png`` there."""
        mask = self._create_mask(mask_type, verbose=False, ebv_threshold=ebv_threshold)
        if mask is None:
            raise ValueError("Could not create mask. Check mask_type parameter.")
        try:
            import matplotlib.pyplot as plt
        except ImportError as e:
            raise ImportError("plot_stream_in_mask needs matplotlib, which is not installed; "
                              "the mask itself is available from _create_mask()") from e
        ra = np.asarray(data["ra"], dtype=float)
        dec = np.asarray(data["dec"], dtype=float)
        pad = 5.0
        lon = np.linspace(max(ra.min() - pad, 0.0), min(ra.max() + pad, 360.0), 600)
        lat = np.linspace(max(dec.min() - pad, -90.0), min(dec.max() + pad, 90.0), 400)
        glon, glat = np.meshgrid(lon, lat)
        pix = hp.ang2pix(hp.get_nside(mask), glon.ravel(), glat.ravel())
        inside = np.asarray(mask)[np.asarray(pix)].reshape(glon.shape)
        fig, ax = plt.subplots(figsize=(8, 5))
        ax.imshow(inside, origin="lower", aspect="auto", cmap="Greys", alpha=0.35,
                  extent=(lon[0], lon[-1], lat[0], lat[-1]))
        ax.scatter(ra, dec, s=1.0, color="tab:red")
        ax.set_xlabel("RA [deg]")
        ax.set_ylabel("Dec [deg]")
        ax.invert_xaxis()
        folder = kwargs.get("output_folder")
        if folder:
            os.makedirs(folder, exist_ok=True)
            fig.savefig(os.path.join(folder, "stream_in_mask.png"), dpi=150)
        return fig, ax

    @classmethod
    def clear_mask_cache(cls):
        cls.mask_cache.clear()
        print("✓ Mask cache cleared")

    @classmethod
    def list_cached_masks(cls):
        if not cls.mask_cache:
            print("No masks cached")
            return []
        print(f"Cached masks ({len(cls.mask_cache)}):")
        for survey_name, mask_types, ebv_thresh in cls.mask_cache.keys():
            extra = f", ebv_threshold={ebv_thresh}" if ebv_thresh is not None else ""
            print(f"  - {survey_name}: {list(mask_types)}{extra}")
        return list(cls.mask_cache.keys())
