"""Survey container, per-star getters and loader with the reference's names
(``stream_sim/surveys.py``).

The ``Survey`` dataclass holds dense RING HEALPix maps (NaN = unseen) and the
tabulated efficiency / photo-error curves.  The per-star getters
(``get_maglim``, ``get_extinction``, ``get_photo_error``, ``get_efficiency`` and
wrappers; reference :173-493) evaluate on the GPU.  ``SurveyFactory`` reproduces
the YAML-driven loader and cache (:496-1490) for the file formats that can be
read without healpy/healsparse (``.npy`` maps, CSV curves; HEALPix FITS and HealSparse
``.hsp`` maps through the native readers in ``fitsio.py``).
"""
from __future__ import annotations

import copy
import os
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from . import _lib
from . import healpix as hp


class TabulatedFunction:
    """Piecewise-linear curve with constant fill outside its nodes: the object
    ``set_completeness`` / ``set_photo_error`` return (the reference returns a
    ``scipy.interpolate.interp1d``; same ``x``, ``y``, ``fill_value`` attributes,
    same call semantics).  Host calls are for set-up scalars and plotting; the
    per-star evaluation is the CUDA ``lin_eval``."""

    def __init__(self, x, y, fill_value):
        x = np.asarray(x, dtype=float)
        y = np.asarray(y, dtype=float)
        order = np.argsort(x, kind="mergesort")
        self.x, self.y = x[order], y[order]
        self.fill_value = float(fill_value)

    def __call__(self, x_new):
        x_new = np.asarray(x_new, dtype=float)
        y = np.interp(x_new, self.x, self.y)
        return np.where((x_new < self.x[0]) | (x_new > self.x[-1]), self.fill_value, y)

    def describe(self, packer):
        f = _lib.SSFunc()
        f.kind = _lib.FN_LINEAR
        for k, v in packer.add_linear(self.x, self.y).items():
            setattr(f, k, v)
        f.xmin, f.xmax = -np.inf, np.inf
        f.p[0] = self.fill_value
        for k, v in packer.add_direct_lookup(self.x, self.y, self.fill_value).items():
            setattr(f, k, v)
        return f

    @classmethod
    def coerce(cls, obj):
        """Accept a TabulatedFunction or an interp1d-like object (x, y, fill_value)."""
        if obj is None or isinstance(obj, cls):
            return obj
        if hasattr(obj, "x") and hasattr(obj, "y"):
            fill = np.atleast_1d(np.asarray(getattr(obj, "fill_value", np.nan), dtype=float))
            return cls(obj.x, obj.y, fill.flat[0])
        raise TypeError("efficiency / photo-error curves must be tabulated (x, y, fill_value)")


@dataclass
class Survey:
    """Survey properties and data; fields as in the reference (surveys.py:95-116),
    plus the detection-only / classification-only curves that upstream attaches
    dynamically (surveys.py:1247)."""

    name: str
    release: Optional[str] = None
    bands: list = None
    maglim_maps: dict = None
    coeff_extinc: dict = None
    saturation: dict = None
    sys_error: dict = None
    completeness: Optional[Callable] = None
    completeness_band: Optional[str] = None
    delta_saturation: Optional[float] = None
    log_photo_error: Optional[Callable] = None
    ebv_map: Optional[np.ndarray] = None
    coverage: Optional[np.ndarray] = None
    efficiency_detection: Optional[Callable] = None
    efficiency_classification: Optional[Callable] = None

    def __post_init__(self):
        self.bands = [] if self.bands is None else self.bands
        self.maglim_maps = {} if self.maglim_maps is None else self.maglim_maps
        self.coeff_extinc = {} if self.coeff_extinc is None else self.coeff_extinc
        self.saturation = {} if self.saturation is None else self.saturation
        self.sys_error = {} if self.sys_error is None else self.sys_error

    @classmethod
    def load(cls, survey, release=None, config_file=None, **kwargs):
        """Load (or fetch from cache) a survey -- ``SurveyFactory.create_survey``."""
        return SurveyFactory.create_survey(survey, release, config_file, **kwargs)

    @classmethod
    def synthetic(cls, nside_maglim=128, nside_ebv=512, device=None, name="synthetic",
                  release=None, completeness_band="r", first=-10.4):
        """A fully synthetic LSST-like survey (see ``synthetic.py``); with
        ``device`` the maps are generated directly in HBM as torch tensors."""
        from . import synthetic as syn
        props = syn.LSST_PROPS
        eff = syn.efficiency_csv_columns(first=first)
        err = syn.photo_error_csv_columns(first=first)
        dsat = props["delta_saturation"]
        sv = cls(name=name, release=release, bands=["g", "r"],
                 maglim_maps=dict(g=syn.healpix_map("maglim_g", nside_maglim, device),
                                  r=syn.healpix_map("maglim_r", nside_maglim, device)),
                 coeff_extinc=dict(props["coeff_extinc"]), saturation=dict(props["saturation"]),
                 sys_error=dict(props["sys_error"]), completeness_band=completeness_band,
                 delta_saturation=dsat, ebv_map=syn.healpix_map("ebv", nside_ebv, device))
        sv.completeness = SurveyFactory.completeness_from_arrays(
            eff["delta_mag"], eff["classification_detection_eff"], dsat)
        sv.efficiency_detection = SurveyFactory.completeness_from_arrays(
            eff["delta_mag"], eff["detection_eff"], dsat)
        sv.efficiency_classification = SurveyFactory.completeness_from_arrays(
            eff["delta_mag"], eff["classifiction_eff"], dsat)
        sv.log_photo_error = SurveyFactory.photo_error_from_arrays(
            err["delta_mag"], err["log_mag_err"], dsat)
        return sv

    # ------------------------------------------------------------------ getters
    def _gather(self, m, pixel):
        import torch
        if isinstance(m, torch.Tensor):
            idx = torch.as_tensor(np.asarray(pixel), device=m.device)
            return m[idx].cpu().numpy() if idx.ndim else float(m[idx])
        return m[pixel]

    def get_maglim(self, band, pixel=None):
        """Magnitude limit map of ``band`` (or its values at ``pixel``)."""
        if band not in self.maglim_maps:
            raise ValueError(f"Band '{band}' not available. Available: {self.bands}")
        if pixel is None:
            return self.maglim_maps[band]
        return self._gather(self.maglim_maps[band], pixel)

    def get_extinction(self, band, pixel=None):
        """A_band = coeff * E(B-V).  Only the requested pixels are multiplied (the
        reference scales the whole map per call, surveys.py:228; same values)."""
        if band not in self.coeff_extinc:
            raise ValueError(f"Band '{band}' not available. Available: {self.bands}")
        if self.ebv_map is None:
            raise ValueError("E(B-V) map not loaded")
        if pixel is None:
            return self.coeff_extinc[band] * self.ebv_map
        return self.coeff_extinc[band] * self._gather(self.ebv_map, pixel)

    def _curve_call(self, which, band, magnitude, maglim, delta_saturation):
        """Run ss_photo_error / ss_efficiency on the GPU for arbitrary array shapes."""
        import ctypes as C
        import torch
        from .engine import BlobPacker, require_device
        if band not in self.saturation:
            raise KeyError(band)
        device = require_device()
        lib = _lib.load()
        scalar = np.isscalar(magnitude) and np.isscalar(maglim)
        mag, ml = np.broadcast_arrays(np.asarray(magnitude, dtype=float),
                                      np.asarray(maglim, dtype=float))
        shape = mag.shape
        p = _lib.SSParams()
        packer = BlobPacker()
        b = 0
        p.saturation[b] = float(self.saturation[band])
        p.sys_error[b] = float(self.sys_error.get(band, 0.005))
        p.delta_saturation = float(delta_saturation)
        if which == "photo_error":
            curve = TabulatedFunction.coerce(self.log_photo_error)
            p.photo_error = curve.describe(packer)
            p.log_err_at_dsat = float(curve(delta_saturation))
            p.err_bright = float(10 ** curve(delta_saturation - 1))
        else:
            p.completeness = TabulatedFunction.coerce(which).describe(packer)
        blob = packer.finalize()
        blob_t = torch.from_numpy(blob).to(device)
        t = _lib.SSTables()
        t.blob, t.blob_bytes = blob_t.data_ptr(), blob_t.numel()
        mag_t = torch.from_numpy(np.ascontiguousarray(mag).ravel()).to(device)
        ml_t = torch.from_numpy(np.ascontiguousarray(ml).ravel()).to(device)
        out = torch.empty_like(mag_t)
        stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        with torch.cuda.device(device):
            if which == "photo_error":
                rc = lib.ss_photo_error(C.byref(p), C.byref(t), b, mag_t.data_ptr(),
                                        ml_t.data_ptr(), out.data_ptr(), mag_t.numel(), stream)
            else:
                rc = lib.ss_efficiency(C.byref(p), C.byref(t), b, 0, mag_t.data_ptr(),
                                       ml_t.data_ptr(), out.data_ptr(), mag_t.numel(), stream)
        _lib.check(rc)
        res = out.cpu().numpy().reshape(shape)
        return float(res) if scalar else res

    def get_photo_error(self, band, magnitude, maglim, **kwargs):
        """Total photometric error (statistical (+) systematic), reference :234-292."""
        if self.log_photo_error is None:
            raise ValueError("Photo error model not loaded")
        dsat = kwargs.get("delta_saturation", self.delta_saturation)
        return self._curve_call("photo_error", band, magnitude, maglim, dsat)

    def get_completeness(self, band, magnitude, maglim, **kwargs):
        return self.get_efficiency(band, magnitude, maglim, type="completeness", **kwargs)

    def get_detection_efficiency(self, band, magnitude, maglim, **kwargs):
        return self.get_efficiency(band, magnitude, maglim, type="detection_efficiency", **kwargs)

    def get_classification_efficiency(self, band, magnitude, maglim, **kwargs):
        return self.get_efficiency(band, magnitude, maglim, type="classification_efficiency",
                                   **kwargs)

    def get_efficiency(self, band, magnitude, maglim, type="completeness", **kwargs):
        """Selection efficiency in [0, 1] (reference :399-493): 0 for saturated
        sources and unseen / too-shallow pixels, 1 between saturation and the
        bright end of the tabulated curve, the curve elsewhere."""
        if type == "completeness":
            func = self.completeness
        elif type == "detection_efficiency":
            func = self.efficiency_detection
        elif type == "classification_efficiency":
            func = self.efficiency_classification
        else:
            raise ValueError(f"Efficiency type '{type}' not recognized.")
        if func is None:
            raise ValueError(f"Efficiency function for type '{type}' not loaded")
        dsat = kwargs.get("delta_saturation", self.delta_saturation)
        return self._curve_call(func, band, magnitude, maglim, dsat)


class SurveyFactory:
    """Creates, populates and caches ``Survey`` objects (reference :496-1490)."""

    _cached_surveys = {}

    @classmethod
    def create_survey(cls, survey, release=None, config_file=None, **kwargs):
        verbose = kwargs.get("verbose", True)
        uniform_survey = kwargs.pop("uniform_survey", False)
        key = f"{survey}_{release}" if release else survey
        if key in cls._cached_surveys:
            if verbose:
                print(f"✓ Using cached survey data for '{key}'")
        else:
            if verbose:
                print(f"Loading survey data for '{key}'...")
            config = copy.deepcopy(config_file) if config_file else cls._load_config(
                survey, release, **kwargs)
            config.update(**kwargs)
            obj = Survey(name=survey, release=release)
            cls._load_survey_data(obj, config, **kwargs)
            cls._cached_surveys[key] = obj
            if verbose:
                print(f"✓ Survey '{key}' loaded and cached successfully")
        if uniform_survey:
            cls._save_uniform_survey(cls._cached_surveys[key], **kwargs)
        return cls._cached_surveys[key]

    @classmethod
    def clear_cache(cls, survey=None, release=None, **kwargs):
        verbose = kwargs.get("verbose", True)
        if survey is None:
            n = len(cls._cached_surveys)
            cls._cached_surveys.clear()
            if verbose:
                print(f"✓ Cleared {n} cached survey(s)")
            return
        if release is None:
            keys = [k for k in cls._cached_surveys if k.startswith(f"{survey}_")]
            keys += [survey] if survey in cls._cached_surveys else []
        else:
            keys = [f"{survey}_{release}"]
        for k in keys:
            if k in cls._cached_surveys:
                del cls._cached_surveys[k]
                if verbose:
                    print(f"✓ Cleared cached survey '{k}'")
            elif verbose:
                print(f"⚠ Survey '{k}' not found in cache")

    @classmethod
    def list_cached_surveys(cls):
        return list(cls._cached_surveys.keys())

    # ------------------------------------------------------------------ uniform variant
    @classmethod
    def _save_uniform_survey(cls, survey, **kwargs):
        """Cache a copy whose depth maps are a single median value (NaNs kept)."""
        verbose = kwargs.get("verbose", True)
        key = f"{survey.name}_{survey.release}_uniform" if survey.release else f"{survey.name}_uniform"
        if key in cls._cached_surveys:
            if verbose:
                print(f"✓ Uniform survey '{key}' already cached")
            return
        flat = copy.deepcopy(survey)
        for band, m in flat.maglim_maps.items():
            m = np.asarray(m)
            level = cls._get_median_value(m, survey=flat, **kwargs)
            flat.maglim_maps[band] = np.where(np.isnan(m), np.nan, level)
            if verbose:
                print(f"  {band}-band: uniform maglim = {level:.2f} mag")
        cls._cached_surveys[key] = flat
        if verbose:
            print(f"✓ Uniform survey cached as '{key}'")

    @classmethod
    def _get_median_value(cls, maglim_map, **kwargs):
        mask_type = kwargs.get("uniform_survey_mask_type", ["nan", "dust"])
        mask_type = [mask_type] if isinstance(mask_type, str) else mask_type
        valid = np.ones_like(maglim_map, dtype=bool)
        for kind in mask_type:
            if kind == "nan":
                valid &= ~np.isnan(maglim_map)
            elif kind == "dust":
                survey = kwargs.get("survey")
                if survey is None:
                    raise ValueError("Survey object required for 'dust' mask type.")
                if getattr(survey, "ebv_map", None) is None:
                    raise ValueError("E(B-V) map not available for 'dust' mask type.")
                ebv = np.asarray(survey.ebv_map)
                if hp.get_nside(maglim_map) != hp.get_nside(ebv):
                    ebv = hp.ud_grade(ebv, hp.get_nside(maglim_map))
                valid &= ebv < 0.2
        if kwargs.get("verbose", True):
            print(f"  Computing median from {np.sum(valid)}/{len(maglim_map)} pixels "
                  f"(mask: {mask_type})")
        return np.nanmedian(maglim_map[valid])

    # ------------------------------------------------------------------ config + files
    @classmethod
    def _config_dir(cls):
        here = os.path.dirname(os.path.abspath(__file__))
        return os.path.join(here, "..", "config", "surveys")

    @classmethod
    def _load_config(cls, survey, release=None, **kwargs):
        name = f"{survey}_{release}" if release else survey
        cdir = cls._config_dir()
        path = None
        for ext in (".yaml", ".json"):
            cand = os.path.join(cdir, name + ext)
            if os.path.exists(cand):
                path = cand
                break
        if path is None:
            raise FileNotFoundError(f"Configuration file not found for '{name}'\n"
                                    f"Searched in: {cdir}\nExpected: {name}.yaml or {name}.json")
        if kwargs.get("verbose", True):
            print(f"  Loading config from: {os.path.basename(path)}")
        with open(path, "r") as handle:
            if path.endswith(".yaml"):
                import yaml
                data = yaml.safe_load(handle)
            else:
                import json
                data = json.load(handle)
        if survey != data.get("name"):
            raise ValueError(f"Survey name mismatch:\n  Requested: '{survey}'\n"
                             f"  In config: '{data.get('name')}'")
        if release != data.get("release", None):
            raise ValueError(f"Release mismatch for survey '{survey}':\n  Requested: '{release}'\n"
                             f"  In config: '{data.get('release', None)}'")
        return data

    @classmethod
    def _load_survey_data(cls, survey, config, **kwargs):
        """Populate ``survey`` from its config (reference :891-1148)."""
        verbose = kwargs.get("verbose", True)
        files = config.get("survey_files", {})
        props = config.get("survey_properties", {})
        here = os.path.dirname(os.path.abspath(__file__))
        tag = f"{survey.name}_{survey.release}" if survey.release else survey.name
        shared = os.path.join(here, "..", "data", "others")
        root = files.get("file_path", "") or os.path.join(here, "..", "data", "surveys", tag)
        if verbose:
            print(f"Survey data directory: {root}")
            print(f"Fallback directory for shared data files: {shared}")

        bands = [k[len("maglim_map_"):] for k in files if k.startswith("maglim_map_")]
        bands = [b for i, b in enumerate(bands) if b and b not in bands[:i]]
        bands += [b for b in props.get("bands", []) if b not in bands]
        survey.bands = sorted(bands or ["g", "r"])

        defaults = {"g": 3.303, "r": 2.285, "i": 1.682, "z": 1.322, "y": 1.087}
        for b in survey.bands:
            survey.coeff_extinc[b] = props.get(f"coeff_extinc_{b}", defaults.get(b))
            survey.saturation[b] = props.get(f"saturation_{b}", props.get("saturation", 16.0))
            survey.sys_error[b] = props.get(f"sys_error_{b}", props.get("sys_error", 0.005))

        for b in survey.bands:
            fname = files.get(f"maglim_map_{b}")
            if fname is None:
                if verbose:
                    print(f"  ⚠ Warning: 'maglim_map_{b}' not specified in config (skipping {b}-band)")
                continue
            path = cls._find_file(fname, root, shared)
            if path is None:
                if verbose:
                    print(f"  ✗ {b}-band magnitude limit: File not found - {fname}")
                continue
            try:
                survey.maglim_maps[b] = cls.read_map(path, map_min=survey.saturation[b])
                if verbose:
                    print(f"  ✓ Success for {b}-band magnitude limit")
            except Exception as e:                      # as upstream: swallow, record None
                if verbose:
                    print(f"    ✗ Failed to load {b}-band maglim map: {e}")
                survey.maglim_maps[b] = None

        survey.completeness_band = files.get("completeness_band", "r")
        survey.delta_saturation = props.get("delta_saturation", -10.4)
        dsat = survey.delta_saturation
        if "completeness" in files:
            for attr, sel, label in (("completeness", "both", "Completeness/efficiency function"),
                                     ("efficiency_detection", "detected", "Detection efficiency function"),
                                     ("efficiency_classification", "classified",
                                      "Classification efficiency function")):
                try:
                    cls._load_file(survey, files, attr, label,
                                   lambda f, s=sel: cls.set_completeness(f, delta_saturation=dsat,
                                                                         selection=s),
                                   root, shared, filename=files.get("completeness"), **kwargs)
                except Exception:
                    if verbose:
                        print(f"No {label.lower()} found, skipping.")
        if "log_photo_error" in files:
            cls._load_file(survey, files, "log_photo_error", "Photometric error model",
                           lambda f: cls.set_photo_error(f, delta_saturation=dsat), root, shared,
                           **kwargs)
        for attr, label in (("ebv_map", "E(B-V) extinction map"), ("coverage", "Survey coverage map")):
            cls._load_file(survey, files, attr, label, cls.read_map, root, shared, **kwargs)
        if survey.coverage is None:
            cls._build_coverage_map(survey, **kwargs)
        if verbose:
            for b in survey.bands:
                print(f"  {b}-band: extinction {survey.coeff_extinc[b]}, saturation "
                      f"{survey.saturation[b]}, systematic error {survey.sys_error[b]}")

    @classmethod
    def _find_file(cls, filename, data_path_survey, data_path_others):
        for root in (data_path_survey, data_path_others):
            path = os.path.join(root, filename)
            if os.path.exists(path):
                return path
        return None

    @classmethod
    def _load_file(cls, survey, config, attr_name, description, loader_func, data_path_survey,
                   data_path_others=None, filename=None, **kwargs):
        verbose = kwargs.get("verbose", True)
        filename = config.get(attr_name) if filename is None else filename
        if filename is None:
            if verbose:
                print(f"  ⚠ Warning: '{attr_name}' not specified in config (skipping)")
            return
        path = cls._find_file(filename, data_path_survey, data_path_others or data_path_survey)
        if path is None:
            if verbose:
                print(f"  ✗ {description}: File not found - {filename}")
            return
        try:
            setattr(survey, attr_name, loader_func(path))
            if verbose:
                print(f"  ✓ {description}: {filename}")
        except Exception as e:
            if verbose:
                print(f"    ✗ Failed to load {attr_name}: {e}")
            setattr(survey, attr_name, None)

    @staticmethod
    def read_map(path, map_min=None, map_max=None):
        """Dense RING map from ``.npy`` or HEALPix FITS (``.fits``, ``.fits.gz``;
        native reader, ``fitsio.py``) or HealSparse ``.hsp`` (native reader too).  ``.npy`` and
        ``.hsp`` maps get UNSEEN -> NaN; FITS maps are returned as ``hp.read_map``
        returns them (reference surveys.py:1019,1117-1118)."""
        lower = path.lower()
        if lower.endswith(".npy"):
            m = np.load(path).astype(np.float64)
            return np.where(m == hp.UNSEEN, np.nan, m)
        if lower.endswith(".hsp"):
            return SurveyFactory.open_map_sparse(path, map_min=map_min, map_max=map_max)
        from .fitsio import read_healpix_map
        return read_healpix_map(path)

    @staticmethod
    def open_map_sparse(file_path, **kwargs):
        """HealSparse file -> dense RING map with NaN for unseen pixels (reference :1255-1284)."""
        from .fitsio import read_healsparse_map       # native reader: healsparse not needed
        m = read_healsparse_map(file_path, nest=False)
        m = np.where(m == hp.UNSEEN, np.nan, m)
        lo, hi = kwargs.get("map_min"), kwargs.get("map_max")
        if lo is not None:
            m = np.where(m < lo, np.nan, m)
        if hi is not None:
            m = np.where(m > hi, np.nan, m)
        return m

    # ------------------------------------------------------------------ curves
    @staticmethod
    def completeness_from_arrays(delta_mags, efficiencies, delta_saturation=-10.4):
        """Zero-pad an efficiency curve at the bright (saturation) and faint (+1 mag)
        ends and tabulate it with fill 0 (reference :1344-1366)."""
        dm = np.array(delta_mags, dtype=float)
        eff = np.array(efficiencies, dtype=float)
        if dm.min() > delta_saturation:
            dm, eff = np.insert(dm, 0, delta_saturation), np.insert(eff, 0, 0.0)
        elif dm.min() == delta_saturation:
            dm, eff = np.insert(dm, 0, delta_saturation - 1e-5), np.insert(eff, 0, 0.0)
        else:
            eff[dm <= delta_saturation] = 0.0
        dm, eff = np.append(dm, dm[-1] + 1), np.append(eff, 0.0)
        return TabulatedFunction(dm, eff, 0.0)

    @staticmethod
    def photo_error_from_arrays(delta_mags, log_errors, delta_saturation=-10.4):
        """Extend log10(error) flat to the saturation offset, fill 1.0 (10 mag)
        outside on both sides (reference :1406-1417)."""
        dm = np.array(delta_mags, dtype=float)
        le = np.array(log_errors, dtype=float)
        if dm.min() > delta_saturation:
            dm, le = np.insert(dm, 0, delta_saturation), np.insert(le, 0, le[0])
        return TabulatedFunction(dm, le, 1.0)

    @staticmethod
    def set_completeness(filename, delta_saturation=-10.4, selection="both"):
        """Efficiency curve from a CSV with columns delta_mag, detection_eff,
        classifiction_eff [sic], classification_detection_eff (reference :1286-1368)."""
        data = np.genfromtxt(filename, delimiter=",", names=True)
        column = {"detected": "detection_eff", "classified": "classifiction_eff",
                  "both": "classification_detection_eff"}.get(selection)
        if column is None:
            raise ValueError(f"Invalid selection '{selection}'. Must be 'detected', 'classified', "
                             "or 'both'")
        return SurveyFactory.completeness_from_arrays(data["delta_mag"], data[column],
                                                      delta_saturation)

    @staticmethod
    def set_photo_error(filename, delta_saturation=-10.4):
        """log10(mag error) curve from a CSV with columns delta_mag, log_mag_err
        (reference :1370-1419)."""
        data = np.genfromtxt(filename, delimiter=",", names=True)
        return SurveyFactory.photo_error_from_arrays(data["delta_mag"], data["log_mag_err"],
                                                     delta_saturation)

    @classmethod
    def _build_coverage_map(cls, survey, **kwargs):
        """Coverage = AND over bands of (maglim finite and above saturation), at the
        coarsest depth-map resolution (reference :1421-1490)."""
        verbose = kwargs.get("verbose", True)
        maps = {b: np.asarray(m) for b, m in survey.maglim_maps.items() if m is not None}
        if not maps:
            if verbose:
                print("  ⚠ Warning: No magnitude limit maps found, cannot build coverage")
            return None
        nside = min(hp.get_nside(m) for m in maps.values())
        cover = np.ones(hp.nside2npix(nside), dtype=bool)
        for b, m in maps.items():
            if hp.get_nside(m) != nside:
                m = hp.ud_grade(m, nside)
            cover &= np.isfinite(m) & (m > survey.saturation[b])
        survey.coverage = cover.astype(float)
        if verbose:
            print(f"  ✓ Built coverage map (nside={nside}, {np.sum(cover)} pixels covered)")
        return survey.coverage
