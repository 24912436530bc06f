"""Synthetic, fully deterministic survey inputs for tests, smoke and benchmarks.

The reference's survey data (LSST/DES depth maps, E(B-V) map, efficiency and
photo-error CSVs, PARSEC isochrones) are downloaded from Zenodo
(reference bin/download_data.py:24-33) and are not available offline, so every
workload here is generated from closed-form integer hashes: the same pixel
index always yields the same value, on the host (NumPy) and on the device
(torch), with no libm call involved -- the values are exact rationals in fp64.

Schema follows what ``SurveyFactory`` hands to ``Survey`` in the reference:
dense RING-ordered float64 HEALPix arrays with NaN for unseen pixels
(reference surveys.py:1272-1284) and the CSV column names of
surveys.py:1329-1338,1402-1404.
"""
from __future__ import annotations

import numpy as np

_M32 = 0xFFFFFFFF


def _mix_np(pix, salt):
    x = (np.asarray(pix, dtype=np.uint64) + np.uint64(salt)) & np.uint64(_M32)
    x = (x * np.uint64(0x9E3779B1)) & np.uint64(_M32)
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x85EBCA77)) & np.uint64(_M32)
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE3D)) & np.uint64(_M32)
    x ^= x >> np.uint64(16)
    return x


def _mix_torch(pix, salt):
    x = (pix + salt) & _M32
    x = (x * 0x9E3779B1) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x85EBCA77) & _M32
    x = x ^ (x >> 13)
    x = (x * 0xC2B2AE3D) & _M32
    x = x ^ (x >> 16)
    return x


def _map_from_hash(h_val, h_hole, kind, xp):
    """kind: 'maglim_g' | 'maglim_r' | 'ebv'.  xp = numpy or torch namespace."""
    u = h_val.to(xp.float64) if xp.__name__ == "torch" else h_val.astype(np.float64)
    u = u * (1.0 / 4294967296.0)
    if kind == "ebv":
        return 0.02 + 0.1 * u
    base = 25.4 if kind == "maglim_g" else 25.1
    m = base + 1.2 * u
    r = h_hole % 200
    nan = float("nan")
    if xp.__name__ == "torch":
        m = xp.where(r < 4, xp.full_like(m, nan), m)        # 2 % unseen
        m = xp.where(r == 4, xp.full_like(m, 15.0), m)      # 0.5 % shallower than saturation
    else:
        m = np.where(r < 4, nan, m)
        m = np.where(r == 4, 15.0, m)
    return m


_SALT = dict(maglim_g=(11, 12), maglim_r=(21, 22), ebv=(31, 32))


def healpix_map(kind, nside, device=None, chunk=1 << 26):
    """Dense RING map of ``kind`` at ``nside``; NumPy array, or a torch tensor on
    ``device`` (built in chunks so nside=4096 never holds more than a few
    temporaries of ``chunk`` pixels)."""
    npix = 12 * nside * nside
    s_val, s_hole = _SALT[kind]
    if device is None:
        pix = np.arange(npix, dtype=np.uint64)
        return _map_from_hash(_mix_np(pix, s_val), _mix_np(pix, s_hole), kind, np)
    import torch
    out = torch.empty(npix, dtype=torch.float64, device=device)
    for lo in range(0, npix, chunk):
        hi = min(npix, lo + chunk)
        pix = torch.arange(lo, hi, dtype=torch.int64, device=device)
        out[lo:hi] = _map_from_hash(_mix_torch(pix, s_val), _mix_torch(pix, s_hole), kind, torch)
    return out


def efficiency_csv_columns(first=-10.4, last=1.5, n=120):
    """Columns of a stellar-efficiency CSV (logistic roll-off around the depth)."""
    d = first + (last - first) * np.arange(n) / (n - 1)
    det = 0.98 / (1.0 + np.exp((d + 0.3) / 0.25))
    cls = 0.97 / (1.0 + np.exp((d + 0.9) / 0.35))
    return dict(delta_mag=d, detection_eff=det, classifiction_eff=cls,
                classification_detection_eff=det * cls)


def photo_error_csv_columns(first=-10.4, last=1.5, n=100):
    """Columns of a photo-error CSV: log10 of a sky-limited error curve."""
    d = first + (last - first) * np.arange(n) / (n - 1)
    return dict(delta_mag=d, log_mag_err=np.log10(0.2171 * 10 ** (0.4 * d) + 0.003))


def isochrone_table(n=256):
    """A smooth old metal-poor-like isochrone: main sequence, turn-off, giant
    branch.  Returns dict(mass_init, mag_1 (g), mag_2 (r)) of absolute mags."""
    t = np.arange(n) / (n - 1.0)
    mass = 0.1 + 0.7 * t ** 0.85
    ms_r = 4.6 - 9.5 * np.log10(mass / 0.8)                     # main sequence
    up = np.clip((mass - 0.74) / 0.06, 0.0, 1.0)                # turn-off -> RGB tip
    mag_r = ms_r - 7.5 * up ** 2.2
    colour = 1.45 - 1.15 * (mass - 0.1) / 0.7 + 0.55 * up ** 1.5
    return dict(mass_init=mass, mag_1=mag_r + colour, mag_2=mag_r)


# LSST-like constants, reference config/surveys/lsst_dc2.yaml:27-41
LSST_PROPS = dict(coeff_extinc=dict(g=3.6605664439892625, r=2.70136780871597),
                  saturation=dict(g=16.0, r=16.0), sys_error=dict(g=0.005, r=0.005),
                  delta_saturation=-10.4, completeness_band="r")

# Great-circle frame accepted in reference notebooks/tutorial_inject_stream.ipynb
# (cell 14, seed 42, trial 2): the two endpoints drawn by observed.py:695-698.
NOTEBOOK_ENDPOINTS = ((351.22404658923216, -54.25699497364132),
                      (282.9831498997034, 31.485273979529264))
