"""HEALPix helpers standing in for the handful of ``healpy`` calls the reference
makes (``hp.get_nside``, ``hp.nside2npix``, ``hp.UNSEEN``, ``hp.ud_grade``,
``hp.ang2pix`` -- reference observed.py:147-176,716-823, surveys.py:788-791,
1452-1475).  healpy itself is not a dependency.

``ang2pix`` runs on the GPU (csrc/device_math.cuh ``hp_ring`` / ``hp_nest``).
``ud_grade`` and the RING<->NEST index conversions are one-off, per-map host
operations used when a sky mask is assembled; they are not on the per-star path.
"""
from __future__ import annotations

import numpy as np

UNSEEN = -1.6375e30


def nside2npix(nside):
    return 12 * int(nside) * int(nside)


def npix2nside(npix):
    nside = int(round(np.sqrt(npix / 12.0)))
    if 12 * nside * nside != npix:
        raise ValueError("Wrong pixel number (it is not 12*nside**2)")
    return nside


def get_nside(m):
    return npix2nside(len(m))


def ang2pix(nside, lon, lat, nest=False, lonlat=True, device=None):
    """Pixel index of sky positions; GPU-computed, returned as a NumPy int64 array
    (scalar for scalar input).  ``lonlat=False`` takes (theta, phi) in radians."""
    from .engine import ang2pix as _dev
    scalar = np.isscalar(lon) and np.isscalar(lat)
    lon, lat = np.broadcast_arrays(np.atleast_1d(np.asarray(lon, dtype=float)),
                                   np.atleast_1d(np.asarray(lat, dtype=float)))
    if not lonlat:
        theta, phi = lon, lat
        lon, lat = np.degrees(phi), 90.0 - np.degrees(theta)
    pix = _dev(nside, lon, lat, nest=nest, device=device).cpu().numpy()
    if (pix < 0).any():
        raise ValueError("THETA is out of range [0,pi]")
    return int(pix[0]) if scalar else pix


# ---------------------------------------------------------------- index conversions
_JRLL = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4])
_JPLL = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7])


def _isqrt(v):
    r = np.floor(np.sqrt(v.astype(np.float64))).astype(np.int64)
    r = np.where(r * r > v, r - 1, r)
    r = np.where((r + 1) * (r + 1) <= v, r + 1, r)
    return r


def _spread(v):
    v = v.astype(np.uint64)
    for s, m in ((16, 0x0000FFFF0000FFFF), (8, 0x00FF00FF00FF00FF), (4, 0x0F0F0F0F0F0F0F0F),
                 (2, 0x3333333333333333), (1, 0x5555555555555555)):
        v = (v | (v << np.uint64(s))) & np.uint64(m)
    return v


def _compress(v):
    v = v.astype(np.uint64) & np.uint64(0x5555555555555555)
    for s, m in ((1, 0x3333333333333333), (2, 0x0F0F0F0F0F0F0F0F), (4, 0x00FF00FF00FF00FF),
                 (8, 0x0000FFFF0000FFFF), (16, 0x00000000FFFFFFFF)):
        v = (v | (v >> np.uint64(s))) & np.uint64(m)
    return v


def ring2nest(nside, pix):
    """RING -> NESTED pixel index (healpix_cxx ring2xyf + xyf2nest)."""
    ns = int(nside)
    if ns & (ns - 1):
        raise ValueError("nside must be a power of two for NESTED ordering")
    pix = np.asarray(pix, dtype=np.int64)
    order = ns.bit_length() - 1
    ncap, npix, nl2, nl4 = 2 * ns * (ns - 1), 12 * ns * ns, 2 * ns, 4 * ns
    north, south = pix < ncap, pix >= npix - ncap
    eq = ~(north | south)
    iring = np.zeros_like(pix)
    iphi = np.zeros_like(pix)
    kshift = np.zeros_like(pix)
    nr = np.full_like(pix, ns)
    face = np.zeros_like(pix)
    # north cap
    p = pix[north]
    ir = (1 + _isqrt(1 + 2 * p)) >> 1
    ip = (p + 1) - 2 * ir * (ir - 1)
    iring[north], iphi[north], nr[north] = ir, ip, ir
    face[north] = (ip - 1) // np.maximum(ir, 1)
    # equatorial belt
    p = pix[eq] - ncap
    tmp = p // nl4
    ir = tmp + ns
    ip = p - nl4 * tmp + 1
    iring[eq], iphi[eq] = ir, ip
    kshift[eq] = (ir + ns) & 1
    ire = ir - ns + 1
    irm = nl2 + 2 - ire
    ifm = (ip - ire // 2 + ns - 1) >> order
    ifp = (ip - irm // 2 + ns - 1) >> order
    face[eq] = np.where(ifp == ifm, ifp | 4, np.where(ifp < ifm, ifp, ifm + 8))
    # south cap
    p = npix - pix[south]
    ir = (1 + _isqrt(2 * p - 1)) >> 1
    ip = 4 * ir + 1 - (p - 2 * ir * (ir - 1))
    iphi[south], nr[south] = ip, ir
    iring[south] = 2 * nl2 - ir
    face[south] = 8 + (ip - 1) // np.maximum(ir, 1)

    irt = iring - _JRLL[face] * ns + 1
    ipt = 2 * iphi - _JPLL[face] * nr - kshift - 1
    ipt = np.where(ipt >= nl2, ipt - 8 * ns, ipt)
    ix = (ipt - irt) >> 1
    iy = (-ipt - irt) >> 1
    inter = _spread(ix) | (_spread(iy) << np.uint64(1))
    return (face << (2 * order)) + inter.astype(np.int64)


def reorder_ring_to_nest(m):
    """Map values re-indexed from RING to NESTED ordering."""
    m = np.asarray(m)
    ns = get_nside(m)
    out = np.empty_like(m)
    out[ring2nest(ns, np.arange(len(m), dtype=np.int64))] = m
    return out


def reorder_nest_to_ring(m):
    m = np.asarray(m)
    ns = get_nside(m)
    return m[ring2nest(ns, np.arange(len(m), dtype=np.int64))]


def ud_grade(map_in, nside_out):
    """Up/down-grade a RING map like ``healpy.ud_grade(m, nside_out)`` with its
    defaults: children are averaged over valid (finite, not UNSEEN) pixels, a
    parent with no valid child becomes UNSEEN; upgrading replicates the parent."""
    m = np.asarray(map_in, dtype=np.float64)
    ns_in = get_nside(m)
    ns_out = int(nside_out)
    if ns_out == ns_in:
        return m.copy()
    nest = reorder_ring_to_nest(m)
    if ns_out > ns_in:
        rat2 = (ns_out // ns_in) ** 2
        out = np.repeat(nest, rat2)
    else:
        rat2 = (ns_in // ns_out) ** 2
        mr = nest.reshape(-1, rat2)
        good = np.isfinite(mr) & (np.abs(mr - UNSEEN) > 1e-5 * np.abs(UNSEEN))
        nhit = good.sum(axis=1)
        tot = np.where(good, mr, 0.0).sum(axis=1)
        out = np.where(nhit > 0, tot / np.maximum(nhit, 1), UNSEEN)
    return reorder_nest_to_ring(out)
