"""Great-circle stream frames.

Stands in for ``gala.coordinates.GreatCircleICRSFrame`` + the astropy transform
used by the reference (observed.py:573, 645-667): a frame is the rotation matrix
R whose rows are (origin, pole x origin, pole) in ICRS cartesian coordinates, so
that ICRS = R^T . stream (checked against the printed
rows of the reference tutorial notebook).  Only the matrix is built on the host;
the per-star rotation runs on the GPU.
"""
from __future__ import annotations

import numpy as np


def _unit(ra_deg, dec_deg):
    ra, dec = np.radians(float(ra_deg)), np.radians(float(dec_deg))
    return np.array([np.cos(dec) * np.cos(ra), np.cos(dec) * np.sin(ra), np.sin(dec)])


def _radec(point):
    """(ra, dec) in degrees from a tuple or any object with .ra/.dec (astropy-like)."""
    if hasattr(point, "ra") and hasattr(point, "dec"):
        ra, dec = point.ra, point.dec
        ra = getattr(ra, "deg", getattr(ra, "degree", ra))
        dec = getattr(dec, "deg", getattr(dec, "degree", dec))
        return float(ra), float(dec)
    ra, dec = point
    return float(ra), float(dec)


class SkyPoint:
    """Minimal ICRS position (degrees)."""

    def __init__(self, ra, dec):
        self.ra, self.dec = ra, dec

    def __repr__(self):
        return f"SkyPoint(ra={self.ra!r}, dec={self.dec!r})"


class GreatCircleFrame:
    """Stream-aligned frame: phi1 along a great circle, phi2 across it."""

    def __init__(self, pole, origin):
        pole = np.asarray(pole, dtype=float)
        origin = np.asarray(origin, dtype=float)
        pole = pole / np.linalg.norm(pole)
        origin = origin - pole * np.dot(pole, origin)
        origin = origin / np.linalg.norm(origin)
        self.pole_xyz, self.origin_xyz = pole, origin
        self.R = np.stack([origin, np.cross(pole, origin), pole])

    @classmethod
    def from_endpoints(cls, end1, end2):
        """Great circle through two sky points; origin at their midpoint
        (gala ``GreatCircleICRSFrame.from_endpoints`` default)."""
        c1, c2 = _unit(*_radec(end1)), _unit(*_radec(end2))
        return cls(np.cross(c1, c2), c1 + c2)

    @classmethod
    def from_matrix(cls, R):
        R = np.asarray(R, dtype=float)
        return cls(R[2], R[0])

    @classmethod
    def coerce(cls, obj):
        """Accept a GreatCircleFrame, a 3x3 matrix, or a gala-like frame exposing
        ``pole`` and ``origin`` sky positions."""
        if obj is None or isinstance(obj, cls):
            return obj
        if hasattr(obj, "pole") and hasattr(obj, "origin") and not isinstance(obj, np.ndarray):
            return cls(_unit(*_radec(obj.pole)), _unit(*_radec(obj.origin)))
        return cls.from_matrix(obj)

    def _point(self, v):
        return SkyPoint(float(np.degrees(np.arctan2(v[1], v[0])) % 360.0),
                        float(np.degrees(np.arcsin(np.clip(v[2], -1, 1)))))

    @property
    def pole(self):
        return self._point(self.pole_xyz)

    @property
    def origin(self):
        return self._point(self.origin_xyz)

    def __repr__(self):
        return f"<GreatCircleFrame pole={self.pole} origin={self.origin}>"


class _Angle:
    def __init__(self, deg):
        self.deg = deg
        self.degree = deg

    @property
    def rad(self):
        return np.radians(self.deg)


class SkyPositions:
    """Result of ``phi_to_radec``: quacks like the astropy SkyCoord the reference
    returns for the attributes its callers use (``.icrs.ra.deg`` / ``.icrs.dec.deg``)."""

    def __init__(self, ra_deg, dec_deg, frame=None, phi1=None, phi2=None):
        self.ra, self.dec = _Angle(ra_deg), _Angle(dec_deg)
        self.frame = frame
        self.phi1 = None if phi1 is None else _Angle(phi1)
        self.phi2 = None if phi2 is None else _Angle(phi2)

    @property
    def icrs(self):
        return self

    def transform_to(self, frame):
        if frame != "icrs":
            raise NotImplementedError("only the ICRS frame is available")
        return self

    def __len__(self):
        return len(self.ra.deg)
