"""Probabilistic samplers with the reference's names and arguments
(``stream_sim/samplers.py``: factory :19-52, inverse transform :55-76, classes
:83-267).

A sampler here is a *description* of a distribution plus the host-side table
builder; the draws themselves are made on the GPU by libstreamsim_b200
(Philox4x32-10, one counter per global star index).  ``sample()`` therefore needs
a B200 and raises if the CUDA library or device is missing -- there is no NumPy
fallback.
"""
from __future__ import annotations

import numpy as np
import scipy.stats

from .functions import (Interpolation, FileInterpolation, Sinusoid, CubicSplineInterpolation,
                        FileCubicSplineInterpolation, LinearDensityCubicSplineInterpolation,
                        FileLinearDensityCubicSplineInterpolation)

_REGISTRY = {}


def _register(*names):
    def deco(cls):
        for n in names:
            _REGISTRY[n] = cls
        return cls
    return deco


def sampler_factory(type_, **kwargs):
    """Create the sampler registered under ``type_`` (reference samplers.py:19-52)."""
    try:
        cls = _REGISTRY[type_]
    except KeyError:
        raise Exception(f"Unrecognized sampler: {type_}") from None
    return cls(**kwargs)


def inverse_cdf_tables(vals, pdf):
    """Host tables behind ``inverse_transform_sample`` (reference samplers.py:69-73):
    the normalised cumulative sum and ``interp1d``'s stable (mergesort) ordering
    of it.  Returns ``(cdf_sorted, order)`` with ``order[j]`` the grid index of the
    j-th smallest CDF value."""
    cdf = np.cumsum(pdf)
    cdf /= cdf[-1]
    order = np.argsort(cdf, kind="mergesort")
    return cdf[order], order


def inverse_transform_sample(vals, pdf, size, seed=None, draws=None):
    """Inverse transform sampling of ``vals`` weighted by ``pdf`` on the GPU.

    Same contract as reference samplers.py:55-76 (nearest-grid-node result:
    ``vals[rint(interp(u))]``).  ``seed``/``draws`` are additive keywords.
    """
    from .engine import sample_inverse_cdf
    return sample_inverse_cdf(np.asarray(vals, dtype=float), np.asarray(pdf, dtype=float),
                              int(np.rint(size)), seed=seed, u=draws)


class Sampler(object):
    def __init__(self, *args, **kwargs):
        pass

    def pdf(self, values):
        return self._pdf(values)

    def sample(self, size):
        pass


class ScipySampler(Sampler):
    """Location/scale family.  ``pdf`` is SciPy's; ``sample`` runs on the GPU as
    ``draw * scale + loc`` (what ``rv.rvs`` computes, reference :101-102)."""

    _normal = False

    def __init__(self, rv):
        self._rv = rv

    def pdf(self, x):
        return self._rv.pdf(x)

    def _loc_scale(self):
        kw = self._rv.kwds
        return kw.get("loc", 0.0), kw.get("scale", 1.0)

    def sample(self, size, random_state=None, seed=None, draws=None):
        from .engine import sample_loc_scale
        loc, scale = self._loc_scale()
        if random_state is not None and seed is None:
            seed = random_state if isinstance(random_state, (int, np.integer)) else None
        return sample_loc_scale(loc, scale, int(size), normal=self._normal, seed=seed, draws=draws)


@_register("uniform")
class UniformSampler(ScipySampler):
    """Uniform between xmin and xmax (scalars or per-star arrays)."""

    def __init__(self, xmin, xmax):
        self.xmin, self.xmax = xmin, xmax
        super().__init__(scipy.stats.uniform(loc=xmin, scale=xmax - xmin))


@_register("gaussian")
class GaussianSampler(ScipySampler):
    """Normal with mean mu and standard deviation sigma (scalars or arrays)."""

    _normal = True

    def __init__(self, mu, sigma):
        self.mu, self.sigma = mu, sigma
        super().__init__(scipy.stats.norm(loc=mu, scale=sigma))


@_register("interpolation")
class InterpolationSampler(Sampler):
    """Sample x with probability density given by an interpolated function."""

    def __init__(self, xvals, yvals, **kwargs):
        self.interp = Interpolation(xvals, yvals, **kwargs)

    def grid(self, nsteps=1e5):
        """(xvals, pdf) on the regular grid the reference builds on every call (:159-160)."""
        xvals = np.linspace(self.interp.xmin, self.interp.xmax, int(nsteps))
        return xvals, self.interp(xvals)

    def _pdf(self, values):
        return self.interp(values)

    def sample(self, size, nsteps=1e5, seed=None, draws=None):
        xvals, pdf = self.grid(nsteps)
        return inverse_transform_sample(xvals, pdf, size=size, seed=seed, draws=draws)


@_register("file", "fileinterpolation")
class FileInterpolationSampler(InterpolationSampler):
    def __init__(self, filename, columns=None):
        self.interp = FileInterpolation(filename, columns)


@_register("sinusoid")
class SinusoidSampler(InterpolationSampler):
    def __init__(self, **kwargs):
        self.interp = Sinusoid(**kwargs)


@_register("cubicsplineinterpolation")
class CubicSplineInterpolationSampler(InterpolationSampler):
    def __init__(self, nodes, node_values):
        self.interp = CubicSplineInterpolation(nodes, node_values)


@_register("filecubicsplineinterpolation")
class FileCubicSplineInterpolationSampler(InterpolationSampler):
    def __init__(self, filename, stream_name=None, spline_type=None):
        self.interp = FileCubicSplineInterpolation(filename, stream_name, spline_type)


@_register("lineardensitycubicsplineinterpolation")
class LinearDensityCubicSplineInterpolationSampler(InterpolationSampler):
    def __init__(self, intensity_nodes, intensity_node_values, spread_nodes, spread_node_values):
        self.interp = LinearDensityCubicSplineInterpolation(
            intensity_nodes, intensity_node_values, spread_nodes, spread_node_values)


@_register("filelineardensitycubicsplineinterpolation")
class FileLinearDensityCubicSplineInterpolationSampler(InterpolationSampler):
    def __init__(self, filename, stream_name):
        self.interp = FileLinearDensityCubicSplineInterpolation(filename, stream_name)
