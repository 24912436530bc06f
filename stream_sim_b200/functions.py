"""phi1 -> y evaluators (track centre / width, distance modulus, density pdf).

Same public surface as the reference's ``stream_sim/functions.py`` (factory at
:11-44, classes at :47-400) so configurations and user code carry over.  Each
function has two faces:

* a *host* face, ``f(x)``: NumPy/SciPy evaluation with exactly the calls the
  reference makes (``interp1d``, ``CubicSpline(bc_type='natural')``).  It is used
  to build lookup tables (e.g. the 100000-point pdf grid behind the inverse-CDF
  sampler) and for interactive use, never per star.
* a *device* face, ``f.describe(packer)``: an ``SSFunc`` descriptor plus knot /
  coefficient tables appended to the table blob that the CUDA kernel evaluates
  per star (csrc/device_math.cuh ``fn_eval``).
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import scipy.interpolate

from . import _lib

_REGISTRY = {}


def _register(*names):
    def deco(cls):
        for n in names:
            _REGISTRY[n] = cls
        return cls
    return deco


def function_factory(type_, **kwargs):
    """Create the function registered under ``type_`` (lower-case name).

    Mirrors reference functions.py:11-44, including the bare ``Exception`` for an
    unknown name.
    """
    try:
        cls = _REGISTRY[type_]
    except KeyError:
        raise Exception(f"Unrecognized function: {type_}") from None
    return cls(**kwargs)


class BoundedFunction(object):
    """Function that evaluates to NaN outside ``[xmin, xmax]`` (reference :47-79)."""

    def __init__(self, xmin=-np.inf, xmax=np.inf):
        self.xmin = xmin
        self.xmax = xmax

    def __call__(self, x, **kwargs):
        self.__dict__.update(kwargs)     # reference quirk (:65): kwargs overwrite attributes
        grid = np.atleast_1d(x)
        y = self._evaluate(grid)
        y = np.where((grid >= self.xmin) & (grid <= self.xmax), y, np.nan)
        return y.item() if np.isscalar(x) else y

    def _evaluate(self, xvals):
        pass

    # ---- device face
    def describe(self, packer):
        """Return the SSFunc descriptor, appending tables to ``packer`` if needed."""
        raise NotImplementedError(f"{type(self).__name__} has no device evaluator")

    def _descriptor(self, kind, **fields):
        f = _lib.SSFunc()
        f.kind = kind
        f.xmin, f.xmax = float(self.xmin), float(self.xmax)
        for k, v in fields.items():
            setattr(f, k, v)
        return f


@_register("constant")
class Constant(BoundedFunction):
    """y = value."""

    def __init__(self, value=1.0, **kwargs):
        super().__init__(**kwargs)
        self.value = value

    def _evaluate(self, xvals):
        return self.value * np.ones(len(xvals))

    def describe(self, packer):
        f = self._descriptor(_lib.FN_CONSTANT)
        f.p[0] = float(self.value)
        return f


@_register("line")
class Line(BoundedFunction):
    """y = slope * x + intercept."""

    def __init__(self, slope=0.0, intercept=0.0, **kwargs):
        super().__init__(**kwargs)
        self.slope = slope
        self.intercept = intercept

    def _evaluate(self, xvals):
        return self.slope * xvals + self.intercept

    def describe(self, packer):
        f = self._descriptor(_lib.FN_LINE)
        f.p[0], f.p[1] = float(self.slope), float(self.intercept)
        return f


@_register("sinusoid")
class Sinusoid(BoundedFunction):
    """y = amplitude/2 * cos(2 pi (x - phase) / period) + amplitude/2 + offset."""

    def __init__(self, amplitude=1.0, period=1.0, phase=0.0, offset=0.0, **kwargs):
        super().__init__(**kwargs)
        self.amplitude = amplitude
        self.period = period
        self.phase = phase
        self.offset = offset

    def _evaluate(self, xvals):
        half = self.amplitude / 2.0
        return half * np.cos(2 * np.pi * (xvals - self.phase) / self.period) + half + self.offset

    def describe(self, packer):
        f = self._descriptor(_lib.FN_SINUSOID)
        f.p[0], f.p[1] = float(self.amplitude), float(self.period)
        f.p[2], f.p[3] = float(self.phase), float(self.offset)
        return f


@_register("interpolation")
class Interpolation(BoundedFunction):
    """Piecewise-linear table, NaN outside the tabulated range."""

    def __init__(self, xvals, yvals, **kwargs):
        self.xvals = np.atleast_1d(xvals)
        self.yvals = np.atleast_1d(yvals)
        kwargs.setdefault("xmin", np.min(xvals))
        kwargs.setdefault("xmax", np.max(xvals))
        super().__init__(**kwargs)

    def _evaluate(self, xvals):
        return scipy.interpolate.interp1d(self.xvals, self.yvals, bounds_error=False,
                                          fill_value=np.nan)(xvals)

    def describe(self, packer):
        return self._descriptor(_lib.FN_LINEAR, **packer.add_linear(self.xvals, self.yvals))


@_register("file", "fileinterpolation")
class FileInterpolation(Interpolation):
    """Piecewise-linear table read from two columns of a CSV file."""

    def __init__(self, filename, columns=None, **kwargs):
        self._filename = filename
        self._data = pd.read_csv(self._filename)
        xname, yname = columns if columns else self._data.columns[[0, 1]]
        super().__init__(self._data[xname].values, self._data[yname].values, **kwargs)


@_register("cubicsplineinterpolation")
class CubicSplineInterpolation(BoundedFunction):
    """Natural cubic spline through (nodes, node_values); NaN outside the nodes."""

    def __init__(self, nodes, node_values, **kwargs):
        self.nodes = np.atleast_1d(nodes)
        self.node_values = np.atleast_1d(node_values)
        kwargs.setdefault("xmin", np.min(nodes))
        kwargs.setdefault("xmax", np.max(nodes))
        super().__init__(**kwargs)

    def _spline(self):
        return scipy.interpolate.CubicSpline(self.nodes, self.node_values, bc_type="natural",
                                             extrapolate=False)

    def _evaluate(self, xvals):
        return self._spline()(xvals)

    def describe(self, packer):
        x_off, c_off, n = packer.add_spline(self._spline())
        return self._descriptor(_lib.FN_SPLINE, n=n, x_off=x_off, c_off=c_off)


@_register("filecubicsplineinterpolation")
class FileCubicSplineInterpolation(CubicSplineInterpolation):
    """Natural cubic spline whose nodes come from a (stream, type)-filtered CSV."""

    def __init__(self, filename, stream_name=None, spline_type=None, nodes_name="phi1",
                 node_vals_name="mean", **kwargs):
        self._filename = filename
        table = pd.read_csv(self._filename)
        if spline_type:
            table = table.loc[table["type"] == spline_type, :]
        if stream_name:
            table = table.loc[table["stream"] == stream_name, :]
        self._data = table
        super().__init__(table[nodes_name].values, table[node_vals_name].values, **kwargs)


@_register("lineardensitycubicsplineinterpolation")
class LinearDensityCubicSplineInterpolation():
    """Linear density sqrt(2 pi) * I(x) * sigma(x) from two natural splines
    (peak intensity and spread).  No bounds mask of its own (reference :349-359)."""

    def __init__(self, intensity_nodes, intensity_node_values, spread_nodes, spread_node_values,
                 **kwargs):
        self._peak_intensity = CubicSplineInterpolation(intensity_nodes, intensity_node_values)
        self._spread = CubicSplineInterpolation(spread_nodes, spread_node_values)
        self.xmin = np.max([np.min(intensity_nodes), np.min(spread_nodes)])
        self.xmax = np.min([np.max(intensity_nodes), np.max(spread_nodes)])

    def __call__(self, x, **kwargs):
        self.__dict__.update(kwargs)
        return self._evaluate(np.atleast_1d(x))

    def _evaluate(self, xvals):
        return np.sqrt(2 * np.pi) * self._peak_intensity(xvals) * self._spread(xvals)

    def describe(self, packer):
        x1, c1, n1 = packer.add_spline(self._peak_intensity._spline())
        x2, c2, n2 = packer.add_spline(self._spread._spline())
        f = _lib.SSFunc()
        f.kind = _lib.FN_SPLINE_PRODUCT
        f.xmin, f.xmax = float(self.xmin), float(self.xmax)
        f.n, f.x_off, f.c_off = n1, x1, c1
        f.n2, f.x2_off, f.c2_off = n2, x2, c2
        return f


@_register("filelineardensitycubicsplineinterpolation")
class FileLinearDensityCubicSplineInterpolation(LinearDensityCubicSplineInterpolation):
    """Linear density whose two splines are read from a per-stream CSV."""

    def __init__(self, filename, stream_name=None, **kwargs):
        self._filename = filename
        table = pd.read_csv(self._filename)
        if stream_name:
            table = table.loc[table["stream"] == stream_name]
        self._data = table
        peak = table["type"] == "peak_intensity"
        width = table["type"] == "spread"
        super().__init__(intensity_nodes=table.loc[peak, "phi1"].values,
                         intensity_node_values=table.loc[peak, "mean"].values,
                         spread_nodes=table.loc[width, "phi1"].values,
                         spread_node_values=table.loc[width, "mean"].values, **kwargs)
