"""Host engine: packs lookup tables, owns device memory (PyTorch tensors) and
launches the sm_100a kernels of libstreamsim_b200 through its C ABI.

PyTorch is plumbing only (allocation, streams, ``torch.distributed``); every
per-star number is produced by the hand-written CUDA library.  Nothing here
falls back to NumPy: without the library or a B200 the calls raise.

Layout in HBM (all fp64 unless noted):
  * ``blob``      -- one packed buffer of small tables (knots, spline coefficients,
                     isochrone rows, IMF CDF, efficiency / photo-error curves,
                     int32 search guides); each CTA copies it to shared memory.
  * ``cdf_pairs`` -- [100000][2] (sorted CDF value, slope) of the phi1 sampler.
  * maps          -- dense RING HEALPix arrays, NaN = unseen (reference
                     surveys.py:1272-1284), replicated on every GPU.
  * outputs       -- structure of arrays, one contiguous column per quantity.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib

COLUMNS_FLOAT = ("phi1", "phi2", "dist", "mag_g", "mag_r", "ra", "dec",
                 "mag_g_obs", "mag_r_obs", "magerr_g", "magerr_r")
GENERATE_COLUMNS = ("phi1", "phi2", "dist", "mag_g", "mag_r")
ROTATE_COLUMNS = ("ra", "dec")
SNR_MIN = 5.0            # reference observed.py:267
GUIDE_ONE, GUIDE_MANY = 1 << 31, 1 << 30      # flag bits of a bucket-guide entry
CDF_GUIDE = 8192
IMF_GUIDE = 4096
ISO_LUT = 1 << 15
GRID_LUT = 1 << 17     # measured: 2^17 (2 MB, partly L1-resident) beats 2^15..2^20
# buckets of the 32-bit word tables (SSTables.grid_wlut / iso_wlut)
GRID_WLUT = 1 << int(os.environ.get("SS_GRID_WLUT_LOG2", "18"))
ISO_WLUT = 1 << int(os.environ.get("SS_ISO_WLUT_LOG2", "15"))


def _torch():
    import torch
    return torch


def require_device(device=None):
    """Return a torch CUDA device or raise: the product path has no CPU fallback."""
    torch = _torch()
    _lib.load()
    if not torch.cuda.is_available():
        raise _lib.StreamSimLibraryError(
            "stream_sim_b200 needs an NVIDIA B200 (sm_100a) GPU; no CUDA device is visible "
            "and there is no CPU fallback.")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise ValueError("device must be a CUDA device")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


# ------------------------------------------------------------------ packing
class BlobPacker:
    """Accumulates the small tables of one launch configuration into one buffer of
    doubles.  int32 search guides are embedded in the same buffer (two per
    double), so every offset handed out is in units of doubles from the start."""

    def __init__(self):
        self._d = []
        self._nd = 0
        self._seen = {}

    def add_doubles(self, arr, dedupe=False, align=1):
        """``align``: offset a multiple of this many doubles (2 for records read as double2)."""
        arr = np.ascontiguousarray(arr, dtype=np.float64).ravel()
        key = (arr.tobytes(), align) if dedupe else None
        if key is not None and key in self._seen:
            return self._seen[key]
        if self._nd % align:
            pad = align - self._nd % align
            self._d.append(np.zeros(pad))
            self._nd += pad
        off = self._nd
        self._d.append(arr)
        self._nd += arr.size
        if key is not None:
            self._seen[key] = off
        return off

    def add_ints(self, arr, dedupe=False):
        """int32 data stored inside the double buffer; returns the offset in doubles."""
        arr = np.ascontiguousarray(arr, dtype=np.int32).ravel()
        if arr.size % 2:
            arr = np.concatenate([arr, np.zeros(1, dtype=np.int32)])
        return self.add_doubles(arr.view(np.float64), dedupe=dedupe)

    def add_bucket_guide(self, xs):
        """Uniform-bucket guide over sorted knots ``xs`` (csrc/device_math.cuh
        ``locate``): returns dict(g_n, g_off, g_x0, g_inv); g_n = 0 for tiny or
        degenerate tables (binary search on the device)."""
        xs = np.asarray(xs, dtype=float)
        n = len(xs)
        span = xs[-1] - xs[0]
        if n < 8 or not np.isfinite(span) or span <= 0:
            return dict(g_n=0, g_off=0, g_x0=0.0, g_inv=0.0)
        g_n = 1 << int(np.ceil(np.log2((4.0 if n <= 512 else 1.5) * n)))
        inv = g_n / span
        # Every x the device can compute into bucket b lies in [edge_b - d, edge_b+1 + d]
        # (d covers the rounding of (x - x0) * inv at a bucket edge).  guide[b] = last knot
        # at or below that interval; GUIDE_ONE marks buckets whose interval holds exactly
        # one more knot (one compare settles it), GUIDE_MANY those with several (scan).
        d = (abs(xs[0]) + span) * 1e-12
        lo_edge = xs[0] + np.arange(g_n) / inv - d
        hi_edge = xs[0] + (np.arange(g_n) + 1) / inv + d
        base = np.maximum(np.searchsorted(xs, lo_edge, side="right") - 1, 0)
        inside = np.searchsorted(xs, hi_edge, side="right") - 1 - base
        flag = np.where(inside == 1, GUIDE_ONE, np.where(inside > 1, GUIDE_MANY, 0))
        guide = (base.astype(np.int64) | flag).astype(np.uint32).view(np.int32)
        return dict(g_n=g_n, g_off=self.add_ints(guide, dedupe=True), g_x0=float(xs[0]),
                    g_inv=float(inv))

    def add_linear(self, x, y):
        """interp1d(kind='linear') tables: stable sort by x (scipy's
        ``argsort(kind='mergesort')``), values followed by np.interp's slopes, and a
        search guide.  Returns a dict of SSFunc fields."""
        x = np.asarray(x, dtype=float)
        y = np.asarray(y, dtype=float)
        order = np.argsort(x, kind="mergesort")
        xs, ys = x[order], y[order]
        with np.errstate(divide="ignore", invalid="ignore"):
            slope = (ys[1:] - ys[:-1]) / (xs[1:] - xs[:-1])
        fields = dict(n=len(xs), x_off=self.add_doubles(xs, dedupe=True),
                      c_off=self.add_doubles(np.concatenate([ys, slope, [0.0]])),
                      p=(0.0, 0.0, float(xs[0]), float(xs[-1])))   # p[2:4] = table range (lin_eval)
        fields.update(self.add_bucket_guide(xs))
        return fields

    def add_direct_lookup(self, x, y, fill):
        """Direct look-up form of an ``interp1d(kind='linear', bounds_error=False,
        fill_value=fill)`` table (``SSFunc.q_*``, include/streamsim_b200.h): framed segment
        records + a uniform bucket table whose entries settle the interval with at most two
        compares.  Returns the q_* fields, or {} when the table does not lend itself to it
        (duplicate abscissae, or knots too crowded for 4096 buckets)."""
        x = np.asarray(x, dtype=float)
        y = np.asarray(y, dtype=float)
        order = np.argsort(x, kind="mergesort")
        xs, ys = x[order], y[order]
        n = len(xs)
        span = xs[-1] - xs[0]
        if n < 2 or not np.all(np.isfinite(xs)) or np.any(np.diff(xs) <= 0) or not np.isfinite(span):
            return {}
        slope = (ys[1:] - ys[:-1]) / (xs[1:] - xs[:-1])
        if not np.all(np.isfinite(slope) | np.isnan(ys[1:] - ys[:-1])):
            return {}
        fill = float(fill)
        top = np.nextafter(xs[-1], np.inf)
        # lower ends of the framed intervals, and their records
        knots = np.concatenate([[-np.inf], xs, [top]])
        seg = np.zeros((n + 2, 4))
        seg[0] = (0.0, fill, 0.0, 0.0)
        seg[1:n, 0], seg[1:n, 1], seg[1:n, 2] = xs[:-1], ys[:-1], slope
        seg[n] = (xs[-1], ys[-1], 0.0, 0.0)
        seg[n + 1] = (0.0, fill, 0.0, 0.0)
        for nb in (256, 512, 1024, 2048, 4096):
            inv = nb / span
            d = (abs(xs[0]) + span) * 1e-12            # rounding of (x - x0) * inv at a bucket edge
            lo_edge = np.concatenate([[-np.inf], xs[0] + np.arange(nb + 1) / inv - d])
            hi_edge = np.concatenate([xs[0] + np.arange(nb + 1) / inv + d, [np.inf]])
            base = np.searchsorted(knots, lo_edge, side="right") - 1       # last knot <= every x of the bucket
            inside = np.searchsorted(knots, hi_edge, side="right") - 1 - base
            if inside.max() <= 2:
                break
        else:
            return {}
        padded = np.concatenate([knots, [np.inf, np.inf]])
        rec = np.zeros((nb + 2, 4))
        rec[:, 0], rec[:, 1] = padded[base + 1], padded[base + 2]
        rec[:, 2] = base.astype(np.int64).view(np.float64)                 # int32 in the low word
        return dict(q_n=nb, q_off=self.add_doubles(rec, dedupe=True, align=2),
                    q_seg_off=self.add_doubles(seg, align=2), q_nseg=n + 2, q_x0=float(xs[0]),
                    q_inv=float(inv))

    def add_spline(self, ppoly):
        """scipy PPoly/CubicSpline: breakpoints and c[4, n-1] (highest power first)."""
        x = np.asarray(ppoly.x, dtype=float)
        c = np.asarray(ppoly.c, dtype=float)
        assert c.shape == (4, len(x) - 1)
        x_off = self.add_doubles(x, dedupe=True)
        c_off = self.add_doubles(np.ascontiguousarray(c.T))
        return x_off, c_off, len(x)

    def finalize(self):
        """-> uint8 array whose size is a multiple of 16 bytes."""
        d = np.concatenate(self._d) if self._d else np.zeros(2)
        if d.size % 2:
            d = np.concatenate([d, [0.0]])
        return d.view(np.uint8)


def direct_lookup_eval(blob, q, x):
    """Host model of the device's direct look-up evaluation (csrc ``q_locate`` / ``q_value``)
    on a finalized blob: what the production kernel computes for a curve at ``x``."""
    d = np.asarray(blob).view(np.float64)
    nb, nseg = q["q_n"], q["q_nseg"]
    rec = d[q["q_off"]:q["q_off"] + 4 * (nb + 2)].reshape(-1, 4)
    seg = d[q["q_seg_off"]:q["q_seg_off"] + 4 * nseg].reshape(-1, 4)
    x = np.asarray(x, dtype=float)
    with np.errstate(invalid="ignore", over="ignore"):
        t = np.floor((x - q["q_x0"]) * q["q_inv"])
        b = np.clip(np.where(np.isnan(t), 0.0, t), -1, nb).astype(np.int64) + 1
        j = rec[b, 2].copy().view(np.int64).astype(np.int32) + (x >= rec[b, 0]) + (x >= rec[b, 1])
        return seg[j, 2] * (x - seg[j, 0]) + seg[j, 1]


def _rebase(f, base):
    """Shift the blob offsets of an SSFunc packed into the cold region."""
    if f.kind in (_lib.FN_LINEAR, _lib.FN_SPLINE, _lib.FN_SPLINE_PRODUCT):
        f.x_off += base
        f.c_off += base
        if f.g_n > 0:
            f.g_off += base
        if f.kind == _lib.FN_SPLINE_PRODUCT:
            f.x2_off += base
            f.c2_off += base


def search_guide(sorted_vals, guide_n):
    """guide[g] = last index j with sorted_vals[j] <= g/guide_n (0 if none)."""
    edges = np.arange(guide_n + 1, dtype=float) / guide_n
    return np.maximum(np.searchsorted(sorted_vals, edges, side="right") - 1, 0).astype(np.int32)


def invcdf_index(cs, u, order=None):
    """Sample index of ``inverse_transform_sample`` for uniforms ``u`` given the sorted CDF
    ``cs`` and interp1d's ordering ``order`` (None = identity):
    rint(interp1d(cs, order, fill_value=0)(u)) (reference samplers.py:71-75)."""
    u = np.asarray(u, dtype=float)
    fp = np.arange(len(cs), dtype=float) if order is None else np.asarray(order, dtype=float)
    y = np.interp(u, cs, fp)
    y = np.where((u < cs[0]) | (u > cs[-1]), 0.0, y)
    with np.errstate(invalid="ignore"):                    # NaN CDFs (degenerate pdf) -> garbage,
        return np.rint(y).astype(np.int64)                 # rejected by the callers' validation


def invcdf_segments(cs, order=None):
    """The sample index as an explicit step function of u on [0, 1]: returns ``(thr, val)``
    with ``index(u) = val[#{k : thr[k] <= u}]`` reproducing the reference arithmetic
    (``invcdf_index``) exactly, or None if that cannot be established.

    Between two consecutive sorted CDF knots the interpolated value moves monotonically from
    ``order[j]`` to ``order[j+1]`` (up, or down where interp1d's sort interleaved the grid:
    spline densities that dip below zero), so its rounding passes through every integer in
    between; each passage happens at one exact double, found by bisection on the bit
    patterns of the doubles against the reference arithmetic itself."""
    cs = np.asarray(cs, dtype=float)
    n = len(cs)
    fp = np.arange(n, dtype=np.int64) if order is None else np.asarray(order, dtype=np.int64)
    keep = cs[:-1] < 1.0                                   # u < 1: later intervals are never reached
    a, b = fp[:-1][keep], fp[1:][keep]
    steps = np.abs(b - a)
    total = int(steps.sum())
    if total > 4 * n + 1024:                               # pathological interleaving: generic path
        return None
    j = np.repeat(np.flatnonzero(keep), steps)             # interval of every passage
    within = np.arange(total) - np.repeat(np.cumsum(steps) - steps, steps)
    sign = np.repeat(np.sign(b - a), steps)
    target = np.repeat(a, steps) + sign * (within + 1)     # index value after the passage
    lo = cs[j].view(np.int64).copy()                       # f(lo) is on the near side
    hi = cs[j + 1].view(np.int64).copy()
    for _ in range(64):
        todo = hi - lo > 1
        if not todo.any():
            break
        mid = lo + (hi - lo) // 2
        f = invcdf_index(cs, mid.view(np.float64), fp)
        passed = np.where(sign > 0, f >= target, f <= target)
        hi = np.where(todo & passed, mid, hi)
        lo = np.where(todo & ~passed, mid, lo)
    thr = hi.view(np.float64).copy()
    val = np.concatenate([invcdf_index(cs, np.zeros(1), fp), target])
    if cs[0] > 0.0 and val[0] != fp[0]:                    # fill value 0 below the first knot
        thr = np.concatenate([[cs[0]], thr])
        val = np.concatenate([val[:1], [fp[0]], target])
    if np.any(np.diff(thr) < 0):
        return None
    # validate against the reference arithmetic, including the thresholds themselves
    rng = np.random.default_rng(12345)
    probe = np.concatenate([rng.uniform(size=200000), thr, np.nextafter(thr, -1.0),
                            rng.uniform(thr.min(), min(thr.max(), 1.0), size=50000)
                            if len(thr) else [], [0.0, np.nextafter(1.0, 0.0)]])
    probe = probe[(probe >= 0.0) & (probe < 1.0)]
    if not np.array_equal(val[np.searchsorted(thr, probe, side="right")],
                          invcdf_index(cs, probe, fp)):
        return None
    return thr, val


def invcdf_thresholds(cs):
    """Monotone case of :func:`invcdf_segments`: thr[k] = smallest u with index(u) >= k+1
    (0.0 for indices below index(0)); None when the index is not monotone in u."""
    seg = invcdf_segments(cs)
    if seg is None:
        return None
    thr, val = seg
    if np.any(np.diff(val) != 1):
        return None
    out = np.zeros(len(cs) - 1)
    out[val[0]:val[0] + len(thr)] = thr
    return out


def bucket_lut(thr, buckets):
    """``SSTables.grid_lut`` / ``iso_lut``: one 16-byte entry per bucket [b, b+1)/buckets of
    u -- (step at the left edge | thresholds inside << 32, the first of those thresholds).
    ``thr`` ascending, finite."""
    thr = np.asarray(thr, dtype=np.float64)
    edges = np.searchsorted(thr, np.arange(buckets + 1) / buckets, side="right")
    lo, k = edges[:-1].astype(np.int64), np.diff(edges).astype(np.int64)
    padded = np.concatenate([thr, [np.inf]])
    lut = np.empty((buckets, 2), dtype=np.float64)
    lut[:, 0] = (lo | (k << 32)).view(np.float64)          # two int32 in one double
    lut[:, 1] = padded[lo]
    return lut


WLUT_K_CAP = 4095          # k field of a word-table entry saturates here
WLUT_NEVER = 0xFFFFFFFF


def word_thresholds(thr):
    """A step function's thresholds for a 32-bit draw ``u = w * 2**-32``:
    ``thr[k] <= u  <=>  w > wthr[k]`` with ``wthr[k] = ceil(thr[k] * 2**32) - 1`` (exact: the
    scaling is a power of two).  Returns ``(wthr, n_below)``: thresholds at or below zero are
    passed by every word and are dropped (``n_below`` of them, a constant added to every
    step); thresholds no word reaches become 0xFFFFFFFF."""
    thr = np.asarray(thr, dtype=np.float64)
    if np.any(np.diff(thr) < 0) or np.any(np.isnan(thr)):
        raise ValueError("thresholds must be ascending")
    scaled = np.ceil(np.minimum(thr, 2.0) * 4294967296.0)   # integers, exact in double
    n_below = int(np.count_nonzero(scaled <= 0.0))
    w = np.clip(scaled[n_below:] - 1.0, 0.0, float(WLUT_NEVER))
    return w.astype(np.uint32), n_below


def bucket_wlut(wthr, n_below, buckets):
    """``SSTables.grid_wlut`` / ``iso_wlut`` (see the header): one 16-byte entry per bucket of
    the word, ``(lo | min(k, 4095) << 20, t0, t1, t2)``."""
    shift = 32 - int(buckets).bit_length() + 1
    if buckets & (buckets - 1) or not 1 <= buckets <= 1 << 32:
        raise ValueError("buckets must be a power of two")
    # w > wthr[k] for every word of buckets >= b  <=>  wthr[k] < b << shift
    starts = (np.arange(buckets + 1, dtype=np.uint64) << np.uint64(shift))
    w64 = wthr.astype(np.uint64)
    edges = np.searchsorted(w64, starts, side="left")     # thresholds passed at the left edge
    edges[-1] = np.count_nonzero(w64 < WLUT_NEVER)         # 0xFFFFFFFF is never passed
    lo, k = edges[:-1].astype(np.int64), np.diff(edges).astype(np.int64)
    if lo.max(initial=0) + n_below >= 1 << 20:
        raise ValueError("too many steps for a word table")
    padded = np.concatenate([w64, np.full(3, WLUT_NEVER, dtype=np.uint64)])
    lut = np.empty((buckets, 4), dtype=np.uint32)
    lut[:, 0] = ((lo + n_below) | (np.minimum(k, WLUT_K_CAP) << 20)).astype(np.uint32)
    for c in range(3):
        lut[:, 1 + c] = np.where(k > c, padded[lo + c], WLUT_NEVER).astype(np.uint32)
    return lut


def wlut_step(lut, wthr, n_below, w):
    """Host model of the device look-up (csrc ``step_from_words``): the step of word ``w``."""
    w = np.asarray(w, dtype=np.uint64)
    buckets = len(lut)
    shift = 32 - buckets.bit_length() + 1
    e = lut[(w >> np.uint64(shift)).astype(np.int64)].astype(np.int64)
    lo, k = e[:, 0] & 0xFFFFF, e[:, 0] >> 20
    step = lo + (w > e[:, 1]) + (w > e[:, 2]) + (w > e[:, 3])
    more = (k > 3) & (w > e[:, 3].astype(np.uint64))
    for i in np.flatnonzero(more):
        first = int(lo[i]) - n_below + 3                   # index into wthr
        last = len(wthr) if k[i] == WLUT_K_CAP else int(lo[i]) - n_below + int(k[i])
        seg = wthr[first:last].astype(np.uint64)
        step[i] = lo[i] + 3 + np.searchsorted(seg, w[i], side="left")
    return step


def snr_error_threshold(snr_min=SNR_MIN):
    """Largest magerr e with fl(1/e) >= snr_min: turns the reference's
    ``1.0/magerr >= 5.0`` (observed.py:278-279) into one compare, bit-exactly."""
    e = 1.0 / snr_min
    while 1.0 / np.nextafter(e, np.inf) >= snr_min:
        e = np.nextafter(e, np.inf)
    while not (1.0 / e >= snr_min):
        e = np.nextafter(e, -np.inf)
    return float(e)


def snr_loge_threshold(sys_error, snr_min=SNR_MIN):
    """The SNR cut decided on x = log10(statistical error) (``SSParams.snr_loge_max``):
    ``magerr(x) = sqrt((10**x)**2 + sys**2)`` is monotone in x, so
    ``1.0/magerr >= snr_min`` (reference observed.py:278-279 on surveys.py:274-290) holds
    for x <= x*, found by bisection against NumPy's own arithmetic.  NaN when no x passes."""
    def passes(x):
        with np.errstate(over="ignore", divide="ignore"):
            stat = 10 ** np.atleast_1d(np.asarray(x, dtype=float))
            return 1.0 / np.sqrt(stat ** 2 + sys_error ** 2) >= snr_min
    lo, hi = -300.0, 300.0
    if not passes(lo)[0]:
        return float("nan")
    if passes(hi)[0]:
        return float("inf")
    # bit patterns of doubles are ordered like the doubles on each side of zero: bisect the
    # value first, then walk the last bits
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if mid == lo or mid == hi:
            break
        if passes(mid)[0]:
            lo = mid
        else:
            hi = mid
    # Around the crossing the rounding of 10**x, the square and the square root can make
    # the verdict flicker over a few neighbouring doubles (libm / SIMD dependent: the
    # reference itself is only defined up to that).  Take the last double below the FIRST
    # failure of a window around the crossing, and insist that the window's ends are clean.
    near = lo + np.arange(-512, 513) * abs(np.spacing(lo))
    ok = passes(near)
    if not ok[:256].all() or ok[-256:].any():
        raise AssertionError("the SNR verdict is not settled 256 ulps away from its threshold")
    return float(near[np.argmin(ok) - 1])


def frame_matrix(frame):
    """3x3 ICRS->stream rotation from a GreatCircleFrame, ndarray or None."""
    if frame is None:
        return np.eye(3)
    R = getattr(frame, "R", frame)
    R = np.asarray(R, dtype=float)
    if R.shape != (3, 3):
        raise ValueError("frame must be a GreatCircleFrame or a 3x3 rotation matrix")
    return R


# ------------------------------------------------------------------ the engine
class StarEngine:
    """Device-resident configuration of one generate / rotate / observe pipeline.

    Parameters
    ----------
    stream : stream_sim_b200.model.StreamModel or None
        Needed for the generate stage.
    survey : stream_sim_b200.surveys.Survey or None
        Needed for the observe stage.
    frame : GreatCircleFrame or 3x3 array, optional
        Needed for the rotate stage.
    bands, nside, dust_correction, perfect_galstarsep, detection_mag_cut :
        ``StreamInjector.inject`` keywords (reference observed.py:95-131).
    """

    def __init__(self, stream=None, survey=None, frame=None, device=None, bands=("r", "g"),
                 nside=4096, dust_correction=True, perfect_galstarsep=False,
                 detection_mag_cut=("g",), nest=False, hist=None, cdf_steps=1e5,
                 grid_table=True, math="f64", iso_direct=True, fused_kernel=True,
                 packed_maps="auto", normals="f32", map_dtype="f64"):
        torch = _torch()
        self.lib = _lib.load()
        self.device = require_device(device)
        self.params = _lib.SSParams()
        self.tables = _lib.SSTables()
        self.maps = _lib.SSMaps()
        self.has_generate = stream is not None
        self.has_observe = survey is not None
        self._keep = []                    # tensors referenced by raw pointers
        # two regions of the table blob: `hot` is staged into shared memory by every CTA;
        # `cold` holds the stream-model tables (track / distance-modulus knots, inverse-CDF
        # guide), which the kernel only touches on the generic path once a phi1 grid table
        # exists -- shared memory not spent on them is L1 cache for the gathers
        hot, cold = BlobPacker(), BlobPacker()
        p = self.params
        self._grid_table = bool(grid_table)
        self._iso_direct = bool(iso_direct)
        if math not in ("f64", "fast"):
            raise ValueError("math must be 'f64' or 'fast'")
        self.params.math_mode = _lib.MATH_FAST if math == "fast" else _lib.MATH_F64
        self.params.fused_kernel = int(bool(fused_kernel))
        if map_dtype not in ("f64", "f32"):
            raise ValueError("map_dtype must be 'f64' (the reference's) or 'f32'")
        self._map_f32 = map_dtype == "f32"
        if normals not in ("f32", "f64"):
            raise ValueError("normals must be 'f32' (fp32 Box-Muller, the default) or 'f64'")
        self.params.normals_f64 = int(normals == "f64")
        self._packed_maps = packed_maps
        self._describe_hist(hist)          # before the stream: the phi1 grid rows carry their bin
        if survey is not None:
            self._describe_survey(survey, hot, bands, nside, dust_correction,
                                  perfect_galstarsep, detection_mag_cut, nest)
        if stream is not None:
            self._describe_stream(stream, hot, cold, int(cdf_steps))
        self.set_frame(frame)
        hot_blob, cold_blob = hot.finalize(), cold.finalize()
        base = hot_blob.size // 8
        for f in (p.track_center, p.track_spread, p.dist_center, p.dist_spread):
            _rebase(f, base)
        p.cdf_guide_off += base
        blob = np.concatenate([hot_blob, cold_blob])
        self.blob = torch.from_numpy(blob).to(self.device)
        self.tables.blob = self.blob.data_ptr()
        self.tables.blob_hot_bytes = (hot_blob.size if p.density_kind == _lib.DENSITY_GRID
                                      else blob.size)
        self.tables.blob_bytes = self.blob.numel()

    # ---- generate-side description (reference model.py:57-104, samplers.py)
    def _describe_stream(self, stream, hot, packer, cdf_steps):
        """``packer`` (the cold region) receives the density guide and the track /
        distance-modulus tables, ``hot`` the two-stage isochrone tables."""
        from . import samplers as S
        torch = _torch()
        p = self.params
        dens = getattr(stream.density, "density", None)   # DensityModel(None) has no sampler
        if dens is None:
            p.density_kind = _lib.DENSITY_NONE
        elif isinstance(dens, S.UniformSampler):
            p.density_kind = _lib.DENSITY_UNIFORM
            p.density_lo, p.density_hi = float(dens.xmin), float(dens.xmax)
        elif isinstance(dens, S.GaussianSampler):
            p.density_kind = _lib.DENSITY_GAUSSIAN
            p.density_lo, p.density_hi = float(dens.mu), float(dens.sigma)
        elif isinstance(dens, S.InterpolationSampler):
            p.density_kind = _lib.DENSITY_INVCDF
            xg, pdf = dens.grid(cdf_steps)
            cs, order = S.inverse_cdf_tables(xg, pdf)
            self._set_cdf(xg, cs, order)
            p.cdf_guide_off = packer.add_ints(search_guide(cs, CDF_GUIDE))
            p.cdf_guide_n = CDF_GUIDE
            if self._grid_table and not self.tables.grid:
                self._set_grid_table(stream, xg, cs, order)
        else:
            raise Exception(f"Unrecognized sampler: {type(dens).__name__}")

        def sampler_kind(track):
            kind = track._config.get("sampler", "Gaussian").lower()
            if kind == "gaussian":
                return _lib.SAMPLER_GAUSSIAN
            if kind == "uniform":
                return _lib.SAMPLER_UNIFORM
            raise Exception(f"Unrecognized sampler: {kind}")

        p.track_center = stream.track.center.describe(packer)
        p.track_spread = stream.track.spread.describe(packer)
        p.track_sampler = sampler_kind(stream.track)
        if stream.distance_modulus is not None:
            p.has_dist = 1
            p.dist_center = stream.distance_modulus.center.describe(packer)
            p.dist_spread = stream.distance_modulus.spread.describe(packer)
            p.dist_sampler = sampler_kind(stream.distance_modulus)
        iso = getattr(stream, "isochrone", None)
        if iso is not None:
            tab = iso.table
            p.has_iso = 1
            if self._iso_direct:
                self._set_iso_direct(tab)
            else:
                self._pack_iso_tables(tab, hot)

    def _set_iso_direct(self, tab):
        """Composite mags(u) table + bucket lookup (csrc ``iso_sample_direct``)."""
        torch = _torch()
        p = self.params
        U, G, sg, R, sr = tab.composite()
        n = len(U)
        # mag_b(u) = fma(slope_b, u, a_b): the intercepts are folded in extended precision
        ld = np.longdouble
        rows = np.zeros((n, 4))
        rows[:, 0] = (ld(1) * G - ld(1) * sg * U).astype(np.float64)
        rows[:, 1] = sg
        rows[:, 2] = (ld(1) * R - ld(1) * sr * U).astype(np.float64)
        rows[:, 3] = sr
        thr = np.concatenate([U[1:], [np.inf]])               # pass from row k to k+1
        lut = bucket_lut(thr[:-1], ISO_LUT)
        self.iso_rows = torch.from_numpy(rows).to(self.device)
        self.iso_lut = torch.from_numpy(lut).to(self.device)
        self.iso_thr = torch.from_numpy(thr).to(self.device)
        self.tables.iso_rows = self.iso_rows.data_ptr()
        self.tables.iso_lut = self.iso_lut.data_ptr()
        self.tables.iso_thr = self.iso_thr.data_ptr()
        p.iso_direct, p.iso_n, p.iso_lut_n = 1, n, ISO_LUT
        self._set_word_table("iso", thr[:-1], ISO_WLUT)


    def _pack_iso_tables(self, tab, packer):
        """Two-stage tables in the shared-memory blob (csrc ``iso_sample``)."""
        p = self.params
        if True:
            mg, mr = tab.mags_gr()
            p.iso_n = len(tab.mass_init)
            with np.errstate(divide="ignore", invalid="ignore"):
                dm = tab.mass_init[1:] - tab.mass_init[:-1]
                sg = np.concatenate([(mg[1:] - mg[:-1]) / dm, [0.0]])
                sr = np.concatenate([(mr[1:] - mr[:-1]) / dm, [0.0]])
            p.iso_mass_off = packer.add_doubles(tab.mass_init)
            p.iso_mag_off[_lib.BAND_G] = packer.add_doubles(np.concatenate([mg, sg]))
            p.iso_mag_off[_lib.BAND_R] = packer.add_doubles(np.concatenate([mr, sr]))
            g = packer.add_bucket_guide(tab.mass_init)
            p.iso_guide_n, p.iso_guide_off = g["g_n"], g["g_off"]
            p.iso_guide_x0, p.iso_guide_inv = g["g_x0"], g["g_inv"]
            n = len(tab.cdf)
            lo, hi = float(tab.mass_grid[0]), float(tab.mass_grid[-1])
            step = (hi - lo) / (n - 1)
            ref = np.arange(n) * step + lo
            ref[-1] = hi
            if not np.array_equal(ref, tab.mass_grid):
                raise AssertionError("IMF mass grid is not reproducible as lo + j*step")
            p.imf_n, p.imf_mass_lo, p.imf_mass_hi, p.imf_mass_step = n, lo, hi, step
            p.imf_cdf_off = packer.add_doubles(tab.cdf)
            p.imf_guide_off = packer.add_ints(search_guide(tab.cdf, IMF_GUIDE))
            p.imf_guide_n = IMF_GUIDE

    def _set_cdf(self, xg, cs, order):
        """Upload the phi1 inverse-CDF tables (reference samplers.py:69-76)."""
        torch = _torch()
        p = self.params
        n = len(cs)
        perm = order.astype(float)
        with np.errstate(divide="ignore", invalid="ignore"):
            slope = np.concatenate([(perm[1:] - perm[:-1]) / (cs[1:] - cs[:-1]), [0.0]])
        pairs = np.ascontiguousarray(np.stack([cs, slope], axis=1))
        self.cdf_pairs = torch.from_numpy(pairs).to(self.device)
        self.tables.cdf_pairs = self.cdf_pairs.data_ptr()
        p.cdf_n = n
        p.cdf_first, p.cdf_last = float(cs[0]), float(cs[-1])
        p.cdf_identity = int(np.array_equal(order, np.arange(n)))
        if not p.cdf_identity:
            self.cdf_perm = torch.from_numpy(order.astype(np.int32)).to(self.device)
            self.tables.cdf_perm = self.cdf_perm.data_ptr()
        lo, hi = float(xg[0]), float(xg[-1])
        step = (hi - lo) / (n - 1)
        ref = np.arange(n) * step + lo
        ref[-1] = hi
        p.density_lo, p.density_hi, p.grid_step = lo, hi, step
        if not np.array_equal(ref, xg):      # arbitrary `vals`: ship the grid itself
            self.grid = torch.from_numpy(np.ascontiguousarray(xg, dtype=float)).to(self.device)
            self.tables.grid = self.grid.data_ptr()

    def _set_grid_table(self, stream, xg, cs, order):
        """SS_DENSITY_GRID: the sample index as an exact step function of u (thresholds +
        one 64-byte row per step holding every phi1-only quantity of the generate/rotate
        stages, evaluated with the host (reference) functions).  Works for monotone CDFs
        (one step per grid point) and for the non-monotone ones of spline densities that
        dip below zero (atlas, phoenix: the index also steps down)."""
        torch = _torch()
        seg = invcdf_segments(cs, order)
        if seg is None:
            return
        thr, val = seg
        idx = np.concatenate([val, [0]])                   # last row: interp1d's fill value 0
        x = xg[idx]
        with np.errstate(invalid="ignore"):
            c, s = stream.track.center(x), stream.track.spread(x)
            if stream.distance_modulus is not None:
                dc, ds = stream.distance_modulus.center(x), stream.distance_modulus.spread(x)
            else:
                dc = ds = np.zeros(len(x))
        rad = np.radians(x)
        thr_full = np.concatenate([thr, [np.inf]])
        # column 0 is not read as a number by any kernel: its low 32 bits carry the phi1
        # histogram bin of the row's x (the production kernel never bins phi1 itself)
        col0 = np.zeros(len(x), dtype=np.int64)
        col0[:] = self.hist_bin("phi1", x).astype(np.int64) & 0xFFFFFFFF
        rows = np.stack([col0.view(np.float64), x, c, s, dc, ds, np.sin(rad), np.cos(rad)], axis=1)
        lut = bucket_lut(thr, GRID_LUT)
        self.grid_rows = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float64)).to(self.device)
        self.grid_lut = torch.from_numpy(lut).to(self.device)
        self.grid_thr = torch.from_numpy(thr_full).to(self.device)
        self.tables.grid_rows = self.grid_rows.data_ptr()
        self.tables.grid_lut = self.grid_lut.data_ptr()
        self.tables.grid_thr = self.grid_thr.data_ptr()
        self.params.grid_lut_n = GRID_LUT
        self.params.grid_nthr = len(thr)
        if cs[-1] >= 1.0:                                  # no draw in [0, 1) takes the fill row
            self._set_word_table("grid", thr, GRID_WLUT)
        self.params.density_kind = _lib.DENSITY_GRID

    def _set_word_table(self, which, thr, buckets):
        """The 32-bit-word form of a step table (``SSTables.<which>_wlut / _wthr``)."""
        torch = _torch()
        try:
            wthr, n_below = word_thresholds(thr)
            lut = bucket_wlut(wthr, n_below, buckets)
        except ValueError:
            return
        # indexed by step on the device: the dropped leading thresholds keep their slots
        full = np.concatenate([np.zeros(n_below, dtype=np.uint32), wthr, np.full(4, WLUT_NEVER, np.uint32)])
        wl = torch.from_numpy(lut.view(np.int32)).to(self.device)
        wt = torch.from_numpy(full.view(np.int32)).to(self.device)
        setattr(self, which + "_wlut", wl)
        setattr(self, which + "_wthr", wt)
        setattr(self.tables, which + "_wlut", wl.data_ptr())
        setattr(self.tables, which + "_wthr", wt.data_ptr())
        setattr(self.params, which + "_wlut_n", buckets)

    def model_tensors(self):
        """Device tensors that make up the model/survey tables (not the maps)."""
        names = ("blob", "cdf_pairs", "cdf_perm", "grid", "grid_rows", "grid_lut", "grid_thr",
                 "iso_rows", "iso_lut", "iso_thr", "grid_wlut", "grid_wthr", "iso_wlut", "iso_wthr")
        return [getattr(self, k) for k in names if getattr(self, k, None) is not None]

    # ---- rotate
    def set_frame(self, frame):
        R = frame_matrix(frame)
        for k in range(9):
            self.params.R[k] = float(R.flat[k])
        self.frame = frame

    # ---- observe-side description (reference surveys.py getters, observed.py:146-291)
    def _describe_survey(self, survey, packer, bands, nside, dust_correction,
                         perfect_galstarsep, detection_mag_cut, nest):
        torch = _torch()
        from .healpix import npix2nside
        p = self.params
        for band in bands:
            if band not in ("r", "g"):
                raise ValueError("Currently only 'r' and 'g' bands are supported.")
        cb = survey.completeness_band
        if cb not in bands:
            raise ValueError(f"Detection flag requires '{cb}' band to be in bands.")
        if survey.ebv_map is None:
            raise ValueError("E(B-V) map not loaded")
        if survey.log_photo_error is None:
            raise ValueError("Photo error model not loaded")
        if survey.completeness is None:
            raise ValueError("Efficiency function for type 'completeness' not loaded")
        self.bands = tuple(bands)
        p.nside_pix = int(nside)
        p.nest = int(bool(nest))
        p.dust_correction = int(bool(dust_correction))
        p.perfect_galstarsep = int(bool(perfect_galstarsep))
        p.completeness_band = _lib.BAND_INDEX[cb]
        p.delta_saturation = float(survey.delta_saturation)
        p.snr_err_max = snr_error_threshold(SNR_MIN)

        def device_map(arr):
            if isinstance(arr, torch.Tensor):
                t = arr.to(device=self.device, dtype=torch.float64).contiguous()
            else:
                t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(self.device)
            if self._map_f32:                  # map_dtype="f32": every kernel sees the rounded values
                t = t.to(torch.float32).to(torch.float64)
            self._keep.append(t)
            return t

        ebv = device_map(survey.ebv_map)
        self._map_tensors = {"ebv": ebv}
        p.nside_ebv = npix2nside(ebv.numel())
        self.maps.ebv = ebv.data_ptr()
        for band in ("g", "r"):
            b = _lib.BAND_INDEX[band]
            if band not in bands:
                continue
            if band not in survey.maglim_maps or survey.maglim_maps[band] is None:
                raise ValueError(f"Band '{band}' not available. Available: {survey.bands}")
            if band not in survey.coeff_extinc:
                raise ValueError(f"Band '{band}' not available. Available: {survey.bands}")
            ml = device_map(survey.maglim_maps[band])
            self._map_tensors[band] = ml
            p.band_present[b] = 1
            p.nside_maglim[b] = npix2nside(ml.numel())
            self.maps.maglim[b] = ml.data_ptr()
            p.coeff_extinc[b] = float(survey.coeff_extinc[band])
            p.saturation[b] = float(survey.saturation[band])
            p.sys_error[b] = float(survey.sys_error.get(band, 0.005))
            p.snr_cut[b] = int(band in detection_mag_cut)
            p.snr_loge_max[b] = snr_loge_threshold(p.sys_error[b])
        from .surveys import TabulatedFunction
        pe = TabulatedFunction.coerce(survey.log_photo_error)
        p.photo_error = pe.describe(packer)
        dsat = survey.delta_saturation
        p.log_err_at_dsat = float(pe(dsat))
        p.err_bright = float(10 ** pe(dsat - 1))
        for b in (_lib.BAND_G, _lib.BAND_R):       # saturated stars carry this constant error
            with np.errstate(divide="ignore"):
                tot = np.sqrt(np.float64(p.err_bright) ** 2 + np.float64(p.sys_error[b]) ** 2)
                p.snr_bright_ok[b] = int(1.0 / tot >= SNR_MIN)
        self._pack_maps()
        p.completeness = TabulatedFunction.coerce(survey.completeness).describe(packer)
        if perfect_galstarsep:
            det = getattr(survey, "efficiency_detection", None)
            if det is None:
                raise ValueError("Efficiency function for type 'detection_efficiency' not loaded")
            p.detection = TabulatedFunction.coerce(det).describe(packer)

    def _pack_maps(self):
        """Interleave the three survey maps as [npix][4] = (ebv, maglim_g, maglim_r, 0): the
        production kernel then gathers ONE 32-byte sector per star instead of three.  Only
        when the three maps share one nside; ``packed_maps='auto'`` also wants the copy
        (4/3 of the separate maps) to fit comfortably in free device memory."""
        torch = _torch()
        p = self.params
        want = self._packed_maps
        same = (p.band_present[0] and p.band_present[1] and not p.nest and
                p.nside_maglim[0] == p.nside_ebv and p.nside_maglim[1] == p.nside_ebv)
        if not want or not same:
            return
        npix = 12 * p.nside_ebv ** 2
        if self._map_f32:
            packed = torch.zeros((npix, 4), dtype=torch.float32, device=self.device)
            for col, key in enumerate(("ebv", "g", "r")):
                packed[:, col] = self._map_tensors[key].to(torch.float32)
            self.packed_maps = packed
            self.maps.packed32 = packed.data_ptr()
            self.maps.nside_packed = p.nside_ebv
            return
        if want == "auto":
            free, _ = torch.cuda.mem_get_info(self.device)
            if 32 * npix > 0.25 * free:
                return
        packed = torch.zeros((npix, 4), dtype=torch.float64, device=self.device)
        for col, key in enumerate(("ebv", "g", "r")):
            packed[:, col] = self._map_tensors[key]
        self.packed_maps = packed
        self.maps.packed = packed.data_ptr()
        self.maps.nside_packed = p.nside_ebv

    def hist_bin(self, name, values):
        """Histogram bin of ``values`` exactly as the kernels compute it (fp32, two roundings;
        -1 = outside or NaN)."""
        k = _lib.HIST_NAMES.index(name)
        lo = np.float32(self.params.hist_lo[k])
        inv = np.float32(self.params.hist_inv_width[k])
        with np.errstate(invalid="ignore", over="ignore"):
            t = (np.asarray(values, dtype=np.float64).astype(np.float32) - lo) * inv
            ok = (t >= 0) & (t < np.float32(_lib.HIST_BINS))
            return np.where(ok, np.where(ok, t, 0).astype(np.int32), np.int32(-1)).astype(np.int32)

    def _describe_hist(self, hist):
        p = self.params
        ranges = dict(phi1=(-20.0, 20.0), phi2=(-5.0, 5.0), dist=(10.0, 25.0),
                      mag_g=(10.0, 35.0), g_minus_r=(-0.5, 2.5))
        if hist:
            ranges.update(hist if isinstance(hist, dict) else {})
        p.hist_enable = int(bool(hist))
        for k, name in enumerate(_lib.HIST_NAMES):
            lo, hi = ranges[name]
            p.hist_lo[k] = lo
            p.hist_inv_width[k] = _lib.HIST_BINS / (hi - lo)
        self.hist_ranges = ranges

    # ---- launches
    def _stream_ptr(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def new_counters(self):
        return _torch().zeros(_lib.COUNTERS_LEN, dtype=_torch().int64, device=self.device)

    def allocate(self, n, columns, dtype="f64", pix=False, draws=False):
        """Structure-of-arrays output buffers for ``columns`` (``draws=True`` adds the
        [8][n] export of the random draws the kernel used, rows ``_lib.DRAW_NAMES``)."""
        torch = _torch()
        ft = torch.float64 if dtype == "f64" else torch.float32
        out = {}
        for c in columns:
            if c in COLUMNS_FLOAT:
                out[c] = torch.empty(n, dtype=ft, device=self.device)
            elif c in ("flag_observed", "flag_perfect_galstarsep"):
                out[c] = torch.empty(n, dtype=torch.uint8, device=self.device)
        if pix:
            out["pix"] = torch.empty(n, dtype=torch.int64, device=self.device)
        if draws:
            out["draws"] = torch.full((_lib.N_DRAWS, n), float("nan"), dtype=torch.float64,
                                      device=self.device)
        return out

    def _out_struct(self, out, counters, dtype):
        o = _lib.SSOut()
        o.dtype = _lib.DTYPE_F64 if dtype == "f64" else _lib.DTYPE_F32
        ptr = lambda k: out[k].data_ptr() if k in out and out[k] is not None else None
        o.phi1, o.phi2, o.dist = ptr("phi1"), ptr("phi2"), ptr("dist")
        o.mag[_lib.BAND_G], o.mag[_lib.BAND_R] = ptr("mag_g"), ptr("mag_r")
        o.ra, o.dec = ptr("ra"), ptr("dec")
        o.mag_obs[_lib.BAND_G], o.mag_obs[_lib.BAND_R] = ptr("mag_g_obs"), ptr("mag_r_obs")
        o.magerr[_lib.BAND_G], o.magerr[_lib.BAND_R] = ptr("magerr_g"), ptr("magerr_r")
        o.flag_observed = ptr("flag_observed")
        o.flag_perfect = ptr("flag_perfect_galstarsep")
        o.pix = ptr("pix")
        o.draws_out = ptr("draws")
        o.counters = counters.data_ptr() if counters is not None else None
        return o

    def _as_device_f64(self, arr):
        torch = _torch()
        if arr is None:
            return None
        if isinstance(arr, torch.Tensor):
            t = arr.to(device=self.device, dtype=torch.float64).contiguous()
        else:
            t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(self.device)
        return t

    def _check_out(self, out, n, dtype, counters, host):
        """Every output column must hold n elements of the launch's dtype, contiguous, on the
        engine's device (pinned/CPU for the host variants): the kernel sees raw pointers."""
        torch = _torch()
        if dtype not in ("f64", "f32"):
            raise ValueError("dtype must be 'f64' or 'f32'")
        ft = torch.float64 if dtype == "f64" else torch.float32
        want = {c: ft for c in COLUMNS_FLOAT}
        want.update(flag_observed=torch.uint8, flag_perfect_galstarsep=torch.uint8, pix=torch.int64)
        for name, t in out.items():
            if t is None or name == "counters":
                continue
            if name == "draws":
                if t.dtype != torch.float64 or t.numel() != _lib.N_DRAWS * n or not t.is_contiguous():
                    raise ValueError("out['draws'] must be a contiguous float64 [8][n] tensor")
            elif name in want:
                if t.dtype != want[name]:
                    raise ValueError(f"out['{name}'] has dtype {t.dtype}, expected {want[name]}")
                if t.numel() < n or not t.is_contiguous():
                    raise ValueError(f"out['{name}'] must be contiguous with at least {n} elements")
            else:
                continue
            if host:
                if t.device.type != "cpu":
                    raise ValueError(f"out['{name}'] must be a host tensor")
            elif t.device != self.device:
                raise ValueError(f"out['{name}'] is on {t.device}, the engine on {self.device}")
        cnt = counters if counters is not None else None
        if cnt is not None:
            if cnt.dtype != torch.int64 or cnt.numel() != _lib.COUNTERS_LEN or not cnt.is_contiguous():
                raise ValueError(f"counters must be a contiguous int64[{_lib.COUNTERS_LEN}] tensor")
            if (cnt.device.type != "cpu") if host else (cnt.device != self.device):
                raise ValueError("counters live on the wrong device")

    def _draws_struct(self, draws, hold, n=None):
        d = _lib.SSDraws()
        if not draws:
            return d

        def get(*names):
            for nm in names:
                if nm in draws and draws[nm] is not None:
                    t = self._as_device_f64(draws[nm])
                    if n is not None and t.numel() != n:
                        raise ValueError(f"draws['{nm}'] has {t.numel()} elements, the launch "
                                         f"covers {n} stars")
                    hold.append(t)
                    return t.data_ptr()
            return None
        p = self.params
        d.u_phi1 = get("z_phi1") if p.density_kind == _lib.DENSITY_GAUSSIAN else get("u_phi1")
        d.n_phi2 = get("z_phi2") if p.track_sampler == _lib.SAMPLER_GAUSSIAN else get("u_phi2")
        d.n_dist = get("z_dist") if p.dist_sampler == _lib.SAMPLER_GAUSSIAN else get("u_dist")
        d.u_mass = get("u_mass")
        d.z_flux[_lib.BAND_G], d.z_flux[_lib.BAND_R] = get("z_g"), get("z_r")
        d.u_det, d.u_det2 = get("u_det"), get("u_det2")
        return d

    def run(self, stages, n, out, catalog=None, draws=None, seed=0, offset=0, counters=None,
            dtype="f64"):
        """Launch ``stages`` for ``n`` stars into the column tensors of ``out``."""
        n = int(n)
        if n == 0:
            return out
        hold = []
        cat = _lib.SSCatalog()

        def column(arr, what):
            t = self._as_device_f64(arr)
            if t.numel() != n:
                raise ValueError(f"{what} has {t.numel()} elements, the launch covers {n} stars")
            hold.append(t)
            return t.data_ptr()
        if catalog:
            for name in ("phi1", "phi2", "dist", "ra", "dec"):
                if catalog.get(name) is not None:
                    setattr(cat, name, column(catalog[name], f"catalog column '{name}'"))
            for band in ("g", "r"):
                if catalog.get("mag_" + band) is not None:
                    cat.mag[_lib.BAND_INDEX[band]] = column(catalog["mag_" + band],
                                                            f"catalog column 'mag_{band}'")
        d = self._draws_struct(draws, hold, n)
        self._check_out(out, n, dtype, counters, host=False)
        o = self._out_struct(out, counters, dtype)
        with _torch().cuda.device(self.device):
            rc = self.lib.ss_run(stages, C.byref(self.params), C.byref(self.tables),
                                 C.byref(self.maps), C.byref(cat), C.byref(d), C.byref(o),
                                 int(n), int(offset), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                 self._stream_ptr())
        _lib.check(rc)
        # tensors in `hold` are only read by work already queued on the current stream;
        # the caching allocator reuses their memory in stream order, so dropping them is safe
        return out

    def generate(self, n, seed=0, offset=0, draws=None, counters=None, dtype="f64", out=None):
        cols = [c for c in GENERATE_COLUMNS
                if not (c == "dist" and not self.params.has_dist)
                and not (c.startswith("mag_") and not self.params.has_iso)]
        out = out if out is not None else self.allocate(n, cols, dtype)
        return self.run(_lib.STAGE_GENERATE, n, out, None, draws, seed, offset, counters, dtype)

    def rotate(self, phi1, phi2, counters=None, dtype="f64", out=None):
        n = len(phi1)
        out = out if out is not None else self.allocate(n, ROTATE_COLUMNS, dtype)
        return self.run(_lib.STAGE_ROTATE, n, out, dict(phi1=phi1, phi2=phi2), None, 0, 0,
                        counters, dtype)

    def _observe_columns(self, pix):
        cols = []
        for b in ("g", "r"):
            if self.params.band_present[_lib.BAND_INDEX[b]]:
                cols += [f"mag_{b}_obs", f"magerr_{b}"]
        cols.append("flag_observed")
        if self.params.perfect_galstarsep:
            cols.append("flag_perfect_galstarsep")
        return cols

    def observe(self, catalog, seed=0, offset=0, draws=None, counters=None, dtype="f64",
                out=None, pix=False, export_draws=False):
        """catalog: dict with (ra, dec) or (phi1, phi2), and mag_g / mag_r."""
        has_radec = catalog.get("ra") is not None and catalog.get("dec") is not None
        n = len(catalog["ra"] if has_radec else catalog["phi1"])
        stages = _lib.STAGE_OBSERVE if has_radec else (_lib.STAGE_ROTATE | _lib.STAGE_OBSERVE)
        cols = self._observe_columns(pix) + ([] if has_radec else list(ROTATE_COLUMNS))
        out = out if out is not None else self.allocate(n, cols, dtype, pix=pix, draws=export_draws)
        return self.run(stages, n, out, catalog, draws, seed, offset, counters, dtype)

    def generate_observe(self, n, seed=0, offset=0, draws=None, counters=None, dtype="f64",
                         out=None, pix=False, columns=None, export_draws=False):
        if out is None:
            cols = columns if columns is not None else (
                list(GENERATE_COLUMNS) + list(ROTATE_COLUMNS) + self._observe_columns(pix))
            out = self.allocate(n, cols, dtype, pix=pix, draws=export_draws)
        st = _lib.STAGE_GENERATE | _lib.STAGE_ROTATE | _lib.STAGE_OBSERVE
        return self.run(st, n, out, None, draws, seed, offset, counters, dtype)

    def run_summary(self, n_total, seed=0, offset=0, chunk=1 << 28, counters=None):
        """Counters + histograms only: the fused path over ``n_total`` stars with no
        per-star column kept (BASELINE config 5: 1e9-1e11 stars).  Launches are
        queued back to back on the current stream; returns the counters tensor."""
        counters = self.new_counters() if counters is None else counters
        done = 0
        while done < n_total:
            n = min(chunk, n_total - done)
            self.generate_observe(n, seed=seed, offset=offset + done, counters=counters, out={})
            done += n
        return counters

    # ---- host-buffer entry (pinned host memory in, pinned host memory out)
    def host_workspace(self, chunk, dtype="f64"):
        torch = _torch()
        nbytes = self.lib.ss_host_workspace_bytes(int(chunk),
                                                  _lib.DTYPE_F64 if dtype == "f64" else _lib.DTYPE_F32)
        return torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def allocate_host(self, n, columns, dtype="f64", pix=False, counters=True):
        """Pinned host column buffers for the *_host entry points."""
        torch = _torch()
        ft = torch.float64 if dtype == "f64" else torch.float32
        out = {}
        for c in columns:
            if c in COLUMNS_FLOAT:
                out[c] = torch.empty(n, dtype=ft).pin_memory()
            elif c in ("flag_observed", "flag_perfect_galstarsep"):
                out[c] = torch.empty(n, dtype=torch.uint8).pin_memory()
        if pix:
            out["pix"] = torch.empty(n, dtype=torch.int64).pin_memory()
        if counters:
            out["counters"] = torch.zeros(_lib.COUNTERS_LEN, dtype=torch.int64).pin_memory()
        return out

    def generate_observe_host(self, n, out_host, workspace, chunk, seed=0, offset=0, dtype="f64"):
        self._check_out(out_host, int(n), dtype, out_host.get("counters"), host=True)
        o = self._out_struct(out_host, out_host.get("counters"), dtype)
        with _torch().cuda.device(self.device):
            rc = self.lib.ss_generate_observe_host(
                C.byref(self.params), C.byref(self.tables), C.byref(self.maps), C.byref(o),
                int(n), int(offset), int(seed) & 0xFFFFFFFFFFFFFFFF,
                C.c_void_p(workspace.data_ptr()), workspace.numel(), int(chunk))
        _lib.check(rc)
        return out_host

    def observe_host(self, catalog_host, out_host, workspace, chunk, seed=0, offset=0,
                     dtype="f64"):
        """catalog_host: dict of (pinned) CPU float64 tensors ra, dec, mag_g, mag_r."""
        cat = _lib.SSCatalog()
        cat.ra, cat.dec = catalog_host["ra"].data_ptr(), catalog_host["dec"].data_ptr()
        for band in ("g", "r"):
            if catalog_host.get("mag_" + band) is not None:
                cat.mag[_lib.BAND_INDEX[band]] = catalog_host["mag_" + band].data_ptr()
        n = catalog_host["ra"].numel()
        for name, t in catalog_host.items():
            if t is not None and (t.device.type != "cpu" or t.dtype != _torch().float64
                                  or t.numel() != n or not t.is_contiguous()):
                raise ValueError(f"catalog_host['{name}'] must be a contiguous float64 host "
                                 f"tensor of {n} elements")
        self._check_out(out_host, n, dtype, out_host.get("counters"), host=True)
        o = self._out_struct(out_host, out_host.get("counters"), dtype)
        with _torch().cuda.device(self.device):
            rc = self.lib.ss_observe_host(
                C.byref(self.params), C.byref(self.tables), C.byref(self.maps), C.byref(cat),
                C.byref(o), int(n), int(offset), int(seed) & 0xFFFFFFFFFFFFFFFF,
                C.c_void_p(workspace.data_ptr()), workspace.numel(), int(chunk))
        _lib.check(rc)
        return out_host

    # ---- summaries
    @staticmethod
    def exported_draws(out, sampler_kinds=None):
        """The ``draws`` export of a launch as the dict the oracle / ``draws=`` accept."""
        d = out["draws"].cpu().numpy()
        named = {k: d[i] for i, k in enumerate(_lib.DRAW_NAMES)}
        res = dict(u_phi1=named["u_phi1"], u_mass=named["u_mass"], z_g=named["z_g"],
                   z_r=named["z_r"], u_det=named["u_det"], u_det2=named["u_det2"])
        # the track slots hold a normal or a uniform depending on the sampler
        res["z_phi2"] = res["u_phi2"] = named["n_phi2"]
        res["z_dist"] = res["u_dist"] = named["n_dist"]
        return res

    @staticmethod
    def summarize(counters):
        """Counter tensor -> dict (one device->host copy)."""
        c = counters.detach().cpu().numpy().astype(np.int64)
        names = ("n_total", "n_observed", "n_perfect_galstarsep", "n_badmag_g", "n_badmag_r",
                 "n_unseen", "n_domain_error", "n_bad_coord", "n_inside_mask")
        out = {k: int(c[i]) for i, k in enumerate(names)}
        for h, name in enumerate(_lib.HIST_NAMES):
            lo = _lib.N_COUNTERS + h * _lib.HIST_BINS
            out["hist_" + name] = c[lo:lo + _lib.HIST_BINS].copy()
        return out

    @staticmethod
    def raise_for_counters(summary):
        """Re-raise, after the fact, what SciPy / healpy would have raised per call."""
        if summary["n_domain_error"]:
            raise ValueError("Domain error in arguments. The `scale` parameter must be positive "
                             "for all distributions, and many distributions have restrictions "
                             "on shape parameters. (%d stars)" % summary["n_domain_error"])
        if summary["n_bad_coord"]:
            raise ValueError("THETA is out of range [0,pi] (%d stars)" % summary["n_bad_coord"])


# ------------------------------------------------------------------ small ops
def _small_call(fn, device, *args):
    torch = _torch()
    with torch.cuda.device(device):
        rc = fn(*args, C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
    _lib.check(rc)


def ang2pix(nside, lon, lat, nest=False, device=None):
    """HEALPix pixel of (lon, lat) in degrees -- ``healpy.ang2pix(..., lonlat=True)``.
    Returns an int64 torch tensor on the device."""
    torch = _torch()
    device = require_device(device)
    lib = _lib.load()
    to = lambda a: (a.to(device=device, dtype=torch.float64).contiguous()
                    if isinstance(a, torch.Tensor)
                    else torch.from_numpy(np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64)).to(device))
    lon_t, lat_t = to(lon), to(lat)
    pix = torch.empty(lon_t.numel(), dtype=torch.int64, device=device)
    _small_call(lib.ss_ang2pix, device, int(nside), int(bool(nest)), lon_t.data_ptr(),
                lat_t.data_ptr(), pix.data_ptr(), lon_t.numel())
    return pix


MATH_FN = {"log10": 0, "exp10": 1, "atan2": 2, "sqrt": 3}


def math_probe(fn, x, y=None, device=None):
    """The kernel's own double-precision log10 / 10**x / atan2(x, y) / sqrt, element-wise
    (``ss_math_probe``); a float64 torch tensor on the device.  Test hook."""
    torch = _torch()
    device = require_device(device)
    lib = _lib.load()
    to = lambda a: torch.from_numpy(np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64)).to(device)
    x_t = to(x)
    y_t = to(y) if y is not None else None
    out = torch.empty_like(x_t)
    _small_call(lib.ss_math_probe, device, MATH_FN[fn], x_t.data_ptr(),
                y_t.data_ptr() if y_t is not None else None, out.data_ptr(), x_t.numel())
    return out


def sample_inverse_cdf(vals, pdf, size, seed=None, u=None, device=None):
    """GPU implementation of ``inverse_transform_sample`` (reference samplers.py:55-76)."""
    from .samplers import inverse_cdf_tables
    torch = _torch()
    eng = StarEngine.__new__(StarEngine)
    eng.lib = _lib.load()
    eng.device = require_device(device)
    eng.params, eng.tables, eng.maps = _lib.SSParams(), _lib.SSTables(), _lib.SSMaps()
    eng._keep = []
    p = eng.params
    cs, order = inverse_cdf_tables(vals, pdf)
    p.density_kind = _lib.DENSITY_INVCDF
    eng._set_cdf(vals, cs, order)
    packer = BlobPacker()
    p.cdf_guide_off = packer.add_ints(search_guide(cs, CDF_GUIDE))
    p.cdf_guide_n = CDF_GUIDE
    p.track_center.kind = p.track_spread.kind = _lib.FN_CONSTANT
    p.track_center.xmin = p.track_spread.xmin = -np.inf
    p.track_center.xmax = p.track_spread.xmax = np.inf
    blob = packer.finalize()
    eng.blob = torch.from_numpy(blob).to(eng.device)
    eng.tables.blob, eng.tables.blob_bytes = eng.blob.data_ptr(), eng.blob.numel()
    eng.set_frame(None)
    eng._describe_hist(None)
    out = eng.allocate(size, ["phi1"])
    seed = _fresh_seed() if seed is None else seed
    eng.run(_lib.STAGE_GENERATE, size, out, None, dict(u_phi1=u) if u is not None else None, seed)
    return out["phi1"].cpu().numpy()


def sample_loc_scale(loc, scale, size, normal, seed=None, draws=None, device=None):
    """``draw * scale + loc`` per star on the GPU (scipy ``rv.rvs`` semantics,
    reference samplers.py:101-102), scalar or per-star loc/scale -- one launch of
    ``ss_loc_scale``; the draws are injected or the stars' own Philox draws."""
    torch = _torch()
    device = require_device(device)
    lib = _lib.load()
    size = int(size)
    loc_a, scale_a = np.asarray(loc, dtype=float), np.asarray(scale, dtype=float)
    if not np.all(scale_a >= 0):
        raise ValueError("Domain error in arguments. The `scale` parameter must be positive for "
                         "all distributions, and many distributions have restrictions on shape "
                         "parameters.")
    seed = _fresh_seed() if seed is None else seed

    def per_star(a):
        if a.ndim == 0:
            return None
        return torch.from_numpy(np.ascontiguousarray(np.broadcast_to(a, (size,)))).to(device)
    loc_t, scale_t = per_star(loc_a), per_star(scale_a)
    draws_t = None
    if draws is not None:
        draws_t = torch.from_numpy(np.ascontiguousarray(draws, dtype=np.float64)).to(device)
        if draws_t.numel() != size:
            raise ValueError(f"draws has {draws_t.numel()} elements, size is {size}")
    out = torch.empty(size, dtype=torch.float64, device=device)
    ptr = lambda t: t.data_ptr() if t is not None else None
    _small_call(lib.ss_loc_scale, device, int(bool(normal)), ptr(loc_t), ptr(scale_t),
                float(loc_a) if loc_a.ndim == 0 else 0.0,
                float(scale_a) if scale_a.ndim == 0 else 0.0, ptr(draws_t), out.data_ptr(),
                size, 0, int(seed) & 0xFFFFFFFFFFFFFFFF)
    return out.cpu().numpy()


def noisy_magnitudes(mag_true, mag_err, z=None, seed=None, device=None):
    """``StreamInjector.sample_measured_magnitudes`` on the GPU (``ss_noisy_magnitudes``):
    float64 array, NaN where the noisy flux is not positive."""
    torch = _torch()
    device = require_device(device)
    lib = _lib.load()
    to = lambda a: torch.from_numpy(np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64)).to(device)
    m, e = to(mag_true), to(mag_err)
    if e.numel() != m.numel():
        e = to(np.broadcast_to(np.asarray(mag_err, dtype=float), (m.numel(),)))
    z_t = to(z) if z is not None else None
    if z_t is not None and z_t.numel() != m.numel():
        raise ValueError("z must have one entry per magnitude")
    out = torch.empty_like(m)
    seed = _fresh_seed() if seed is None else seed
    _small_call(lib.ss_noisy_magnitudes, device, m.data_ptr(), e.data_ptr(),
                z_t.data_ptr() if z_t is not None else None, out.data_ptr(), m.numel(), 0,
                int(seed) & 0xFFFFFFFFFFFFFFFF)
    return out.cpu().numpy()


def bernoulli(prob, u=None, seed=None, device=None):
    """``u <= prob`` per element on the GPU (``ss_bernoulli``); bool array."""
    torch = _torch()
    device = require_device(device)
    lib = _lib.load()
    to = lambda a: torch.from_numpy(np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64)).to(device)
    pr = to(prob)
    u_t = to(u) if u is not None else None
    if u_t is not None and u_t.numel() != pr.numel():
        raise ValueError("u must have one entry per probability")
    out = torch.empty(pr.numel(), dtype=torch.uint8, device=device)
    seed = _fresh_seed() if seed is None else seed
    _small_call(lib.ss_bernoulli, device, pr.data_ptr(),
                u_t.data_ptr() if u_t is not None else None, out.data_ptr(), pr.numel(), 0,
                int(seed) & 0xFFFFFFFFFFFFFFFF)
    return out.cpu().numpy().astype(bool)


def _fresh_seed():
    return int.from_bytes(os.urandom(8), "little")
