"""ctypes binding of libstreamsim_b200.so (C ABI: include/streamsim_b200.h).

There is deliberately no fallback: if the shared library is missing, or a call
fails, an exception is raised.  Nothing in the product path computes per-star
quantities on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

ABI_VERSION = 15

# ---- constants mirrored from the header
STAGE_GENERATE, STAGE_ROTATE, STAGE_OBSERVE = 1, 2, 4
FN_NONE, FN_CONSTANT, FN_LINE, FN_SINUSOID, FN_LINEAR, FN_SPLINE, FN_SPLINE_PRODUCT = range(7)
SAMPLER_GAUSSIAN, SAMPLER_UNIFORM = 0, 1
DENSITY_UNIFORM, DENSITY_INVCDF, DENSITY_GAUSSIAN, DENSITY_NONE, DENSITY_GRID = 0, 1, 2, 3, 4
BAND_G, BAND_R = 0, 1
BAND_INDEX = {"g": BAND_G, "r": BAND_R}
DTYPE_F64, DTYPE_F32 = 0, 1
MATH_F64, MATH_FAST = 0, 1
(CNT_TOTAL, CNT_OBSERVED, CNT_PERFECT, CNT_BADMAG_G, CNT_BADMAG_R, CNT_UNSEEN, CNT_DOMAIN,
 CNT_BADCOORD, CNT_INSIDE) = range(9)
N_COUNTERS, HIST_BINS, N_HIST = 16, 256, 5
N_DRAWS = 8
DRAW_NAMES = ("u_phi1", "n_phi2", "n_dist", "u_mass", "z_g", "z_r", "u_det", "u_det2")
COUNTERS_LEN = N_COUNTERS + N_HIST * HIST_BINS
HIST_NAMES = ("phi1", "phi2", "dist", "mag_g", "g_minus_r")

_i32, _f64, _vp, _i64, _u64 = C.c_int32, C.c_double, C.c_void_p, C.c_int64, C.c_uint64


class SSFunc(C.Structure):
    _fields_ = [("kind", _i32), ("n", _i32), ("x_off", _i32), ("c_off", _i32),
                ("xmin", _f64), ("xmax", _f64), ("p", _f64 * 4),
                ("n2", _i32), ("x2_off", _i32), ("c2_off", _i32), ("g_n", _i32),
                ("g_off", _i32), ("pad_", _i32), ("g_x0", _f64), ("g_inv", _f64),
                ("q_n", _i32), ("q_off", _i32), ("q_seg_off", _i32), ("q_nseg", _i32),
                ("q_x0", _f64), ("q_inv", _f64)]


class SSParams(C.Structure):
    _fields_ = [
        ("density_kind", _i32), ("cdf_n", _i32), ("cdf_guide_n", _i32), ("cdf_identity", _i32),
        ("density_lo", _f64), ("density_hi", _f64), ("grid_step", _f64),
        ("cdf_first", _f64), ("cdf_last", _f64),
        ("track_center", SSFunc), ("track_spread", SSFunc),
        ("dist_center", SSFunc), ("dist_spread", SSFunc),
        ("track_sampler", _i32), ("dist_sampler", _i32), ("has_dist", _i32), ("has_iso", _i32),
        ("iso_n", _i32), ("iso_mass_off", _i32), ("iso_mag_off", _i32 * 2),
        ("iso_guide_n", _i32), ("iso_guide_off", _i32), ("imf_n", _i32), ("imf_cdf_off", _i32),
        ("imf_guide_n", _i32), ("imf_guide_off", _i32), ("cdf_guide_off", _i32), ("iso_direct", _i32),
        ("imf_mass_lo", _f64), ("imf_mass_hi", _f64), ("imf_mass_step", _f64),
        ("iso_guide_x0", _f64), ("iso_guide_inv", _f64),
        ("R", _f64 * 9),
        ("nside_pix", _i32), ("nside_ebv", _i32), ("nside_maglim", _i32 * 2),
        ("band_present", _i32 * 2), ("completeness_band", _i32),
        ("dust_correction", _i32), ("perfect_galstarsep", _i32), ("snr_cut", _i32 * 2),
        ("nest", _i32),
        ("coeff_extinc", _f64 * 2), ("saturation", _f64 * 2), ("sys_error", _f64 * 2),
        ("delta_saturation", _f64), ("snr_err_max", _f64),
        ("snr_loge_max", _f64 * 2), ("snr_bright_ok", _i32 * 2),
        ("log_err_at_dsat", _f64), ("err_bright", _f64),
        ("photo_error", SSFunc), ("completeness", SSFunc), ("detection", SSFunc),
        ("hist_enable", _i32), ("grid_lut_n", _i32),
        ("hist_lo", _f64 * N_HIST), ("hist_inv_width", _f64 * N_HIST),
        ("math_mode", _i32), ("iso_lut_n", _i32), ("grid_nthr", _i32), ("fused_kernel", _i32),
        ("grid_wlut_n", _i32), ("iso_wlut_n", _i32), ("normals_f64", _i32)]


class SSTables(C.Structure):
    _fields_ = [("blob", _vp), ("blob_bytes", _i64), ("blob_hot_bytes", _i64), ("cdf_pairs", _vp), ("cdf_perm", _vp),
                ("grid", _vp), ("grid_rows", _vp), ("grid_lut", _vp), ("grid_thr", _vp),
                ("iso_rows", _vp), ("iso_lut", _vp), ("iso_thr", _vp),
                ("grid_wlut", _vp), ("grid_wthr", _vp), ("iso_wlut", _vp), ("iso_wthr", _vp)]


class SSMaps(C.Structure):
    _fields_ = [("ebv", _vp), ("maglim", _vp * 2), ("mask", _vp), ("nside_mask", _i32),
                ("nside_packed", _i32), ("packed", _vp), ("packed32", _vp)]


class SSCatalog(C.Structure):
    _fields_ = [("phi1", _vp), ("phi2", _vp), ("dist", _vp), ("ra", _vp), ("dec", _vp), ("mag", _vp * 2)]


class SSDraws(C.Structure):
    _fields_ = [("u_phi1", _vp), ("n_phi2", _vp), ("n_dist", _vp), ("u_mass", _vp),
                ("z_flux", _vp * 2), ("u_det", _vp), ("u_det2", _vp)]


class SSOut(C.Structure):
    _fields_ = [("dtype", _i32), ("pad_", _i32), ("phi1", _vp), ("phi2", _vp), ("dist", _vp),
                ("mag", _vp * 2), ("ra", _vp), ("dec", _vp), ("mag_obs", _vp * 2),
                ("magerr", _vp * 2), ("flag_observed", _vp), ("flag_perfect", _vp),
                ("pix", _vp), ("counters", _vp), ("draws_out", _vp)]


_STRUCTS = (SSFunc, SSParams, SSTables, SSMaps, SSCatalog, SSDraws, SSOut)

# every symbol include/streamsim_b200.h declares: name -> (restype, argtypes)
_P = C.POINTER
SYMBOLS = {
    "ss_abi_version": (C.c_int, []),
    "ss_last_error": (C.c_char_p, []),
    "ss_struct_sizes": (C.c_int, [_P(_i32)]),
    "ss_launch_info": (C.c_int, [_i64, _u64, _P(_i32), _P(_i32), _P(_i32)]),
    "ss_run": (C.c_int, [C.c_uint32, _P(SSParams), _P(SSTables), _P(SSMaps), _P(SSCatalog),
                         _P(SSDraws), _P(SSOut), _u64, _u64, _u64, _vp]),
    "ss_generate": (C.c_int, [_P(SSParams), _P(SSTables), _P(SSDraws), _P(SSOut), _u64, _u64,
                              _u64, _vp]),
    "ss_rotate": (C.c_int, [_P(SSParams), _P(SSCatalog), _P(SSOut), _u64, _vp]),
    "ss_observe": (C.c_int, [_P(SSParams), _P(SSTables), _P(SSMaps), _P(SSCatalog), _P(SSDraws),
                             _P(SSOut), _u64, _u64, _u64, _vp]),
    "ss_generate_observe": (C.c_int, [_P(SSParams), _P(SSTables), _P(SSMaps), _P(SSDraws),
                                      _P(SSOut), _u64, _u64, _u64, _vp]),
    "ss_frame_fraction": (C.c_int, [_P(SSParams), _P(SSMaps), _P(SSCatalog), _vp, _u64, _vp]),
    "ss_photo_error": (C.c_int, [_P(SSParams), _P(SSTables), _i32, _vp, _vp, _vp, _u64, _vp]),
    "ss_efficiency": (C.c_int, [_P(SSParams), _P(SSTables), _i32, _i32, _vp, _vp, _vp, _u64, _vp]),
    "ss_eval_function": (C.c_int, [_P(SSFunc), _P(SSTables), _vp, _vp, _u64, _vp]),
    "ss_ang2pix": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _u64, _vp]),
    "ss_math_probe": (C.c_int, [_i32, _vp, _vp, _vp, _u64, _vp]),
    "ss_loc_scale": (C.c_int, [_i32, _vp, _vp, _f64, _f64, _vp, _vp, _u64, _u64, _u64, _vp]),
    "ss_noisy_magnitudes": (C.c_int, [_vp, _vp, _vp, _vp, _u64, _u64, _u64, _vp]),
    "ss_bernoulli": (C.c_int, [_vp, _vp, _vp, _u64, _u64, _u64, _vp]),
    "ss_host_workspace_bytes": (_i64, [_u64, _i32]),
    "ss_generate_observe_host": (C.c_int, [_P(SSParams), _P(SSTables), _P(SSMaps), _P(SSOut),
                                           _u64, _u64, _u64, _vp, _i64, _u64]),
    "ss_observe_host": (C.c_int, [_P(SSParams), _P(SSTables), _P(SSMaps), _P(SSCatalog),
                                  _P(SSOut), _u64, _u64, _u64, _vp, _i64, _u64]),
}

LIB_PATH = os.environ.get("STREAMSIM_B200_LIB") or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), "libstreamsim_b200.so")
_lib = None


class StreamSimLibraryError(RuntimeError):
    """The CUDA library is missing or reported a failure."""


def load():
    """Load the shared library once; verify ABI version and struct layouts."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StreamSimLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C stream_sim_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype, fn.argtypes = res, args
    if lib.ss_abi_version() != ABI_VERSION:
        raise StreamSimLibraryError(
            f"ABI mismatch: library {lib.ss_abi_version()}, binding {ABI_VERSION}; rebuild")
    sizes = (_i32 * 7)()
    lib.ss_struct_sizes(sizes)
    for s, n in zip(_STRUCTS, sizes):
        if C.sizeof(s) != n:
            raise StreamSimLibraryError(f"struct {s.__name__}: binding {C.sizeof(s)} B, library {n} B")
    _lib = lib
    return lib


def check(rc):
    """Translate a C return code into the exception the reference would raise."""
    if rc == 0:
        return
    msg = load().ss_last_error().decode()
    if rc == -1:
        raise ValueError(msg)
    raise StreamSimLibraryError(f"libstreamsim_b200 error {rc}: {msg}")
