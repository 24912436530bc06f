"""stream_sim_b200 -- B200-native implementation of the per-star
generate-and-observe path of LSSTDESC/stream_sim, behind the reference's own
Python API (modules ``functions``, ``samplers``, ``model``, ``surveys``,
``observed``, ``utils``).  ``import stream_sim`` resolves to this package through
the alias package at the repository root, so reference scripts run unchanged.

The per-star work runs in hand-written sm_100a CUDA (``csrc/``) behind the C ABI
of ``include/streamsim_b200.h``; there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (constants + loader; the .so itself is loaded lazily)
