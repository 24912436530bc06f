"""Drop-in alias: ``stream_sim`` -> ``stream_sim_b200``.

Reference user code does ``from stream_sim.model import StreamModel``,
``from stream_sim import surveys, observed``, ``from stream_sim.utils import
parse_config`` ... (reference bin/generate_stream.py:15-19, notebooks).  This
package makes those imports resolve to the B200 implementation.
"""
import importlib
import sys

import stream_sim_b200 as _impl

for _name in ("functions", "samplers", "model", "surveys", "observed", "utils", "healpix",
              "frames", "engine", "isochrone", "synthetic"):
    _mod = importlib.import_module("stream_sim_b200." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    globals()[_name] = _mod

__version__ = _impl.__version__
