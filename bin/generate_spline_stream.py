#!/usr/bin/env python
"""Generate a spline-defined stream (config/atlas_spline_config.yaml schema) plus
background on the GPU; CLI of the reference's bin/generate_spline_stream.py:43-68
(which cannot run upstream: SplineStreamModel raises, model.py:670)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from generate_stream import main                      # noqa: E402
from stream_sim.model import SplineStreamModel        # noqa: E402

if __name__ == "__main__":
    main(SplineStreamModel)
