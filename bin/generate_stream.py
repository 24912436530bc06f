#!/usr/bin/env python
"""Generate a mock stream plus background on the GPU and write it to CSV.

Same command line as the reference's bin/generate_stream.py:51-76:

    generate_stream.py CONFIG [-o stream_out.csv] [-p] [--seed N]

An output name ending in .parquet / .pq selects the Parquet sink instead of CSV.
"""
import argparse
import os
import sys

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from stream_sim.model import BackgroundModel, StreamModel   # noqa: E402
from stream_sim.utils import parse_config                   # noqa: E402


def generate_stream(config, stream_cls=StreamModel, seed=None):
    """Stream + background stars with a ``flag`` column (1 = stream, 0 = background)."""
    frames, flags = [], []
    for section, cls, flag in (("stream", stream_cls, 1), ("background", BackgroundModel, 0)):
        if section not in config:
            continue
        print(f"Generating {section}...")
        df = cls(config[section]).sample(config[section]["nstars"],
                                         seed=None if seed is None else seed + flag)
        print(f"  generated {len(df)} {section} stars.")
        frames.append(df)
        flags.append(np.full(len(df), flag, dtype=int))
    out = pd.concat(frames)
    out["flag"] = np.hstack(flags)
    return out


def main(stream_cls=StreamModel):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("config", help="configuration file")
    ap.add_argument("-o", "--outfile", default="stream_out.csv", help="output file")
    ap.add_argument("-p", "--plot", action="store_true", help="plot stream and background")
    ap.add_argument("--seed", type=int, default=None, help="Philox seed (default: fresh entropy)")
    args = ap.parse_args()
    print(f"Reading config: {args.config}")
    stars = generate_stream(parse_config(args.config), stream_cls, args.seed)
    print(f"Writing {args.outfile}")
    if args.outfile.endswith((".parquet", ".pq")):      # columnar sink (not in the reference)
        stars.to_parquet(args.outfile, index=False)
    else:
        stars.to_csv(args.outfile, index=False)
    if args.plot:
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except ImportError:
            print("matplotlib is not installed; skipping the plot.")
            return
        fig, ax = plt.subplots(figsize=(10, 3))
        ax.hist2d(stars["phi1"], stars["phi2"], bins=(200, 60))
        ax.set_xlabel("phi1 (deg)")
        ax.set_ylabel("phi2 (deg)")
        png = os.path.splitext(args.outfile)[0] + ".png"
        print(f"Writing {png}")
        fig.savefig(png, facecolor="white")


if __name__ == "__main__":
    main()
