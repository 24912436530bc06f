import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _repo_cwd():
    """Stream configs use ./data/... relative paths, like the reference's."""
    old = os.getcwd()
    os.chdir(ROOT)
    yield
    os.chdir(old)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_generate_case(name):
    g = load_golden(f"generate_{name}.npz")
    cfg = json.loads(str(g["config"]))
    draws = {k[5:]: v for k, v in g.items() if k.startswith("draw_")}
    out = {k[4:]: v for k, v in g.items() if k.startswith("out_")}
    return cfg, draws, out


def golden_inject_case(tag, dust):
    g = load_golden(f"inject_{tag}_dust{int(dust)}.npz")
    cat = {k[4:]: v for k, v in g.items() if k.startswith("cat_")}
    draws = {k[5:]: v for k, v in g.items() if k.startswith("draw_")}
    out = {k[4:]: v for k, v in g.items() if k.startswith("out_")}
    meta = dict(seed=int(g["seed"]), compl_band=str(g["compl_band"]), first=float(g["first"]),
                nside_maglim=int(g["nside_maglim"]), nside_ebv=int(g["nside_ebv"]),
                nside=int(g["nside"]))
    return cat, draws, out, meta


@pytest.fixture(scope="session")
def known_answers():
    with open(os.path.join(GOLDEN, "known_answers.json")) as f:
        return json.load(f)


GENERATE_CASES = ("toy1", "toy2", "pal5", "toy1_background", "atlas_spline", "phoenix_spline")
INJECT_CASES = [(t, d) for t in "ABC" for d in (True, False)]
