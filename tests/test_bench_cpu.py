"""bench.py's CPU arm (the reference leg of the driver's contract): one JSON line on stdout with
the contract's keys, nothing from ranks other than 0.  Runs the oracle port on the host cores
for one short step -- no GPU."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(extra_env=None, *args):
    env = dict(os.environ, **(extra_env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                           "--steps", "1", "--warmup", "1", *args],
                          capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                      # ONE line on the real stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "observed stars/sec" and d["unit"] == "stars/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 1e5 and d["ms_per_step"] > 0
    assert d["config"]["workload"] == "lsst_injection_nside4096_pal5"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "stars/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    ref = d["cpu_baseline_reference"]
    assert ref["kind"] == "reference" and ref["cores"] == 1 and ref["value"] > 0


def test_reference_arm_other_ranks_stay_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""
