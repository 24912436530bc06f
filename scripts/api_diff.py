#!/usr/bin/env python
"""Compare the public API surface of the reference (imported unmodified from /root/reference,
third-party imports stubbed) with this package's drop-in modules: every public class, function,
method and their parameter names / defaults.  Build-container tool (the reference does not
travel to the GPU box).  Exit code 0 = nothing missing."""
import importlib
import inspect
import os
import subprocess
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
MODULES = ("functions", "samplers", "model", "utils", "surveys", "observed")


def surface(prefix):
    out = {}
    for m in MODULES:
        mod = importlib.import_module(f"{prefix}.{m}")
        for name, obj in vars(mod).items():
            if name.startswith("_") or getattr(obj, "__module__", None) != mod.__name__:
                continue
            if inspect.isclass(obj):
                out[f"{m}.{name}"] = "class"
                for mn, meth in vars(obj).items():
                    if mn.startswith("__") and mn != "__init__":
                        continue
                    fn = meth.__func__ if isinstance(meth, (staticmethod, classmethod)) else meth
                    if callable(fn):
                        out[f"{m}.{name}.{mn}"] = sig(fn)
            elif callable(obj):
                out[f"{m}.{name}"] = sig(obj)
    return out


def sig(fn):
    try:
        ps = inspect.signature(fn).parameters.values()
    except (TypeError, ValueError):
        return "?"
    return ", ".join(p.name if p.default is p.empty else f"{p.name}={p.default!r}" for p in ps
                     if p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL)) + \
        (", **" if any(p.kind == p.VAR_KEYWORD for p in ps) else "")


def child(which):
    if which == "ref":
        sys.path = [p for p in sys.path if os.path.abspath(p or ".") != REPO]
        sys.path.insert(0, REF)
        hp = types.ModuleType("healpy")
        hp.UNSEEN = -1.6375e30
        sys.modules["healpy"] = hp
        for name in ("healsparse", "astropy", "astropy.coordinates", "astropy.units", "gala",
                     "gala.coordinates", "pylab", "skyproj", "matplotlib", "matplotlib.pyplot"):
            sys.modules[name] = types.ModuleType(name)
        sys.modules["astropy"].coordinates = sys.modules["astropy.coordinates"]
        sys.modules["astropy"].units = sys.modules["astropy.units"]
        sys.modules["gala"].coordinates = sys.modules["gala.coordinates"]
        os.chdir(REF)
        s = surface("stream_sim")
    else:
        sys.path.insert(0, REPO)
        os.chdir(REPO)
        s = surface("stream_sim_b200")
    import json
    print(json.dumps(s))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
        sys.exit(0)
    import json
    got = {}
    for which in ("ref", "ours"):
        r = subprocess.run([sys.executable, __file__, which], capture_output=True, text=True)
        if r.returncode:
            print(r.stderr[-2000:])
            sys.exit(2)
        got[which] = json.loads(r.stdout.strip().splitlines()[-1])
    missing = [k for k in got["ref"] if k not in got["ours"]]
    differ = []
    for k, v in got["ref"].items():
        if k in got["ours"] and v not in ("class", "?") and got["ours"][k] != v:
            ref_names = [a.split("=")[0] for a in v.split(", ") if a]
            our_names = [a.split("=")[0] for a in got["ours"][k].split(", ") if a]
            # additive keywords at the end are allowed; the reference's parameters must be a prefix
            if our_names[:len([n for n in ref_names if n != "**"])] != [n for n in ref_names if n != "**"]:
                differ.append((k, v, got["ours"][k]))
    print(f"reference surface: {len(got['ref'])} names; ours: {len(got['ours'])}")
    print("MISSING:", *missing, sep="\n  ")
    print("SIGNATURE ORDER DIFFERS:")
    for k, a, b in differ:
        print(f"  {k}\n     ref : {a}\n     ours: {b}")
    sys.exit(1 if missing else 0)
