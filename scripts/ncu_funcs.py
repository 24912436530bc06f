#!/usr/bin/env python
"""Aggregate an ncu report's executed instructions per enclosing C++ function (inlined code
is attributed to the function whose source lines it came from):
python scripts/ncu_funcs.py rep.ncu-rep stars"""
import collections, csv, re, subprocess, sys, os
rep = sys.argv[1]; stars = float(sys.argv[2]) if len(sys.argv) > 2 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
funcs = {}
def load(path):
    if path in funcs: return funcs[path]
    starts = []
    try:
        for i, l in enumerate(open(path), 1):
            m = re.match(r"^(?:template.*\n)?(?:__device__|__global__|static|inline|int |void ).*?(\w+)\s*\(", l)
            if m and not l.startswith(" "): starts.append((i, m.group(1)))
            m2 = re.match(r"^k_fused|^k_stars", l)
    except OSError:
        pass
    funcs[path] = starts
    return starts
def func_of(path, line):
    name = "?"
    for s, n in load(path):
        if s <= line: name = n
        else: break
    return name
cur = None; agg = collections.Counter(); samp = collections.Counter(); total = 0; stot = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1]; continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name") or r[0] == "": continue
    try: inst, sa = int(r[7]), int(r[6])
    except ValueError: continue
    key = os.path.basename(cur) + ":" + func_of(cur, int(r[0]))
    agg[key] += inst; samp[key] += sa; total += inst; stot += sa
print("total warp instructions", total, "samples", stot)
for k, v in agg.most_common(40):
    per = f"{32*v/stars:7.1f} instr/star" if stars else ""
    print(f"{100*v/total:5.1f}% inst {100*samp[k]/max(stot,1):5.1f}% samp {per}  {k}")
