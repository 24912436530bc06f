#!/usr/bin/env python
"""Aggregate an ncu report per CUDA source line: python scripts/ncu_lines.py rep.ncu-rep [top]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None; agg = collections.OrderedDict(); total = 0; stot = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name") or r[0] == "":
        continue
    try:
        inst, samp = int(r[7]), int(r[6])
    except ValueError:
        continue
    a = agg.setdefault((cur, int(r[0]), r[1].strip()[:100]), [0, 0]); a[0] += inst; a[1] += samp
    total += inst; stot += samp
print("total warp instructions", total, "samples", stot)
for (f, l, s), (i, sa) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*i/total:5.1f}% inst {100*sa/max(stot,1):5.1f}% samp  {f}:{l}  {s}")
