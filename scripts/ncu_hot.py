#!/usr/bin/env python
"""Hot-loop view of an ncu report: the executed SASS stream of the star loop with its stall
samples and reasons, the dynamic opcode mix per star, and the top stall sites.

  python scripts/ncu_hot.py rep.ncu-rep <stars per launch> [top]
"""
import collections, csv, re, subprocess, sys
rep, stars = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]; ix = {n: i for i, n in enumerate(hdr)}
names = ['stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_math', 'stall_dispatch', 'stall_mio',
         'stall_lg', 'stall_not_selected', 'stall_branch_resolving', 'stall_barrier', 'stall_no_inst']
warps = stars / 32
hot, ops, tot = [], collections.Counter(), 0
for r in rows[2:]:
    try:
        n = int(r[ix["Instructions Executed"]]); sa = int(r[ix["# Samples"]])
    except (ValueError, IndexError):
        continue
    t = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip())
    op = t.split()[0]
    base = op.split(".")[0]
    ops[".".join(op.split(".")[:2]) if base in ("MUFU", "F2F", "F2I", "I2F", "IMAD") else base] += n
    tot += n
    if n >= 0.5 * warps:
        hot.append((sa, len(hot), n / warps, r))
print("thread instructions per star: %.1f" % (tot * 32 / stars))
print("opcode mix (per star):", ", ".join("%s %.0f" % (k, v * 32 / stars) for k, v in ops.most_common(24)))
allsa = sum(h[0] for h in hot)
agg = collections.Counter()
for sa, k, _, r in hot:
    for n in names:
        agg[n] += int(r[ix[n]])
print("stall samples in the loop: %d;" % allsa, ", ".join("%s %.1f%%" % (n.replace("stall_", ""), 100 * v / allsa) for n, v in agg.most_common()))
for sa, k, w, r in sorted(hot, key=lambda x: -x[0])[:top]:
    d = {n: int(r[ix[n]]) for n in names}
    why = " ".join("%s=%d" % (a.replace("stall_", ""), b) for a, b in sorted(d.items(), key=lambda kv: -kv[1])[:2] if b)
    print("%5d %5.1f%% %-64s %s" % (k, 100 * sa / allsa, r[ix["Source"]].strip()[:64], why))
