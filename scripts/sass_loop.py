#!/usr/bin/env python
"""Static look at a kernel's SASS: instruction count and opcode mix of its largest loop
(the persistent per-star loop), from `cuobjdump -sass`.  No GPU needed.

  python scripts/sass_loop.py stream_sim_b200/libstreamsim_b200.so k_fusedIdLi0ELb1ELb1
"""
import collections
import re
import subprocess
import sys


def functions(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, out = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            out[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return out


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    for name, ins in functions(lib).items():
        if pat not in name:
            continue
        # largest backward branch = the star loop
        best = None
        for addr, text in ins:
            m = re.search(r"\bBRA\S*\s+(?:!?U?P\d,? )?`\(\.L_x_\d+\)|BRA.*0x([0-9a-f]+)", text)
            m2 = re.search(r"0x([0-9a-f]+)", text) if text.split()[0].lstrip("@!UP0123456789 ").startswith("BRA") or " BRA" in text else None
            if m2:
                tgt = int(m2.group(1), 16)
                if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                    best = (tgt, addr)
        print(name, "total instructions:", len(ins))
        if best is None:
            continue
        loop = [t for a, t in ins if best[0] <= a <= best[1]]
        ops = collections.Counter()
        for t in loop:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            ops[t.split()[0].split(".")[0]] += 1
        print("  loop 0x%x..0x%x: %d instructions" % (best[0], best[1], len(loop)))
        grp = collections.Counter()
        for k, v in ops.items():
            g = ("fp64" if k in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX") else
                 "mufu" if k == "MUFU" else
                 "fp32" if k in ("FFMA", "FADD", "FMUL", "FSETP", "FSEL", "FMNMX") else
                 "conv" if k in ("F2F", "I2F", "F2I", "I2FP", "F2FP") else
                 "ldg" if k in ("LDG",) else "lds" if k == "LDS" else
                 "st" if k in ("STG", "STS", "STL") else "ldl" if k == "LDL" else
                 "atom" if k in ("ATOMS", "ATOMG", "RED", "REDUX", "VOTE", "VOTEU") else
                 "ldc" if k in ("LDC", "LDCU", "ULDC") else
                 "imad" if k in ("IMAD", "IADD3", "IADD", "LEA", "SHF", "LOP3", "ISETP", "SEL", "MOV", "PRMT", "IABS", "IMNMX", "VIADD", "VIMNMX", "IADD64", "PLOP3") else
                 "uniform" if k.startswith("U") else
                 "branch" if k in ("BRA", "BSSY", "BSYNC", "CALL", "RET", "EXIT", "WARPSYNC", "BAR", "NOP") else "other")
            grp[g] += v
        print("  groups:", dict(grp.most_common()))
        print("  top ops:", ops.most_common(25))


main()
