#!/usr/bin/env python
"""Instruction mix of the descent loop of traceKernel<0> in a built libb200rt.so (no GPU needed):
    python tools/sass_loop.py [lib.so] [kernel-template-arg]
Finds the innermost backward branch that encloses the node load (LDG.E.64) and classifies the instructions of that range by
issue pipe (ALU = the half-rate integer/logic/compare/select pipe that bounds this kernel, profiles/r1f_*)."""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "libyafaray_b200/libb200rt.so"
q = sys.argv[2] if len(sys.argv) > 2 else "0"
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
blocks = out.split("Function : ")
body = next(b for b in blocks if b.startswith(f"_ZN6b200rt11traceKernelILi{q}ELb0EEE"))
ins = []
for line in body.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
node = next(i for i, (_, t) in enumerate(ins) if "LDG.E.64" in t)
best = None
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA\s+(?:`\(\.L_x_\d+\)|0x([0-9a-f]+))", t)
    if m and m.group(1):
        tgt = int(m.group(1), 16)
        if tgt in addr and addr[tgt] <= node <= i and (best is None or i - addr[tgt] < best[1] - best[0]):
            best = (addr[tgt], i)
lo, hi = best
FMA = ("FFMA", "FADD", "FMUL", "IMAD", "HFMA2", "FSWZADD")
LSU = ("LDG", "LDS", "STS", "STG", "LDC", "LDL", "STL", "ATOM", "RED")
CTL = ("BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "NOP", "BREAK", "CALL", "RET")
XU = ("MUFU", "POPC", "FLO", "I2F", "F2I", "I2I", "F2F")
mix = {"ALU": 0, "FMA": 0, "LSU": 0, "CTL": 0, "XU": 0, "OTHER": 0}
for a, t in ins[lo:hi + 1]:
    op = re.sub(r"^@!?U?P\d\s+", "", t).split()[0].split(".")[0]
    k = "FMA" if op in FMA else "LSU" if op in LSU else "CTL" if op in CTL else "XU" if op in XU else "ALU"
    mix[k] += 1
    if len(sys.argv) > 3:
        print(f"{a:05x} {k:4s} {t}")
print(f"descent loop {ins[lo][0]:#x}..{ins[hi][0]:#x}: {hi - lo + 1} instructions  " + "  ".join(f"{k} {v}" for k, v in mix.items() if v))
