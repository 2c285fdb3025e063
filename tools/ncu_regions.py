#!/usr/bin/env python
"""Per-region instruction / stall breakdown of traceKernel from an .ncu-rep's source page (SASS view).

    python tools/ncu_regions.py gpurun_out/prof.ncu-rep [kernel-index]

Regions are delimited by SASS landmarks of the kernel (the LDG of the node, the three LDG.128 of a leaf record,
the ray LDG.EF, ...) so that it survives recompiles; prints warp instructions, thread instructions, samples."""
import csv, io, subprocess, sys

def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    # split per kernel
    blocks, cur = [], None
    for line in out.splitlines():
        if line.startswith('"Kernel Name"'):
            cur = [line]; blocks.append(cur)
        elif cur is not None:
            cur.append(line)
    blk = blocks[which]
    print("#", blk[0][:160])
    rows = list(csv.reader(io.StringIO("\n".join(blk[1:]))))
    hdr = rows[0]
    ci = {k: hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed", "Predicated-On Thread Instructions Executed", "stall_long_sb", "stall_wait", "stall_math", "stall_not_selected", "stall_selected", "stall_branch_resolving", "stall_short_sb", "stall_mio")}
    data = []
    for r in rows[1:]:
        if len(r) < len(hdr): continue
        def f(k):
            try: return float(r[ci[k]])
            except ValueError: return 0.0
        try: addr = int(r[ci["Address"]], 16) if not r[ci["Address"]].isdigit() else int(r[ci["Address"]])
        except ValueError: addr = 0
        data.append(dict(addr=addr, src=r[ci["Source"]], samples=f("# Samples"), inst=f("Instructions Executed"), tinst=f("Thread Instructions Executed"), pinst=f("Predicated-On Thread Instructions Executed"),
                         long_sb=f("stall_long_sb"), wait=f("stall_wait"), math=f("stall_math"), notsel=f("stall_not_selected"), sel=f("stall_selected"), br=f("stall_branch_resolving"), ssb=f("stall_short_sb"), mio=f("stall_mio")))
    tot_i = sum(d["inst"] for d in data); tot_s = sum(d["samples"] for d in data)
    if len(sys.argv) > 3 and sys.argv[3] == "dump":
        for i, d in enumerate(data):
            print(f"{i:5d} {d['inst']/1e6:9.2f}M {d['tinst']/max(d['inst'],1):5.1f} {d['pinst']/max(d['inst'],1):5.1f} s={d['samples']:7.0f} lsb={d['long_sb']:6.0f} {d['src'][:90]}")
        return
    # landmarks
    def find(pred, start=0):
        for i in range(start, len(data)):
            if pred(data[i]["src"]): return i
        return len(data)
    import re
    i_node = find(lambda s: "LDG.E.64" in s)
    i_tri = find(lambda s: "LDG.E.128.CONSTANT" in s)
    i_exit = find(lambda s: s.strip().startswith("EXIT"))
    nodes = [i for i, d in enumerate(data) if "LDG.E.64" in d["src"]]
    addr_of = [d["addr"] for d in data]
    # the descent loop ends at the backward branch that follows the last (unrolled) node load and targets the first one
    i_loop_end = nodes[-1]
    for i in range(nodes[-1], len(data)):
        m = re.search(r"BRA\s+(?:P\d,\s*)?0x([0-9a-f]+)", data[i]["src"])
        if m and int(m.group(1), 16) <= addr_of[i_node] and int(m.group(1), 16) >= addr_of[max(i_node - 14, 0)]:
            i_loop_end = i + 1
            break
    if i_tri < i_node:
        # r1h layout: refill, votes, leaf phase, descent loop, stop-reason / restart / result write
        regions = [("refill+setup+votes", 0, i_tri - 6), ("leaf test + pop", i_tri - 6, i_node - 3), ("descent loop", i_node - 3, i_loop_end),
                   ("stop reason + restart + result write", i_loop_end, i_exit + 1), ("cold paths (div slow path, BRA.DIV)", i_exit + 1, len(data))]
    else:
        regions = [("refill+setup", 0, i_node - 14), ("descent loop", i_node - 14, i_tri - 6), ("leaf test + pop + write", i_tri - 6, i_exit + 1), ("cold paths (div slow path, BRA.DIV)", i_exit + 1, len(data))]
    print(f"{'region':40s} {'warp inst':>12s} {'share':>7s} {'thr/inst':>9s} {'pred-on/inst':>12s} {'samples':>9s} {'share':>7s} {'long_sb':>8s} {'wait':>7s} {'math':>7s} {'notsel':>7s} {'branch':>7s}")
    for name, a, b in regions:
        seg = data[max(a, 0):b]
        i = sum(d["inst"] for d in seg); t = sum(d["tinst"] for d in seg); p = sum(d["pinst"] for d in seg); s = sum(d["samples"] for d in seg)
        print(f"{name:40s} {i:12.0f} {100*i/tot_i:6.1f}% {t/max(i,1):9.2f} {p/max(i,1):12.2f} {s:9.0f} {100*s/max(tot_s,1):6.1f}% {sum(d['long_sb'] for d in seg):8.0f} {sum(d['wait'] for d in seg):7.0f} {sum(d['math'] for d in seg):7.0f} {sum(d['notsel'] for d in seg):7.0f} {sum(d['br'] for d in seg):7.0f}")
    print(f"{'total':40s} {tot_i:12.0f}")

main()
