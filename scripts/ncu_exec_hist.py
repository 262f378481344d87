#!/usr/bin/env python
"""Where the executed warp-instructions of a kernel go: groups SASS instructions by their
execution count (loop nest level) and prints instruction-class mix per group.
usage: ncu_exec_hist.py report.ncu-rep kernel-substring"""
import csv, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, data = None, None, {}
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = r[1]; data[cur] = []; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if cur and hdr and len(r) == len(hdr):
        data[cur].append(dict(zip(hdr, r)))
for k, lst in data.items():
    if pat not in k: continue
    tot = sum(int(x["Instructions Executed"]) for x in lst)
    print(k[:70], "total warp-instr", tot)
    groups = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    for x in lst:
        e = int(x["Instructions Executed"])
        if e == 0: continue
        import math
        key = round(math.log10(e) * 4) / 4
        g = groups[key]
        g[0] += 1; g[1] += e
        op = x["Source"].strip().split()
        op = [o for o in op if not o.startswith("@")][0].split(".")[0]
        g[2][op] += e
    for key in sorted(groups, reverse=True):
        n, e, c = groups[key]
        if e < tot * 0.005: continue
        print(f"  exec~1e{key:.2f}: {n:4d} instrs, {100*e/tot:5.1f}% of executed; top: " +
              ", ".join(f"{o}:{100*v/e:.0f}%" for o, v in c.most_common(12)))
    break
