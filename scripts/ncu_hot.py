#!/usr/bin/env python
"""Top stall sites of a kernel from `ncu --page source --csv` (SASS view).
usage: ncu_hot.py report.ncu-rep kernel-substring [N]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
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
    tot = sum(int(x["# Samples"]) for x in lst)
    print(k[:80], "samples", tot, "instr", len(lst))
    stall_keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {s: sum(int(x[s]) for x in lst) for s in stall_keys}
    print("  totals:", {s: v for s, v in sorted(agg.items(), key=lambda t: -t[1]) if v > tot * 0.01})
    idx = sorted(range(len(lst)), key=lambda i: -int(lst[i]["# Samples"]))[:N]
    for i in sorted(idx):
        x = lst[i]
        top = sorted(((int(x[s]), s) for s in stall_keys), reverse=True)[:2]
        print(f"  {i:5d} {int(x['# Samples']):7d} {100*int(x['# Samples'])/tot:5.1f}%  exec={x['Instructions Executed']:>10}  {x['Source'].strip()[:70]:70s} {top}")
    break
