"""Wall-clock of the whole HBT analysis from results/particle_samples.gz to the .dat files:
the reference binary, the drop-in binary (reference reader + our class) and hbt_fast_analysis.e
(our reader + driver) on the same input.  Usage: python tests/tools/e2e_files.py [groups] [events/group]"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hadronic_afterburner_toolkit_b200 import synth  # noqa: E402
from hadronic_afterburner_toolkit_b200.params import C3  # noqa: E402

ngrp = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nev = int(sys.argv[2]) if len(sys.argv) > 2 else 10
skip_ref = len(sys.argv) > 3 and sys.argv[3] == "noref"
mult = 1500
HOST = os.path.join(ROOT, "hadronic_afterburner_toolkit_b200", "host", "build")
EXES = {"reference": os.path.join(ROOT, "oracle", "_ref", "hadronic_afterburner_tools.e"),
        "drop-in": os.path.join(HOST, "hadronic_afterburner_tools_b200.e"),
        "fast": os.path.join(HOST, "hbt_fast_analysis.e")}
if skip_ref:
    del EXES["reference"]
tmp = tempfile.mkdtemp(prefix="hbt_e2e_")
gz = os.path.join(tmp, "input.gz")
t0 = time.time()
synth.write_iss_gz(gz, synth.make_batches(20260003, ngrp, nev))
n = nev * mult
pairs = ngrp * (n * (n - 1) // 2 + nev * (nev // 2 + 1) * mult * mult)
print(f"input: {ngrp} groups x {nev} events x {mult} pi+, {os.path.getsize(gz) / 1e6:.1f} MB gz, {pairs:.3e} pairs "
      f"(written in {time.time() - t0:.1f} s)", flush=True)
text = C3.parameters_dat(event_buffer_size=n)
out = {"groups": ngrp, "events_per_group": nev, "pairs": pairs}
files = {}
for name, exe in EXES.items():
    wd = os.path.join(tmp, name)
    os.makedirs(os.path.join(wd, "EOS"))
    os.makedirs(os.path.join(wd, "results"))
    shutil.copy(os.path.join(ROOT, "oracle", "_ref", "EOS", "pdg.dat"), os.path.join(wd, "EOS", "pdg.dat"))
    shutil.copy(gz, os.path.join(wd, "results", "particle_samples.gz"))
    open(os.path.join(wd, "parameters.dat"), "w").write(text)
    t0 = time.time()
    r = subprocess.run([exe], cwd=wd, capture_output=True, text=True, env=dict(os.environ, HBT_B200_DEVICES="1"))
    dt = time.time() - t0
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out[name] = {"wall_s": dt, "pairs_per_s": pairs / dt}
    res = os.path.join(wd, "results")
    files[name] = {f: open(os.path.join(res, f)).read() for f in sorted(os.listdir(res)) if f.startswith("HBT_")}
    tail = [l for l in r.stdout.splitlines() if "hbt_fast_analysis" in l]
    print(f"{name}: {dt:.2f} s  ({pairs / dt:.3e} pairs/s)" + (("  " + tail[-1]) if tail else ""), flush=True)


def worst_deviation(a, b):
    """largest relative difference between two sets of output files (same layout required)"""
    assert a.keys() == b.keys()
    worst = 0.0
    for f in a:
        la, lb = a[f].splitlines(), b[f].splitlines()
        assert len(la) == len(lb), f
        for x, y in zip(la, lb):
            if x == y:
                continue
            for u, v in zip(x.split(), y.split()):
                if u != v:
                    worst = max(worst, abs(float(u) - float(v)) / max(abs(float(u)), 1e-300))
    return worst


names = list(files)
for k in names[1:]:
    dev = worst_deviation(files[names[0]], files[k])
    print(f"{k} vs {names[0]}: {len(files[k])} files, largest relative difference of a printed number {dev:.2e}")
    out[k]["max_rel_diff_vs_" + names[0]] = dev
print(json.dumps(out))
shutil.rmtree(tmp)
