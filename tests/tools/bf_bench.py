"""BalanceFunction row: GPU pair kernel rate on a production-size batch (100 events x 1500 pi+ and
1500 pi-), and the unmodified reference binary vs the drop-in binary, file to file, on a smaller input."""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hadronic_afterburner_toolkit_b200.balance_function import BalanceFunction  # noqa: E402
from hadronic_afterburner_toolkit_b200.hbt_correlation import Random  # noqa: E402
from hadronic_afterburner_toolkit_b200.params import HBTParams  # noqa: E402

rng = np.random.default_rng(1)


def species(nev, n):
    out = []
    for _ in range(nev):
        pT = rng.gamma(2.0, 0.3, n)
        out.append({"pT": pT, "phi": rng.uniform(-np.pi, np.pi, n), "rap_y": rng.normal(0, 1.3, n), "rap_eta": rng.normal(0, 1.5, n)})
    return out


out = {}
plus, minus = species(100, 1500), species(100, 1500)
lists = {"a": plus, "abar": minus, "b": minus, "bbar": plus}
bf = BalanceFunction(211, -211, 21, 2.0, 0.2, 3.0, 1, ran_gen=Random(3))
bf.calculate_balance_function(lists)  # warm-up
t0 = bf.timers()
w0 = time.time()
for _ in range(5):
    bf.calculate_balance_function(lists)
wall = time.time() - w0
t1 = bf.timers()
pairs = t1["pairs"] - t0["pairs"]
ms = t1["kernel_ms"] - t0["kernel_ms"]
out["gpu"] = {"pairs": int(pairs), "kernel_ms": ms, "kernel_pairs_per_s": pairs / (ms * 1e-3), "wall_pairs_per_s": pairs / wall}
print(f"GPU: {pairs:.3e} pairs, kernels {ms:.2f} ms -> {pairs / ms / 1e6:.3e} Gpairs/s... {pairs / (ms * 1e-3):.3e} pairs/s; "
      f"with host gathers and draws {pairs / wall:.3e} pairs/s", flush=True)

# file to file: reference binary vs drop-in binary
tmp = tempfile.mkdtemp(prefix="bf_bench_")
gz = os.path.join(tmp, "in.gz")
nev, n = 40, 1200
with gzip.open(gz, "wt", compresslevel=1) as f:
    for _ in range(nev):
        f.write(f"{2 * n}\n")
        for q in (211, -211):
            pT = rng.gamma(2.0, 0.3, n); phi = rng.uniform(-np.pi, np.pi, n); y = rng.normal(0, 1.3, n)
            mT = np.sqrt(0.13957 ** 2 + pT * pT)
            for k in range(n):
                f.write("%d 0.13957 1 0 0 0 %.17g %.17g %.17g %.17g\n" % (q, mT[k] * np.cosh(y[k]), pT[k] * np.cos(phi[k]),
                                                                            pT[k] * np.sin(phi[k]), mT[k] * np.sinh(y[k])))
text = HBTParams(randomSeed=5).parameters_dat(analyze_HBT=0, analyze_balance_function=1, event_buffer_size=10 * 2 * n,
                                               particle_alpha=211, particle_beta=-211, Bnpts=21, Brap_max=2.0, BpT_min=0.2,
                                               BpT_max=3.0, rap_type=1)
HOST = os.path.join(ROOT, "hadronic_afterburner_toolkit_b200", "host", "build")
files = {}
for name, exe in (("reference", os.path.join(ROOT, "oracle", "_ref", "hadronic_afterburner_tools.e")),
                  ("drop-in", os.path.join(HOST, "hadronic_afterburner_tools_b200.e"))):
    wd = os.path.join(tmp, name)
    os.makedirs(os.path.join(wd, "EOS")); os.makedirs(os.path.join(wd, "results"))
    shutil.copy(os.path.join(ROOT, "oracle", "_ref", "EOS", "pdg.dat"), os.path.join(wd, "EOS", "pdg.dat"))
    shutil.copy(gz, os.path.join(wd, "results", "particle_samples.gz"))
    open(os.path.join(wd, "parameters.dat"), "w").write(text)
    t0 = time.time()
    r = subprocess.run([exe], cwd=wd, capture_output=True, text=True)
    dt = time.time() - t0
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = os.path.join(wd, "results")
    files[name] = {f: open(os.path.join(res, f)).read() for f in sorted(os.listdir(res)) if f.endswith(".dat")}
    out[name] = {"wall_s": dt}
    print(f"{name}: {dt:.2f} s", flush=True)
out["identical_files"] = files["reference"] == files["drop-in"]
out["file_pairs"] = 8 * nev * n * n
print("identical files:", out["identical_files"], f"({out['file_pairs']:.3e} pairs of which ~80 % inside the pT cut)")
print(json.dumps(out))
shutil.rmtree(tmp)
