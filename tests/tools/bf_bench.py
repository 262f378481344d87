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
HOST = os.path.join(ROOT, "hadronic_afterburner_toolkit_b200", "host", "build")
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "hadronic_afterburner_tools.e")
tmp = tempfile.mkdtemp(prefix="bf_bench_")
nev = 40


def write_input(path, n):
    with gzip.open(path, "wt", compresslevel=1) as f:
        for _ in range(nev):
            f.write(f"{2 * n}\n")
            for q in (211, -211):
                pT = rng.gamma(2.0, 0.3, n); phi = rng.uniform(-np.pi, np.pi, n); y = rng.normal(0, 1.3, n)
                mT = np.sqrt(0.13957 ** 2 + pT * pT)
                for k in range(n):
                    f.write("%d 0.13957 1 0 0 0 %.17g %.17g %.17g %.17g\n" % (q, mT[k] * np.cosh(y[k]), pT[k] * np.cos(phi[k]),
                                                                                pT[k] * np.sin(phi[k]), mT[k] * np.sinh(y[k])))


def run(name, exe, gz, n):
    text = HBTParams(randomSeed=5).parameters_dat(analyze_HBT=0, analyze_balance_function=1, event_buffer_size=10 * 2 * n,
                                                   particle_alpha=211, particle_beta=-211, Bnpts=21, Brap_max=2.0, BpT_min=0.2,
                                                   BpT_max=3.0, rap_type=1)
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
    return dt, {f: open(os.path.join(res, f)).read() for f in sorted(os.listdir(res)) if f.endswith(".dat")}


n = 1200
gz = os.path.join(tmp, "in.gz")
write_input(gz, n)
gz_half = os.path.join(tmp, "in_half.gz")
write_input(gz_half, n // 2)
files = {}
for name, exe, g, nn in (("reference", REF_EXE, gz, n), ("reference, half multiplicity", REF_EXE, gz_half, n // 2),
                         ("drop-in", os.path.join(HOST, "hadronic_afterburner_tools_b200.e"), gz, n)):
    dt, files[name] = run(name.replace(" ", "_").replace(",", ""), exe, g, nn)
    out[name] = {"wall_s": dt}
    print(f"{name}: {dt:.2f} s", flush=True)
out["identical_files"] = files["reference"] == files["drop-in"]
out["file_pairs"] = 8 * nev * n * n
# reading scales with the multiplicity n, the operator with n^2: T(n) = R n + P n^2 from the two reference runs
T1, T2 = out["reference"]["wall_s"], out["reference, half multiplicity"]["wall_s"]
P = (T1 / n - T2 / (n // 2)) / (n - n // 2)
cpu_s = P * n * n
out["cpu_operator"] = {"seconds": cpu_s, "pairs_per_s_one_core": out["file_pairs"] / cpu_s, "cores": 1,
                       "note": "unmodified reference binary, BalanceFunction operator alone: the n^2 term of T(n) = R n + P n^2 "
                               "fitted to runs at multiplicity n and n/2 (the rest is its text reader)"}
print(f"reference BalanceFunction operator alone: {cpu_s:.2f} s of {T1:.2f} s for {out['file_pairs']:.3e} pair visits "
      f"-> {out['file_pairs'] / cpu_s:.3e} pairs/s on one core; GPU kernel / CPU core = "
      f"{out['gpu']['kernel_pairs_per_s'] / (out['file_pairs'] / cpu_s):.0f}x")
print("identical files:", out["identical_files"], f"({out['file_pairs']:.3e} pairs of which ~80 % inside the pT cut)")
print(json.dumps(out))
shutil.rmtree(tmp)
