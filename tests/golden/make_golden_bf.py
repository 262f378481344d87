"""Golden vectors for the BalanceFunction row: runs the UNMODIFIED reference
binary (oracle/_ref/hadronic_afterburner_tools.e, built from /root/reference by oracle/Makefile) with
analyze_balance_function = 1 on a small synthetic particle_samples.gz and commits the input and
the reference's three output files per case.  Run here (needs /root/reference for the build):

    python tests/golden/make_golden_bf.py
"""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from hadronic_afterburner_toolkit_b200.params import HBTParams  # noqa: E402

REF_EXE = os.path.join(ROOT, "oracle", "_ref", "hadronic_afterburner_tools.e")
PDG = os.path.join(ROOT, "oracle", "_ref", "EOS", "pdg.dat")
SPECIES = [(211, 0.13957), (-211, 0.13957), (321, 0.49368), (-321, 0.49368), (2212, 0.93827), (-2212, 0.93827)]

CASES = {
    # name: (alpha, beta, Bnpts, Brap_max, BpT_min, BpT_max, rap_type, event_buffer_size)
    "bf_pions": (211, -211, 21, 2.0, 0.2, 3.0, 1, 1300),
    "bf_kaons_eta": (321, -321, 11, 1.6, 0.1, 2.0, 0, 100000),
}


def write_input(path, seed=20260030, nev=9, mult=420):
    rng = np.random.default_rng(seed)
    with gzip.open(path, "wt", compresslevel=6) as f:
        for _ in range(nev):
            n = int(mult + rng.integers(-40, 40))
            f.write(f"{n}\n")
            for _ in range(n):
                mv, m = SPECIES[int(rng.choice(len(SPECIES), p=[0.36, 0.36, 0.09, 0.09, 0.06, 0.04]))]
                pT = rng.gamma(2.0, 0.28)
                phi = rng.uniform(-np.pi, np.pi)
                y = rng.normal(0.0, 1.1)
                mT = np.sqrt(m * m + pT * pT)
                px, py, pz, E = pT * np.cos(phi), pT * np.sin(phi), mT * np.sinh(y), mT * np.cosh(y)
                t, x, yy, z = rng.gamma(4, 2.5), rng.normal(0, 4), rng.normal(0, 4), rng.normal(0, 5)
                f.write("%d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n" % (mv, m, t, x, yy, z, E, px, py, pz))


def params_text(case):
    alpha, beta, Bnpts, Brap_max, BpT_min, BpT_max, rap_type, buf = case
    return HBTParams(randomSeed=4711).parameters_dat(
        analyze_HBT=0, analyze_balance_function=1, event_buffer_size=buf, particle_alpha=alpha, particle_beta=beta,
        Bnpts=Bnpts, Brap_max=Brap_max, BpT_min=BpT_min, BpT_max=BpT_max, rap_type=rap_type)


def main():
    gz = os.path.join(HERE, "bf_input.gz")
    write_input(gz)
    meta = {}
    for name, case in CASES.items():
        wd = tempfile.mkdtemp(prefix="bf_golden_")
        os.makedirs(os.path.join(wd, "EOS"))
        os.makedirs(os.path.join(wd, "results"))
        shutil.copy(PDG, os.path.join(wd, "EOS", "pdg.dat"))
        shutil.copy(gz, os.path.join(wd, "results", "particle_samples.gz"))
        open(os.path.join(wd, "parameters.dat"), "w").write(params_text(case))
        r = subprocess.run([REF_EXE], cwd=wd, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        files = sorted(f for f in os.listdir(os.path.join(wd, "results")) if f.endswith(".dat"))
        assert len(files) == 3, files
        for fn in files:
            shutil.copy(os.path.join(wd, "results", fn), os.path.join(HERE, f"{name}.{fn}"))
        meta[name] = {"case": list(case), "files": files, "randomSeed": 4711}
        shutil.rmtree(wd)
        print(name, files)
    json.dump(meta, open(os.path.join(HERE, "bf_cases.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
