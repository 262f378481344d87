#!/usr/bin/env python
"""Golden vectors for the downstream consumers, produced by running the
reference's OWN scripts in this container (they cannot travel to the GPU box):

  * /root/reference/ebe_scripts/average_event_HBT_correlation_function.py on a small tree of
    event folders  ->  tests/golden/ebe_avg.npz (inputs + the script's output tables)
  * /root/reference/ebe_scripts/fit_HBT_radii.py on synthetic 8-column tables
    ->  tests/golden/ebe_fit.npz (inputs + the HBT_radii_KT_*.dat datasets it stored)

h5py is not installed here, and the script's `string_` is gone from numpy 2: the fit script is
executed unmodified against a dict-backed stand-in for `h5py` and with `string_` supplied
through builtins.  Neither shim touches the numerics (curve_fit on the arrays it is handed).

    python tests/golden/make_golden_ebe.py
"""
import builtins
import os
import runpy
import subprocess
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/ebe_scripts"
HBARC = 0.19733


def synthetic_5col(rng, nq=5):
    q = np.linspace(-0.1, 0.1, nq)
    ql, qo, qs = np.meshgrid(q, q, q, indexing="ij")  # q_long outermost, like the writer
    den = rng.integers(0, 4000, size=qo.size).astype(float)
    num = den * (1.0 + 0.6 * np.exp(-((qo.ravel() ** 2 + qs.ravel() ** 2 + ql.ravel() ** 2) * (5.0 / HBARC) ** 2))) \
        + rng.normal(0, 3, size=qo.size)
    return np.column_stack([qo.ravel() + rng.normal(0, 1e-4, qo.size), qs.ravel(), ql.ravel(), num, den])


def make_avg():
    rng = np.random.default_rng(20260417)
    names = ["HBT_correlation_function_KT_0.15_0.25.dat", "HBT_correlation_function_KT_0.25_0.35.dat"]
    with tempfile.TemporaryDirectory() as td:
        work, out = os.path.join(td, "work"), os.path.join(td, "avg")
        inputs = {}
        for ev in range(3):
            d = os.path.join(work, f"UrQMD_{ev}", "UrQMD_results")
            os.makedirs(d)
            for n in names:
                t = synthetic_5col(rng)
                # the analysis writes 8 significant digits after the point in scientific notation
                np.savetxt(os.path.join(d, n), t, fmt="%18.8e", delimiter="    ")
                inputs[f"in/{ev}/{n}"] = np.loadtxt(os.path.join(d, n))
        subprocess.run([sys.executable, os.path.join(REF, "average_event_HBT_correlation_function.py"), work, out],
                       check=True, stdout=subprocess.DEVNULL)
        outputs = {f"out/{n}": np.loadtxt(os.path.join(out, n)) for n in names}
        text = {f"text/{n}": np.frombuffer(open(os.path.join(out, n), "rb").read(), dtype=np.uint8) for n in names}
    np.savez_compressed(os.path.join(HERE, "ebe_avg.npz"), **inputs, **outputs, **text)
    print("ebe_avg.npz:", len(inputs), "inputs,", len(outputs), "outputs")


def synthetic_8col(rng, radii, lam, nq=15, qmax=0.14):
    q = np.linspace(-qmax, qmax, nq)
    ql, qo, qs = np.meshgrid(q, q, q, indexing="ij")
    qo, qs, ql = qo.ravel(), qs.ravel(), ql.ravel()
    ro, rs, rl, ros, rol = [r / HBARC for r in radii]
    c = lam * np.exp(-((ro * qo) ** 2 + (rs * qs) ** 2 + (rl * ql) ** 2 + 2 * qo * qs * ros ** 2 + 2 * qo * ql * rol ** 2))
    err = 0.01 + 0.02 * rng.random(c.size)
    t = np.zeros((c.size, 8))
    t[:, 0], t[:, 1], t[:, 2] = qo + rng.normal(0, 2e-4, c.size), qs + rng.normal(0, 2e-4, c.size), ql + rng.normal(0, 2e-4, c.size)
    t[:, 3] = rng.integers(100, 5000, c.size)
    t[:, 6] = c + err * rng.normal(0, 1, c.size)
    t[:, 7] = err
    return t


class _Attrs(dict):
    def create(self, k, v):
        self[k] = v


class _Dataset:
    def __init__(self, data):
        self.data = np.array(data)
        self.attrs = _Attrs()


class _Group(dict):
    def get(self, k):
        return self[k].data if isinstance(self[k], _Dataset) else self[k]

    def create_dataset(self, name, data=None, **kw):
        self[name] = _Dataset(data)
        return self[name]


class _File(dict):
    def close(self):
        pass


def make_fit():
    rng = np.random.default_rng(20260418)
    truth = {"0_0.2": ((6.1, 5.4, 7.2, 1.1, 0.8), 0.72), "0.2_0.4": ((5.2, 4.9, 6.0, 0.9, 0.6), 0.65),
             "0.4_0.6": ((4.4, 4.3, 4.9, 0.7, 0.5), 0.61), "0.6_0.8": ((3.7, 3.8, 4.0, 0.5, 0.4), 0.58)}
    db = _File()
    db["event_0"] = _Group()
    inputs = {}
    for cut, (radii, lam) in truth.items():
        t = synthetic_8col(rng, radii, lam)
        db["event_0"][f"HBT_correlation_function_KT_{cut}.dat"] = t
        inputs[f"in/{cut}"] = t
    fake = types.ModuleType("h5py")
    fake.File = lambda name, *a, **k: db
    sys.modules["h5py"] = fake
    builtins.string_ = np.bytes_  # numpy 2 dropped the alias the script uses for the header attribute
    argv = sys.argv
    sys.argv = ["fit_HBT_radii.py", "database.h5"]
    try:
        runpy.run_path(os.path.join(REF, "fit_HBT_radii.py"), run_name="__main__")
    finally:
        sys.argv = argv
        del builtins.string_
        del sys.modules["h5py"]
    outputs = {f"out/{cut}": db["event_0"][f"HBT_radii_KT_{cut}.dat"].data for cut in truth}
    header = db["event_0"]["HBT_radii_KT_0_0.2.dat"].attrs["header"]
    np.savez_compressed(os.path.join(HERE, "ebe_fit.npz"), header=np.frombuffer(bytes(header), dtype=np.uint8),
                        **inputs, **outputs)
    for cut in truth:
        print(cut, "truth", truth[cut][0][:3], "fit@0.1", outputs[f"out/{cut}"][2, [3, 5, 7]])


if __name__ == "__main__":
    make_avg()
    make_fit()
