#!/usr/bin/env python
"""A/B timing of library variants (scripts/build_variant.sh) on one C5-shape group, device resident:
the two loops as separate kernels (one lane) and the fused launch.  One subprocess per variant
(the library path is fixed at import).  Prints one JSON line per (variant, environment).

    python scripts/ab_variants.py base dbg1 dbg2 ... [--occ 18,12] [--shape C5] [--events 100]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(a):
    import numpy as np
    import torch

    from hadronic_afterburner_toolkit_b200 import synth
    from hadronic_afterburner_toolkit_b200.hbt_correlation import HBT_correlation, Random, _check, gather_rapidity
    from hadronic_afterburner_toolkit_b200.params import C2, C3, C4, C5, KAON_MASS, PION_MASS

    P = {"C2": C2, "C3": C3, "C4": C4, "C4_31": C4.with_(qnpts=31), "C5": C5, "C5_QINV": C5.with_(invariant_radius_flag=1)}[a.shape]
    nev, mult = a.events, a.multiplicity
    mass = KAON_MASS if a.kaons else PION_MASS
    arr = synth.make_group(20260005, 0, nev, mass, mult).reshape(nev * mult, 8)
    flat = np.ascontiguousarray(np.concatenate([gather_rapidity(P, arr[e * mult:(e + 1) * mult]) for e in range(nev)]))
    off = np.arange(nev + 1, dtype=np.int64) * mult
    d = torch.from_numpy(flat).cuda()
    h = HBT_correlation(P)
    L, hh = h._L, h._h
    ids, cs = Random(P.randomSeed).mixed_plan(nev, nev)
    nmix = ids.shape[1]
    n = flat.shape[0]
    psi = 0.3 if P.azimuthal_flag else 0.0
    out = {"variant": a.worker, "occ": os.environ.get("HBT_B200_OCC"), "kernel": os.environ.get("HBT_B200_KERNEL"), "shape": a.shape, "n": n,
           "pairs_same": n * (n - 1) // 2, "pairs_mixed": nev * mult * nmix * mult}

    def run(kind, reps):
        t0 = h.timers()
        for _ in range(reps):
            if kind == "same":
                _check(hh, L.hbt_accumulate_same_dev(hh, d.data_ptr(), n, psi))
            elif kind == "mixed":
                _check(hh, L.hbt_accumulate_mixed_dev(hh, d.data_ptr(), off.ctypes.data, nev, None, None, 0, ids.ctypes.data,
                                                      cs.ctypes.data, nmix, psi))
            else:
                _check(hh, L.hbt_accumulate_batch_dev(hh, d.data_ptr(), off.ctypes.data, nev, ids.ctypes.data, cs.ctypes.data,
                                                      nmix, psi))
            h.synchronize()
        t1 = h.timers()
        return (t1["same_ms"] + t1["mixed_ms"] - t0["same_ms"] - t0["mixed_ms"]) / reps

    _check(hh, L.hbt_set_option(hh, 4, 1))  # one lane: launches do not overlap
    for kind in (("same", "mixed") if a.no_fused else ("same", "mixed", "fused")):
        run(kind, 1)
        out[kind + "_ms"] = round(run(kind, a.reps), 4)
    st = h.stage_counters()
    out["accepted_same_per_launch"] = int(st[4]) // (2 * (a.reps + 1))
    out["accepted_mixed_per_launch"] = int(st[10]) // (2 * (a.reps + 1))
    print(json.dumps(out), flush=True)
    h.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("variants", nargs="*")
    ap.add_argument("--worker")
    ap.add_argument("--occ", default="")
    ap.add_argument("--shape", default="C5")
    ap.add_argument("--events", type=int, default=100)
    ap.add_argument("--multiplicity", type=int, default=1500)
    ap.add_argument("--kaons", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-fused", action="store_true")
    ap.add_argument("--env", action="append", default=[], help="KEY=VALUE for the workers")
    a = ap.parse_args()
    if a.worker:
        return worker(a)
    pkg = os.path.join(ROOT, "hadronic_afterburner_toolkit_b200")
    for v in a.variants:
        lib = os.path.join(pkg, "libhbt_b200.so") if v == "base" else os.path.join(pkg, "variants", f"libhbt_b200_{v}.so")
        for occ in (a.occ.split(",") if a.occ else [""]):
            env = dict(os.environ, HBT_B200_LIB=lib)
            for kv in a.env:
                k, v_ = kv.split("=", 1)
                env[k] = v_
            if occ:
                env["HBT_B200_OCC"] = occ
            cmd = [sys.executable, os.path.abspath(__file__), "--worker", v, "--shape", a.shape, "--events", str(a.events),
                   "--multiplicity", str(a.multiplicity), "--reps", str(a.reps)] + (["--kaons"] if a.kaons else []) + (["--no-fused"] if a.no_fused else [])
            r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
            sys.stdout.write(r.stdout if r.returncode == 0 else json.dumps({"variant": v, "occ": occ, "error": r.stderr[-800:]}) + "\n")
            sys.stdout.flush()


if __name__ == "__main__":
    main()
