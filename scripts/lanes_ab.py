#!/usr/bin/env python
"""A/B of HBT_OPT_LANES (one compute stream vs two that take batches in turn) on the group
shapes of BASELINE.json's configs, through the host-buffer C-ABI call (hbt_accumulate_batch,
plans pre-drawn so that the Python side costs little).  Prints one JSON line per (shape, lanes):
device stopwatch time of the whole sequence (hbt_timer_start/stop: uploads the kernels wait for,
sort/cull helpers, pair kernels), the launch timers' union, and the host wall clock.

    python scripts/lanes_ab.py [--shapes C2,C3,C4,C5] [--groups 40]
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hadronic_afterburner_toolkit_b200 import synth  # noqa: E402
from hadronic_afterburner_toolkit_b200.hbt_correlation import HBT_correlation, Random, _check, gather_rapidity  # noqa: E402
from hadronic_afterburner_toolkit_b200.params import C2, C3, C4, C5, PION_MASS  # noqa: E402

SHAPES = {  # params, events per group, mixed?, default groups
    "C2": (C2, 10, False), "C3": (C3, 10, True), "C4": (C4.with_(qnpts=31), 50, True), "C5": (C5, 100, True),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="C2,C3,C4,C5")
    ap.add_argument("--groups", type=int, default=40)
    ap.add_argument("--repeat", type=int, default=3)
    a = ap.parse_args()
    for name in a.shapes.split(","):
        P, nev, mixed = SHAPES[name]
        ng = a.groups if nev <= 10 else max(4, a.groups // (nev // 5))
        nd = min(ng, 8)  # distinct groups, cycled
        rng = Random(P.randomSeed)
        groups = []
        for g in range(nd):
            arr = synth.make_group(20260000 + int(name[1]), g, nev, PION_MASS, 1500).reshape(nev * 1500, 8)
            cut = [gather_rapidity(P, arr[e * 1500:(e + 1) * 1500]) for e in range(nev)]
            flat = np.ascontiguousarray(np.concatenate(cut))
            off = np.zeros(nev + 1, dtype=np.int64)
            off[1:] = np.cumsum([len(c) for c in cut])
            ids, cs = rng.mixed_plan(nev, nev)
            groups.append((flat, off, ids, cs))
        n = groups[0][0].shape[0]
        nmix = groups[0][2].shape[1]
        pairs = ng * (n * (n - 1) // 2 + (sum(len(c) for c in cut) * nmix * 1500 if mixed else 0))
        for lanes in (1, 2):
            h = HBT_correlation(P, lanes=lanes)
            L, hh = h._L, h._h
            best = None
            for rep in range(a.repeat + 1):  # rep 0 = warm-up (allocations)
                _check(hh, L.hbt_synchronize(hh))
                t0 = time.perf_counter()
                _check(hh, L.hbt_timer_start(hh))
                tm0 = h.timers()
                for g in range(ng):
                    flat, off, ids, cs = groups[g % nd]
                    _check(hh, L.hbt_accumulate_batch(hh, flat.ctypes.data, off.ctypes.data, nev, None, None, 0,
                                                      ids.ctypes.data, cs.ctypes.data, nmix, 0.0, 1, 1 if mixed else 0))
                ms = ctypes.c_double()
                _check(hh, L.hbt_timer_stop(hh, ctypes.byref(ms)))
                _check(hh, L.hbt_synchronize(hh))
                wall = time.perf_counter() - t0
                tm1 = h.timers()
                kern = tm1["same_ms"] + tm1["mixed_ms"] - tm0["same_ms"] - tm0["mixed_ms"]
                if rep and (best is None or ms.value < best[0]):
                    best = (ms.value, kern, wall)
            h.close()
            print(json.dumps({"shape": name, "lanes": lanes, "groups": ng, "particles_per_group": int(n), "pairs": int(pairs),
                              "device_ms": best[0], "launch_timers_ms": best[1], "wall_ms": 1e3 * best[2],
                              "pairs_per_s_device": pairs / (best[0] * 1e-3), "pairs_per_s_wall": pairs / best[2]}), flush=True)


if __name__ == "__main__":
    main()
