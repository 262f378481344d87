#!/usr/bin/env python
"""Production launches only (no instrumented pass), for ncu: `--launches` C5-shape groups through
hbt_accumulate_batch_dev.  HBT_B200_FUSE=0 gives the two loops as separate kernels.

    ncu --set full --clock-control none --import-source on -k regex:hbt_pairs_v3 -s 1 -c 1 \\
        -o gpurun_out/fused python scripts/profile_step.py --launches 2
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402  (device buffer only)

from hadronic_afterburner_toolkit_b200 import synth  # noqa: E402
from hadronic_afterburner_toolkit_b200.hbt_correlation import HBT_correlation, Random, _check, gather_rapidity  # noqa: E402
from hadronic_afterburner_toolkit_b200.params import C2, C3, C4, C5, PION_MASS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--launches", type=int, default=2)
ap.add_argument("--events", type=int, default=100)
ap.add_argument("--shape", default="C5")
a = ap.parse_args()
P = {"C2": C2, "C3": C3, "C4": C4.with_(qnpts=31), "C4_41": C4, "C5": C5}[a.shape]
nev, mult = a.events, 1500
arr = synth.make_group(20260005, 0, nev, PION_MASS, mult).reshape(nev * mult, 8)
flat = np.ascontiguousarray(np.concatenate([gather_rapidity(P, arr[e * mult:(e + 1) * mult]) for e in range(nev)]))
off = np.arange(nev + 1, dtype=np.int64) * mult
d = torch.from_numpy(flat).cuda()
h = HBT_correlation(P)
ids, cs = Random(P.randomSeed).mixed_plan(nev, nev)
for _ in range(a.launches):
    _check(h._h, h._L.hbt_accumulate_batch_dev(h._h, d.data_ptr(), off.ctypes.data, nev, ids.ctypes.data, cs.ctypes.data,
                                               ids.shape[1], 0.3 if P.azimuthal_flag else 0.0))
h.synchronize()
t = h.timers()
print({k: v for k, v in t.items()})
h.close()
