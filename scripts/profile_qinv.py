#!/usr/bin/env python
"""q_inv mode (invariant_radius_flag = 1), production launches for ncu: one group through hbt_accumulate_batch_dev.

    ncu --set full --clock-control none --import-source on -k regex:hbt_pairs_v3 -s 2 -c 2 -o gpurun_out/qinv \\
        python scripts/profile_qinv.py --events 40
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
from hadronic_afterburner_toolkit_b200.params import C5, PION_MASS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--launches", type=int, default=2)
ap.add_argument("--events", type=int, default=40)
a = ap.parse_args()
P = C5.with_(invariant_radius_flag=1)
nev, mult = a.events, 1500
arr = synth.make_group(20260005, 0, nev, PION_MASS, mult).reshape(nev * mult, 8)
flat = np.ascontiguousarray(np.concatenate([gather_rapidity(P, arr[e * mult:(e + 1) * mult]) for e in range(nev)]))
off = np.arange(nev + 1, dtype=np.int64) * mult
d = torch.from_numpy(flat).cuda()
h = HBT_correlation(P)
ids, cs = Random(P.randomSeed).mixed_plan(nev, nev)
for _ in range(a.launches):
    _check(h._h, h._L.hbt_accumulate_batch_dev(h._h, d.data_ptr(), off.ctypes.data, nev, ids.ctypes.data, cs.ctypes.data,
                                               ids.shape[1], 0.0))
h.synchronize()
print(h.timers())
acc = h.accumulators()
print("q_inv accepted same / mixed:", int(acc.qinv_count.sum()), int(acc.qinv_den.sum()), " 3-D accepted:", int(acc.num_count.sum()),
      int(acc.den_count.sum()), " pairs:", h.pairs_same, h.pairs_mixed)
h.close()
