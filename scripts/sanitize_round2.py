#!/usr/bin/env python
"""Small runs of what round 2 added, for compute-sanitizer: q_inv mode on the tuned kernels (+ the replica fold), a
one-sided q window, small batches in one launch, page-locked caller buffers, the BalanceFunction kernel."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from hadronic_afterburner_toolkit_b200 import hbtio, synth  # noqa: E402
from hadronic_afterburner_toolkit_b200.balance_function import BalanceFunction  # noqa: E402
from hadronic_afterburner_toolkit_b200.hbt_correlation import HBT_correlation, Random  # noqa: E402
from hadronic_afterburner_toolkit_b200.params import HBTParams  # noqa: E402

for name, P in (("qinv", HBTParams(invariant_radius_flag=1, qnpts=21)),
                ("positive window", HBTParams(qnpts=13, q_min=0.02, q_max=0.14)),
                ("3-D, small batches together, pinned", HBTParams(qnpts=21))):
    h = HBT_correlation(P)
    h.pin_host = name.endswith("pinned")
    for b in synth.make_batches(7, 5, 3, multiplicity=300):
        h.calculate_HBT_correlation_function(b)
    acc = h.accumulators()
    print(name, int(acc.num_count.sum()), int(acc.den_count.sum()), int(np.sum(acc.qinv_count)) if P.invariant_radius_flag else "")
    h.close()

rng = np.random.default_rng(1)


def species(nev, n):
    return [{"pT": rng.gamma(2.0, 0.3, n), "phi": rng.uniform(-np.pi, np.pi, n), "rap_y": rng.normal(0, 1.3, n),
             "rap_eta": rng.normal(0, 1.5, n)} for _ in range(nev)]


plus, minus = species(4, 300), species(4, 280)
bf = BalanceFunction(211, -211, 21, 2.0, 0.2, 3.0, 1, ran_gen=Random(3))
bf.calculate_balance_function({"a": plus, "abar": minus, "b": minus, "bbar": plus})
print("bf", int(bf.histograms().sum()))
bf.close()
