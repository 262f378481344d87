#!/usr/bin/env python
"""Small runs of what round 2, session 3 added, for compute-sanitizer: whole batches as two kernels next to each other
(hbt_pairs_v3<same-event> on the lane's stream, hbt_pairs_v4_mixed on its side stream, late launches of both), the v4
kernel's parked pairs, 16-bit survivor queues, swizzled / padded list-1 tiles; small batches together (page-locked caller
buffers, staged launch tables) and whole K_phi batches (not coalesced)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from hadronic_afterburner_toolkit_b200 import synth  # noqa: E402
from hadronic_afterburner_toolkit_b200.hbt_correlation import HBT_correlation  # noqa: E402
from hadronic_afterburner_toolkit_b200.params import HBTParams  # noqa: E402

for name, P, mult in (("3-D, small batches together, pinned", HBTParams(qnpts=21), 300),
                      ("K_phi, whole batches", HBTParams(qnpts=15, azimuthal_flag=1, n_Kphi=4), 400),
                      ("3-D, KT_min = 0 (error floor of the prefilter), pageable", HBTParams(qnpts=21, KT_min=0.0, KT_max=1.0, n_KT=3), 300)):
    h = HBT_correlation(P)
    h.pin_host = name.endswith("pinned")
    for b in synth.make_batches(11, 4, 4, multiplicity=mult):
        h.calculate_HBT_correlation_function(b)
    acc = h.accumulators()
    print(name, int(acc.num_count.sum()), int(acc.den_count.sum()), "deferred", h.deferred_pairs() if hasattr(h, "deferred_pairs") else "")
    h.close()
