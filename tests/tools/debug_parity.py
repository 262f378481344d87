"""Developer aid: one small same+mixed batch through the product and the oracle, every
accumulator compared separately (prints the worst deviation instead of stopping at the first)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hadronic_afterburner_toolkit_b200 import synth  # noqa: E402
from hadronic_afterburner_toolkit_b200.hbt_correlation import HBT_correlation  # noqa: E402
from hadronic_afterburner_toolkit_b200.params import C3  # noqa: E402
from oracle import oracle_py  # noqa: E402

P = C3.with_(qnpts=21)
batches = synth.make_batches(20260001, 1, 4, multiplicity=600)
for stats in (False, True):
    h = HBT_correlation(P, device=0, stage_counters=stats)
    o = oracle_py.Oracle(P)
    for b in batches:
        h.calculate_HBT_correlation_function(b)
        o.process_batch(b)
    got, ref = h.accumulators(), o.accumulators()
    print("stats" if stats else "production", "stage ref", list(map(int, ref.stage)), "got", list(map(int, got.stage)))
    for name in ("num_count", "den_count", "npairs_num", "npairs_den"):
        a, b = np.asarray(getattr(ref, name), dtype=np.float64), np.asarray(getattr(got, name), dtype=np.float64)
        print(f"  {name}: sum ref {a.sum():.0f} got {b.sum():.0f}, entries differing {int((a != b).sum())}")
    cnt = np.asarray(ref.num_count, dtype=np.float64)
    for name in ("num_cos", "sum_qo", "sum_qs", "sum_ql"):
        a, b = np.asarray(getattr(ref, name)), np.asarray(getattr(got, name))
        d = np.abs(a - b)
        k = int(d.argmax())
        print(f"  {name}: max abs dev {d.max():.3e} at bin {k} (ref {a.flat[k]:.17g} got {b.flat[k]:.17g} count {cnt.flat[k]:.0f})")
