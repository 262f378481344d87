"""The launch structures of a whole batch give the same histograms: the same-event kernel and the v4 mixed-event kernel
next to each other on two streams (default), one after the other (HBT_B200_CORUN=0), other co-run splits, the mixed-event
loops on the v3 kernel (HBT_B200_MIXED4=0) and the single fused kernel of earlier versions (HBT_B200_SPLIT=0).  The
switches are read at hbt_create.  Integers bit-exact (every pair decides its bins with the reference's own results
whichever kernel evaluates it); floating sums only differ by the order of the atomic additions."""
import os

import numpy as np
import pytest

from hadronic_afterburner_toolkit_b200 import hbtio, synth
from hadronic_afterburner_toolkit_b200.hbt_correlation import HBT_correlation
from hadronic_afterburner_toolkit_b200.params import C4, HBTParams
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu

VARIANTS = {
    "default": {},
    "one_after_the_other": {"HBT_B200_CORUN": "0"},
    "corun_6_16": {"HBT_B200_CORUN_SAME": "6", "HBT_B200_CORUN_MIXED": "16"},
    "mixed_on_v3": {"HBT_B200_MIXED4": "0"},
    "fused_v3": {"HBT_B200_SPLIT": "0"},
}

CASES = {
    # whole batches (K_phi grids are never coalesced): launch_fused -> launch_split_pair
    "az": (C4.with_(qnpts=15, n_KT=4, n_Kphi=4), 2, 6, 900),
    # small batches together (flush_pending -> launch_split_pair), several launches
    "3d_small_batches": (HBTParams(qnpts=21), 12, 4, 500),
    # KT_min = 0: the prefilter's error floor is active (the FLOOR instantiation of both pair loops)
    "kt_min_zero": (HBTParams(qnpts=21, n_KT=3, KT_min=0.0, KT_max=1.0), 3, 5, 600),
}


def run(P, batches, env):
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        h = HBT_correlation(P)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    for b in batches:
        h.calculate_HBT_correlation_function(b)
    acc = h.accumulators()
    h.close()
    return acc


@pytest.mark.parametrize("case", sorted(CASES))
def test_launch_structures_agree(case):
    P, nb, nev, mult = CASES[case]
    batches = synth.make_batches(20260300 + len(case), nb, nev, multiplicity=mult)
    o = O.Oracle(P)
    for b in batches:
        o.process_batch(b)
    ref = o.accumulators()
    base = None
    for name, env in VARIANTS.items():
        acc = run(P, batches, env)
        hbtio.compare(ref, acc, rtol=1e-10, check_stage="cheap")  # each of them against the oracle
        if base is None:
            base = acc
            continue
        for f in ("num_count", "den_count"):
            assert np.array_equal(np.asarray(getattr(base, f)), np.asarray(getattr(acc, f))), f"{case}/{name}: {f} differs from the default launch"
