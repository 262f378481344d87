"""hbt_group_*: one analysis over several engine contexts (one per GPU; or several on one GPU, which exercises the
same code on a single-GPU box).  Batches go to the contexts in turn; the needed_number_of_pairs cap, cumulative over
the batches IN ORDER (src/HBT_correlation.cpp:402-406, :651-655), must cut at the reference's pair whichever context
holds the batch that crosses it.  Checked against the CPU oracle: integers bit-exact, sums within 1e-10."""
import numpy as np
import pytest

from hadronic_afterburner_toolkit_b200 import capi, hbtio, synth
from hadronic_afterburner_toolkit_b200.hbt_correlation import HBT_correlation
from hadronic_afterburner_toolkit_b200.params import C3, C4, HBTParams
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu
RTOL = 1e-10

CASES = {
    # name: (params, groups, events/group, multiplicity, cap must engage)
    "uncapped": (C3.with_(qnpts=21), 6, 4, 400, False),
    "cap_first_batches": (HBTParams(qnpts=21, needed_number_of_pairs=4000.0), 6, 4, 400, True),
    # many small batches: far from the cap for the first ones (they run asynchronously), then it engages
    "cap_mid_run": (HBTParams(qnpts=21, needed_number_of_pairs=40000.0), 220, 3, 60, True),
    "cap_zero": (HBTParams(qnpts=11, needed_number_of_pairs=0.0), 4, 3, 150, True),
    "cap_az": (C4.with_(qnpts=11, n_KT=4, n_Kphi=4, needed_number_of_pairs=900.0), 5, 4, 400, True),
    "cap_qinv": (HBTParams(qnpts=21, invariant_radius_flag=1, needed_number_of_pairs=700.0), 6, 4, 400, True),
    "cap_real_mixed": (HBTParams(qnpts=21, needed_number_of_pairs=3000.0), 5, 4, 400, True),
}


def _batches(name):
    P, ngrp, nev, mult, _ = CASES[name]
    b = synth.make_batches(20260011, ngrp, nev, multiplicity=mult)
    if name == "cap_real_mixed":  # separate mixed-event lists (read_in_real_mixed_events = 1)
        m = synth.make_batches(20260012, ngrp, nev + 1, multiplicity=mult - 50)
        b = [hbtio.Batch(x.same, y.same) for x, y in zip(b, m)]
    return b


def _devices(kind):
    ndev = capi.lib().hbt_device_count()
    if kind == "gpus":
        if ndev < 2:
            pytest.skip("needs >= 2 GPUs")
        return list(range(ndev))
    return [0] * int(kind)


@pytest.mark.parametrize("kind", ["2", "3", "gpus"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_group_against_oracle(name, kind):
    P, ngrp, nev, mult, capped = CASES[name]
    batches = _batches(name)
    o = O.Oracle(P)
    for b in batches:
        o.process_batch(b)
    ref = o.accumulators()
    h = HBT_correlation(P, devices=_devices(kind))
    for b in batches:
        h.calculate_HBT_correlation_function(b)
    acc = h.accumulators()
    hbtio.compare(ref, acc, rtol=RTOL, check_stage="cheap")
    if capped:
        lim = int(P.needed_number_of_pairs) + 1
        assert int(acc.npairs_num.max()) == lim and int(acc.npairs_den.max()) == lim
        assert h.ordered_batches() > 0
    else:
        assert h.ordered_batches() == 0
    if name == "cap_mid_run":
        assert h.ordered_batches() < ngrp  # the first batches ran asynchronously, far from the cap
    h.close()


def test_group_of_one_is_the_plain_context():
    P, ngrp, nev, mult, _ = CASES["cap_first_batches"]
    batches = _batches("cap_first_batches")
    a = HBT_correlation(P)
    b = HBT_correlation(P, devices=[0])
    for x in batches:
        a.calculate_HBT_correlation_function(x)
        b.calculate_HBT_correlation_function(x)
    ra, rb = a.accumulators(), b.accumulators()
    hbtio.compare(ra, rb, rtol=RTOL, check_stage="cheap")  # the very same launches (the order of the atomic adds may differ)
