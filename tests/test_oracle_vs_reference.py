"""Oracle restatement vs the unmodified reference executed here (oracle/_ref, built from
/root/reference by oracle/Makefile).  Skipped where the compiled reference is absent."""
import numpy as np
import pytest

from hadronic_afterburner_toolkit_b200 import synth
from hadronic_afterburner_toolkit_b200.params import HBTParams
from oracle import oracle_py as O

pytestmark = pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref not built")

CASES = {
    "az0": (HBTParams(), 2, 4, 250),
    "az1": (HBTParams(azimuthal_flag=1, n_KT=9, KT_max=0.95, qnpts=21), 2, 4, 250),
    "qinv": (HBTParams(invariant_radius_flag=1, qnpts=21), 1, 3, 250),
    "noboost": (HBTParams(long_comoving_boost=0, qnpts=21), 1, 3, 200),
    "cap": (HBTParams(needed_number_of_pairs=2000, qnpts=21), 2, 4, 250),
    "cap_az": (HBTParams(needed_number_of_pairs=300, azimuthal_flag=1, qnpts=21), 2, 4, 250),
    "one_event": (HBTParams(qnpts=21), 2, 1, 300),  # mixed_nev == 1: self-pairing allowed
    "asym_window": (HBTParams(qnpts=16, q_min=-0.05, q_max=0.25), 1, 4, 250),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_bit_identical(name):
    P, ngrp, nev, mult = CASES[name]
    batches = synth.make_batches(11, ngrp, nev, multiplicity=mult)
    ref = O.run_reference(P, batches)
    o = O.Oracle(P)
    for b in batches:
        o.process_batch(b)
    acc = o.accumulators()
    for k in ("num_count", "num_cos", "sum_qo", "sum_qs", "sum_ql", "den_count", "npairs_num", "npairs_den"):
        assert np.array_equal(getattr(ref, k), getattr(acc, k)), k
    if P.invariant_radius_flag == 1:
        for k in ("qinv_count", "qinv_mean", "qinv_num", "qinv_den", "npairs_num_qinv", "npairs_den_qinv"):
            assert np.array_equal(getattr(ref, k), getattr(acc, k)), k
    assert ref.psi_ref == acc.psi_ref


def test_rng_stream():
    ints, reals = O.reference_rng(20260017, 3000)
    o = O.Oracle(HBTParams(randomSeed=20260017))
    assert [o.rand_int_uniform() for _ in range(3000)] == ints
    assert [o.rand_uniform() for _ in range(3000)] == reals


def test_same_only_mode():
    P = HBTParams(qnpts=21)
    batches = synth.make_batches(5, 2, 3, multiplicity=200)
    ref = O.run_reference(P, batches, same_only=True)
    o = O.Oracle(P)
    for b in batches:
        o.process_batch(b, do_mixed=False)
    acc = o.accumulators()
    assert np.array_equal(ref.num_count, acc.num_count) and np.array_equal(ref.num_cos, acc.num_cos)
    assert not acc.den_count.any() and not ref.den_count.any()
