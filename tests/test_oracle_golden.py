"""The CPU restatement (oracle/hbt_oracle.c) against the golden vectors produced by the
unmodified reference (tests/golden/make_golden.py) — this is what pins the oracle where
/root/reference is absent.  Everything is bit-exact: the oracle performs the same IEEE
operations in the same order with the same libm."""
import numpy as np
import pytest

from conftest import GOLDEN_NAMES, load_golden
from oracle import oracle_py as O

INT_FIELDS = ("num_count", "den_count", "npairs_num", "npairs_den")
SUM_FIELDS = ("num_cos", "sum_qo", "sum_qs", "sum_ql")


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_matches_reference_golden(name):
    P, batches, ref, meta = load_golden(name)
    o = O.Oracle(P)
    for b in batches:
        o.process_batch(b)
    acc = o.accumulators()
    for k in INT_FIELDS:
        assert np.array_equal(getattr(ref, k), getattr(acc, k)), k
    for k in SUM_FIELDS:
        assert np.array_equal(getattr(ref, k), getattr(acc, k)), k
    if P.invariant_radius_flag == 1:
        for k in ("qinv_count", "qinv_mean", "qinv_num", "qinv_den", "npairs_num_qinv", "npairs_den_qinv"):
            assert np.array_equal(getattr(ref, k), getattr(acc, k)), k
    if P.azimuthal_flag == 1:
        assert acc.psi_ref == ref.psi_ref
    # same-event pair count as the reference logs it (src/HBT_correlation.cpp:282-284)
    assert int(acc.stage[0]) == meta["pairs_same"]
    # stage populations are nested
    s = acc.stage
    assert all(s[i] >= s[i + 1] for i in range(5)) and all(s[i] >= s[i + 1] for i in range(6, 11))
    assert int(s[5]) == int(np.sum(acc.npairs_num)) and int(s[11]) == int(np.sum(acc.npairs_den))


def test_psi_ref_known_answers():
    """unit_tests/HBT_unittest.cc:39,43,47 — the only numerical goldens the reference's own
    tests hold for this path (zipped-UrQMD fixture, 116+97 pi+, tolerance 1e-8 there)."""
    P, batches, ref, meta = load_golden("c1_urqmd_gz")
    assert meta["events_per_batch"] == [[116, 97]]
    p, _ = batches[0].flat()
    for n, want in ((1, -1.6751628499713109), (2, -0.85236384292521539), (3, 0.06674083818478263)):
        assert abs(O.Oracle.psi_ref(p, n) - want) < 1e-8


def test_reader_counts_of_the_fixtures(golden_cases):
    """particleSamples_unittest.cc known answers: pi+ 116/97 (UrQMD), 619/635 (iSS gz),
    K+ 11/14 (OSCAR)."""
    assert golden_cases["unit_urqmd_txt"]["events_per_batch"] == [[116, 97]]
    assert golden_cases["c1_iss_gz"]["events_per_batch"] == [[619, 635]]
    assert golden_cases["unit_oscar_kplus"]["events_per_batch"] == [[11, 14]]


def test_rng_mapping_known_answers():
    """First draws of RandomUtil::Random(12345) on libstdc++ 13 (captured from the reference
    build: oracle/_ref/ref_driver rng 12345 5)."""
    from hadronic_afterburner_toolkit_b200.params import HBTParams

    o = O.Oracle(HBTParams(randomSeed=12345))
    assert [o.rand_int_uniform() for _ in range(5)] == [1996335345, 1911592690, 679411342, 280691776, 394962642]
    o = O.Oracle(HBTParams(randomSeed=12345))
    for _ in range(5):
        o.rand_int_uniform()
    want = [0.20456027938978735, 0.56772502647397016, 0.59554470273015969, 0.96451452163893492,
            0.65317709638316335]
    assert [o.rand_uniform() for _ in range(5)] == want
