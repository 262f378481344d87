"""BalanceFunction row on the GPU, through the C ABI (hbt_bf_*): the same
integers as the oracle in all eight histograms, the same files as the unmodified reference binary."""
import os

import numpy as np
import pytest

from bf_common import CASES, GOLDEN, batches_of, golden_text, read_all_species, species_lists
from hadronic_afterburner_toolkit_b200.balance_function import BalanceFunction
from hadronic_afterburner_toolkit_b200.hbt_correlation import Random
from hadronic_afterburner_toolkit_b200.params import HBTParams
from oracle import bf_oracle, oracle_py

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_files_of_the_reference_binary(name, tmp_path):
    alpha, beta, Bnpts, Brap_max, BpT_min, BpT_max, rap_type, buf = CASES[name]["case"]
    bf = BalanceFunction(alpha, beta, Bnpts, Brap_max, BpT_min, BpT_max, rap_type, path=str(tmp_path),
                         ran_gen=Random(CASES[name]["randomSeed"]))
    events = read_all_species(os.path.join(GOLDEN, "bf_input.gz"))
    for batch in batches_of(events, buf):
        bf.calculate_balance_function(species_lists(batch, alpha, beta))
    bf.output_balance_function()
    for fn, text in golden_text(name).items():
        assert open(tmp_path / fn).read() == text, fn
    bf.close()


def synthetic_lists(seed, nev, n_plus, n_minus):
    rng = np.random.default_rng(seed)

    def species(n, m):
        out = []
        for _ in range(nev):
            k = int(n + rng.integers(-n // 10, n // 10 + 1))
            pT = rng.gamma(2.0, 0.3, k)
            phi = rng.uniform(-np.pi, np.pi, k)
            y = rng.normal(0.0, 1.3, k)
            mT = np.sqrt(m * m + pT * pT)
            pt_, ph_, ry, re = bf_oracle.kinematics(pT * np.cos(phi), pT * np.sin(phi), mT * np.sinh(y), mT * np.cosh(y),
                                                    np.full(k, m))
            out.append({"pT": pt_, "phi": ph_, "rap_y": ry, "rap_eta": re})
        return out
    plus, minus = species(n_plus, 0.13957), species(n_minus, 0.13957)
    return {"a": plus, "abar": minus, "b": minus, "bbar": plus}


@pytest.mark.parametrize("rap_type", [1, 0])
def test_seeded_batches_against_oracle(rap_type):
    """several batches, ragged events (tile edges: 128-particle tiles), identical particles in both lists
    (|Delta y| < 1e-10 rejects self pairs in C_abbar / C_abarb), an empty event"""
    args = (21, 2.0, 0.2, 3.0, rap_type)
    seed = 99
    o = bf_oracle.BFOracle(*args, oracle_py.Oracle(HBTParams(randomSeed=seed)))
    bf = BalanceFunction(211, -211, *args, ran_gen=Random(seed))
    for g in range(3):
        lists = synthetic_lists(1000 + g, 6, 700, 650)
        if g == 1:
            for k in lists:  # an event without any particle of either species
                lists[k][2] = {f: v[:0] for f, v in lists[k][2].items()}
        o.calculate_balance_function(lists)
        bf.calculate_balance_function(lists)
    got, want = bf.histograms(), o.histograms()
    assert want.sum() > 1e6
    assert np.array_equal(got, want)
    assert (bf.N_b, bf.N_bbar) == (o.N_b, o.N_bbar)
    assert bf.timers()["pairs"] == o.pairs
    bf.close()


def test_float_decision_bands_hold_next_to_bin_edges():
    """The kernel bins a pair in binary32 when it is farther than a proven error band from every bin edge and takes
    the reference's binary64 chain otherwise.  Pairs are placed 1e-9 ... 1e-3 bin widths on either side of rapidity
    and azimuth bin edges (and of the two ends of the rapidity range, and at |Delta y| around the 1e-10 self-pair
    test), at small and large |y| / |phi|: the histograms must equal the oracle's integers."""
    Bnpts, Brap_max = 21, 2.0
    args = (Bnpts, Brap_max, 0.2, 3.0, 1)
    drap = 2. * Brap_max / (Bnpts - 1)
    rap_min = -Brap_max - 0.5 * drap
    dphi = 2. * np.pi / 20
    rng = np.random.default_rng(7)
    eps = np.concatenate([[0.0], 10.0 ** np.arange(-9.0, -2.5, 0.5)])
    eps = np.concatenate([eps, -eps])
    nev, n = 4, 640
    def lists_for(rot_free):
        ev_a, ev_b = [], []
        for e in range(nev):
            yb = rng.uniform(-3.0, 3.0, n) * (30.0 if e == 3 else 1.0)   # e == 3: large |y| (wide bands)
            pb = rng.uniform(-np.pi, np.pi, n)
            k = rng.integers(0, Bnpts + 1, n)            # rapidity edges 0 .. Bnpts (both ends of the range included)
            m = rng.integers(-14, 15, n)                 # azimuth edges
            ya = yb + rap_min + (k + rng.choice(eps, n)) * drap
            pa = pb - np.pi / 2 + (m + rng.choice(eps, n)) * dphi
            pa = np.clip(pa, -np.pi, np.pi)
            if e == 2:                                   # |Delta y| around the self-pair threshold
                ya = yb + rng.choice([0.0, 5e-11, 1.5e-10, -5e-11, -1.5e-10, 1e-9, 1e-7], n)
            pT = np.full(n, 1.0)
            ev_a.append({"pT": pT, "phi": pa, "rap_y": ya, "rap_eta": ya})
            ev_b.append({"pT": pT, "phi": pb, "rap_y": yb, "rap_eta": yb})
        return {"a": ev_a, "abar": ev_b, "b": ev_b, "bbar": ev_a}
    seed = 3
    o = bf_oracle.BFOracle(*args, oracle_py.Oracle(HBTParams(randomSeed=seed)))
    bf = BalanceFunction(211, -211, *args, ran_gen=Random(seed))
    for _ in range(2):
        lists = lists_for(True)
        o.calculate_balance_function(lists)
        bf.calculate_balance_function(lists)
    got, want = bf.histograms(), o.histograms()
    assert want.sum() > 1e6
    assert np.array_equal(got, want)
    bf.close()


def test_bad_arguments():
    import ctypes

    from hadronic_afterburner_toolkit_b200 import capi
    L = capi.lib()
    h = ctypes.c_void_p()
    assert L.hbt_bf_create(1, 2.0, 0, ctypes.byref(h)) == -1       # Bnpts < 2
    assert L.hbt_bf_create(21, 2.0, 99, ctypes.byref(h)) == -1     # no such device
    bf = BalanceFunction(211, -211, 21, 2.0, 0.2, 3.0, 1)
    lists = synthetic_lists(5, 2, 50, 50)
    a = bf._flat(lists["a"])
    with pytest.raises(capi.HBTError):
        bf._call(8, a, a, np.zeros(2, dtype=np.int32), np.zeros(2))    # histogram index out of range
    with pytest.raises(capi.HBTError):
        bf._call(0, a, a, np.array([0, 5], dtype=np.int32), np.zeros(2))  # partner event out of range
    bf.close()
