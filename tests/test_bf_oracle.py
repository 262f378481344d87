"""Pins the BalanceFunction oracle (oracle/bf_oracle.py) and the output writer against the files the
unmodified reference binary wrote (tests/golden/bf_*.dat, made by tests/golden/make_golden_bf.py):
same text, character for character.  CPU only."""
import os

import numpy as np
import pytest

from bf_common import CASES, GOLDEN, batches_of, golden_text, read_all_species, species_lists
from hadronic_afterburner_toolkit_b200 import balance_function
from hadronic_afterburner_toolkit_b200.params import HBTParams
from oracle import bf_oracle, oracle_py


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_the_reference_files(name, tmp_path):
    alpha, beta, Bnpts, Brap_max, BpT_min, BpT_max, rap_type, buf = CASES[name]["case"]
    rng = oracle_py.Oracle(HBTParams(randomSeed=CASES[name]["randomSeed"]))  # the reference's mt19937 stream
    o = bf_oracle.BFOracle(Bnpts, Brap_max, BpT_min, BpT_max, rap_type, rng)
    events = read_all_species(os.path.join(GOLDEN, "bf_input.gz"))
    for batch in batches_of(events, buf):
        o.calculate_balance_function(species_lists(batch, alpha, beta))
    assert o.pairs > 50000
    balance_function.write_outputs(o.histograms().astype(np.float64), alpha, beta, Bnpts, o.Brap_min, o.drap, str(tmp_path))
    want = golden_text(name)
    for fn, text in want.items():
        assert open(tmp_path / fn).read() == text, fn


def test_cxx_formatting():
    f = balance_function.cxx_sci
    assert f(-2.0, 18) == "   -2.00000000e+00"
    assert f(6497.0) == "6.49700000e+03"
    assert f(np.float64(0.0) / np.float64(1.0)) == "0.00000000e+00"
    with np.errstate(invalid="ignore"):
        assert f(np.float64(0.0) / np.float64(0.0)) == "-nan"  # x86 default NaN, as the reference prints it
    assert f(float("inf")) == "inf"
