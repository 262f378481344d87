"""Full-size parity under the driver (`-m gpu`): BASELINE.json's configurations at the densities they are quoted at,
against (a) the CPU oracle run inside the test where that takes seconds (a full config-2/3 group, a config-4-shape
group on the 9 x 8 x 41^3 grid) and (b) accumulators the UNMODIFIED reference produced offline for config 5
(tests/golden/make_golden_fullsize.py: groups 0, 99 and 199 of the 200-group run, each at its position in the RNG
stream, and the complete run reduced to 2 000 events).  Integers bit-exact; sums within 1e-10 (hbtio.compare's
conditioning floor); every test prints how many bins only the floor accepted and the largest |delta| per pair."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from hadronic_afterburner_toolkit_b200 import hbtio, synth
from hadronic_afterburner_toolkit_b200.hbt_correlation import HBT_correlation
from hadronic_afterburner_toolkit_b200.params import C3, C4, C5, KAON_MASS, PION_MASS
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu
RTOL = 1e-10
SEED_C5 = 20260005


def _summary(name, rep):
    line = {"case": name}
    for k, v in rep.items():
        line[k] = {"max_rel": v["max_rel"], "floor_bins": v["floor_bins"], "bins": v["bins"], "max_abs_per_pair": v["max_abs_per_pair"]}
    print("FULLSIZE_PARITY " + json.dumps(line))


def test_full_config3_group_against_oracle():
    """One whole group of configs 2/3: 10 events x 1500 pi+ = 15 000 particles, 1.125e8 same-event + 1.35e8
    mixed-event pairs, production kernels (sort, culling, FP32 prefilter, fused launch) against the oracle."""
    batch = synth.make_batches(20260003, 1, 10, PION_MASS)[0]
    o = O.Oracle(C3)
    o.process_batch(batch)
    ref = o.accumulators()
    h = HBT_correlation(C3)
    h.calculate_HBT_correlation_function(batch)
    acc = h.accumulators()
    rep = hbtio.compare(ref, acc, rtol=RTOL, check_stage="cheap")
    assert int(acc.stage[0]) == 15000 * 14999 // 2 and int(acc.stage[6]) == 10 * 1500 * 6 * 1500
    _summary("C3 group, 15000 pi+", rep)
    h.close()


@pytest.mark.parametrize("species,mass", [("pi+", PION_MASS), ("K+", KAON_MASS)])
def test_config4_shape_group_on_the_41_cubed_grid(species, mass):
    """Config 4 at its surveyed histogram size: 9 K_T edges x 8 K_phi bins x 41^3 = 4.96e6 bins per accumulator
    (238 MB for the six: the histograms do not fit the L2), 14 events x 1500 = 21 000 particles."""
    P = C4
    assert P.qnpts == 41 and P.n_KT == 9 and P.n_Kphi == 8
    batch = synth.make_batches(20260004, 1, 14, mass)[0]
    o = O.Oracle(P)
    o.process_batch(batch)
    ref = o.accumulators()
    h = HBT_correlation(P)
    h.calculate_HBT_correlation_function(batch)
    acc = h.accumulators()
    rep = hbtio.compare(ref, acc, rtol=RTOL, check_stage="cheap")
    assert int(acc.stage[0]) == 21000 * 20999 // 2
    _summary(f"C4-shape group at 41^3, 21000 {species}", rep)
    h.close()


def _compare_compact(name, gold, acc):
    """Compact golden (make_golden_fullsize.py): counts of every bin, sums on every 8th bin, per-slab totals."""
    assert np.array_equal(gold["num_count"].astype(np.uint64), np.asarray(acc.num_count).astype(np.uint64)), "num_count differs"
    assert np.array_equal(gold["den_count"].astype(np.uint64), np.asarray(acc.den_count).astype(np.uint64)), "den_count differs"
    assert np.array_equal(gold["npairs_num"], np.asarray(acc.npairs_num, dtype=np.uint64))
    assert np.array_equal(gold["npairs_den"], np.asarray(acc.npairs_den, dtype=np.uint64))
    nb = gold["num_count"].size
    sel = np.arange(int(gold["sel_offset"]), nb, int(gold["sel_stride"]))
    cnt = gold["num_count"][sel].astype(np.float64)
    q3 = acc.qnpts ** 3
    line = {"case": name}
    for k, scale in (("num_cos", 1.0), ("sum_qo", 0.25), ("sum_qs", 0.25), ("sum_ql", 0.25)):
        a, b = gold[k + "_sel"], np.asarray(getattr(acc, k))[sel]
        d = np.abs(a - b)
        tol = RTOL * np.maximum(np.abs(a), 1e-4 * cnt * scale)
        assert np.all(d <= tol), f"{k}: {int(np.sum(d > tol))} sampled bins exceed the tolerance (max {d.max():.3e})"
        nz = cnt > 0
        full = np.asarray(getattr(acc, k))
        slab = np.array([np.sum(full[s * q3:(s + 1) * q3].astype(np.longdouble)) for s in range(nb // q3)], dtype=np.float64)
        assert np.all(np.abs(slab - gold[k + "_slab"]) <= RTOL * gold[k + "_slab_abs"]), f"{k}: slab totals differ"
        line[k] = {"max_rel": float((d[nz] / np.maximum(np.abs(a[nz]), 1e-300)).max()),
                   "floor_bins": int(np.sum(nz & (d > RTOL * np.abs(a)))), "bins": int(nz.sum()),
                   "max_abs_per_pair": float((d[nz] / cnt[nz]).max())}
    print("FULLSIZE_PARITY " + json.dumps(line))


def _golden(name):
    f = os.path.join(GOLDEN, name + ".fs.npz")
    if not os.path.exists(f):
        pytest.skip(f"{name}.fs.npz not generated (tests/golden/make_golden_fullsize.py)")
    return dict(np.load(f))


@pytest.mark.parametrize("g", [0, 99, 199])
def test_config5_sampled_groups_against_the_reference(g):
    """Group g of the 200-group config-5 run (100 events x 1500 pi+, 2.27e10 pairs) at its own position in the RNG
    stream: the draws of the groups before it are replayed, as a rank that does not own them would."""
    gold = _golden(f"c5_group{g:03d}")
    h = HBT_correlation(C5)
    for _ in range(g):
        h.ran_gen.skip_batch(100, 100)
    batch = synth.make_batches(SEED_C5, 1, 100, PION_MASS, first_group=g)[0]
    h.calculate_HBT_correlation_function(batch)
    acc = h.accumulators()
    assert int(acc.stage[0]) == 150000 * 149999 // 2 and int(acc.stage[6]) == 100 * 1500 * 51 * 1500
    _compare_compact(f"C5 group {g}", gold, acc)
    h.close()


def test_config5_complete_run_reduced_to_2000_events():
    """The whole config-5 pipeline on 2 000 events (20 groups, 4.5e11 pairs) on one GPU."""
    gold = _golden("c5_2000ev")
    h = HBT_correlation(C5)
    for g in range(20):
        h.calculate_HBT_correlation_function(synth.make_batches(SEED_C5, 1, 100, PION_MASS, first_group=g)[0])
    acc = h.accumulators()
    assert int(acc.stage[0]) == 20 * (150000 * 149999 // 2)
    _compare_compact("C5 complete, 2000 events", gold, acc)
    h.close()
