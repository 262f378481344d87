"""Parity tests proper: the CUDA path, called through the C ABI (libhbt_b200.so), against
(1) the golden vectors produced by the unmodified reference and (2) the CPU oracle on the
same seeded inputs.  Tolerances (north_star): every integer accumulator bit-exact; floating
sums within 1e-10 relative (see hbtio.compare for the conditioning floor)."""
import ctypes
import gzip
import os

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_NAMES, load_golden
from hadronic_afterburner_toolkit_b200 import capi, hbtio, synth
from hadronic_afterburner_toolkit_b200.hbt_correlation import HBT_correlation, Random
from hadronic_afterburner_toolkit_b200.params import C3, C4, HBTParams, KAON_MASS
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu
RTOL = 1e-10



def run_product(P, batches, do_mixed=True, stats=False, coalesce=None):
    h = HBT_correlation(P, stage_counters=stats, coalesce=coalesce)
    for b in batches:
        h.calculate_HBT_correlation_function(b, do_mixed=do_mixed)
    acc = h.accumulators()
    return h, acc


def run_oracle(P, batches, do_mixed=True):
    o = O.Oracle(P)
    for b in batches:
        o.process_batch(b, do_mixed=do_mixed)
    return o.accumulators()


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_golden_reference_vectors(name):
    """Includes the cases in which needed_number_of_pairs is reached (unit_iss_gz_cap,
    synth_c3_cap): the ordered cap must reproduce the reference's first-come-first-served
    acceptance exactly."""
    P, batches, ref, meta = load_golden(name)
    h, acc = run_product(P, batches)
    hbtio.compare(ref, acc, rtol=RTOL)
    assert acc.psi_ref == ref.psi_ref  # host glibc path: bit-identical
    assert h.pairs_same == meta["pairs_same"] == int(acc.stage[0])


CAPPED = {
    # needed_number_of_pairs small enough to close slabs at different batches / phases
    "cap_az0": (HBTParams(qnpts=21, needed_number_of_pairs=4000.0), 4, 4, 400),
    "cap_az0_tiny": (HBTParams(qnpts=11, needed_number_of_pairs=7.0), 3, 3, 150),
    "cap_zero": (HBTParams(qnpts=11, needed_number_of_pairs=0.0), 2, 3, 150),
    "cap_az1": (C4.with_(qnpts=11, n_KT=4, n_Kphi=4, needed_number_of_pairs=900.0), 3, 4, 400),
    "cap_first_batch_only": (HBTParams(qnpts=21, needed_number_of_pairs=50000.0), 3, 4, 500),
    # q_inv mode: the 1-D histograms stop at 50 x needed per K_T bin, independently of the 3-D slabs
    "cap_qinv_first_batch": (HBTParams(qnpts=21, invariant_radius_flag=1, needed_number_of_pairs=20.0), 3, 4, 400),
    "cap_qinv_late": (HBTParams(qnpts=21, invariant_radius_flag=1, needed_number_of_pairs=700.0), 4, 4, 400),
    "cap_qinv_zero": (HBTParams(qnpts=11, invariant_radius_flag=1, needed_number_of_pairs=0.0), 2, 3, 150),
}


@pytest.mark.parametrize("name", sorted(CAPPED))
def test_ordered_cap_against_oracle(name):
    P, ngrp, nev, mult = CAPPED[name]
    batches = synth.make_batches(20260007, ngrp, nev, multiplicity=mult)
    ref = run_oracle(P, batches)
    h, acc = run_product(P, batches)
    hbtio.compare(ref, acc, rtol=RTOL, check_stage="cheap")
    lim = int(P.needed_number_of_pairs) + 1
    assert int(acc.npairs_num.max()) <= lim and int(acc.npairs_den.max()) <= lim
    if name != "cap_first_batch_only":
        assert int(acc.npairs_num.max()) == lim  # the cap really engaged
    if P.invariant_radius_flag == 1:
        qlim = 50 * int(P.needed_number_of_pairs)
        assert int(acc.npairs_num_qinv.max()) <= qlim and int(acc.npairs_den_qinv.max()) <= qlim
        assert int(acc.npairs_num_qinv.max()) == qlim and int(acc.npairs_den_qinv.max()) == qlim  # engaged in both loops


def test_full_size_group_production_equals_literal_kernels():
    """Size-independent property at BASELINE size (one C3 group, 15 000 pi+, 2.5e8 pairs): the
    production path (sort + culling + guarded fast path) and the literal v1 kernels give the
    same integers in every bin, and the sums agree to 1e-10."""
    batch = synth.make_batches(20260003, 1, 10)[0]
    res = []
    for kernel in (1, 2):
        h = HBT_correlation(C3, kernel=kernel)
        h.calculate_HBT_correlation_function(batch)
        res.append(h.accumulators())
        h.close()
    hbtio.compare(res[0], res[1], rtol=RTOL, check_stage="cheap")
    assert int(res[1].stage[0]) == 15000 * 14999 // 2
    # splitting invariance: pairs(A u B) = pairs(A) + pairs(B) + cross pairs, the cross term being
    # a mixed-event pass with the identity rotation (same kernels, different tiling)
    A, B = batch.same[:5], batch.same[5:]
    hall = HBT_correlation(C3); hall.set_particle_list(batch); hall.combine_and_bin_particle_pairs(list(range(10)))
    hA = HBT_correlation(C3); hA.set_particle_list(hbtio.Batch(A)); hA.combine_and_bin_particle_pairs(list(range(5)))
    hB = HBT_correlation(C3); hB.set_particle_list(hbtio.Batch(B)); hB.combine_and_bin_particle_pairs(list(range(5)))
    hX = HBT_correlation(C3)
    pa = np.ascontiguousarray(np.concatenate(A)); pb = np.ascontiguousarray(np.concatenate(B))
    offa = np.array([0, len(pa)], dtype=np.int64); offb = np.array([0, len(pb)], dtype=np.int64)
    ids = np.zeros((1, 1), dtype=np.int32); cs = np.array([[[1.0, 0.0]]])
    from hadronic_afterburner_toolkit_b200.hbt_correlation import _check
    _check(hX._h, hX._L.hbt_accumulate_mixed(hX._h, pa.ctypes.data, offa.ctypes.data, 1, pb.ctypes.data, offb.ctypes.data, 1,
                                             ids.ctypes.data, cs.ctypes.data, 1, 0.0))
    tot = hall.accumulators().num_count
    parts = hA.accumulators().num_count + hB.accumulators().num_count + hX.accumulators().den_count
    assert np.array_equal(tot, parts)


def test_batch_beyond_the_unit_encoding_is_refused():
    """The same-event loop pairs every particle of the batch with every other one; a batch of more
    than ~4.19e6 particles (8.8e12 pairs, 2^31 work units) does not fit the unit encoding and is
    refused with HBT_ERR_INVALID before any kernel runs — never truncated."""
    from hadronic_afterburner_toolkit_b200.capi import HBTError
    batch = synth.make_batches(20260011, 1, 21500, multiplicity=200)[0]
    assert sum(len(e) for e in batch.same) > (1 << 22)
    h = HBT_correlation(C3)
    h.set_particle_list(batch)
    with pytest.raises(HBTError, match="batch too large"):
        h.combine_and_bin_particle_pairs(list(range(len(batch.same))))
    assert int(h.accumulators().num_count.sum()) == 0
    h.close()


def test_fused_and_separate_kernels_agree_and_device_resident_batch():
    """A whole batch runs as one kernel over the interleaved same-event and mixed-event units
    (HBT_OPT_FUSE, default) or as one kernel per loop: same integers, same oracle.  The
    device-resident whole-batch entry point (hbt_accumulate_batch_dev) takes the same path."""
    import torch
    from hadronic_afterburner_toolkit_b200.hbt_correlation import gather_rapidity

    P = C3.with_(qnpts=21)
    batches = synth.make_batches(20260012, 2, 5, multiplicity=700)
    ref = run_oracle(P, batches)
    res = []
    for fuse in (True, False):
        h = HBT_correlation(P, fuse=fuse, coalesce=False)  # one launch (fused) / two launches per batch
        for b in batches:
            h.calculate_HBT_correlation_function(b)
        res.append(h.accumulators())
        t = h.timers()
        assert t["same_launches"] == t["mixed_launches"] == len(batches)
        h.close()
    hbtio.compare(ref, res[0], rtol=RTOL, check_stage="cheap")
    hbtio.compare(res[0], res[1], rtol=RTOL, check_stage="cheap")
    h = HBT_correlation(P)  # default: the two small batches go out as ONE fused launch
    for b in batches:
        h.calculate_HBT_correlation_function(b)
    hbtio.compare(ref, h.accumulators(), rtol=RTOL, check_stage="cheap")
    assert h.timers()["same_launches"] == 1
    h.close()
    # device-resident list, same plan as the oracle's draws (the class shares the RNG stream)
    h = HBT_correlation(P)
    keep = []
    for b in batches:
        nev = len(b.same)
        cut = [gather_rapidity(P, ev) for ev in b.same]
        flat = np.ascontiguousarray(np.concatenate(cut))
        off = np.zeros(nev + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(ev) for ev in cut])
        ids, cs = h.ran_gen.mixed_plan(nev, nev)
        d = torch.from_numpy(flat).cuda()
        keep.append((d, off, ids, cs))
        rc = h._L.hbt_accumulate_batch_dev(h._h, d.data_ptr(), off.ctypes.data, nev, ids.ctypes.data, cs.ctypes.data,
                                           ids.shape[1], 0.0)
        assert rc == 0, h._L.hbt_last_error(h._h)
    hbtio.compare(ref, h.accumulators(), rtol=RTOL, check_stage="cheap")
    h.close()


@pytest.mark.parametrize("same_only", [False, True])
def test_batches_in_flight_on_two_lanes(same_only):
    """Production batches alternate between two compute streams (HBT_OPT_LANES, default 2), each
    with its own sort / cull scratch, so consecutive batches overlap on the device.  Batches
    commute (atomic adds), so one lane and two lanes must give the oracle's integers; the batch
    sizes grow and shrink so that a lane's scratch is reallocated while the other lane is busy,
    and more batches than staging slots are submitted without a synchronize in between.  The
    launch timers count overlapping launches once: their sum cannot exceed the stopwatch."""
    P = C3.with_(qnpts=21)
    batches = []
    for g, (nev, mult) in enumerate([(3, 300), (4, 900), (2, 200), (5, 1200), (3, 500), (6, 1500), (2, 100), (4, 800), (3, 400)]):
        batches += synth.make_batches(20260020 + g, 1, nev, multiplicity=mult)
    ref = run_oracle(P, batches, do_mixed=not same_only)
    res = []
    for lanes in (2, 1):
        h = HBT_correlation(P, lanes=lanes, coalesce=False)  # one launch per batch: the lanes take them in turn
        assert h._L.hbt_timer_start(h._h) == 0
        for b in batches:
            if same_only:
                h.set_particle_list(b)
                h.combine_and_bin_particle_pairs(list(range(len(b.same))))
            else:
                h.calculate_HBT_correlation_function(b)
        ms = ctypes.c_double()
        assert h._L.hbt_timer_stop(h._h, ctypes.byref(ms)) == 0
        res.append(h.accumulators())
        t = h.timers()
        assert t["same_launches"] == len(batches)
        assert 0.0 < t["same_ms"] + t["mixed_ms"] <= ms.value * 1.02 + 0.05
        h.close()
    hbtio.compare(ref, res[0], rtol=RTOL, check_stage="cheap")
    hbtio.compare(ref, res[1], rtol=RTOL, check_stage="cheap")
    h = HBT_correlation(P)
    assert h._L.hbt_set_option(h._h, 4, 3) == -1  # at most two lanes
    h.close()


@pytest.mark.parametrize("case", ["c3", "c4_az", "ragged", "real_mixed", "outliers", "kt_from_0"])
def test_mixed_loops_on_the_pt_sorted_copy(case):
    """HBT_OPT_PTSORT: the production mixed-event loops read a copy in which every event is sorted by
    pT and skip the list-2 particles whose pT is farther than the q_out window from the sub-tile's pT
    range (|q_out| >= |pT_i - pT_j|).  Forced on (2: the default sorts only batches with >= 5e8
    mixed-event pairs) and off (0), the accumulators are the oracle's."""
    rng = np.random.default_rng(11)
    if case == "real_mixed":
        P, batches, ref, _ = load_golden("unit_iss_gz_realmixed")
    else:
        P = {"c3": C3.with_(qnpts=21), "c4_az": C4.with_(qnpts=11, n_KT=4, n_Kphi=4), "ragged": C3.with_(qnpts=21),
             "outliers": HBTParams(qnpts=21), "kt_from_0": HBTParams(qnpts=21, KT_min=0.0, KT_max=1.0, n_KT=5)}[case]
        batches = synth.make_batches(20260040, 2, 6, multiplicity=700)
        if case == "ragged":  # events of very different sizes, an empty one, a single particle
            for b in batches:
                b.same[1] = b.same[1][:0]
                b.same[2] = b.same[2][:1]
                b.same[3] = b.same[3][:137]
        if case == "outliers":
            for b in batches:
                for ev in b.same:
                    k = rng.choice(len(ev), size=5, replace=False)
                    ev[k, 0:3] *= 300.0
                    ev[:, 3] = np.sqrt(0.138 ** 2 + (ev[:, 0:3] ** 2).sum(axis=1))
        ref = run_oracle(P, batches)
    for ptsort in (2, 0):
        h = HBT_correlation(P, ptsort=ptsort)
        for b in batches:
            h.calculate_HBT_correlation_function(b)
        hbtio.compare(ref, h.accumulators(), rtol=RTOL, check_stage="cheap")
        assert h._L.hbt_set_option(h._h, 5, 3) == -1
        h.close()


def _craft_edge_pairs(P, rng_plan, rng, n, which):
    """Two events of n pi+ such that the mixed-event pair (event 0 particle i, event 1 particle i
    rotated by the first partner angle of event 0) has q_out (which=0) or q_long (which=1) at a
    log-uniform distance 1e-7 ... 2e-3 bin widths from a bin edge, on either side; the other two
    components sit at a bin centre and K_T inside the cut."""
    ids, cs = rng_plan.mixed_plan(2, 2)
    assert ids[0, 0] == 1
    c, s = cs[0, 0]
    m = 0.138
    dq = (P.q_max - P.q_min) / (P.qnpts - 1)
    q_base = P.q_min - dq / 2
    pt = rng.uniform(0.22, 0.5, n)
    ph = rng.uniform(0, 2 * np.pi, n)
    y = rng.uniform(-0.4, 0.4, n)
    ax, ay = pt * np.cos(ph), pt * np.sin(ph)
    mTa = np.sqrt(m * m + ax * ax + ay * ay)
    az, aE = mTa * np.sinh(y), mTa * np.cosh(y)
    if which == 2:
        # K_T = |a_T + b'_T| / 2 at a relative distance 1e-9 ... 1e-4 from a K_T bin edge (or from the cut):
        # b'_T = a_T + 0.03 a_T/|a_T| (collinear: q_side = 0, q_out = -0.03, both bin centres), same rapidity
        dKT = (P.KT_max - P.KT_min) / (P.n_KT - 1)
        kedge = P.KT_min + dKT * rng.integers(0, P.n_KT, n)
        kt = kedge * (1.0 + rng.choice([-1.0, 1.0], n) * 10.0 ** rng.uniform(-9, -4, n))
        pt = kt - 0.015
        ax, ay = pt * np.cos(ph), pt * np.sin(ph)
        mTa = np.sqrt(m * m + pt * pt)
        az, aE = mTa * np.sinh(y), mTa * np.cosh(y)
        bx, by = (pt + 0.03) * np.cos(ph), (pt + 0.03) * np.sin(ph)
        mTb = np.sqrt(m * m + (pt + 0.03) ** 2)
        bz, bE = mTb * np.sinh(y), mTb * np.cosh(y)
        ux, uy = c * bx + s * by, -s * bx + c * by
        pos = lambda k: rng.normal(0, 4, k)
        ev0 = np.column_stack([ax, ay, az, aE, pos(n), pos(n), pos(n), np.abs(pos(n)) + 5])
        ev1 = np.column_stack([ux, uy, bz, bE, pos(n), pos(n), pos(n), np.abs(pos(n)) + 5])
        return hbtio.Batch([ev0, ev1]), kt
    edge = rng.integers(P.qnpts // 2 - 6, P.qnpts // 2 + 7, n).astype(np.float64)
    target = edge + rng.choice([-1.0, 1.0], n) * 10.0 ** rng.uniform(-7, np.log10(2e-3), n)

    def forward(lam):  # b' (already rotated) as a function of the scan parameter; returns (b', u)
        if which == 0:
            bx, by = ax - lam * np.cos(ph), ay - lam * np.sin(ph)
            mTb = np.sqrt(m * m + bx * bx + by * by)
            bz, bE = mTb * np.sinh(y), mTb * np.cosh(y)
            Kx, Ky = 0.5 * (ax + bx), 0.5 * (ay + by)
            Kp = np.sqrt(Kx * Kx + Ky * Ky)
            q = (ax - bx) * (Kx / Kp) + (ay - by) * (Ky / Kp)
        else:
            bx, by = ax.copy(), ay.copy()
            bz, bE = mTa * np.sinh(y - lam), mTa * np.cosh(y - lam)
            Kz, KE = 0.5 * (az + bz), 0.5 * (aE + bE)
            Mt = np.sqrt(KE * KE - Kz * Kz)
            q = (KE * (az - bz) - Kz * (aE - bE)) / Mt
        return (bx, by, bz, bE), (q - q_base) / dq

    lo, hi = np.full(n, -0.3), np.full(n, 0.3)  # u is increasing in lam on this range
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        _, u = forward(mid)
        up = u < target
        lo, hi = np.where(up, mid, lo), np.where(up, hi, mid)
    (bx, by, bz, bE), u = forward(0.5 * (lo + hi))
    assert np.max(np.abs(u - target)) < 1e-9
    # undo the rotation the mixed-event loop applies to the partner (src :522-523)
    ux, uy = c * bx + s * by, -s * bx + c * by
    pos = lambda k: rng.normal(0, 4, k)
    ev0 = np.column_stack([ax, ay, az, aE, pos(n), pos(n), pos(n), np.abs(pos(n)) + 5])
    ev1 = np.column_stack([ux, uy, bz, bE, pos(n), pos(n), pos(n), np.abs(pos(n)) + 5])
    return hbtio.Batch([ev0, ev1]), target


@pytest.mark.parametrize("which", [0, 1, 2])
def test_fp32_decision_bands_hold_next_to_bin_edges(which):
    """Mixed-event survivors are binned in FP32 when every component is farther from every edge
    than a bound on the float evaluation error (v3_mixed_f32), in FP64 otherwise.  Thousands of
    pairs placed 1e-7 ... 2e-3 bin widths from an edge of q_out / q_long, on both sides — inside the
    float band, just outside it, and well outside — must all land in the reference's bins; likewise
    pairs whose K_T is 1e-9 ... 1e-4 (relative) from a K_T bin edge or from the K_T cut (which=2)."""
    P = C3.with_(qnpts=41)
    rng = np.random.default_rng(77 + which)
    plan = Random(P.randomSeed)
    batches, targets = [], []
    for _ in range(3):
        b, t = _craft_edge_pairs(P, plan, rng, 700, which)
        batches.append(b)
        targets.append(t)
    ref = run_oracle(P, batches)
    for env in ("1", "0"):  # with the float path and with every survivor through the FP64 path
        os.environ["HBT_B200_F32MIX"] = env
        try:
            _, acc = run_product(P, batches)
        finally:
            del os.environ["HBT_B200_F32MIX"]
        hbtio.compare(ref, acc, rtol=RTOL, check_stage="cheap")
    # the crafted pairs are really there: the central q_side (and q_long / q_out) row holds them
    assert int(ref.den_count.sum()) > 3 * 700


@pytest.mark.parametrize("scale,ktmin", [(30.0, 0.15), (3000.0, 0.15), (1.0, 0.0), (1e-3, 0.0)])
def test_prefilter_margins_with_outliers(scale, ktmin):
    """The float prefilter's margins are derived from the largest pT^2 of the two tiles.  Momentum
    outliers (a few particles scaled up by 30x / 3000x), a K_T range that starts at 0 and a sample
    of tiny momenta must all still give the reference's integers (the prefilter may only get less
    selective, never drop an accepted pair)."""
    P = HBTParams(qnpts=21, KT_min=ktmin, KT_max=0.55 if ktmin else 1.0, n_KT=5)
    batches = synth.make_batches(20260009, 1, 5, multiplicity=600)
    rng = np.random.default_rng(3)
    for b in batches:
        for ev in b.same:
            if scale >= 1.0:
                k = rng.choice(len(ev), size=6, replace=False)
                ev[k, 0:3] *= scale
            else:
                ev[:, 0:3] *= scale  # everything tiny: K_T << dKT, q << delta_q
            ev[:, 3] = np.sqrt(0.138 ** 2 + (ev[:, 0:3] ** 2).sum(axis=1))
    ref = run_oracle(P, batches)
    for stats in (False, True):
        _, acc = run_product(P, batches, stats=stats)
        hbtio.compare(ref, acc, rtol=RTOL, check_stage=True if stats else "cheap", q_scale=0.25)


def test_options_and_literal_kernels_agree():
    P = HBTParams(qnpts=21)
    batches = synth.make_batches(20260010, 2, 4, multiplicity=500)
    ref = run_oracle(P, batches)
    h = HBT_correlation(P, kernel=1)  # HBT_OPT_KERNEL = literal v1 kernels
    for b in batches:
        h.calculate_HBT_correlation_function(b)
    hbtio.compare(ref, h.accumulators(), rtol=RTOL, check_stage=True)
    assert h._L.hbt_set_option(h._h, 99, 0) == -1  # unknown option


def test_device_resident_entry_points_refuse_near_the_cap():
    import ctypes

    P = HBTParams(qnpts=11, needed_number_of_pairs=10.0)
    h = HBT_correlation(P)
    x = np.zeros((64, 8))
    rc = h._L.hbt_accumulate_same_dev(h._h, ctypes.c_void_p(1), 64, 0.0)
    assert rc == -4  # HBT_ERR_CAP: only the host-buffer calls replay the cap in order


SEEDED = {
    # name: (params, groups, events/group, multiplicity, mass)
    "c2_same_only": (HBTParams(), 1, 4, 1500, 0.138),
    "c3_same_mixed": (C3, 2, 4, 1000, 0.138),
    "c4_az_pions": (C4.with_(qnpts=21), 1, 6, 800, 0.138),
    "c4_az_kaons": (C4.with_(qnpts=21), 1, 6, 800, KAON_MASS),
    "qinv": (HBTParams(invariant_radius_flag=1, qnpts=31), 1, 4, 800, 0.138),
    # q_inv mode on the tuned kernels (v3_qinv_pair): windows that start below / above zero, kaons, with K_phi slabs
    "qinv_positive_window": (HBTParams(invariant_radius_flag=1, qnpts=13, q_min=0.02, q_max=0.14), 1, 4, 600, 0.138),
    "qinv_wide_kaons": (HBTParams(invariant_radius_flag=1, qnpts=41, q_min=-0.4, q_max=0.4, KT_min=0.1, KT_max=1.2, n_KT=7), 1, 4, 700, KAON_MASS),
    "qinv_az": (C4.with_(qnpts=11, n_KT=4, n_Kphi=4, invariant_radius_flag=1), 1, 5, 500, 0.138),
    "noboost": (HBTParams(long_comoving_boost=0), 1, 3, 800, 0.138),
    "ragged_tiles": (HBTParams(qnpts=21), 2, 3, 257, 0.138),   # tile edges: 257 = 2*128 + 1
    "single_event": (HBTParams(qnpts=21), 1, 1, 700, 0.138),   # mixed_nev == 1: self pairing
    # one-sided windows run on the tuned kernels too (the prefilter tests against max(|q_lo|, |q_hi|))
    "asym_window": (HBTParams(qnpts=16, q_min=-0.05, q_max=0.25), 1, 4, 600, 0.138),
    "positive_window": (HBTParams(qnpts=13, q_min=0.02, q_max=0.14), 1, 4, 600, 0.138),
    "negative_window_noboost": (HBTParams(qnpts=11, q_min=-0.2, q_max=-0.01, long_comoving_boost=0), 1, 4, 600, 0.138),
    "kt_from_zero": (HBTParams(qnpts=21, KT_min=0.0, KT_max=1.0, n_KT=6), 1, 3, 600, 0.138),
}


@pytest.mark.parametrize("mode", ["production", "one_launch_per_batch", "instrumented"])
@pytest.mark.parametrize("name", sorted(SEEDED))
def test_seeded_against_oracle(name, mode):
    """production: Morton-sorted same-event list with tile culling, small batches launched together;
    one_launch_per_batch: the same without coalescing; instrumented: every pair through the prefilter, all six
    stage populations exact."""
    stats = mode == "instrumented"
    P, ngrp, nev, mult, mass = SEEDED[name]
    batches = synth.make_batches(20260002, ngrp, nev, mass=mass, multiplicity=mult)
    do_mixed = name != "c2_same_only"
    ref = run_oracle(P, batches, do_mixed)
    h, acc = run_product(P, batches, do_mixed, stats=stats, coalesce=False if mode == "one_launch_per_batch" else None)
    rep = hbtio.compare(ref, acc, rtol=RTOL, check_stage=True if stats else "cheap")
    assert int(acc.stage[0]) == h.pairs_same and int(acc.stage[6]) == h.pairs_mixed
    print(name, rep, "deferred", h.deferred_pairs())


@pytest.mark.parametrize("coalesce", [False, None], ids=["one_launch_per_batch", "small_batches_together"])
@pytest.mark.parametrize("name", ["c3_same_mixed", "c4_az_pions", "qinv"])
def test_page_locked_caller_buffers_are_uploaded_directly(name, coalesce):
    """hbt_accumulate_batch hands page-locked caller buffers to the DMA engine without its staging copy (and returns
    once they are free again); separate mixed-event lists included.  Same integers as the oracle."""
    P, ngrp, nev, mult, mass = SEEDED[name]
    batches = synth.make_batches(20260002, ngrp, nev, mass=mass, multiplicity=mult)
    other = synth.make_batches(20260009, ngrp, nev + 1, mass=mass, multiplicity=mult - 37)
    batches = batches + [hbtio.Batch(x.same, y.same) for x, y in zip(batches, other)]  # read_in_real_mixed_events = 1
    ref = run_oracle(P, batches, True)
    if coalesce is None:  # more, smaller batches: several of them per launch, each uploaded as it arrives
        more = synth.make_batches(20260010, 6, 3, mass=mass, multiplicity=300)
        ref = None
        batches = more[:3] + batches + more[3:]
        ref = run_oracle(P, batches, True)
    h = HBT_correlation(P, coalesce=coalesce)
    h.pin_host = True
    for b in batches:
        h.calculate_HBT_correlation_function(b)
    acc = h.accumulators()
    hbtio.compare(ref, acc, rtol=RTOL, check_stage="cheap")
    h.close()


def test_qinv_edges_decided_like_the_reference():
    """q_inv mode on the tuned kernels decides the q_inv window and bin by comparing s = -(q.q), evaluated with the
    reference's operation order, with thresholds found on the host.  Partners are placed so that q_inv lands
    1e-13 ... 1e-4 (relative) on either side of every bin edge and of the window's upper end, in both loops; all
    integers must be the oracle's."""
    P = HBTParams(invariant_radius_flag=1, qnpts=21)
    dq = (P.q_max - P.q_min) / (P.qnpts - 1)
    q_base = P.q_min - dq / 2
    rng = np.random.default_rng(11)
    m = 0.138
    evs = []
    for e in range(4):
        n = 500
        base = synth.make_group(4242, e, 1, multiplicity=n // 2)[0]
        part = base.copy()
        k = rng.integers(P.qnpts // 2, P.qnpts + 1, n // 2)          # edges at q_inv >= 0 up to the window's end
        eps = rng.choice(np.concatenate([[0.0], 10.0 ** np.arange(-13.0, -3.5, 0.5)]), n // 2) * rng.choice([-1.0, 1.0], n // 2)
        target = np.maximum((q_base + k * dq) * (1.0 + eps), 1e-6)
        # partner = the particle with its momentum scaled so that q_inv hits the target (bisection in the scale)
        p = base[:, 0:3]
        lo, hi = np.ones(n // 2), np.full(n // 2, 3.0)
        for _ in range(200):
            mid = 0.5 * (lo + hi)
            pm = p * mid[:, None]
            Em = np.sqrt(m * m + (pm * pm).sum(1))
            dqv = p - pm
            qinv = np.sqrt(np.maximum((dqv * dqv).sum(1) - (base[:, 3] - Em) ** 2, 0.0))
            lo = np.where(qinv < target, mid, lo)
            hi = np.where(qinv < target, hi, mid)
        part[:, 0:3] = p * hi[:, None]
        part[:, 3] = np.sqrt(m * m + (part[:, 0:3] ** 2).sum(1))
        evs.append(np.concatenate([base, part]))
    batches = [hbtio.Batch(evs)]
    ref = run_oracle(P, batches, True)
    assert int(ref.qinv_count.sum()) > 5000
    for coalesce in (None,):
        h, acc = run_product(P, batches, True, coalesce=coalesce)
        hbtio.compare(ref, acc, rtol=RTOL, check_stage="cheap")
        h.close()


@pytest.mark.parametrize("resident", [False, True], ids=["host_buffers", "device_resident"])
@pytest.mark.parametrize("do_mixed", [True, False], ids=["same_mixed", "same_only"])
def test_small_batches_launched_together(do_mixed, resident):
    """HBT_OPT_COALESCE: 40 small batches of ragged sizes (event multiplicities that are no multiples of the 64-particle
    work units, one-particle events, a batch of a single event) go out as two launches of up to 32 batches: same
    integers as the oracle in every bin, batch by batch in the reference's RNG order."""
    import torch

    P = HBTParams(qnpts=21)
    rng = np.random.default_rng(5)
    batches = []
    for k in range(40):
        nev = int(rng.integers(1, 6))
        evs = [synth.make_group(900 + k, e, 1, multiplicity=int(rng.integers(1, 330)))[0] for e in range(nev)]
        batches.append(hbtio.Batch(evs))
    ref = run_oracle(P, batches, do_mixed)
    h = HBT_correlation(P)
    keep = []
    for b in batches:
        if not resident:
            h.calculate_HBT_correlation_function(b, do_mixed=do_mixed)
            continue
        flat, off = b.flat("same")
        d = torch.from_numpy(np.ascontiguousarray(flat)).cuda()
        keep.append(d)  # device buffers stay valid until the context is synchronised
        nev = len(b.same)
        if do_mixed:
            ids, cs = h.ran_gen.mixed_plan(nev, nev)
            from hadronic_afterburner_toolkit_b200.hbt_correlation import _check
            _check(h._h, h._L.hbt_accumulate_batch_dev(h._h, d.data_ptr(), off.ctypes.data, nev, ids.ctypes.data, cs.ctypes.data,
                                                       ids.shape[1], 0.0))
        else:
            from hadronic_afterburner_toolkit_b200.hbt_correlation import _check
            _check(h._h, h._L.hbt_accumulate_same_dev(h._h, d.data_ptr(), int(off[-1]), 0.0))
    acc = h.accumulators()
    hbtio.compare(ref, acc, rtol=RTOL)
    t = h.timers()
    # 40 batches in 2 launches when device resident, a launch per 8 host batches (one more if a batch did not qualify)
    assert t["same_launches"] <= (3 if resident else 7), t
    h.close()


def test_empty_and_tiny_batches():
    P = HBTParams(qnpts=11)
    batches = [hbtio.Batch([]),                        # the reader's trailing empty batch
               hbtio.Batch([np.zeros((0, 8))]),        # one event, no particle of the species
               synth.make_batches(1, 1, 1, multiplicity=1)[0],  # one particle: no same-event pair, but
                                                                # mixed_nev == 1 pairs it with its own rotated copy
               synth.make_batches(2, 1, 3, multiplicity=2)[0]]
    h = HBT_correlation(P, stage_counters=True)
    o = O.Oracle(P)
    for b in batches:
        h.calculate_HBT_correlation_function(b)
        o.process_batch(b)
    acc = h.accumulators()
    hbtio.compare(o.accumulators(), acc, rtol=RTOL, check_stage=True)
    assert int(acc.stage[0]) == 15 and int(acc.stage[6]) == 1 + 3 * 2 * 2 * 2
    # the RNG stream advanced exactly like the oracle's (an empty batch draws nothing)
    assert h.ran_gen.rand_int_uniform() == o.rand_int_uniform()


def test_rapidity_cut_applied():
    P = HBTParams(qnpts=21, HBTrap_min=-0.2, HBTrap_max=0.3)
    batches = synth.make_batches(5, 1, 4, multiplicity=500)
    ref = run_oracle(P, batches)
    _, acc = run_product(P, batches, stats=True)
    hbtio.compare(ref, acc, rtol=RTOL, check_stage=True)
    _, acc = run_product(P, batches)
    hbtio.compare(ref, acc, rtol=RTOL, check_stage="cheap")


def test_real_mixed_event_lists():
    P = HBTParams(qnpts=21)
    a = synth.make_batches(8, 1, 4, multiplicity=400)[0]
    b = synth.make_batches(9, 1, 5, multiplicity=350)[0]
    batch = hbtio.Batch(a.same, b.same)
    ref = run_oracle(P, [batch])
    _, acc = run_product(P, [batch], stats=True)
    hbtio.compare(ref, acc, rtol=RTOL, check_stage=True)
    _, acc = run_product(P, [batch])
    hbtio.compare(ref, acc, rtol=RTOL, check_stage="cheap")


@pytest.mark.parametrize("P", [HBTParams(qnpts=21, needed_number_of_pairs=3000.0),
                               HBTParams(qnpts=21, invariant_radius_flag=1, needed_number_of_pairs=300.0),
                               C4.with_(qnpts=11, n_KT=4, n_Kphi=4, needed_number_of_pairs=700.0)],
                         ids=["3d", "qinv", "az"])
def test_ordered_cap_with_real_mixed_event_lists(P):
    """The cap replay when list 2 is a separate buffer (read_in_real_mixed_events = 1): the literal passes must read
    list 2 at its offset in the staged buffer, like the optimistic pass and the host's row evaluation do."""
    batches = []
    for k in range(3):
        a = synth.make_batches(80 + k, 1, 4, multiplicity=400)[0]
        b = synth.make_batches(90 + k, 1, 5, multiplicity=350)[0]
        batches.append(hbtio.Batch(a.same, b.same))
    ref = run_oracle(P, batches)
    _, acc = run_product(P, batches)
    hbtio.compare(ref, acc, rtol=RTOL, check_stage="cheap")
    lim = int(P.needed_number_of_pairs) + 1
    assert int(acc.npairs_den.max()) == lim  # the denominator cap engaged on the separate list


def test_per_method_interface_matches_batched_call():
    """combine_and_bin_particle_pairs + per-event combine_and_bin_particle_pairs_mixed_events
    (the reference's public methods) give the same result as the one-submission batch call."""
    P = HBTParams(qnpts=21)
    batch = synth.make_batches(4, 1, 4, multiplicity=300)[0]
    ref = run_oracle(P, [batch])
    h = HBT_correlation(P)
    h.set_particle_list(batch)
    nev = len(batch.same)
    h.combine_and_bin_particle_pairs(list(range(nev)))
    nmix = nev // 2 + 1
    for iev in range(nev):
        ids = []
        while len(ids) < nmix:  # src/HBT_correlation.cpp:206-215
            k = h.ran_gen.rand_int_uniform() % nev
            while k == iev and nev != 1:
                k = h.ran_gen.rand_int_uniform() % nev
            ids.append(k)
        h.combine_and_bin_particle_pairs_mixed_events(iev, ids)
    hbtio.compare(ref, h.accumulators(), rtol=RTOL)


@pytest.mark.parametrize("name", ["c1_iss_gz_rap10", "unit_iss_gz_inv", "unit_iss_gz_az"])
def test_output_files_match_reference_text(name, tmp_path):
    """The written .dat files (format frozen by src/HBT_correlation.cpp:694-855) equal the
    reference's own output for the same input, line by line; a numeric field may differ in the
    last printed digit where a <=1e-10 difference straddles a rounding boundary."""
    P, batches, ref, meta = load_golden(name)
    h = HBT_correlation(P, path=str(tmp_path))
    for b in batches:
        h.calculate_HBT_correlation_function(b)
    files = {os.path.basename(f): f for f in h.output_HBTcorrelation()}
    assert sorted(files) == sorted(meta["dat_files"])
    checked = 0
    for fn in os.listdir(GOLDEN):
        if not fn.startswith(name + ".HBT_") or not fn.endswith(".gz"):
            continue
        base = fn[len(name) + 1:-3]
        want = gzip.open(os.path.join(GOLDEN, fn), "rt").read().splitlines()
        got = open(files[base]).read().splitlines()
        assert len(want) == len(got)
        for lw, lg in zip(want, got):
            if lw == lg:
                continue
            assert len(lw) == len(lg)
            fw, fg = lw.split(), lg.split()
            for a, b in zip(fw, fg):
                if a != b:
                    assert abs(float(a) - float(b)) <= 2e-8 * max(abs(float(a)), 1e-300), (lw, lg)
        checked += 1
    assert checked >= 1


def _random_case(k):
    """a seeded random parameters.dat: grid sizes, windows, K_T / K_phi binning, boost, rapidity cut,
    pair cap, species mass and sample size all vary"""
    r = np.random.default_rng(7000 + k)
    qmax = float(r.choice([0.05, 0.1, 0.2, 0.3, 0.4]))
    az = int(r.random() < 0.35)
    ktmin = float(r.choice([0.0, 0.05, 0.15, 0.3]))
    cap = float(r.choice([1e15, 1e15, 3000.0, 150.0]))
    P = HBTParams(qnpts=int(r.integers(5, 34)), q_min=-qmax, q_max=qmax, n_KT=int(r.integers(2, 8)), KT_min=ktmin,
                  KT_max=ktmin + float(r.choice([0.2, 0.4, 0.9])), n_Kphi=int(r.integers(1, 13)) if az else 8, azimuthal_flag=az,
                  invariant_radius_flag=int(r.random() < 0.15), long_comoving_boost=int(r.random() < 0.8),
                  HBTrap_min=-float(r.choice([0.2, 0.5, 1.0])), HBTrap_max=float(r.choice([0.2, 0.5, 1.0])),
                  needed_number_of_pairs=cap, randomSeed=int(r.integers(1, 10 ** 6)))
    mass = float(r.choice([0.138, KAON_MASS, 0.938]))
    return P, int(r.integers(1, 4)), int(r.integers(1, 6)), int(r.integers(150, 700)), mass


@pytest.mark.parametrize("k", range(16))
def test_random_parameter_sets_against_oracle(k):
    """16 seeded random parameter files (every switch of the class varies) through the production path
    and, for the uncapped ones, the instrumented path: all integers equal, sums within 1e-10."""
    P, ngrp, nev, mult, mass = _random_case(k)
    batches = synth.make_batches(20261000 + k, ngrp, nev, mass=mass, multiplicity=mult)
    ref = run_oracle(P, batches)
    _, acc = run_product(P, batches)
    hbtio.compare(ref, acc, rtol=RTOL, check_stage="cheap", q_scale=P.q_max)
    if P.needed_number_of_pairs > 1e14:
        _, acc = run_product(P, batches, stats=True)
        hbtio.compare(ref, acc, rtol=RTOL, check_stage=True, q_scale=P.q_max)
