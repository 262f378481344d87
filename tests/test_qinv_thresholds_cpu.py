"""hbt_qinv_thresholds (csrc/hbt_host.cpp): the tuned kernels decide the reference's tests on q_inv = sqrt(s)
(src/HBT_correlation.cpp:342-345, :597-599) by comparing the double s with thresholds found on the host.  Here every
threshold is checked to be the exact step of the reference's expression — evaluated with numpy's correctly rounded
sqrt and IEEE division — one ulp below it the test fails / the bin is lower, at it the test holds / the bin is reached;
and the K_T thresholds of hbt_host_derive_grid the same way."""
import ctypes

import numpy as np
import pytest

from hadronic_afterburner_toolkit_b200 import capi
from hadronic_afterburner_toolkit_b200.params import HBTParams

HBT_MAX_KT = 64


class Grid(ctypes.Structure):  # HbtGrid of csrc/hbt_common.h
    _fields_ = [("nq", ctypes.c_int32), ("nKT", ctypes.c_int32), ("nKphi", ctypes.c_int32), ("az", ctypes.c_int32),
                ("qinv", ctypes.c_int32), ("boost", ctypes.c_int32), ("nslab", ctypes.c_int32), ("pad0", ctypes.c_int32),
                ("nbins", ctypes.c_int64), ("KT_min", ctypes.c_double), ("dKT", ctypes.c_double), ("KT_min_sq", ctypes.c_double),
                ("KT_max_sq", ctypes.c_double), ("q_base", ctypes.c_double), ("q_lo", ctypes.c_double), ("q_hi", ctypes.c_double),
                ("dq", ctypes.c_double), ("dKphi", ctypes.c_double), ("two_pi", ctypes.c_double), ("hbarc_inv", ctypes.c_double),
                ("rap_lo", ctypes.c_double), ("rap_hi", ctypes.c_double), ("needed", ctypes.c_uint64),
                ("kt_thr_sq", ctypes.c_double * HBT_MAX_KT), ("inv_dq", ctypes.c_double), ("slack", ctypes.c_double * 8)]


def derive(P):
    L = capi.lib()
    g = Grid()
    err = ctypes.create_string_buffer(256)
    cp = P.to_c()
    assert L.hbt_host_derive_grid(ctypes.byref(cp), ctypes.byref(g), err, 256) == 0, err.value
    assert g.nq == P.qnpts and g.nKT == P.n_KT and g.dq == (P.q_max - P.q_min) / (P.qnpts - 1)  # the mirror matches the layout
    return L, g


def below(x):
    return np.nextafter(x, -np.inf)


CASES = [HBTParams(invariant_radius_flag=1), HBTParams(invariant_radius_flag=1, qnpts=31),
         HBTParams(invariant_radius_flag=1, qnpts=13, q_min=0.02, q_max=0.14),
         HBTParams(invariant_radius_flag=1, qnpts=41, q_min=-0.4, q_max=0.4),
         HBTParams(invariant_radius_flag=1, qnpts=7, q_min=0.0, q_max=0.3),
         HBTParams(invariant_radius_flag=1, qnpts=16, q_min=-0.05, q_max=0.25)]


@pytest.mark.parametrize("P", CASES, ids=[f"q{p.qnpts}_{p.q_min}_{p.q_max}" for p in CASES])
def test_qinv_thresholds_are_the_exact_steps(P):
    L, g = derive(P)
    nq = P.qnpts
    s_lo, s_hi = ctypes.c_double(), ctypes.c_double()
    thr = (ctypes.c_double * (nq + 1))()
    L.hbt_qinv_thresholds(ctypes.byref(g), ctypes.byref(s_lo), ctypes.byref(s_hi), thr)
    thr = np.array(thr[:])
    q_lo, q_hi, q_base, dq = g.q_lo, g.q_hi, g.q_base, g.dq

    def idx(s):
        return int((np.sqrt(np.float64(s)) - q_base) / dq)

    # window: q_inv > q_lo <=> s >= s_lo ; q_inv < q_hi <=> s < s_hi
    assert np.sqrt(np.float64(s_lo.value)) > q_lo
    if s_lo.value > 0.0:
        assert not (np.sqrt(below(np.float64(s_lo.value))) > q_lo)
    assert not (np.sqrt(np.float64(s_hi.value)) < q_hi) and np.sqrt(below(np.float64(s_hi.value))) < q_hi
    assert thr[0] == s_lo.value
    # bins: thr[k] is the first s (inside the window) whose index is >= k
    finite = [k for k in range(1, nq + 1) if np.isfinite(thr[k])]
    for k in finite:
        assert thr[k] >= s_lo.value and thr[k] < s_hi.value
        assert idx(thr[k]) >= k
        if thr[k] > s_lo.value:
            assert idx(below(thr[k])) < k
    assert all(thr[a] <= thr[b] for a, b in zip(finite, finite[1:]))
    # random s inside the window: counting thresholds gives the reference's index
    rng = np.random.default_rng(1)
    s = rng.uniform(s_lo.value, below(np.float64(s_hi.value)), 20000)
    s = np.concatenate([s, thr[np.isfinite(thr)], below(thr[np.isfinite(thr) & (thr > s_lo.value)])])
    want = ((np.sqrt(s) - q_base) / dq).astype(np.int64)
    got = np.searchsorted(thr[1:], s, side="right")  # number of k >= 1 with s >= thr[k]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("P", [HBTParams(), HBTParams(n_KT=9, KT_min=0.15, KT_max=0.95), HBTParams(n_KT=6, KT_min=0.0, KT_max=1.0),
                               HBTParams(n_KT=33, KT_min=0.1, KT_max=2.0)], ids=["c5", "c4", "from_zero", "n33"])
def test_kt_thresholds_are_the_exact_steps(P):
    _, g = derive(P)
    thr = np.array(g.kt_thr_sq[:P.n_KT])

    def idx(x):
        return int((np.sqrt(np.float64(x)) - g.KT_min) / g.dKT)

    assert thr[0] == g.KT_min_sq
    for k in range(1, P.n_KT):
        if not np.isfinite(thr[k]):
            assert idx(g.KT_max_sq) < k
            continue
        assert idx(thr[k]) >= k and idx(below(thr[k])) < k
    rng = np.random.default_rng(2)
    x = rng.uniform(g.KT_min_sq, g.KT_max_sq, 20000)
    want = ((np.sqrt(x) - g.KT_min) / g.dKT).astype(np.int64)
    got = np.searchsorted(thr[1:][np.isfinite(thr[1:])], x, side="right")
    assert np.array_equal(got, want)
