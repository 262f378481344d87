"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares,
fails loudly without a GPU, and its host helpers (gather, psi_ref, RNG plan) reproduce the
oracle / the reference's known answers.  No pair-loop compute is called here."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
from hadronic_afterburner_toolkit_b200 import capi, synth
from hadronic_afterburner_toolkit_b200.hbt_correlation import Random, gather_rapidity, psi_ref
from hadronic_afterburner_toolkit_b200.params import HBTParams
from oracle import oracle_py as O


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(capi.LIB_PATH):
        capi.build()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hbt_b200.h")).read()
    declared = set(re.findall(r"\b(hbt_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    L = capi.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.hbt_version()


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = capi.lib()
    cp = HBTParams().to_c()
    h = ctypes.c_void_p()
    rc = L.hbt_create(ctypes.byref(cp), 0, ctypes.byref(h))
    assert rc == -3 and not h.value  # HBT_ERR_NO_DEVICE
    assert b"no CPU path" in L.hbt_last_error(None)


def test_invalid_parameters_rejected():
    L = capi.lib()
    cp = HBTParams(qnpts=1).to_c()
    h = ctypes.c_void_p()
    rc = L.hbt_create(ctypes.byref(cp), 0, ctypes.byref(h))
    assert rc in (-1, -3)


def test_psi_ref_known_answers():
    _, batches, _, _ = load_golden("c1_urqmd_gz")
    p, _ = batches[0].flat()
    for n, want in ((1, -1.6751628499713109), (2, -0.85236384292521539), (3, 0.06674083818478263)):
        assert abs(psi_ref(p, n) - want) < 1e-8
        assert psi_ref(p, n) == O.Oracle.psi_ref(p, n)


def test_rng_stream_matches_oracle():
    r = Random(12345)
    o = O.Oracle(HBTParams(randomSeed=12345))
    assert [r.rand_int_uniform() for _ in range(500)] == [o.rand_int_uniform() for _ in range(500)]
    assert [r.rand_uniform() for _ in range(500)] == [o.rand_uniform() for _ in range(500)]
    assert Random(12345).rand_int_uniform() == 1996335345  # reference build, ref_driver rng 12345


@pytest.mark.parametrize("nev", [1, 2, 5, 10])
def test_mixed_plan_matches_oracle(nev):
    P = HBTParams(qnpts=5, randomSeed=99)
    batches = synth.make_batches(3, 2, nev, multiplicity=20)
    o = O.Oracle(P)
    r = Random(P.randomSeed)
    for b in batches:
        o.process_batch(b)
        ids_o, ang_o = o.last_plan()
        ids, cs, ang = r.mixed_plan(nev, nev, want_angles=True)
        assert np.array_equal(ids, ids_o)
        assert np.array_equal(ang, ang_o)
        assert cs[..., 0].ravel().tolist() == [math.cos(a) for a in ang_o.ravel()]  # glibc, not numpy SIMD
        assert cs[..., 1].ravel().tolist() == [math.sin(a) for a in ang_o.ravel()]
    # skipping a batch consumes exactly the same draws
    r1, r2 = Random(7), Random(7)
    r1.mixed_plan(nev, nev)
    r2.skip_batch(nev, nev)
    assert r1.rand_int_uniform() == r2.rand_int_uniform()


def test_gather_rapidity_cut():
    _, batches, _, _ = load_golden("c1_iss_gz")
    P = HBTParams(HBTrap_min=-0.5, HBTrap_max=0.5)
    p, _ = batches[0].flat()
    got = gather_rapidity(P, p)
    r = p[:, 2] / p[:, 3]
    want = p[(r > np.tanh(-0.5)) & (r < np.tanh(0.5))]
    assert np.array_equal(got, want) and 0 < len(got) < len(p)
