"""TEST INFRASTRUCTURE — Python access to the two checkers.

* ``Oracle``        ctypes wrapper of oracle/libhbt_oracle.so (the plain-C restatement).
* ``run_reference`` runs oracle/_ref/ref_driver (the unmodified reference compiled from
                    /root/reference by oracle/Makefile) on in-memory batches.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference``
legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import tempfile
from typing import List, Optional

import numpy as np

from hadronic_afterburner_toolkit_b200.hbtio import (Accumulators, Batch, read_accumulators,
                                                     write_batches)
from hadronic_afterburner_toolkit_b200.params import CParams, HBTParams

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libhbt_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_DRIVER = os.path.join(REF_DIR, "ref_driver")
REF_EXE = os.path.join(REF_DIR, "hadronic_afterburner_tools.e")


def build(ref: bool = True) -> None:
    """Compile the checkers (``make -C oracle``).  The reference part needs
    /root/reference and is skipped where that tree is absent (the GPU box uses the
    prebuilt oracle/_ref that travelled with the snapshot)."""
    subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-C", HERE, "-j8", "ref"], check=True, capture_output=True)


def have_reference() -> bool:
    return os.path.exists(REF_DRIVER)


_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(ORACLE_SO):
        build(ref=False)
    L = ctypes.CDLL(ORACLE_SO)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    L.oracle_create.restype = vp
    L.oracle_create.argtypes = [ctypes.POINTER(CParams), i32]
    L.oracle_destroy.argtypes = [vp]
    L.oracle_rand_int_uniform.restype = i32
    L.oracle_rand_int_uniform.argtypes = [vp]
    L.oracle_rand_uniform.restype = dbl
    L.oracle_rand_uniform.argtypes = [vp]
    L.oracle_psi_ref.restype = dbl
    L.oracle_psi_ref.argtypes = [vp, i64, i32]
    L.oracle_process_batch.argtypes = [vp, vp, vp, i32, vp, vp, i32, i32]
    L.oracle_skip_batch.argtypes = [vp, i32, i32]
    L.oracle_nbins.restype = i64
    L.oracle_nbins.argtypes = [vp]
    for name in ("num_count", "num_cos", "sum_qo", "sum_qs", "sum_ql", "den_count", "npairs_num",
                 "npairs_den", "qinv_count", "qinv_mean", "qinv_num", "qinv_den", "npairs_num_qinv",
                 "npairs_den_qinv", "stage_counters", "last_partner_ids", "last_angles"):
        f = getattr(L, "oracle_" + name)
        f.restype = vp
        f.argtypes = [vp]
    L.oracle_last_psi_ref.restype = dbl
    L.oracle_last_psi_ref.argtypes = [vp]
    L.oracle_last_nmix.restype = i32
    L.oracle_last_nmix.argtypes = [vp]
    L.oracle_time_same.restype = dbl
    L.oracle_time_same.argtypes = [vp]
    L.oracle_time_mixed.restype = dbl
    L.oracle_time_mixed.argtypes = [vp]
    _lib = L
    return L


def _view(ptr, dtype, n):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


class Oracle:
    """The CPU restatement (oracle/hbt_oracle.c) behind the same verbs as the product."""

    def __init__(self, params: HBTParams):
        self.params = params
        self.L = _load()
        cp = params.to_c()
        self.h = self.L.oracle_create(ctypes.byref(cp), params.randomSeed)
        self.psi_refs: List[float] = []
        self.last_nev = 0

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def rand_int_uniform(self) -> int:
        return self.L.oracle_rand_int_uniform(self.h)

    def rand_uniform(self) -> float:
        return self.L.oracle_rand_uniform(self.h)

    @staticmethod
    def psi_ref(p: np.ndarray, n_order: int) -> float:
        p = np.ascontiguousarray(p, dtype=np.float64)
        return _load().oracle_psi_ref(p.ctypes.data, p.shape[0], n_order)

    def process_batch(self, batch: Batch, do_mixed: bool = True) -> None:
        p, off = batch.flat("same")
        if batch.mixed is not None:
            pm, offm = batch.flat("mixed")
            self.L.oracle_process_batch(self.h, p.ctypes.data, off.ctypes.data, len(batch.same),
                                        pm.ctypes.data, offm.ctypes.data, len(batch.mixed), int(do_mixed))
        else:
            self.L.oracle_process_batch(self.h, p.ctypes.data, off.ctypes.data, len(batch.same),
                                        None, None, 0, int(do_mixed))
        self.last_nev = len(batch.same)
        self.psi_refs.append(self.L.oracle_last_psi_ref(self.h))

    def skip_batch(self, nev: int, nev_mixed: int) -> None:
        self.L.oracle_skip_batch(self.h, nev, nev_mixed)

    def last_plan(self):
        nmix = self.L.oracle_last_nmix(self.h)
        n = self.last_nev * nmix
        ids = _view(self.L.oracle_last_partner_ids(self.h), np.int32, n).reshape(self.last_nev, nmix)
        ang = _view(self.L.oracle_last_angles(self.h), np.float64, n).reshape(self.last_nev, nmix)
        return ids, ang

    def times(self):
        return self.L.oracle_time_same(self.h), self.L.oracle_time_mixed(self.h)

    def accumulators(self) -> Accumulators:
        P, L, h = self.params, self.L, self.h
        nb = L.oracle_nbins(h)
        g = lambda name, dt, n: _view(getattr(L, "oracle_" + name)(h), dt, n)
        acc = Accumulators(
            P.azimuthal_flag, P.invariant_radius_flag, P.n_KT, P.n_Kphi, P.qnpts,
            g("num_count", np.float64, nb), g("num_cos", np.float64, nb), g("sum_qo", np.float64, nb),
            g("sum_qs", np.float64, nb), g("sum_ql", np.float64, nb), g("den_count", np.float64, nb),
            g("npairs_num", np.uint64, P.n_slabs), g("npairs_den", np.uint64, P.n_slabs),
            psi_ref=list(self.psi_refs), stage=g("stage_counters", np.uint64, 12))
        if P.invariant_radius_flag == 1:
            n1 = P.n_KT * P.qnpts
            acc.qinv_count = g("qinv_count", np.float64, n1)
            acc.qinv_mean = g("qinv_mean", np.float64, n1)
            acc.qinv_num = g("qinv_num", np.float64, n1)
            acc.qinv_den = g("qinv_den", np.float64, n1)
            acc.npairs_num_qinv = g("npairs_num_qinv", np.uint64, P.n_KT)
            acc.npairs_den_qinv = g("npairs_den_qinv", np.uint64, P.n_KT)
        acc.t_same, t_mixed = self.times()
        acc.t_total = acc.t_same + t_mixed
        return acc


def run_reference(params: HBTParams, batches: List[Batch], same_only: bool = False,
                  workdir: Optional[str] = None, quiet: bool = True, only=None) -> Accumulators:
    """Push in-memory batches through the unmodified reference (oracle/_ref/ref_driver mem).
    ``only``: indices of the batches to process (the others only advance the RNG stream)."""
    assert have_reference(), "oracle/_ref/ref_driver missing: run `make -C oracle ref` where /root/reference exists"
    with tempfile.TemporaryDirectory(dir=workdir) as td:
        fin, fpar, fout = (os.path.join(td, n) for n in ("batches.bin", "parameters.dat", "out.bin"))
        write_batches(fin, batches)
        with open(fpar, "w") as f:
            f.write(params.parameters_dat())
        cmd = [REF_DRIVER, "mem", fpar, fin, fout] + (["same_only"] if same_only else [])
        if only is not None:
            cmd.append("only=" + ",".join(str(int(i)) for i in only))
        r = subprocess.run(cmd, stdout=subprocess.DEVNULL if quiet else None, stderr=subprocess.PIPE)
        if r.returncode != 0:
            raise RuntimeError(f"ref_driver failed ({r.returncode}): {r.stderr.decode()[-400:]}")
        return read_accumulators(fout)


def reference_rng(seed: int, n: int):
    """n rand_int_uniform() then n rand_uniform() draws of the reference's Random class."""
    out = subprocess.run([REF_DRIVER, "rng", str(seed), str(n)], check=True, capture_output=True).stdout.split()
    return [int(x) for x in out[:n]], [float(x) for x in out[n:]]
