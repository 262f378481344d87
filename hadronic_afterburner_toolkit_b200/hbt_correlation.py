"""Python mirror of the reference's ``HBT_correlation`` operator interface
(``/root/reference/src/HBT_correlation.h:63-85``) on top of the C ABI of libhbt_b200.so.

The production host side is the C++ class in ``host/`` (same public surface, linked into the
reference's own binary).  This mirror exists so that tests and bench.py can drive the very
same C-ABI calls from Python; all arithmetic happens in the library (CUDA kernels for the
pair loops, reference-identical host helpers for the O(N) parts).  No fallbacks.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import List, Optional, Sequence

import numpy as np

from . import capi
from .hbtio import Accumulators, Batch
from .params import HBTParams


def _check(ctx, rc: int) -> None:
    if rc != capi.HBT_OK:
        msg = capi.lib().hbt_last_error(ctx)
        raise capi.HBTError(rc, msg.decode() if msg else "")


class Random:
    """``RandomUtil::Random`` (``src/Random.h:11-24``): the shared mt19937 stream."""

    def __init__(self, seed: int):
        self._L = capi.lib()
        h = ctypes.c_void_p()
        _check(None, self._L.hbt_rng_create(seed, ctypes.byref(h)))
        self._h = h

    def rand_int_uniform(self) -> int:
        return self._L.hbt_rng_int_uniform(self._h)

    def rand_uniform(self) -> float:
        return self._L.hbt_rng_uniform(self._h)

    def mixed_plan(self, nev: int, nev_mixed: int, want_angles: bool = False):
        """Draws of one batch in the reference's order (``src/HBT_correlation.cpp:202-217,
        493-497``): partner ids [nev, nmix], (cos, sin) [nev, nmix, 2] (and the angles)."""
        nmix = nev_mixed // 2 + 1 if nev_mixed > 0 else 0
        ids = np.zeros((nev, nmix), dtype=np.int32)
        cs = np.zeros((nev, nmix, 2), dtype=np.float64)
        ang = np.zeros((nev, nmix), dtype=np.float64) if want_angles else None
        got = self._L.hbt_rng_mixed_plan(self._h, nev, nev_mixed, ids.ctypes.data, cs.ctypes.data,
                                         ang.ctypes.data if want_angles else None)
        assert got == nmix
        return (ids, cs, ang) if want_angles else (ids, cs)

    def skip_batch(self, nev: int, nev_mixed: int) -> None:
        """Fast-forward the stream past a batch another rank processes."""
        self._L.hbt_rng_mixed_plan(self._h, nev, nev_mixed, None, None, None)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.hbt_rng_destroy(self._h)
            self._h = None


def gather_rapidity(params: HBTParams, p: np.ndarray) -> np.ndarray:
    """Single-particle rapidity cut of the gathers (``src/HBT_correlation.cpp:255-266``)."""
    p = np.ascontiguousarray(p, dtype=np.float64).reshape(-1, 8)
    out = np.empty_like(p)
    cp = params.to_c()
    n = capi.lib().hbt_gather_rapidity(ctypes.byref(cp), p.ctypes.data, p.shape[0], out.ctypes.data)
    return out[:n]


def psi_ref(p: np.ndarray, n_order: int) -> float:
    p = np.ascontiguousarray(p, dtype=np.float64).reshape(-1, 8)
    return capi.lib().hbt_psi_ref(p.ctypes.data, p.shape[0], n_order)


class HBT_correlation:
    """Same verbs as the reference class; ``particle_list`` is a :class:`Batch` (what
    ``particleSamples`` holds for one read: the filtered events and, optionally, real mixed
    events)."""

    def __init__(self, params: HBTParams, path: str = ".", ran_gen: Optional[Random] = None, device: int = 0,
                 stage_counters: Optional[bool] = None, kernel: Optional[int] = None, fuse: Optional[bool] = None,
                 lanes: Optional[int] = None, ptsort: Optional[int] = None, devices: Optional[Sequence[int]] = None,
                 coalesce: Optional[bool] = None):
        """``devices``: run the analysis on a GROUP of contexts, one per listed CUDA device (``hbt_group_*``; what the
        C++ class does for ``HBT_B200_DEVICES`` > 1): batches go to the devices in turn, the ordered pair cap stays
        exact, results are summed when read.  A device may be listed twice (several contexts on one GPU)."""
        self.params = params
        self.path_ = path
        self.ran_gen = ran_gen if ran_gen is not None else Random(params.randomSeed)
        self._L = capi.lib()
        self._cp = params.to_c()
        self._g = None
        self._hs = []
        if devices is not None:
            g = ctypes.c_void_p()
            devs = (ctypes.c_int32 * len(devices))(*devices)
            _check(None, self._L.hbt_group_create(ctypes.byref(self._cp), len(devices), devs, ctypes.byref(g)))
            self._g = g
            self._hs = [ctypes.c_void_p(self._L.hbt_group_ctx(g, i)) for i in range(len(devices))]
            h = self._hs[0]
        else:
            h = ctypes.c_void_p()
            _check(None, self._L.hbt_create(ctypes.byref(self._cp), device, ctypes.byref(h)))
            self._hs = [h]
        self._h = h
        for h in self._hs:
            if stage_counters is not None:  # HBT_OPT_STAGE_COUNTERS: instrumented run, exact stage populations
                _check(h, self._L.hbt_set_option(h, 1, int(stage_counters)))
            if kernel is not None:  # HBT_OPT_KERNEL
                _check(h, self._L.hbt_set_option(h, 2, int(kernel)))
            if fuse is not None:  # HBT_OPT_FUSE: one kernel for both loops of a batch (default) or one per loop
                _check(h, self._L.hbt_set_option(h, 3, int(fuse)))
            if lanes is not None:  # HBT_OPT_LANES: compute streams that take production batches in turn
                _check(h, self._L.hbt_set_option(h, 4, int(lanes)))
            if ptsort is not None:  # HBT_OPT_PTSORT: per-event pT-sorted copy for the mixed-event loops (0 never, 1 auto, 2 always)
                _check(h, self._L.hbt_set_option(h, 5, int(ptsort)))
            if coalesce is not None:  # HBT_OPT_COALESCE: small batches wait for each other and go out as one launch
                _check(h, self._L.hbt_set_option(h, 6, int(coalesce)))
        h = self._h
        self.psi_ref = 0.0
        self.psi_refs: List[float] = []
        self.particle_list: Optional[Batch] = None
        self.pairs_same = 0
        self.pairs_mixed = 0

    # -- lifetime ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "_g", None):
            self._L.hbt_group_destroy(self._g)
            self._g = None
            self._h = None
        elif getattr(self, "_h", None):
            self._L.hbt_destroy(self._h)
            self._h = None

    def _submit_batch(self, *args) -> None:
        """hbt_accumulate_batch on the single context, or hbt_group_accumulate_batch on the group."""
        if self._g is not None:
            rc = self._L.hbt_group_accumulate_batch(self._g, *args)
            if rc != capi.HBT_OK:
                raise capi.HBTError(rc, (self._L.hbt_group_last_error(self._g) or b"").decode())
        else:
            _check(self._h, self._L.hbt_accumulate_batch(self._h, *args))

    def _reduce(self) -> None:
        if self._g is not None:
            rc = self._L.hbt_group_reduce(self._g)
            if rc != capi.HBT_OK:
                raise capi.HBTError(rc, (self._L.hbt_group_last_error(self._g) or b"").decode())

    def ordered_batches(self) -> int:
        n = ctypes.c_uint64()
        if self._g is not None:
            self._L.hbt_group_ordered_batches(self._g, ctypes.byref(n))
        return n.value

    def __del__(self):
        self.close()

    # -- reference interface ----------------------------------------------------------
    def get_psi_ref(self) -> float:
        return self.psi_ref

    def set_particle_list(self, particle_list: Batch) -> None:
        self.particle_list = particle_list

    def calculate_flow_event_plane_angle(self, n_order: int) -> None:
        p, _ = self.particle_list.flat("same")
        self.psi_ref = psi_ref(p, n_order)

    def _gather(self, events: Sequence[np.ndarray]):
        cut = [gather_rapidity(self.params, ev) for ev in events]
        off = np.zeros(len(cut) + 1, dtype=np.int64)
        if cut:
            off[1:] = np.cumsum([len(c) for c in cut])
        flat = np.ascontiguousarray(np.concatenate(cut)) if cut else np.zeros((0, 8))
        if getattr(self, "pin_host", False) and flat.size:
            # page-locked gather buffers (the library then uploads them without a staging copy)
            import torch
            t = torch.from_numpy(flat).pin_memory()
            self._pinned = getattr(self, "_pinned", [])[-7:] + [t]  # (kept alive past the call)
            flat = t.numpy()
        return flat, off

    def calculate_HBT_correlation_function(self, particle_list: Batch, do_mixed: bool = True) -> None:
        """``src/HBT_correlation.cpp:177-218`` for one batch, as ONE library submission."""
        self.set_particle_list(particle_list)
        nev = len(particle_list.same)
        if self.params.azimuthal_flag == 1:
            self.calculate_flow_event_plane_angle(2)
        self.psi_refs.append(self.psi_ref)
        if nev == 0:
            return
        p1, off1 = self._gather(particle_list.same)
        if particle_list.mixed is not None:
            p2, off2 = self._gather(particle_list.mixed)
            nev2 = len(particle_list.mixed)
            a2 = (p2.ctypes.data, off2.ctypes.data, nev2)
        else:
            p2, off2, nev2 = None, off1, nev
            a2 = (None, None, 0)
        n = int(off1[-1])
        self.pairs_same += n * (n - 1) // 2
        if do_mixed:
            ids, cs = self.ran_gen.mixed_plan(nev, nev2)
            nmix = ids.shape[1]
            n2 = np.diff(off2)
            self.pairs_mixed += int(np.sum(np.diff(off1)[:, None] * n2[ids]))
            self._submit_batch(p1.ctypes.data, off1.ctypes.data, nev, a2[0], a2[1], a2[2], ids.ctypes.data,
                               cs.ctypes.data, nmix, self.psi_ref, 1, 1)
        else:
            self._submit_batch(p1.ctypes.data, off1.ctypes.data, nev, a2[0], a2[1], a2[2], None, None, 0,
                               self.psi_ref, 1, 0)

    def combine_and_bin_particle_pairs(self, event_list: Sequence[int]) -> None:
        """``src/HBT_correlation.cpp:251-462``: same-event pairs of the listed events merged."""
        p, _ = self._gather([self.particle_list.same[e] for e in event_list])
        self.pairs_same += len(p) * (len(p) - 1) // 2
        _check(self._h, self._L.hbt_accumulate_same(self._h, p.ctypes.data, p.shape[0], self.psi_ref))

    def combine_and_bin_particle_pairs_mixed_events(self, event_id: int, mixed_event_list: Sequence[int]) -> None:
        """``src/HBT_correlation.cpp:464-692``: one event against the listed partners, each
        rotated by a fresh ``rand_uniform()*2*pi`` angle drawn here, like the reference."""
        src = self.particle_list.mixed if self.particle_list.mixed is not None else self.particle_list.same
        p1, off1 = self._gather([self.particle_list.same[event_id]])
        p2, off2 = self._gather(src)
        ids = np.asarray(mixed_event_list, dtype=np.int32).reshape(1, -1)
        # math.cos / math.sin are glibc's (numpy may dispatch to a SIMD libm that differs by an ulp)
        ang = [self.ran_gen.rand_uniform() * 2 * math.pi for _ in range(ids.shape[1])]
        cs = np.array([[math.cos(a), math.sin(a)] for a in ang], dtype=np.float64).reshape(1, -1, 2)
        self.pairs_mixed += int(len(p1) * np.sum(np.diff(off2)[ids[0]]))
        _check(self._h, self._L.hbt_accumulate_mixed(
            self._h, p1.ctypes.data, off1.ctypes.data, 1, p2.ctypes.data, off2.ctypes.data, len(src),
            ids.ctypes.data, cs.ctypes.data, ids.shape[1], self.psi_ref))

    # -- results ----------------------------------------------------------------------
    def synchronize(self) -> None:
        for h in self._hs:
            _check(h, self._L.hbt_synchronize(h))

    def accumulators(self) -> Accumulators:
        P = self.params
        self._reduce()
        nb, ns = P.n_bins, P.n_slabs
        u = lambda n: np.zeros(n, dtype=np.uint64)
        d = lambda n: np.zeros(n, dtype=np.float64)
        num_count, den_count, kn, kd = u(nb), u(nb), u(ns), u(ns)
        num_cos, qo, qs, ql = d(nb), d(nb), d(nb), d(nb)
        _check(self._h, self._L.hbt_read(self._h, num_count.ctypes.data, num_cos.ctypes.data, qo.ctypes.data,
                                         qs.ctypes.data, ql.ctypes.data, den_count.ctypes.data,
                                         kn.ctypes.data, kd.ctypes.data))
        st = u(12)
        _check(self._h, self._L.hbt_get_stage_counters(self._h, st.ctypes.data, st[6:].ctypes.data))
        acc = Accumulators(P.azimuthal_flag, P.invariant_radius_flag, P.n_KT, P.n_Kphi, P.qnpts,
                           num_count, num_cos, qo, qs, ql, den_count, kn, kd,
                           psi_ref=list(self.psi_refs), stage=st)
        if P.invariant_radius_flag == 1:
            n1 = P.n_KT * P.qnpts
            acc.qinv_count, acc.qinv_den = u(n1), u(n1)
            acc.qinv_mean, acc.qinv_num = d(n1), d(n1)
            acc.npairs_num_qinv, acc.npairs_den_qinv = u(P.n_KT), u(P.n_KT)
            _check(self._h, self._L.hbt_read_qinv(
                self._h, acc.qinv_count.ctypes.data, acc.qinv_mean.ctypes.data, acc.qinv_num.ctypes.data,
                acc.qinv_den.ctypes.data, acc.npairs_num_qinv.ctypes.data, acc.npairs_den_qinv.ctypes.data))
        return acc

    def stage_counters(self) -> np.ndarray:
        st = np.zeros(12, dtype=np.uint64)
        _check(self._h, self._L.hbt_get_stage_counters(self._h, st.ctypes.data, st[6:].ctypes.data))
        return st

    def timers(self):
        s, m = ctypes.c_double(), ctypes.c_double()
        ns, nm = ctypes.c_uint64(), ctypes.c_uint64()
        _check(self._h, self._L.hbt_get_timers(self._h, ctypes.byref(s), ctypes.byref(m), ctypes.byref(ns),
                                               ctypes.byref(nm)))
        return {"same_ms": s.value, "mixed_ms": m.value, "same_launches": ns.value, "mixed_launches": nm.value}

    def deferred_pairs(self) -> int:
        n = ctypes.c_uint64()
        _check(self._h, self._L.hbt_get_deferred_pairs(self._h, ctypes.byref(n)))
        return n.value

    # -- output (format frozen by the reference, src/HBT_correlation.cpp:694-855) -------
    def output_HBTcorrelation(self) -> List[str]:
        acc = self.accumulators()
        files = []
        if self.params.invariant_radius_flag == 1:
            files += write_correlation_function_inv(self.path_, self.params, acc)
        if self.params.azimuthal_flag == 0:
            files += write_correlation_function(self.path_, self.params, acc)
        else:
            files += write_correlation_function_Kphi_differential(self.path_, self.params, acc)
        return files


def _sci(v) -> str:
    # what `ostream << std::scientific << setprecision(8)` prints, including glibc's "-nan"
    v = float(v)
    if v != v:
        return "-nan" if np.signbit(v) else "nan"
    return "%.8e" % v


def _fmt_row(vals) -> str:
    # std::scientific << setw(18) << setprecision(8): the width applies to the first field
    # only, 4 blanks between fields (src/HBT_correlation.cpp:772-776)
    return _sci(vals[0]).rjust(18) + "".join("    " + _sci(v) for v in vals[1:]) + "\n"


def _grid(P: HBTParams):
    dq = (P.q_max - P.q_min) / (P.qnpts - 1)
    q = [P.q_min + i * dq for i in range(P.qnpts)]
    dKT = (P.KT_max - P.KT_min) / (P.n_KT - 1)
    KT = [P.KT_min + i * dKT for i in range(P.n_KT)]
    dKphi = 2 * np.pi / P.n_Kphi
    Kphi = [i * dKphi for i in range(P.n_Kphi)]
    return q, KT, Kphi


def _write_slab(fn, P, q, num_count, num_cos, sqo, sqs, sql, den, ratio, eco):
    nq = P.qnpts
    shape = (nq, nq, nq)
    num_count, num_cos, sqo, sqs, sql, den = (a.reshape(shape) for a in (num_count, num_cos, sqo, sqs, sql, den))
    with open(fn, "w") as f:
        for il in range(nq):  # q_long outer, q_out middle, q_side inner (:735-737)
            for io in range(nq):
                for is_ in range(nq):
                    nn, nd = int(num_count[io, is_, il]), int(den[io, is_, il])
                    if nn < 2 or nd < 2:
                        row = (q[io], q[is_], q[il], 0.0, float(nd))
                    else:
                        row = (sqo[io, is_, il] / nn, sqs[io, is_, il] / nn, sql[io, is_, il] / nn,
                               num_cos[io, is_, il], ratio * float(den[io, is_, il]))
                    f.write(_fmt_row(row[3:] if eco else row))


def write_correlation_function(path, P: HBTParams, acc: Accumulators, eco: bool = False) -> List[str]:
    """``output_correlation_function`` (``src/HBT_correlation.cpp:726-783``)."""
    q, KT, _ = _grid(P)
    nb = P.qnpts ** 3
    out = []
    for iK in range(P.n_KT - 1):
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = float(np.float64(acc.npairs_num[iK]) / np.float64(acc.npairs_den[iK]))
        fn = os.path.join(path, "HBT_correlation_function_KT_%g_%g.dat" % (KT[iK], KT[iK + 1]))
        sl = slice(iK * nb, (iK + 1) * nb)
        _write_slab(fn, P, q, acc.num_count[sl], acc.num_cos[sl], acc.sum_qo[sl], acc.sum_qs[sl], acc.sum_ql[sl],
                    acc.den_count[sl], ratio, eco)
        out.append(fn)
    return out


def write_correlation_function_Kphi_differential(path, P: HBTParams, acc: Accumulators, eco: bool = False) -> List[str]:
    """``output_correlation_function_Kphi_differential`` (``src/HBT_correlation.cpp:785-855``)."""
    q, KT, Kphi = _grid(P)
    nb = P.qnpts ** 3
    out = []
    for iK in range(P.n_KT - 1):
        for ip in range(P.n_Kphi):
            slab = iK * P.n_Kphi + ip
            with np.errstate(divide="ignore", invalid="ignore"):
                ratio = float(np.float64(acc.npairs_num[slab]) / np.float64(acc.npairs_den[slab]))
            fn = os.path.join(path, "HBT_correlation_function_KT_%g_%g_Kphi_%g.dat" % (KT[iK], KT[iK + 1], Kphi[ip]))
            sl = slice(slab * nb, (slab + 1) * nb)
            _write_slab(fn, P, q, acc.num_count[sl], acc.num_cos[sl], acc.sum_qo[sl], acc.sum_qs[sl],
                        acc.sum_ql[sl], acc.den_count[sl], ratio, eco)
            out.append(fn)
    return out


def write_correlation_function_inv(path, P: HBTParams, acc: Accumulators, eco: bool = False) -> List[str]:
    """``output_correlation_function_inv`` (``src/HBT_correlation.cpp:694-724``)."""
    _, KT, _ = _grid(P)
    nq = P.qnpts
    out = []
    for iK in range(P.n_KT - 1):
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = np.float64(acc.npairs_num_qinv[iK]) / np.float64(acc.npairs_den_qinv[iK])
            fn = os.path.join(path, "HBT_correlation_function_inv_KT_%g_%g.dat" % (KT[iK], KT[iK + 1]))
            with open(fn, "w") as f:
                for k in range(nq):
                    i = iK * nq + k
                    qm = np.float64(acc.qinv_mean[i]) / np.float64(acc.qinv_count[i])
                    num = acc.qinv_num[i]
                    den = np.float64(acc.qinv_den[i]) * ratio
                    f.write(_fmt_row((num, den) if eco else (qm, num, den)))
        out.append(fn)
    return out
