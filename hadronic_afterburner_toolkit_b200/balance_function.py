"""Host-side mirror of the reference's ``class BalanceFunction``
(``/root/reference/src/BalanceFunction.h:16-75``) over the C ABI ``hbt_bf_*`` of
``include/hbt_b200.h`` — the "next" row SURVEY.md §8f rank 3.

The pair loops (``src/BalanceFunction.cpp:120-197``) run on the GPU; the host keeps the per-particle
pT cut, the RNG draws of the mixed-event routine (``:167-170``) and the output writers
(``:211-319``, format frozen).  No CPU implementation of the loops here: without the CUDA library
the constructor raises.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Dict, List

import numpy as np

from . import capi
from .hbt_correlation import Random

BNPHI = 20
HISTS = ("C_ab", "C_abarbbar", "C_abbar", "C_abarb", "C_mixed_ab", "C_mixed_abarbbar", "C_mixed_abbar", "C_mixed_abarb")


def cxx_sci(x: float, width: int = 0) -> str:
    """``ostream << std::scientific << std::setprecision(8) [<< std::setw(width)] << x``"""
    x = float(x)
    if math.isnan(x):
        s = "-nan" if math.copysign(1.0, x) < 0 else "nan"
    elif math.isinf(x):
        s = "-inf" if x < 0 else "inf"
    else:
        s = "%.8e" % x
    return s.rjust(width)


class BalanceFunction:
    """Same public surface as the reference class: constructor(parameters, path, random generator),
    ``calculate_balance_function(lists)``, ``output_balance_function()``.  ``lists`` maps "a", "b",
    "abar", "bbar" to per-event dicts of arrays ``pT, phi, rap_y, rap_eta`` — what
    ``particleSamples::get_balance_function_particle_list_*`` hold (``particle_info.pT, phi_p, rap_y,
    rap_eta``)."""

    def __init__(self, particle_alpha: int, particle_beta: int, Bnpts: int, Brap_max: float, BpT_min: float,
                 BpT_max: float, rap_type: int, path: str = ".", ran_gen: Random = None, device: int = 0):
        self.particle_monval_a, self.particle_monval_b = particle_alpha, particle_beta
        self.Bnpts, self.BpT_min, self.BpT_max, self.rap_type = Bnpts, BpT_min, BpT_max, rap_type
        self.drap = 2. * abs(Brap_max) / (Bnpts - 1)          # src/BalanceFunction.cpp:31-35
        self.Brap_min = -abs(Brap_max) - 0.5 * self.drap
        self.dphi = 2. * math.pi / BNPHI
        self.Bphi_min = -math.pi / 2.
        self.path_ = path
        self.ran_gen = ran_gen if ran_gen is not None else Random(-1)
        self._L = capi.lib()
        h = ctypes.c_void_p()
        rc = self._L.hbt_bf_create(Bnpts, Brap_max, device, ctypes.byref(h))
        if rc != 0:
            raise capi.HBTError(rc, (self._L.hbt_bf_last_error(None) or b"").decode())
        self._h = h
        self.N_b = 0
        self.N_bbar = 0

    def close(self):
        if self._h:
            self._L.hbt_bf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- gathers ------------------------------------------------------------------------------
    def _flat(self, plist: List[dict]):
        """(phi, rapidity) of the particles inside the pT cut (:127, :129), flat + per-event offsets"""
        off = np.zeros(len(plist) + 1, dtype=np.int64)
        parts = []
        for i, ev in enumerate(plist):
            keep = ~((ev["pT"] < self.BpT_min) | (ev["pT"] > self.BpT_max))
            rap = ev["rap_y"] if self.rap_type != 0 else ev["rap_eta"]
            parts.append(np.stack([ev["phi"][keep], rap[keep]], axis=1))
            off[i + 1] = off[i] + int(keep.sum())
        flat = np.ascontiguousarray(np.concatenate(parts) if parts else np.zeros((0, 2)), dtype=np.float64)
        return flat, off

    def _call(self, hist: int, a, b, partner: np.ndarray, rotation: np.ndarray):
        (fa, oa), (fb, ob) = a, b
        rc = self._L.hbt_bf_accumulate(self._h, hist, fa.ctypes.data, oa.ctypes.data, len(oa) - 1, fb.ctypes.data,
                                       ob.ctypes.data, len(ob) - 1, partner.ctypes.data, rotation.ctypes.data)
        if rc != 0:
            raise capi.HBTError(rc, self._L.hbt_bf_last_error(self._h).decode())

    def calculate_balance_function(self, lists: Dict[str, List[dict]]):
        """src/BalanceFunction.cpp:61-98, the eight calls in the reference's order"""
        g = {k: self._flat(lists[k]) for k in ("a", "b", "abar", "bbar")}
        nev = len(lists["a"])
        self.N_b += int(g["b"][1][-1])
        self.N_bbar += int(g["bbar"][1][-1])
        ident = np.arange(nev, dtype=np.int32)
        zero = np.zeros(nev, dtype=np.float64)
        order = (("a", "b"), ("abar", "bbar"), ("a", "bbar"), ("abar", "b"))
        for h, (x, y) in enumerate(order):
            self._call(h, g[x], g[y], ident, zero)
        for h, (x, y) in enumerate(order):  # mixed-event lists alias the batch (src/particleSamples.cpp:528-551)
            nev_mixed = len(lists[y])
            partner = np.zeros(nev, dtype=np.int32)
            rotation = np.zeros(nev, dtype=np.float64)
            for iev in range(nev):  # the draws of :167-170, in order
                partner[iev] = self.ran_gen.rand_int_uniform() % nev_mixed
                rotation[iev] = self.ran_gen.rand_uniform() * 2. * math.pi
            self._call(4 + h, g[x], g[y], partner, rotation)

    def histograms(self) -> np.ndarray:
        out = np.zeros((len(HISTS), self.Bnpts, BNPHI), dtype=np.uint64)
        rc = self._L.hbt_bf_read(self._h, out.ctypes.data)
        if rc != 0:
            raise capi.HBTError(rc, self._L.hbt_bf_last_error(self._h).decode())
        return out

    def timers(self):
        ms, pairs = ctypes.c_double(), ctypes.c_uint64()
        self._L.hbt_bf_get_timers(self._h, ctypes.byref(ms), ctypes.byref(pairs))
        return {"kernel_ms": ms.value, "pairs": pairs.value}

    def output_balance_function(self) -> None:
        write_outputs(self.histograms().astype(np.float64), self.particle_monval_a, self.particle_monval_b, self.Bnpts,
                      self.Brap_min, self.drap, self.path_)


def write_outputs(h: np.ndarray, monval_a: int, monval_b: int, Bnpts: int, Brap_min: float, drap: float, path: str) -> None:
    """``BalanceFunction::output_balance_function``, src/BalanceFunction.cpp:211-319: the three files,
    same sums in the same order, same columns.  ``h``: float64 [8][Bnpts][20]."""
    C = dict(zip(HISTS, h))
    dphi = 2. * math.pi / BNPHI
    Bphi_min = -math.pi / 2.
    f64 = np.float64
    N_OS = N_OS_mixed = N_SS = N_SS_mixed = f64(0.)
    y_OS, y_SS, y_OSm, y_SSm = (np.zeros(Bnpts) for _ in range(4))
    for i in range(Bnpts):
        for j in range(BNPHI):
            y_OS[i] += C["C_ab"][i][j] + C["C_abarbbar"][i][j]
            y_SS[i] += C["C_abbar"][i][j] + C["C_abarb"][i][j]
            y_OSm[i] += C["C_mixed_ab"][i][j] + C["C_mixed_abarbbar"][i][j]
            y_SSm[i] += C["C_mixed_abbar"][i][j] + C["C_mixed_abarb"][i][j]
        N_OS += y_OS[i]; N_SS += y_SS[i]; N_OS_mixed += y_OSm[i]; N_SS_mixed += y_SSm[i]
    p_OS, p_SS, p_OSm, p_SSm = (np.zeros(BNPHI) for _ in range(4))
    for j in range(BNPHI):
        for i in range(Bnpts):
            p_OS[j] += C["C_ab"][i][j] + C["C_abarbbar"][i][j]
            p_SS[j] += C["C_abbar"][i][j] + C["C_abarb"][i][j]
            p_OSm[j] += C["C_mixed_ab"][i][j] + C["C_mixed_abarbbar"][i][j]
            p_SSm[j] += C["C_mixed_abbar"][i][j] + C["C_mixed_abarb"][i][j]
    Delta_y = [Brap_min + (i + 0.5) * drap for i in range(Bnpts)]
    Delta_phi = [Bphi_min + (j + 0.5) * dphi for j in range(BNPHI)]
    tag = f"{monval_a}_{monval_b}"
    with np.errstate(divide="ignore", invalid="ignore"):
        with open(os.path.join(path, f"Balance_function_{tag}_Delta_y.dat"), "w") as f:
            f.write("# DeltaY  Delta_C2  C2(OS)  rho2(OS)  rho1^2(OS)  C2(SS) rho2(SS)  rho1^2(SS)\n")
            for i in range(Bnpts):
                C2_OS = f64(y_OS[i]) / f64(y_OSm[i]) * N_OS_mixed / N_OS
                C2_SS = f64(y_SS[i]) / f64(y_SSm[i]) * N_SS_mixed / N_SS
                f.write(cxx_sci(Delta_y[i], 18) + "   " + cxx_sci(C2_OS - C2_SS) + "  " + cxx_sci(C2_OS) + "  "
                        + cxx_sci(y_OS[i]) + "  " + cxx_sci(y_OSm[i]) + "  " + cxx_sci(C2_SS) + "  " + cxx_sci(y_SS[i])
                        + "  " + cxx_sci(y_SSm[i]) + "\n")
        with open(os.path.join(path, f"Balance_function_{tag}_Delta_phi.dat"), "w") as f:
            f.write("# Delta_phi  Delta_C2  C2(OS)  rho2(OS)  rho1^2(OS)  C2(SS) rho2(SS)  rho1^2(SS)\n")
            for j in range(BNPHI):
                C2_OS = f64(p_OS[j]) / f64(p_OSm[j]) * N_OS_mixed / N_OS
                C2_SS = f64(p_SS[j]) / f64(p_SSm[j]) * N_SS_mixed / N_SS
                f.write(cxx_sci(Delta_phi[j], 18) + "   " + cxx_sci(C2_OS - C2_SS) + "  " + cxx_sci(C2_OS) + "  "
                        + cxx_sci(p_OS[j]) + "  " + cxx_sci(p_OSm[j]) + "  " + cxx_sci(C2_SS) + "  " + cxx_sci(p_SS[j])
                        + "  " + cxx_sci(p_SSm[j]) + "\n")
        with open(os.path.join(path, f"Correlation_function_{tag}_2D.dat"), "w") as f:
            f.write("# DY  Dphi  C2(OS)  rho2(OS)  rho1^2(OS)  C2(SS)  rho2(SS)  rho1^2(SS)\n")
            for i in range(Bnpts):
                for j in range(BNPHI):
                    os_ = f64(C["C_ab"][i][j] + C["C_abarbbar"][i][j])
                    osm = f64(C["C_mixed_ab"][i][j] + C["C_mixed_abarbbar"][i][j])
                    ss_ = f64(C["C_abbar"][i][j] + C["C_abarb"][i][j])
                    ssm = f64(C["C_mixed_abbar"][i][j] + C["C_mixed_abarb"][i][j])
                    C2_OS = os_ / (osm + 1e-15)
                    C2_SS = ss_ / (ssm + 1e-15)
                    f.write(cxx_sci(Delta_y[i], 18) + "  " + cxx_sci(Delta_phi[j]) + "  " + cxx_sci(C2_OS * N_OS_mixed / N_OS)
                            + "  " + cxx_sci(os_) + "  " + cxx_sci(osm) + "  " + cxx_sci(C2_SS * N_SS_mixed / N_SS) + "  "
                            + cxx_sci(ss_) + "  " + cxx_sci(ssm) + "\n")
