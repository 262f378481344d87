"""TEST INFRASTRUCTURE — numpy restatement of the reference's BalanceFunction operator
(/root/reference/src/BalanceFunction.cpp) for the "next" row SURVEY.md §8f rank 3.

Every function cites the reference lines it follows.  The arithmetic per pair is IEEE binary64
add / divide / floor / int cast, which numpy evaluates exactly as the reference's C++ does
(no FMA contraction in either), so the histograms are the reference's integers.  Pinned against
the reference binary's own output files (tests/golden/bf_*.dat, made by
tests/golden/make_golden_bf.py): tests/test_bf_oracle.py.

Only tests/ may import this module; the product never does.
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np

BNPHI = 20  # src/BalanceFunction.cpp:33
HISTS = ("C_ab", "C_abarbbar", "C_abbar", "C_abarb", "C_mixed_ab", "C_mixed_abarbbar", "C_mixed_abbar", "C_mixed_abarb")


def kinematics(px, py, pz, E, mass, rap_shift: float = 0.0):
    """boostParticles, src/particleSamples.cpp:441-470: pT, phi_p, rap_y, rap_eta per particle
    (glibc atan2 / asinh through Python's math module: the same libm the reference links)."""
    ch, sh = math.cosh(rap_shift), math.sinh(rap_shift)
    n = len(px)
    pT = np.sqrt(px * px + py * py)
    mT = np.sqrt(pT * pT + mass * mass)
    pz_s = pz * ch + E * sh
    phi = np.array([math.atan2(py[i], px[i]) for i in range(n)], dtype=np.float64)
    rap_y = np.array([math.asinh(pz_s[i] / mT[i]) for i in range(n)], dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        rap_eta = np.array([math.asinh(pz_s[i] / pT[i]) if pT[i] > 0 else math.copysign(math.inf, pz_s[i])
                            for i in range(n)], dtype=np.float64)
    return pT, phi, rap_y, rap_eta


class BFOracle:
    """class BalanceFunction.  ``rng`` supplies the reference stream: ``rand_int_uniform()`` and
    ``rand_uniform()`` (oracle_py.Oracle has them)."""

    def __init__(self, Bnpts: int, Brap_max: float, BpT_min: float, BpT_max: float, rap_type: int, rng):
        # constructor, :27-36
        self.Bnpts = Bnpts
        self.drap = 2. * abs(Brap_max) / (Bnpts - 1)
        self.Brap_min = -abs(Brap_max) - 0.5 * self.drap
        self.dphi = 2. * math.pi / BNPHI
        self.Bphi_min = -math.pi / 2.
        self.BpT_min, self.BpT_max, self.rap_type = BpT_min, BpT_max, rap_type
        self.rng = rng
        self.h = {k: np.zeros((Bnpts, BNPHI), dtype=np.int64) for k in HISTS}
        self.N_b = 0
        self.N_bbar = 0
        self.pairs = 0

    # one particle list = list over events of dicts {"pT","phi","rap_y","rap_eta"}
    def _cut(self, ev):
        keep = ~((ev["pT"] < self.BpT_min) | (ev["pT"] > self.BpT_max))  # :127, :129
        rap = ev["rap_y"] if self.rap_type != 0 else ev["rap_eta"]  # :141-143
        return ev["phi"][keep], rap[keep]

    def _bin(self, hist, pa, ya, pb, yb, rotation):
        """the body of both loops, :134-151 / :175-192, for one event pair"""
        if len(pa) == 0 or len(pb) == 0:
            return
        self.pairs += len(pa) * len(pb)
        dphi = (pa[:, None] - pb[None, :]) + rotation
        phi_idx = np.floor((dphi - self.Bphi_min) / self.dphi).astype(np.int64)
        phi_idx = np.fmod(phi_idx, BNPHI)  # C++ %: sign of the dividend
        phi_idx = np.where(phi_idx < 0, phi_idx + BNPHI, phi_idx)
        dy = ya[:, None] - yb[None, :]
        ok = ~(np.abs(dy) < 1e-10) & ~(dy < self.Brap_min)
        with np.errstate(invalid="ignore"):
            y_idx = np.trunc((dy - self.Brap_min) / self.drap)
        ok &= (y_idx >= 0) & (y_idx < self.Bnpts)
        np.add.at(hist, (y_idx[ok].astype(np.int64), phi_idx[ok]), 1)

    def combine_and_bin_particle_pairs(self, name, plist_a, plist_b):  # :120-157
        for iev in range(len(plist_a)):
            pa, ya = self._cut(plist_a[iev])
            pb, yb = self._cut(plist_b[iev])
            self._bin(self.h[name], pa, ya, pb, yb, 0.0)

    def combine_and_bin_mixed_particle_pairs(self, name, plist_a, plist_b):  # :159-197
        nev, nev_mixed = len(plist_a), len(plist_b)
        for iev in range(nev):
            iev_mixed = self.rng.rand_int_uniform() % nev_mixed
            rotation = self.rng.rand_uniform() * 2. * math.pi
            pa, ya = self._cut(plist_a[iev])
            pb, yb = self._cut(plist_b[iev_mixed])
            self._bin(self.h[name], pa, ya, pb, yb, rotation)

    def calculate_balance_function(self, lists: Dict[str, List[dict]]):  # :61-98 (mixed lists alias the batch)
        a, b, abar, bbar = lists["a"], lists["b"], lists["abar"], lists["bbar"]
        self.N_b += sum(len(self._cut(ev)[0]) for ev in b)        # :69, get_number_of_particles :199-209
        self.N_bbar += sum(len(self._cut(ev)[0]) for ev in bbar)  # :70
        self.combine_and_bin_particle_pairs("C_ab", a, b)
        self.combine_and_bin_particle_pairs("C_abarbbar", abar, bbar)
        self.combine_and_bin_particle_pairs("C_abbar", a, bbar)
        self.combine_and_bin_particle_pairs("C_abarb", abar, b)
        self.combine_and_bin_mixed_particle_pairs("C_mixed_ab", a, b)
        self.combine_and_bin_mixed_particle_pairs("C_mixed_abarbbar", abar, bbar)
        self.combine_and_bin_mixed_particle_pairs("C_mixed_abbar", a, bbar)
        self.combine_and_bin_mixed_particle_pairs("C_mixed_abarb", abar, b)

    def histograms(self) -> np.ndarray:
        return np.stack([self.h[k] for k in HISTS]).astype(np.uint64)
