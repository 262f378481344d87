"""The parameters.dat switches of the HBT path and the named benchmark configurations.

Field names are the reference's own keys (``/root/reference/parameters.dat:83-119``,
read at ``src/HBT_correlation.cpp:22-46``); ``n_KT`` is the number of K_T *edges*.
"""
from __future__ import annotations

import ctypes
from dataclasses import asdict, dataclass, replace


class CParams(ctypes.Structure):
    """``hbt_params`` of include/hbt_b200.h (also the layout of ``oracle_params``)."""

    _fields_ = [
        ("qnpts", ctypes.c_int32),
        ("n_KT", ctypes.c_int32),
        ("n_Kphi", ctypes.c_int32),
        ("azimuthal_flag", ctypes.c_int32),
        ("invariant_radius_flag", ctypes.c_int32),
        ("long_comoving_boost", ctypes.c_int32),
        ("q_min", ctypes.c_double),
        ("q_max", ctypes.c_double),
        ("KT_min", ctypes.c_double),
        ("KT_max", ctypes.c_double),
        ("HBTrap_min", ctypes.c_double),
        ("HBTrap_max", ctypes.c_double),
        ("needed_number_of_pairs", ctypes.c_double),
    ]


@dataclass(frozen=True)
class HBTParams:
    qnpts: int = 41
    q_min: float = -0.2
    q_max: float = 0.2
    n_KT: int = 5
    KT_min: float = 0.15
    KT_max: float = 0.55
    n_Kphi: int = 8
    azimuthal_flag: int = 0
    invariant_radius_flag: int = 0
    long_comoving_boost: int = 1
    HBTrap_min: float = -0.5
    HBTrap_max: float = 0.5
    needed_number_of_pairs: float = 1e15
    randomSeed: int = 12345
    particle_monval: int = 211

    def with_(self, **kw) -> "HBTParams":
        return replace(self, **kw)

    def to_c(self) -> CParams:
        d = asdict(self)
        return CParams(**{k: d[k] for k, _ in CParams._fields_})

    @property
    def n_slabs(self) -> int:
        return self.n_KT * (self.n_Kphi if self.azimuthal_flag == 1 else 1)

    @property
    def n_bins(self) -> int:
        return self.n_slabs * self.qnpts ** 3

    def parameters_dat(self, **extra) -> str:
        """A complete parameters.dat for the reference binary / ref_driver: every key the
        particleSamples and HBT_correlation constructors read with the no-default getVal
        (``src/particleSamples.cpp:24-125``, ``src/HBT_correlation.cpp:22-46``)."""
        kv = {
            "echo_level": 0,
            "read_in_mode": 10,
            "ecoOutput": 0,
            "analyze_flow": 0,
            "analyze_HBT": 1,
            "analyze_balance_function": 0,
            "analyze_ebe_yield": 0,
            "randomSeed": self.randomSeed,
            "read_in_real_mixed_events": 0,
            "particle_monval": self.particle_monval,
            "distinguish_isospin": 1,
            "resonance_weak_feed_down_flag": 0,
            "resonance_feed_down_flag": 0,
            "resonance_weak_feed_down_Sigma_to_Lambda_flag": 0,
            "select_resonances_flag": 0,
            "net_particle_flag": 0,
            "collect_neutral_particles": 0,
            "event_buffer_size": 1000000,
            "rapidity_shift": 0.0,
            "flag_charge_dependence": 0,
            "long_comoving_boost": self.long_comoving_boost,
            "needed_number_of_pairs": self.needed_number_of_pairs,
            "invariant_radius_flag": self.invariant_radius_flag,
            "azimuthal_flag": self.azimuthal_flag,
            "n_KT": self.n_KT,
            "KT_min": self.KT_min,
            "KT_max": self.KT_max,
            "n_Kphi": self.n_Kphi,
            "HBTrap_min": self.HBTrap_min,
            "HBTrap_max": self.HBTrap_max,
            "qnpts": self.qnpts,
            "q_min": self.q_min,
            "q_max": self.q_max,
        }
        kv.update(extra)
        return "".join(f"{k} = {v!r}\n" for k, v in kv.items())


# BASELINE.json configs.  C1 uses the unit-test fixtures.
C1 = HBTParams(qnpts=31, q_min=-0.15, q_max=0.15, n_KT=2, KT_min=0.0, KT_max=1.0, randomSeed=0)
C2 = HBTParams()  # 41^3, 4 K_T bins, oversampling 10, same-event only
C3 = C2  # + mixed events
C4 = HBTParams(n_KT=9, KT_min=0.15, KT_max=0.95, n_Kphi=8, azimuthal_flag=1)  # oversampling 50
C5 = C2  # oversampling 100, 20k events, same + mixed, sharded

PION_MASS = 0.138
KAON_MASS = 0.494
EVENT_MULTIPLICITY = 1500
