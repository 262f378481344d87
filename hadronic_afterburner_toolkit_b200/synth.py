"""Synthetic Pb+Pb-like single-species samples.

Per event: event-plane angle Psi ~ U[0,2pi); p_x', p_y' ~ N(0,0.37), N(0,0.33) GeV rotated by
Psi; y ~ U(-0.45,0.45) (inside HBTrap = +-0.5, so every particle survives the rapidity cut);
m_T = sqrt(m^2+p_T^2), p_z = m_T sinh y, E = m_T cosh y; source x', y' ~ N(0, 4 fm) rotated by
Psi; tau ~ Gamma(k=4, theta=2.5 fm); eta_s = y + N(0,0.3); t = tau cosh eta_s, z = tau sinh eta_s.
Fixed multiplicity per event, so ``event_buffer_size = oversample * M`` yields groups of
exactly ``oversample`` events (reader rule, ``src/particleSamples.cpp:1256-1284``).

Groups are generated independently (Philox keyed by seed and group index), so any rank can
build just its own shard.
"""
from __future__ import annotations

import gzip
from typing import List

import numpy as np

from .hbtio import Batch
from .params import EVENT_MULTIPLICITY, PION_MASS


def make_group(seed: int, group: int, n_events: int, mass: float = PION_MASS,
               multiplicity: int = EVENT_MULTIPLICITY) -> np.ndarray:
    """One oversample group as a float64 array [n_events, multiplicity, 8] (px,py,pz,E,x,y,z,t)."""
    rng = np.random.Generator(np.random.Philox(key=[seed, group]))
    shp = (n_events, multiplicity)
    psi = rng.uniform(0.0, 2.0 * np.pi, size=(n_events, 1))
    c, s = np.cos(psi), np.sin(psi)
    pxp = rng.normal(0.0, 0.37, shp)
    pyp = rng.normal(0.0, 0.33, shp)
    y = rng.uniform(-0.45, 0.45, shp)
    xp = rng.normal(0.0, 4.0, shp)
    yp = rng.normal(0.0, 4.0, shp)
    tau = rng.gamma(4.0, 2.5, shp)
    eta = y + rng.normal(0.0, 0.3, shp)
    out = np.empty(shp + (8,), dtype=np.float64)
    px = pxp * c - pyp * s
    py = pxp * s + pyp * c
    mT = np.sqrt(mass * mass + px * px + py * py)
    out[..., 0] = px
    out[..., 1] = py
    out[..., 2] = mT * np.sinh(y)
    out[..., 3] = mT * np.cosh(y)
    out[..., 4] = xp * c - yp * s
    out[..., 5] = xp * s + yp * c
    out[..., 6] = tau * np.sinh(eta)
    out[..., 7] = tau * np.cosh(eta)
    return out


def make_batches(seed: int, n_groups: int, oversample: int, mass: float = PION_MASS,
                 multiplicity: int = EVENT_MULTIPLICITY, first_group: int = 0) -> List[Batch]:
    out = []
    for g in range(first_group, first_group + n_groups):
        arr = make_group(seed, g, oversample, mass, multiplicity)
        out.append(Batch([arr[e] for e in range(oversample)]))
    return out


def write_iss_gz(path: str, batches: List[Batch], monval: int = 211, mass: float = PION_MASS) -> None:
    """Write events as the reference's read_in_mode=10 text (``particle_samples.gz``): per event
    a line with the particle count, then ``monval mass t x y z E px py pz`` per particle
    (field order of ``src/particleSamples.cpp:1247-1286``), %.17g so the reader's text parse
    returns the same doubles."""
    with gzip.open(path, "wt", compresslevel=1) as f:
        for b in batches:
            for ev in b.same:
                f.write(f"{len(ev)}\n")
                for p in ev:
                    f.write("%d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n"
                            % (monval, mass, p[7], p[4], p[5], p[6], p[3], p[0], p[1], p[2]))


# UrQMD (particle id, 2 x isospin projection) of the species the writers below can emit
URQMD_IDS = {211: (101, 2), -211: (101, -2), 111: (101, 0), 321: (106, 1), -321: (-106, -1), 2212: (1, 1), 2112: (1, -1),
             3122: (27, 0), 0: (104, 0)}  # 0: an id the reference does not know (rho: dropped, but counted)
_URQMD_CHARGE = {211: 1, -211: -1, 111: 0, 321: 1, -321: -1, 2212: 1, 2112: 0, 3122: 0, 0: 0}


def urqmd_records(events, rng, other_fraction: float = 0.3, others=(321, -211, 2212, 0)):
    """Per event a list of (pdg, mass, p[8]) records: the given particles as pi+ with particles of
    other species (and of an id the reference's table does not hold) interleaved — they count
    for event_buffer_size and are dropped by the species filter."""
    out = []
    for ev in events:
        rows = []
        for p in ev:
            rows.append((211, PION_MASS, p))
            if rng.random() < other_fraction:
                rows.append((int(rng.choice(others)), 0.494, p * rng.uniform(0.5, 1.5)))
        out.append(rows)
    return out


def write_urqmd_gz(path: str, records, trailing_newline: bool = True) -> None:
    """read_in_mode=2 text (gzipped ``particle_list.dat``, ``src/particleSamples.cpp:910-974``):
    per event the particle count, one line the reader skips, then per particle
    ``id iso3 charge n1 n2 process mass t x y z E px py pz`` (%.17g: the text parse returns the
    same doubles)."""
    lines = []
    for rows in records:
        lines.append(f"{len(rows)} ")
        lines.append("65 4 61 0 221 11 0 0 ")
        for pdg, mass, p in rows:
            uid, iso3 = URQMD_IDS[pdg]
            lines.append("%d %d %d %d %d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g"
                         % (uid, iso3, _URQMD_CHARGE[pdg], 1, 6, 99, mass, p[7], p[4], p[5], p[6], p[3], p[0], p[1], p[2]))
    with gzip.open(path, "wt", compresslevel=1) as f:
        f.write("\n".join(lines) + ("\n" if trailing_newline else ""))


def write_urqmd_bin(path: str, records) -> None:
    """read_in_mode=21 binary (``particle_list.bin``, ``src/particleSamples.cpp:976-1059``): per
    event int32 count + 8 int32 that are skipped, per particle 6 int32 (id, iso3, charge, ...)
    and 9 float32 (mass t x y z E px py pz).  The values are rounded to float32 on the way."""
    with open(path, "wb") as f:
        for rows in records:
            f.write(np.array([len(rows), 65, 4, 61, 0, 221, 11, 0, 0], dtype=np.int32).tobytes())
            for pdg, mass, p in rows:
                uid, iso3 = URQMD_IDS[pdg]
                f.write(np.array([uid, iso3, _URQMD_CHARGE[pdg], 1, 6, 99], dtype=np.int32).tobytes())
                f.write(np.array([mass, p[7], p[4], p[5], p[6], p[3], p[0], p[1], p[2]], dtype=np.float32).tobytes())


def write_oscar(path: str, records) -> None:
    """read_in_mode=0 text (``OSCAR.DAT``, OSCAR1997A final_id_p_x; ``src/particleSamples.cpp:680-714``): three
    header lines, then per event ``<event id> <n> 0 0`` and n lines ``<index> <monval> px py pz E mass x y z t``."""
    with open(path, "w") as f:
        f.write("OSC1997A\nfinal_id_p_x\n 3DHydro       1.1  (197,    79)+(197,    79)  eqsp  0.1000E+03         1\n")
        for iev, rows in enumerate(records):
            f.write("%10d  %10d  %8d  %8d\n" % (iev + 1, len(rows), 0, 0))
            for k, (pdg, mass, p) in enumerate(rows):
                f.write("%10d  %10d  %.17g  %.17g  %.17g  %.17g  %.17g  %.17g  %.17g  %.17g  %.17g\n"
                        % (k + 1, pdg if pdg else 113, p[0], p[1], p[2], p[3], mass, p[4], p[5], p[6], p[7]))


def write_urqmd_f13(path: str, records, header_lines=17) -> None:
    """read_in_mode=1 text (UrQMD file-13 style ``particle_list.dat``, ``src/particleSamples.cpp:838-908``): per
    event 17 header lines, ``<n> <time>``, one line the reader skips, then n lines ``r0 rx ry rz p0 px py pz m ityp
    2i3 chg lcl# ncl or t x y z E px py pz`` of which the reader uses m, ityp, 2i3 and the last eight."""
    with open(path, "w") as f:
        for iev, rows in enumerate(records):
            # header_lines: 17 (read_in_mode 1), 14 (read_in_mode 4, UrQMD 3.3p) or 0 (read_in_mode 3)
            if header_lines:
                f.write("UQMD   version:       30400   1000  30400  output_file  13\n")
                for k in range(header_lines - 2):
                    f.write("header line %d of event %d\n" % (k + 2, iev + 1))
                f.write("pvec: r0 rx ry rz p0 px py pz m ityp 2i3 chg lcl# ncl or\n")
            f.write("%12d %11d\n" % (len(rows), 8000))
            f.write("      65       4      61       0     221      11       0       0\n")
            for pdg, mass, p in rows:
                uid, iso3 = URQMD_IDS[pdg]
                f.write(" 0.8E+04 1.0 2.0 3.0 %.8E %.8E %.8E %.8E %.17g %d %d %d %d %d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n"
                        % (p[3], p[0], p[1], p[2], mass, uid, iso3, _URQMD_CHARGE[pdg], 6, 1, 99,
                           p[7], p[4], p[5], p[6], p[3], p[0], p[1], p[2]))


def write_iss_bin(path: str, records) -> None:
    """read_in_mode=9 binary (iSS ``particle_list.bin``, ``src/particleSamples.cpp:1203-1245``): per event int32
    count, per particle int32 Monte-Carlo number and 9 float32 (mass t x y z E px py pz)."""
    with open(path, "wb") as f:
        for rows in records:
            f.write(np.array([len(rows)], dtype=np.int32).tobytes())
            for pdg, mass, p in rows:
                f.write(np.array([pdg if pdg else 113], dtype=np.int32).tobytes())
                f.write(np.array([mass, p[7], p[4], p[5], p[6], p[3], p[0], p[1], p[2]], dtype=np.float32).tobytes())


def write_smash_gz(path: str, records) -> None:
    """read_in_mode=7 text (gzipped SMASH ``particle_list.dat``, ``src/particleSamples.cpp:1061-1102``): per event
    the particle count, then ``pdg charge process mother1 mother2 mass t x y z E px py pz`` per particle."""
    lines = []
    for rows in records:
        lines.append(f"{len(rows)}")
        for pdg, mass, p in rows:
            lines.append("%d %d %d %d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g"
                         % (pdg if pdg else 113, _URQMD_CHARGE[pdg], 0, 0, 0, mass, p[7], p[4], p[5], p[6], p[3], p[0], p[1], p[2]))
    with gzip.open(path, "wt", compresslevel=1) as f:
        f.write("\n".join(lines) + "\n")


def write_jam(path: str, records) -> None:
    """read_in_mode=5 text (JAM ``particle_list.dat``, ``src/particleSamples.cpp:716-756``): one header line, per
    event ``# <event id> <n>`` and n lines ``monval mass px py pz x y z t`` (the reader computes E itself)."""
    with open(path, "w") as f:
        f.write("# JAM style list written by synth.write_jam\n")
        for iev, rows in enumerate(records):
            f.write("# %d %d\n" % (iev + 1, len(rows)))
            for pdg, mass, p in rows:
                f.write("%d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n"
                        % (pdg if pdg else 113, mass, p[0], p[1], p[2], p[4], p[5], p[6], p[7]))


def write_smash_bin(path: str, records, rng, format_version: int = 7) -> None:
    """read_in_mode=8 binary (extended SMASH ``particles_binary.bin``, ``src/particleSamples.cpp:288-323,
    1104-1201``): header ``SMSH``, u16 format version, u16 variant 1, u32 length + version string; per event a
    ``p`` block (u32 n, n records of 128 bytes) and an ``f`` block (u32 event, f64 impact parameter, one more
    byte from format version 7 on).  The particles are written at a later time t > t_last on their straight
    line, as SMASH does; the reader moves them back to t_last."""
    import struct
    with open(path, "wb") as f:
        ver = b"SMASH-synth"
        f.write(b"SMSH" + struct.pack("<HHI", format_version, 1, len(ver)) + ver)
        for iev, rows in enumerate(records):
            f.write(b"p" + struct.pack("<I", len(rows)))
            for k, (pdg, mass, p) in enumerate(rows):
                dt = float(rng.uniform(0.0, 30.0))
                t = p[7] + dt
                x, y, z = (p[4 + c] + p[c] / p[3] * dt for c in range(3))
                f.write(struct.pack("<9d4i2d2id2i", t, x, y, z, mass, p[3], p[0], p[1], p[2], pdg if pdg else 113, k,
                                    _URQMD_CHARGE[pdg], 3, 0.0, 1.0, 17, 5, p[7], 0, 0))
            f.write(b"f" + struct.pack("<Id", iev, 0.0) + (b"\x00" if format_version > 6 else b""))
