#!/usr/bin/env python
"""How often do the accepted same-event pairs of one work unit hit the same histogram bin?  (CPU, numpy.)
The question behind block-privatised shared-memory histograms: a C5-shape group (150 000 pi+) is Morton-sorted in
(p_x, p_y) as the library does, 64 x 64 units near the diagonal of the sorted list are evaluated in binary64, and the
accepted pairs' (K_T, q_out, q_side, q_long) bins are counted per unit, per drain round of 32 pairs, and over a block
of 36 neighbouring units (what the warps of one SM hold at a time).

    python scripts/bin_repeat_stats.py > profiles/r02_bin_repeats.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hadronic_afterburner_toolkit_b200 import synth  # noqa: E402
from hadronic_afterburner_toolkit_b200.params import C5, PION_MASS  # noqa: E402

P = C5
p = np.asarray(synth.make_group(20260005, 0, 100, PION_MASS, 1500)).reshape(-1, 8)
R = np.abs(p[:, :2]).max() * 1.0001
u = ((p[:, :2] + R) * (32767.5 / R)).astype(np.uint32)


def spread(x):
    x = x.astype(np.uint64)
    x = (x | (x << 8)) & 0x00FF00FF
    x = (x | (x << 4)) & 0x0F0F0F0F
    x = (x | (x << 2)) & 0x33333333
    return (x | (x << 1)) & 0x55555555


s = p[np.argsort(spread(u[:, 0]) | (spread(u[:, 1]) << 1), kind="stable")]
dq = (P.q_max - P.q_min) / (P.qnpts - 1)
qb, lo, hi = P.q_min - dq / 2, P.q_min - dq / 2 + 1e-8, P.q_max + dq / 2 - 1e-8
dKT = (P.KT_max - P.KT_min) / (P.n_KT - 1)


def bins(a, b):
    Kx, Ky = 0.5 * (a[:, None, 0] + b[None, :, 0]), 0.5 * (a[:, None, 1] + b[None, :, 1])
    K2 = Kx * Kx + Ky * Ky
    ok = (K2 >= P.KT_min ** 2) & (K2 <= P.KT_max ** 2)
    Kp = np.sqrt(np.where(ok, K2, 1.0))
    iK = ((Kp - P.KT_min) / dKT).astype(int)
    qx, qy, qz, qE = (a[:, None, k] - b[None, :, k] for k in range(4))
    qo, qs = (qx * Kx + qy * Ky) / Kp, (qy * Kx - qx * Ky) / Kp
    Kz, KE = 0.5 * (a[:, None, 2] + b[None, :, 2]), 0.5 * (a[:, None, 3] + b[None, :, 3])
    ql = (KE / np.sqrt(KE * KE - Kz * Kz)) * (qz - (Kz / KE) * qE)
    for q in (qo, qs, ql):
        ok &= (q >= lo) & (q <= hi)
    io, is_, il = (((q - qb) / dq).astype(int) for q in (qo, qs, ql))
    ok &= (io < P.qnpts) & (is_ < P.qnpts) & (il < P.qnpts)
    return ok, ((iK * P.qnpts + io) * P.qnpts + is_) * P.qnpts + il, io * P.qnpts + is_


rng = np.random.default_rng(0)
ntile = len(s) // 64
tri = np.triu(np.ones((64, 64), bool), 1)
acc = dist = units = cols = 0
per_round = []
for t in rng.integers(0, ntile - 40, 60):
    for dt in (0, 1, 3, 8, 20):
        ok, flat, col = bins(s[t * 64:(t + 1) * 64], s[(t + dt) * 64:(t + dt + 1) * 64])
        if dt == 0:
            ok &= tri
        f = flat[ok]
        if len(f) < 8:
            continue
        units += 1
        acc += len(f)
        dist += len(np.unique(f))
        cols += len(np.unique(col[ok]))
        per_round += [len(np.unique(f[k:k + 32])) for k in range(0, len(f) - 31, 32)]
print(f"C5-shape group, {len(s)} pi+, Morton-sorted; {units} sampled 64 x 64 units that hold accepted pairs (tile distances 0, 1, 3, 8, 20)")
print(f"  accepted pairs per unit            {acc / units:8.0f}")
print(f"  distinct bins per unit             {dist / units:8.0f}   -> every bin of a unit is hit {acc / dist:.1f} times on average")
print(f"  distinct (q_out, q_side) columns   {cols / units:8.0f}   (x 41 q_long bins x the K_T slabs the unit reaches)")
print(f"  shared-memory table for one unit   {dist / units * 40 / 1024:8.1f} KB at 40 bytes per bin (count + four sums); a warp owns 11.6 KB today")
print(f"  distinct bins per drain round      {np.mean(per_round):8.1f} of 32 (min {np.min(per_round)})")
t0 = int(rng.integers(0, ntile - 60))
fs = []
for t in range(t0, t0 + 6):
    for dt in range(6):
        ok, flat, _ = bins(s[t * 64:(t + 1) * 64], s[(t + dt) * 64:(t + dt + 1) * 64])
        if dt == 0:
            ok &= tri
        fs.append(flat[ok])
f = np.concatenate(fs)
print(f"36 neighbouring units (one SM's warps at a time): {len(f)} accepted pairs in {len(np.unique(f))} distinct bins "
      f"({len(f) / len(np.unique(f)):.1f} hits per bin); a table for them: {len(np.unique(f)) * 40 / 1024:.0f} KB of the SM's 227 KB")
