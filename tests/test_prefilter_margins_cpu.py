"""The packed-FP32 prefilter of the production kernels (v3_run_unit in csrc/hbt_kernels_v3.cuh) may only
drop a pair that certainly fails the K_T cut or the q_out / q_side window.  Its margins (64 u S on k2, the
relative margin rho on the window, the k2 floor when rho is too large) are re-stated here in numpy — float32
operations rounded once, fma as one rounding, directed roundings of the thresholds as in the kernel — and
checked against the binary64 decision on millions of pairs per case: tiles of 128 x 128 particles of the
benchmark distribution, with momentum outliers in the tile (x30, x3000: S grows, the margins follow), with
tiny momenta and a K_T range that starts at 0 (the floor branch)."""
import numpy as np
import pytest

F = np.float32
U = 2.0 ** -24


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def f32_ru(x):
    y = np.float32(x)
    return np.nextafter(y, np.float32(np.inf)) if float(y) < x else y


def f32_rd(x):
    y = np.float32(x)
    return np.nextafter(y, np.float32(-np.inf)) if float(y) > x else y


def tile(rng, n, scale, outlier):
    px, py = rng.normal(0, 0.37, n) * scale, rng.normal(0, 0.33, n) * scale
    if outlier:
        k = rng.choice(n, 3, replace=False)
        px[k] *= outlier
        py[k] *= outlier
    return px, py


CASES = {
    "benchmark": dict(scale=1.0, outlier=0, ktmin=0.15, ktmax=0.55),
    "outliers_x30": dict(scale=1.0, outlier=30.0, ktmin=0.15, ktmax=0.55),
    "outliers_x3000": dict(scale=1.0, outlier=3000.0, ktmin=0.15, ktmax=0.55),
    "kt_from_0": dict(scale=1.0, outlier=0, ktmin=0.0, ktmax=1.0),
    "tiny_kt_from_0": dict(scale=1e-3, outlier=0, ktmin=0.0, ktmax=1.0),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_prefilter_never_drops_a_pair_the_exact_test_keeps(case):
    c = CASES[case]
    rng = np.random.default_rng(20260600 + sorted(CASES).index(case))
    dq = 0.4 / 40
    q_lo, q_hi = -0.2 - dq / 2 + 1e-8, 0.2 + dq / 2 - 1e-8
    W2 = max(q_lo * q_lo, q_hi * q_hi)
    Wd = np.sqrt(W2)
    k2lo, k2hi = 4 * c["ktmin"] ** 2, 4 * c["ktmax"] ** 2
    kept_exact = dropped_exact = floor_tiles = 0
    for _ in range(60):
        ax, ay = tile(rng, 128, c["scale"], c["outlier"])
        bx, by = tile(rng, 128, c["scale"], c["outlier"])
        at, bt = ax * ax + ay * ay, bx * bx + by * by
        S = at.max() + bt.max()
        # the margins, as the kernel derives them per unit (binary64)
        Ek = 64.0 * U * S
        k2e = max(k2lo - Ek, 0.0)
        rho = 64.0 * U * S / (Wd * np.sqrt(k2e)) + 64.0 * U * S / k2e + 4.0 * U if k2e > 0 else 1.0
        use_floor = not (rho <= 0.03)
        k2_floor = 0.0
        if use_floor:
            a1, a2 = 64.0 * U * S / (0.015 * Wd), 64.0 * U * S / 0.015
            k2_floor = max(a1 * a1, a2) + Ek
            rho = 0.03 + 4.0 * U
            floor_tiles += 1
        Wqf = f32_ru(0.25 * W2 * (1.0 + 4.0 * rho + 16.0 * U))
        klo_f, khi_f, kfloor_f = f32_rd(max(k2lo - Ek, 0.0)), f32_ru(k2hi + Ek), f32_ru(k2_floor)
        # float prefilter on all 128 x 128 pairs
        fax, fay, fat = (v.astype(np.float32)[:, None] for v in (ax, ay, 0.5 * at))
        fbx, fby, fnb = (v.astype(np.float32)[None, :] for v in (bx, by, -0.5 * bt))
        sx, sy = fax + fbx, fay + fby
        k2 = fma32(sy, sy, sx * sx)
        d = fat + fnb
        x = fma32(np.broadcast_to(fbx, k2.shape), np.broadcast_to(fay, k2.shape), (-fax) * fby)
        keep = (k2 >= klo_f) & (k2 <= khi_f)
        inside = np.maximum(d * d, x * x) <= k2 * Wqf
        if use_floor:
            inside |= k2 < kfloor_f
        passed = keep & inside
        # binary64 decision: K_T cut and both windows (what the drain / the reference decide)
        SX, SY, QX, QY = ax[:, None] + bx[None, :], ay[:, None] + by[None, :], ax[:, None] - bx[None, :], ay[:, None] - by[None, :]
        K2 = SX * SX + SY * SY
        with np.errstate(divide="ignore", invalid="ignore"):
            QO, QS = (QX * SX + QY * SY) / np.sqrt(K2), (QY * SX - QX * SY) / np.sqrt(K2)
        exact = (K2 >= k2lo) & (K2 <= k2hi) & (QO > q_lo) & (QO < q_hi) & (QS > q_lo) & (QS < q_hi)
        assert not np.any(exact & ~passed), (case, int(np.sum(exact & ~passed)))
        kept_exact += int(exact.sum())
        dropped_exact += int((~passed).sum())
    assert kept_exact > 1000                       # the sample does contain accepted pairs
    if case == "benchmark":
        assert dropped_exact > 0.5 * 60 * 128 * 128  # ... and the prefilter does discard most of the others
    if case == "tiny_kt_from_0":
        assert floor_tiles > 0                     # the floor branch was exercised


@pytest.mark.parametrize("scale", [1.0, 30.0, 1e-3])
def test_pt_range_restriction_never_excludes_an_accepted_pair(scale):
    """The production mixed-event units only visit the list-2 particles with (pT_min - W)^2 (1 - 1e-9) <= pT^2 <=
    (pT_max + W)^2 (1 + 1e-9), pT_min / pT_max over the unit's 128 list-1 particles and W = sqrt(W2) (1 + 1e-9)
    (v3_run_unit, PTRANGE): every pair with |q_out| inside the window must lie in that stretch, because
    |q_out| = |pT_i^2 - pT_j^2| / (2 K_perp) >= |pT_i - pT_j|."""
    rng = np.random.default_rng(int(scale * 1000) + 5)
    dq = 0.4 / 40
    q_lo, q_hi = -0.2 - dq / 2 + 1e-8, 0.2 + dq / 2 - 1e-8
    W2 = max(q_lo * q_lo, q_hi * q_hi)
    n_in = 0
    for _ in range(40):
        ax, ay = tile(rng, 128, scale, 0)
        ax, ay = (v[np.argsort(ax * ax + ay * ay)[40:168]] if len(v) > 168 else v for v in (ax, ay))
        bx, by = tile(rng, 1500, scale, 0)
        at, bt = ax * ax + ay * ay, bx * bx + by * by
        Wd = np.sqrt(W2) * (1.0 + 1e-9)
        lo, hi = np.sqrt(at.min()) - Wd, np.sqrt(at.max()) + Wd
        pt2_lo = lo * lo * (1.0 - 1e-9) if lo > 0 else 0.0
        pt2_hi = hi * hi * (1.0 + 1e-9)
        visited = (bt >= pt2_lo) & (bt <= pt2_hi)
        SX, SY = ax[:, None] + bx[None, :], ay[:, None] + by[None, :]
        K2 = SX * SX + SY * SY
        with np.errstate(divide="ignore", invalid="ignore"):
            QO = ((ax[:, None] - bx[None, :]) * SX + (ay[:, None] - by[None, :]) * SY) / np.sqrt(K2)
            # the inequality itself, to rounding
            assert np.all(np.abs(QO) >= np.abs(np.sqrt(at)[:, None] - np.sqrt(bt)[None, :]) * (1 - 1e-12) - 1e-300)
        inside = (QO > q_lo) & (QO < q_hi)
        assert not np.any(inside & ~visited[None, :])
        n_in += int(inside.sum())
    assert n_in > 1000


QINV_CASES = {
    # momentum scale, outlier factor, rapidity spread, mass: the margin follows the largest |p_z|, |E| of the tiles
    "pions": dict(scale=1.0, outlier=0, ymax=0.5, m=0.138),
    "pions_wide_rapidity": dict(scale=1.0, outlier=0, ymax=3.0, m=0.138),
    "kaons_outliers": dict(scale=1.0, outlier=30.0, ymax=0.5, m=0.494),
    "soft": dict(scale=0.4, outlier=0, ymax=0.5, m=0.138),
}


@pytest.mark.parametrize("case", sorted(QINV_CASES))
def test_qinv_prefilter_never_drops_a_pair_inside_the_qinv_window(case):
    """q_inv mode (v3_run_unit<.., QINV = true>): a pair is also kept when, in floats,
    4 e + q_z^2 <= q_E^2 + k2 + B with e = pT_i^2/2 + pT_j^2/2 and B = max(q_hi, 0)^2 (1 + 4u) +
    2u (81 S + 6 Z^2 + 6 E^2).  Every pair that passes the K_T cut and has s = -(q.q) < q_hi^2 in binary64 must
    pass (K_T cut with its own margin, as above)."""
    c = QINV_CASES[case]
    rng = np.random.default_rng(20260700 + sorted(QINV_CASES).index(case))
    dq = 0.4 / 40
    q_hi = 0.2 + dq / 2 - 1e-8
    w2f = np.float32(q_hi * q_hi * (1.0 + 2.4e-7))
    k2lo, k2hi = 4 * 0.15 ** 2, 4 * 0.55 ** 2
    n_in = n_dropped = 0
    for _ in range(60):
        def full(n):
            px, py = tile(rng, n, c["scale"], c["outlier"])
            y = rng.uniform(-c["ymax"], c["ymax"], n)
            mt = np.sqrt(c["m"] ** 2 + px * px + py * py)
            return px, py, mt * np.sinh(y), mt * np.cosh(y)
        ax, ay, az, aE = full(128)
        bx, by, bz, bE = full(128)
        at, bt = ax * ax + ay * ay, bx * bx + by * by
        S = at.max() + bt.max()
        Ek = 64.0 * U * S
        klo_f, khi_f = f32_rd(max(k2lo - Ek, 0.0)), f32_ru(k2hi + Ek)
        f = lambda v: v.astype(np.float32)
        Zs = float(np.abs(f(az)).max()) + float(np.abs(f(bz)).max())
        Es = float(np.abs(f(aE)).max()) + float(np.abs(f(bE)).max())
        Bq = f32_ru(float(w2f) * (1.0 + 4.0 * U) + 2.0 * U * (81.0 * S + 6.0 * Zs * Zs + 6.0 * Es * Es))
        fax, fay, fat, faz, faE = (f(v)[:, None] for v in (ax, ay, 0.5 * at, az, aE))
        fbx, fby, fnb, fnz, fnE = (f(v)[None, :] for v in (bx, by, -0.5 * bt, bz, bE))
        fnz, fnE = -fnz, -fnE  # the tile stores -p_z, -E
        sx, sy = fax + fbx, fay + fby
        k2 = fma32(sy, sy, sx * sx)
        qz, qE = faz + fnz, faE + fnE
        e2 = fat + (-fnb)
        lhs = fma32(e2, np.float32(4.0) * np.ones_like(e2), qz * qz)
        rhs = fma32(qE, qE, k2 + Bq)
        passed = (k2 >= klo_f) & (k2 <= khi_f) & (lhs <= rhs)
        # binary64: the reference's s and its window test
        SX, SY = ax[:, None] + bx[None, :], ay[:, None] + by[None, :]
        K2 = SX * SX + SY * SY
        QX, QY, QZ, QE = (u[:, None] - v[None, :] for u, v in ((ax, bx), (ay, by), (az, bz), (aE, bE)))
        s = -(QE * QE - QX * QX - QY * QY - QZ * QZ)
        with np.errstate(invalid="ignore"):
            exact = (K2 >= k2lo) & (K2 <= k2hi) & (np.sqrt(s) < q_hi)
        assert not np.any(exact & ~passed), (case, int(np.sum(exact & ~passed)))
        n_in += int(exact.sum())
        n_dropped += int((~passed).sum())
    assert n_in > 1000
    if case == "pions":
        assert n_dropped > 0.5 * 60 * 128 * 128
