"""The error bounds behind the FP32 decision of mixed-event survivors (v3_mixed_f32 in
hadronic_afterburner_toolkit_b200/csrc/hbt_kernels_v3.cuh), checked on the CPU: the same binary32
chain is evaluated with numpy (every operation rounded to float32, fma as one rounding, the reciprocal square
root perturbed by the 2^-22 the hardware approximation is allowed) on millions of pairs — the benchmark
distribution, momentum outliers, tiny momenta, nearly back-to-back pairs (small K_T), large rapidities — and
compared with the binary64 values.  The kernel's band is TWICE the bound tested here, plus the FP64 path's own
guard."""
import numpy as np
import pytest

U = np.float32(2.0 ** -24)
F = np.float32


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def rsq(x, rng):
    eps = rng.uniform(-2.0 ** -22, 2.0 ** -22, size=x.shape)
    return ((1.0 / np.sqrt(x.astype(np.float64))) * (1.0 + eps)).astype(np.float32)


def sample(rng, n, scale=1.0, ymax=0.5, back_to_back=False, mass=0.138):
    def particles():
        px, py = rng.normal(0, 0.37, n) * scale, rng.normal(0, 0.33, n) * scale
        y = rng.uniform(-ymax, ymax, n)
        mT = np.sqrt((mass * scale) ** 2 + px * px + py * py)
        return px, py, mT * np.sinh(y), mT * np.cosh(y)
    a, b = particles(), particles()
    if back_to_back:  # K_perp tiny against the momenta: the conditioning the bound has to follow
        eps = 10.0 ** rng.uniform(-5, -1, n)
        b = (-a[0] * (1 + eps), -a[1] * (1 - eps), b[2], np.sqrt((mass * scale) ** 2 + (a[0] * (1 + eps)) ** 2
                                                              + (a[1] * (1 - eps)) ** 2 + b[2] ** 2))
    return a, b


CASES = {"benchmark": {}, "outliers_x30": {"scale": 30.0}, "tiny_1e-3": {"scale": 1e-3}, "rapidity_3": {"ymax": 3.0},
         "back_to_back": {"back_to_back": True}, "kaons": {"mass": 0.494}}


@pytest.mark.parametrize("case", sorted(CASES))
def test_float_chain_stays_inside_its_bound(case):
    rng = np.random.default_rng(20260500 + sorted(CASES).index(case))
    n = 400000
    (ax, ay, az, aE), (bx, by, bz, bE) = sample(rng, n, **CASES[case])
    # binary64 reference values (the double chain's own rounding, ~1e-16, is far below the float bounds)
    sx, sy, qx, qy = ax + bx, ay + by, ax - bx, ay - by
    k2 = sx * sx + sy * sy
    qo, qs = (qx * sx + qy * sy) / np.sqrt(k2), (qy * sx - qx * sy) / np.sqrt(k2)
    sz, sE, qz, qE = az + bz, aE + bE, az - bz, aE - bE
    ql = (sE * qz - sz * qE) / np.sqrt((sE - sz) * (sE + sz))
    # the binary32 chain of v3_mixed_f32
    fax, fay, faz, faE, fbx, fby, fbz, fbE = (v.astype(np.float32) for v in (ax, ay, az, aE, bx, by, bz, bE))
    fsx, fsy, fqx, fqy = fax + fbx, fay + fby, fax - fbx, fay - fby
    S = (np.abs(fax) + np.abs(fbx)) + (np.abs(fay) + np.abs(fby))
    S2 = S * S
    fk2 = fma(fsy, fsy, fsx * fsx)
    r = rsq(fk2, rng)
    d, e = fma(fqx, fsx, fqy * fsy), fma(fqy, fsx, -(fqx * fsy))
    fqo, fqs = d * r, e * r
    A = S2 * r
    R = A * r
    c6, A6 = fma(F(2) * np.ones_like(R), R, F(6) * np.ones_like(R)), F(6) * A
    bound_o = U * fma(np.abs(fqo), c6, A6)
    bound_s = U * fma(np.abs(fqs), c6, A6)
    bound_k = U * (F(4) * S2 + F(2) * fk2)
    fsz, fsE, fqz, fqE = faz + fbz, faE + fbE, faz - fbz, faE - fbE
    Z = np.abs(faz) + np.abs(fbz)
    m2 = (fsE - fsz) * (fsE + fsz)
    r2 = rsq(m2, rng)
    t1, t2 = fsE * fqz, fsz * fqE
    fql = (t1 - t2) * r2
    W = ((np.abs(faE) + np.abs(fbE)) + Z) * r2
    bound_l = U * fma(np.abs(fql), fma(F(2) * W, W, F(8) * np.ones_like(W)), F(10) * W * Z)
    ok = np.isfinite(fqo) & np.isfinite(fql) & (k2 > 0)
    assert ok.mean() > 0.999
    worst = {}
    for name, got, ref, bound in (("k2", fk2, k2, bound_k), ("q_out", fqo, qo, bound_o), ("q_side", fqs, qs, bound_s),
                                  ("q_long", fql, ql, bound_l)):
        err = np.abs(got.astype(np.float64) - ref)[ok]
        ratio = err / bound.astype(np.float64)[ok]
        worst[name] = float(ratio.max())
        assert ratio.max() <= 1.0, (case, name, worst)
    # the bounds are not vacuous either: within ~two orders of magnitude of what actually happens
    assert max(worst.values()) > 0.01, worst


def test_band_in_bin_units_on_the_benchmark_sample():
    """What the band costs: on the benchmark distribution twice the bound is a few 1e-4 of a 0.01 GeV bin, so
    ~1e-3 of the survivors fall through to the FP64 path."""
    rng = np.random.default_rng(1)
    (ax, ay, az, aE), (bx, by, bz, bE) = sample(rng, 200000)
    S = (np.abs(ax) + np.abs(bx)) + (np.abs(ay) + np.abs(by))
    k2 = (ax + bx) ** 2 + (ay + by) ** 2
    sel = (k2 > 4 * 0.15 ** 2) & (k2 < 4 * 0.55 ** 2)
    r = 1 / np.sqrt(k2)
    band = 2 * 2.0 ** -24 * (6 * S * S * r + 0.2 * (6 + 2 * S * S * r * r)) / 0.01
    assert 1e-5 < np.median(band[sel]) < 1e-3
