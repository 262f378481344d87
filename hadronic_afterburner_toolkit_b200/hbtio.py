"""Binary interchange formats shared by the product bindings, the tests and bench.py.

HBTIN001 — batches of events ("oversample groups"), each event n x 8 float64
           (px,py,pz,E,x,y,z,t): the 64 hot bytes of ``particle_info``
           (``src/particle_info.h:5-12``) in the order the pair loops read them.
HBTOUT01 — raw accumulators of one run (counts, sums, per-K counters).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


@dataclass
class Batch:
    """One oversample group: a list of events, optionally with real mixed events."""

    same: List[np.ndarray]
    mixed: Optional[List[np.ndarray]] = None

    def flat(self, which: str = "same"):
        evs = self.same if which == "same" else self.mixed
        off = np.zeros(len(evs) + 1, dtype=np.int64)
        if evs:
            off[1:] = np.cumsum([len(e) for e in evs])
            p = np.ascontiguousarray(np.concatenate([e.reshape(-1, 8) for e in evs]), dtype=np.float64)
        else:
            p = np.zeros((0, 8), dtype=np.float64)
        return p, off


def write_batches(path, batches: List[Batch]) -> None:
    with open(path, "wb") as f:
        f.write(b"HBTIN001")
        f.write(struct.pack("<i", len(batches)))
        for b in batches:
            mixed = b.mixed or []
            f.write(struct.pack("<ii", len(b.same), len(mixed)))
            for lst in (b.same, mixed):
                for ev in lst:
                    ev = np.ascontiguousarray(ev, dtype="<f8").reshape(-1, 8)
                    f.write(struct.pack("<i", ev.shape[0]))
                    f.write(ev.tobytes())


def read_batches(path) -> List[Batch]:
    with open(path, "rb") as f:
        data = f.read()
    assert data[:8] == b"HBTIN001", "bad magic"
    pos = 8
    (nb,) = struct.unpack_from("<i", data, pos)
    pos += 4
    out = []
    for _ in range(nb):
        nev, nmx = struct.unpack_from("<ii", data, pos)
        pos += 8
        lists = []
        for cnt in (nev, nmx):
            evs = []
            for _ in range(cnt):
                (n,) = struct.unpack_from("<i", data, pos)
                pos += 4
                evs.append(np.frombuffer(data, dtype="<f8", count=n * 8, offset=pos).reshape(n, 8).copy())
                pos += n * 64
            lists.append(evs)
        out.append(Batch(lists[0], lists[1] if nmx else None))
    return out


@dataclass
class Accumulators:
    """Raw accumulators.  3-D histograms are flat ``[slab][q_out][q_side][q_long]`` with
    ``slab = K`` (azimuthal_flag=0) or ``K*n_Kphi + phi`` (azimuthal_flag=1)."""

    azimuthal_flag: int
    invariant_radius_flag: int
    n_KT: int
    n_Kphi: int
    qnpts: int
    num_count: np.ndarray  # integer-valued
    num_cos: np.ndarray
    sum_qo: np.ndarray
    sum_qs: np.ndarray
    sum_ql: np.ndarray
    den_count: np.ndarray  # integer-valued
    npairs_num: np.ndarray  # uint64 per slab
    npairs_den: np.ndarray
    qinv_count: Optional[np.ndarray] = None
    qinv_mean: Optional[np.ndarray] = None
    qinv_num: Optional[np.ndarray] = None
    qinv_den: Optional[np.ndarray] = None
    npairs_num_qinv: Optional[np.ndarray] = None
    npairs_den_qinv: Optional[np.ndarray] = None
    psi_ref: List[float] = field(default_factory=list)
    stage: Optional[np.ndarray] = None  # uint64[12]
    t_same: float = 0.0
    t_total: float = 0.0
    pairs_same: int = 0


def read_accumulators(path) -> Accumulators:
    """Parse an HBTOUT01 dump written by oracle/ref_driver.cpp."""
    with open(path, "rb") as f:
        data = f.read()
    assert data[:8] == b"HBTOUT01", "bad magic"
    pos = 8
    az, qinv, nK, nP, nq, npsi = struct.unpack_from("<6i", data, pos)
    pos += 24
    psi = list(np.frombuffer(data, "<f8", npsi, pos))
    pos += 8 * npsi
    t_same, t_total = struct.unpack_from("<2d", data, pos)
    pos += 16
    pairs_same, _ = struct.unpack_from("<2Q", data, pos)
    pos += 16

    def take(dtype, n):
        nonlocal pos
        a = np.frombuffer(data, dtype, n, pos).copy()
        pos += a.nbytes
        return a

    cK = take("<u8", 4 * nK)
    num_K, den_K, num_Kq, den_Kq = cK[:nK], cK[nK:2 * nK], cK[2 * nK:3 * nK], cK[3 * nK:]
    if az == 1:
        c2 = take("<u8", 2 * nK * nP)
        npairs_num, npairs_den = c2[:nK * nP], c2[nK * nP:]
        nslab = nK * nP
    else:
        npairs_num, npairs_den = num_K, den_K
        nslab = nK
    nb = nslab * nq ** 3
    arrs = [take("<f8", nb) for _ in range(6)]
    acc = Accumulators(az, qinv, nK, nP, nq, arrs[0], arrs[1], arrs[2], arrs[3], arrs[4], arrs[5],
                       npairs_num, npairs_den, psi_ref=psi, t_same=t_same, t_total=t_total,
                       pairs_same=pairs_same)
    if qinv == 1:
        acc.qinv_count, acc.qinv_mean, acc.qinv_num, acc.qinv_den = (take("<f8", nK * nq) for _ in range(4))
        acc.npairs_num_qinv, acc.npairs_den_qinv = num_Kq, den_Kq
    assert pos == len(data), "trailing bytes in HBTOUT01"
    return acc


def save_accumulators_npz(path, acc: Accumulators) -> None:
    d = {k: v for k, v in acc.__dict__.items() if v is not None}
    d["psi_ref"] = np.asarray(acc.psi_ref, dtype=np.float64)
    np.savez_compressed(path, **d)


def load_accumulators_npz(path) -> Accumulators:
    z = np.load(path)
    kw = {}
    for k in z.files:
        v = z[k]
        kw[k] = v.item() if v.shape == () else v
    kw["psi_ref"] = list(kw.get("psi_ref", []))
    return Accumulators(**kw)


def compare(ref: Accumulators, got: Accumulators, rtol: float = 1e-10, check_stage: bool = False,
            q_scale: float = 0.25):
    """Parity protocol of SURVEY.md §8(c).

    * integer accumulators (bin counts, per-K pair counters, stage counters): bit-exact;
    * floating sums (Σcos, Σq_out, Σq_side, Σq_long): per bin
      ``|got-ref| <= rtol * max(|ref|, 1e-4 * count * scale)`` — a pure relative test at
      ``rtol`` = 1e-10 wherever the sum is not cancelled below 1e-4 of its natural size
      (``count`` terms of magnitude <= ``scale``: 1 for cos, ~q_max for the q sums), and an
      absolute 1e-14*count*scale floor for the ill-conditioned near-zero sums.
    Returns the worst deviations; raises AssertionError on any mismatch."""
    report = {}
    for name in ("num_count", "den_count", "npairs_num", "npairs_den"):
        a, b = np.asarray(getattr(ref, name)), np.asarray(getattr(got, name))
        assert a.shape == b.shape, f"{name}: shape {a.shape} vs {b.shape}"
        bad = np.flatnonzero(a.astype(np.float64) != b.astype(np.float64))
        assert bad.size == 0, (f"{name}: {bad.size} entries differ, first {bad[:5]} "
                               f"ref={a[bad[:5]]} got={b[bad[:5]]}")
    cnt = np.asarray(ref.num_count, dtype=np.float64)
    for name, scale in (("num_cos", 1.0), ("sum_qo", q_scale), ("sum_qs", q_scale), ("sum_ql", q_scale)):
        a, b = np.asarray(getattr(ref, name)), np.asarray(getattr(got, name))
        d = np.abs(a - b)
        tol = rtol * np.maximum(np.abs(a), 1e-4 * cnt * scale)
        nz = cnt > 0
        rel = d[nz] / np.maximum(np.abs(a[nz]), 1e-300)
        # bins that the pure relative test alone would reject and only the conditioning floor accepts, and the
        # deviation per accumulated pair (SURVEY.md 8c asks for both to be reported)
        floor_bins = int(np.sum(nz & (d > rtol * np.abs(a)) & (d <= tol)))
        per_pair = d[nz] / cnt[nz]
        report[name] = {"max_abs": float(d.max()) if d.size else 0.0,
                        "max_rel": float(rel.max()) if rel.size else 0.0,
                        "floor_bins": floor_bins, "bins": int(nz.sum()),
                        "max_abs_per_pair": float(per_pair.max()) if per_pair.size else 0.0}
        if not np.all(d <= tol):
            bad = np.flatnonzero(~(d <= tol).ravel())
            k = bad[0]
            raise AssertionError(f"{name}: {bad.size} bins exceed tolerance, max_abs={d.max():.3e} "
                                 f"max_rel={report[name]['max_rel']:.3e}; first bin {k}: count={cnt.ravel()[k]:.0f} "
                                 f"ref={a.ravel()[k]!r} got={b.ravel()[k]!r} tol={tol.ravel()[k]:.3e}")
    if ref.invariant_radius_flag == 1:
        for name in ("qinv_count", "qinv_den", "npairs_num_qinv", "npairs_den_qinv"):
            a, b = np.asarray(getattr(ref, name)), np.asarray(getattr(got, name))
            assert np.array_equal(a.astype(np.float64), b.astype(np.float64)), f"{name} differs"
        c1 = np.asarray(ref.qinv_count, dtype=np.float64)
        for name, scale in (("qinv_mean", q_scale), ("qinv_num", 1.0)):
            a, b = np.asarray(getattr(ref, name)), np.asarray(getattr(got, name))
            tol = rtol * np.maximum(np.abs(a), 1e-4 * c1 * scale)
            assert np.all(np.abs(a - b) <= tol), f"{name}: exceeds tolerance"
    if check_stage and ref.stage is not None and got.stage is not None:
        # "cheap": the populations the production kernels keep (all pairs, passed q_long, accepted);
        # True: all six per loop (needs HBT_OPT_STAGE_COUNTERS on the product side)
        idx = [0, 4, 5, 6, 10, 11] if check_stage == "cheap" else list(range(12))
        a, b = np.asarray(ref.stage)[idx], np.asarray(got.stage)[idx]
        assert np.array_equal(a, b), f"stage counters differ: ref={ref.stage} got={got.stage}"
    return report


class FastReader:
    """``hbt_reader_*`` of ``include/hbt_b200.h``: the particle samples of ``read_in_mode=10``
    (gzipped iSS text, ``src/particleSamples.cpp:1247-1286``), ``2`` (gzipped UrQMD text, ``:910-974``)
    or ``21`` (UrQMD binary, ``:976-1059``) as batches of events of one
    species, grouped by the reference's ``event_buffer_size`` rule.  Iterating yields
    :class:`Batch` objects (copies); ``all_particles`` holds the all-species count of the last one."""

    def __init__(self, path: str, particle_monval: int, event_buffer_size: int, rap_shift: float = 0.0,
                 rapidity_cut=None, read_in_mode: int = 10):
        import ctypes

        from . import capi

        self._ct = ctypes
        self._L = capi.lib()
        self._cut = rapidity_cut.to_c() if rapidity_cut is not None else None
        h = ctypes.c_void_p()
        rc = self._L.hbt_reader_open(str(path).encode(), read_in_mode, particle_monval, event_buffer_size, rap_shift,
                                     ctypes.byref(self._cut) if self._cut is not None else None, ctypes.byref(h))
        if rc != 0:
            raise capi.HBTError(rc, f"hbt_reader_open({path}, mode {read_in_mode}, monval {particle_monval})")
        self._h = h
        self.all_particles = 0

    def __iter__(self):
        return self

    def __next__(self) -> Batch:
        ct = self._ct
        p, off, allp = ct.POINTER(ct.c_double)(), ct.POINTER(ct.c_int64)(), ct.c_int64()
        nev = self._L.hbt_reader_next(self._h, ct.byref(p), ct.byref(off), ct.byref(allp))
        if nev < 0:
            from . import capi
            raise capi.HBTError(nev, self._L.hbt_reader_error(self._h).decode())
        if nev == 0:
            raise StopIteration
        offs = np.ctypeslib.as_array(off, shape=(nev + 1,)).copy()
        flat = np.ctypeslib.as_array(p, shape=(int(offs[-1]), 8)).copy() if offs[-1] else np.zeros((0, 8))
        self.all_particles = allp.value
        return Batch([flat[offs[i]:offs[i + 1]] for i in range(nev)])

    def bytes_read(self) -> int:
        return int(self._L.hbt_reader_bytes(self._h))

    def close(self) -> None:
        if self._h:
            self._L.hbt_reader_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
