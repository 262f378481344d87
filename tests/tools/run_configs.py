#!/usr/bin/env python
"""Run BASELINE.json's five configurations on the GPU path and, beside them, the UNMODIFIED
reference (oracle/_ref/ref_driver) on the host cores; check parity on the raw accumulators
(integers bit-exact, sums <= 1e-10) and print one JSON line per config plus a markdown table.

    python tests/tools/run_configs.py [--configs C1,C2,C3,C4,C5] [--c5-groups 200] [--out gpurun_out/configs.json]

The reference is run by a pool of single-threaded processes, one per host core, each owning a
subset of the oversample groups (`only=`: the other groups only advance the shared RNG stream).
C5's 4.5e12 pairs would take ~50 core-hours on the CPU: there the reference processes ONE
full-size group (150 000 pi+, 2.3e10 pairs) for parity, started first so that it overlaps the rest.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from hadronic_afterburner_toolkit_b200 import hbtio, synth  # noqa: E402
from hadronic_afterburner_toolkit_b200.hbt_correlation import HBT_correlation  # noqa: E402
from hadronic_afterburner_toolkit_b200.params import C1, C2, C3, C4, C5, KAON_MASS, PION_MASS  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

CORES = len(os.sched_getaffinity(0))
SUM_FIELDS = ("num_count", "den_count", "npairs_num", "npairs_den", "num_cos", "sum_qo", "sum_qs", "sum_ql")


def start_reference(P, batches, td, same_only=False, nproc=None, groups=None):
    """Spawn the reference processes; returns a handle for collect_reference."""
    fin, fpar = os.path.join(td, "batches.bin"), os.path.join(td, "parameters.dat")
    hbtio.write_batches(fin, batches)
    open(fpar, "w").write(P.parameters_dat())
    groups = list(range(len(batches))) if groups is None else list(groups)
    nproc = min(nproc or CORES, len(groups))
    procs, t0 = [], time.perf_counter()
    for r in range(nproc):
        mine = groups[r::nproc]
        out = os.path.join(td, f"out{r}.bin")
        cmd = [O.REF_DRIVER, "mem", fpar, fin, out, "only=" + ",".join(map(str, mine))] + (["same_only"] if same_only else [])
        procs.append((subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL), out))
    return procs, t0


def collect_reference(handle):
    procs, t0 = handle
    total, cpu_s = None, 0.0
    for p, out in procs:
        assert p.wait() == 0, "ref_driver failed"
    wall = time.perf_counter() - t0
    for p, out in procs:
        acc = hbtio.read_accumulators(out)
        cpu_s += acc.t_total
        os.unlink(out)
        if total is None:
            total = acc
        else:
            for k in SUM_FIELDS:
                setattr(total, k, getattr(total, k) + getattr(acc, k))
    return total, wall, cpu_s, len(procs)


def run_gpu(P, batches, do_mixed=True):
    h = HBT_correlation(P)
    t0 = time.perf_counter()
    for b in batches:
        h.calculate_HBT_correlation_function(b, do_mixed=do_mixed)
    h.synchronize()
    wall = time.perf_counter() - t0
    acc = h.accumulators()
    tm = h.timers()
    res = {"pairs_same": h.pairs_same, "pairs_mixed": h.pairs_mixed, "wall_s": wall,
           "same_ms": tm["same_ms"], "mixed_ms": tm["mixed_ms"], "deferred": h.deferred_pairs()}
    h.close()
    return acc, res


def report(name, desc, res, ref=None, acc=None, ref_wall=None, ref_cpu=None, nproc=None, ref_pairs=None, note=""):
    out = {"config": name, "desc": desc, "gpus": 1, **res}
    ps, pm = res["pairs_same"], res["pairs_mixed"]
    out["gpu_same_pairs_per_s"] = ps / (res["same_ms"] * 1e-3) if res["same_ms"] else None
    out["gpu_mixed_pairs_per_s"] = pm / (res["mixed_ms"] * 1e-3) if res["mixed_ms"] else None
    out["gpu_pairs_per_s_kernels"] = (ps + pm) / ((res["same_ms"] + res["mixed_ms"]) * 1e-3)
    out["gpu_pairs_per_s_wall"] = (ps + pm) / res["wall_s"]
    if ref is not None:
        rep = hbtio.compare(ref, acc, rtol=1e-10)
        out["parity"] = {"counts": "bit-exact", "max_rel_sum_err": max(v["max_rel"] for v in rep.values()),
                         "max_abs_sum_err": max(v["max_abs"] for v in rep.values())}
        out["cpu_1core_pairs_per_s"] = ref_pairs / ref_cpu
        out["cpu_ncore_pairs_per_s"] = ref_pairs / ref_wall
        out["cpu_cores"] = nproc
    out["note"] = note
    print(json.dumps(out), flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C1,C2,C3,C4,C5")
    ap.add_argument("--c5-groups", type=int, default=200)
    ap.add_argument("--c4-events", type=int, default=500)
    ap.add_argument("--c4-qnpts", type=int, default=31)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    ap.add_argument("--no-reference", action="store_true",
                    help="GPU timings only (C2-C4): no reference processes, no parity columns")
    a = ap.parse_args()
    want = a.configs.split(",")
    rows = []
    tmp = tempfile.TemporaryDirectory()
    assert a.no_reference or O.have_reference(), "oracle/_ref missing"

    c5_handle = None
    if "C5" in want and not a.no_reference:  # one full-size group on the reference, overlapping everything else
        b0 = synth.make_batches(20260005, 1, 100, PION_MASS)
        td = os.path.join(tmp.name, "c5ref"); os.makedirs(td)
        c5_handle = start_reference(C5, b0, td, nproc=1)

    if "C1" in want:
        import glob
        for f in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "c1_*.ref.npz"))):
            name = os.path.basename(f)[:-8]
            meta = json.load(open(os.path.join(ROOT, "tests", "golden", "cases.json")))[name]
            from hadronic_afterburner_toolkit_b200.params import HBTParams
            P = HBTParams(**meta["params"])
            batches = hbtio.read_batches(os.path.join(ROOT, "tests", "golden", name + ".particles.bin"))
            ref = hbtio.load_accumulators_npz(f)
            acc, res = run_gpu(P, batches)
            rep = hbtio.compare(ref, acc, rtol=1e-10)
            o = report("C1:" + name, "unit-test fixture, reference golden vectors", res, note="tiny: timings are launch latency")
            o["parity"] = {"counts": "bit-exact", "max_rel_sum_err": max(v["max_rel"] for v in rep.values())}
            rows.append(o)

    for name, P, mixed in (("C2", C2, False), ("C3", C3, True)):
        if name not in want:
            continue
        batches = synth.make_batches(20260000 + int(name[1]), 100, 10, PION_MASS)
        td = os.path.join(tmp.name, name); os.makedirs(td)
        desc = "1000 ev x 1500 pi+, oversampling 10, 41^3, 4 K_T bins, same" + ("+mixed" if mixed else " only")
        # the GPU run first: the reference's processes would take the cores away from the submitting thread
        acc, res = run_gpu(P, batches, do_mixed=mixed)
        if a.no_reference:
            rows.append(report(name, desc, res))
            continue
        hd = start_reference(P, batches, td, same_only=not mixed, nproc=CORES - (1 if c5_handle else 0))
        ref, wall, cpu, nproc = collect_reference(hd)
        rows.append(report(name, desc, res, ref, acc, wall, cpu, nproc, res["pairs_same"] + res["pairs_mixed"]))

    if "C4" in want:
        for sp, mass, mon in (("pi+", PION_MASS, 211), ("K+", KAON_MASS, 321)):
            P = C4.with_(qnpts=a.c4_qnpts, particle_monval=mon)
            ng = a.c4_events // 50
            batches = synth.make_batches(20260004 + mon, ng, 50, mass)
            td = os.path.join(tmp.name, "C4" + sp); os.makedirs(td)
            desc = f"{a.c4_events} ev x 1500 {sp}, oversampling 50, 8 K_T x 8 K_phi bins, {a.c4_qnpts}^3, same+mixed"
            acc, res = run_gpu(P, batches)
            if a.no_reference:
                rows.append(report("C4:" + sp, desc, res))
                continue
            hd = start_reference(P, batches, td, nproc=CORES - (1 if c5_handle else 0))
            ref, wall, cpu, nproc = collect_reference(hd)
            rows.append(report("C4:" + sp, desc, res, ref, acc, wall, cpu, nproc, res["pairs_same"] + res["pairs_mixed"]))

    if "C5" in want and not a.no_reference:
        # the full run: groups generated on the fly (1.9 GB would not be kept at once)
        h = HBT_correlation(C5)
        t0 = time.perf_counter(); tgen = 0.0
        for g in range(a.c5_groups):
            tg = time.perf_counter()
            b = synth.make_batches(20260005, 1, 100, PION_MASS, first_group=g)[0]
            tgen += time.perf_counter() - tg
            h.calculate_HBT_correlation_function(b)
        h.synchronize()
        wall = time.perf_counter() - t0  # includes the host-side generation, which overlaps the GPU work
        tm = h.timers()
        res = {"pairs_same": h.pairs_same, "pairs_mixed": h.pairs_mixed, "wall_s": wall, "same_ms": tm["same_ms"],
               "mixed_ms": tm["mixed_ms"], "deferred": h.deferred_pairs()}
        acc_full = h.accumulators()
        st = acc_full.stage
        assert int(st[0]) == h.pairs_same and int(st[6]) == h.pairs_mixed
        assert int(acc_full.num_count.sum()) == int(acc_full.npairs_num.sum()) == int(st[5])
        h.close()
        # parity of the first group against the reference's full-size run of it
        acc0, res0 = run_gpu(C5, synth.make_batches(20260005, 1, 100, PION_MASS))
        ref, rwall, rcpu, nproc = collect_reference(c5_handle)
        o = report("C5", f"{a.c5_groups} groups of 100 ev x 1500 pi+ (oversampling 100), 41^3, 4 K_T bins, same+mixed; "
                   "parity on group 0 at full size (2.27e10 pairs) vs the reference", res, ref, acc0, rwall, rcpu, nproc,
                   res0["pairs_same"] + res0["pairs_mixed"], note=f"wall_s includes {tgen:.1f} s of host-side synthetic generation (overlapped with the GPU work)")
        rows.append(o)

    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rows, open(a.out, "w"), indent=1)
    print("\n| config | same pairs/s | mixed pairs/s | same+mixed pairs/s (kernels) | wall pairs/s | CPU 1-core | CPU N-core | parity |")
    print("|---|---|---|---|---|---|---|---|")
    f = lambda x: "—" if x is None else f"{x:.3g}"
    for o in rows:
        par = o.get("parity", {})
        print(f"| {o['config']} | {f(o['gpu_same_pairs_per_s'])} | {f(o['gpu_mixed_pairs_per_s'])} | {f(o['gpu_pairs_per_s_kernels'])} | "
              f"{f(o['gpu_pairs_per_s_wall'])} | {f(o.get('cpu_1core_pairs_per_s'))} | {f(o.get('cpu_ncore_pairs_per_s'))} ({o.get('cpu_cores', '—')}) | "
              f"{par.get('counts', '—')} / {f(par.get('max_rel_sum_err'))} |")


if __name__ == "__main__":
    main()
