#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Full-size golden accumulators for BASELINE.json's config 5 (and the reduced complete run
SURVEY.md 8d asks for), produced OFFLINE by the UNMODIFIED reference (oracle/_ref/ref_driver, compiled from
/root/reference by oracle/Makefile) in the build container:

  * c5_group000 / c5_group099 / c5_group199 : single oversample groups of the 200-group config-5 run
    (100 events x 1500 pi+, 1.125e10 same-event + 1.1475e10 mixed-event pairs each), each at ITS position in the
    reference's RNG stream (the draws of the other groups are replayed, `only=`);
  * c5_2000ev : the complete config-5 run reduced to 2 000 events (groups 0..19): sum of the 20 per-group dumps
    (integers exact; the f64 sums differ from a single sequential run by ~1e-16 relative, far inside 1e-10).

Every group costs ~15 min of one host core.  Stored compactly (tests/golden/*.fs.npz): the two count histograms in
full (compressed u32), the four f64 sum histograms on every 8th bin (offset = group % 8) plus their per-slab totals.

    python tests/golden/make_golden_fullsize.py [--procs 6]
"""
import argparse
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from hadronic_afterburner_toolkit_b200 import hbtio, synth  # noqa: E402
from hadronic_afterburner_toolkit_b200.params import C5, PION_MASS  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

SEED = 20260005
NGROUPS = 200
SAMPLED = (0, 99, 199)
REDUCED = tuple(range(20))
STRIDE = 8


def compact(acc, offset):
    """What the golden keeps of a full accumulator set."""
    nb = acc.num_count.size
    sel = np.arange(offset % STRIDE, nb, STRIDE)
    q3 = acc.qnpts ** 3
    d = {"num_count": np.asarray(acc.num_count).astype(np.uint32), "den_count": np.asarray(acc.den_count).astype(np.uint32),
         "npairs_num": np.asarray(acc.npairs_num, dtype=np.uint64), "npairs_den": np.asarray(acc.npairs_den, dtype=np.uint64),
         "sel_offset": np.int64(offset % STRIDE), "sel_stride": np.int64(STRIDE)}
    assert np.array_equal(d["num_count"].astype(np.float64), np.asarray(acc.num_count, dtype=np.float64))
    for k in ("num_cos", "sum_qo", "sum_qs", "sum_ql"):
        a = np.asarray(getattr(acc, k), dtype=np.float64)
        d[k + "_sel"] = a[sel]
        d[k + "_slab"] = np.array([np.sum(a[s * q3:(s + 1) * q3].astype(np.longdouble)) for s in range(nb // q3)], dtype=np.float64)
        d[k + "_slab_abs"] = np.array([np.sum(np.abs(a[s * q3:(s + 1) * q3]).astype(np.longdouble)) for s in range(nb // q3)], dtype=np.float64)
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=6)
    a = ap.parse_args()
    assert O.have_reference(), "oracle/_ref missing: run oracle/Makefile where /root/reference is present"
    groups = sorted(set(SAMPLED) | set(REDUCED))
    td = tempfile.mkdtemp(prefix="hbt_fullsize_")
    fpar = os.path.join(td, "parameters.dat")
    open(fpar, "w").write(C5.parameters_dat())
    empty = hbtio.Batch([np.zeros((0, 8)) for _ in range(100)])
    pending, running, done = list(groups), [], {}
    t0 = time.time()
    while pending or running:
        while pending and len(running) < a.procs:
            g = pending.pop(0)
            # all 200 batches so that the reference's RNG stream is at the right position; the others hold empty
            # events (their draws depend on the event count only, src/HBT_correlation.cpp:200-215)
            batches = [empty] * g + synth.make_batches(SEED, 1, 100, PION_MASS, first_group=g)
            fin, fout = os.path.join(td, f"in{g}.bin"), os.path.join(td, f"out{g}.bin")
            hbtio.write_batches(fin, batches)
            p = subprocess.Popen(["nice", "-n", "15", O.REF_DRIVER, "mem", fpar, fin, fout, f"only={g}"],
                                 stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            running.append((g, p, fin, fout))
        time.sleep(5)
        for item in list(running):
            g, p, fin, fout = item
            if p.poll() is None:
                continue
            assert p.returncode == 0, f"ref_driver failed on group {g}"
            running.remove(item)
            done[g] = hbtio.read_accumulators(fout)
            os.unlink(fin)
            print(f"group {g} done after {time.time() - t0:.0f} s ({done[g].t_total:.0f} s in the reference's loops)", flush=True)
            if g in SAMPLED:
                np.savez_compressed(os.path.join(HERE, f"c5_group{g:03d}.fs.npz"), **compact(done[g], g))
    total = None
    for g in REDUCED:
        acc = done[g]
        if total is None:
            total = acc
        else:
            for k in ("num_count", "den_count", "npairs_num", "npairs_den", "num_cos", "sum_qo", "sum_qs", "sum_ql"):
                setattr(total, k, getattr(total, k) + getattr(acc, k))
    np.savez_compressed(os.path.join(HERE, "c5_2000ev.fs.npz"), **compact(total, 0))
    print("cpu seconds in the reference's pair loops:", sum(x.t_total for x in done.values()))


if __name__ == "__main__":
    main()
