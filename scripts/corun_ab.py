#!/usr/bin/env python
"""Co-residency experiment: the same-event kernel and the v4 mixed-event kernel of one C5-shape group on two
streams (lanes) at the same time, each with a capped number of resident warps per SM, against the same two
kernels one after the other.  One subprocess per setting (the caps are read at hbt_create).

    python scripts/corun_ab.py 18:24:1 9:11:2 10:10:2 ...      (occ_same:occ_mixed4:lanes)
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(tag, reps=4):
    import numpy as np
    import torch

    from hadronic_afterburner_toolkit_b200 import synth
    from hadronic_afterburner_toolkit_b200.hbt_correlation import HBT_correlation, Random, _check, gather_rapidity
    from hadronic_afterburner_toolkit_b200.params import C5, PION_MASS

    P = C5
    nev, mult = 100, 1500
    arr = synth.make_group(20260005, 0, nev, PION_MASS, mult).reshape(nev * mult, 8)
    flat = np.ascontiguousarray(np.concatenate([gather_rapidity(P, arr[e * mult:(e + 1) * mult]) for e in range(nev)]))
    off = np.arange(nev + 1, dtype=np.int64) * mult
    d = torch.from_numpy(flat).cuda()
    h = HBT_correlation(P)
    L, hh = h._L, h._h
    ids, cs = Random(P.randomSeed).mixed_plan(nev, nev)
    nmix = ids.shape[1]
    n = flat.shape[0]
    lanes = int(os.environ.get("CORUN_LANES", "2"))
    _check(hh, L.hbt_set_option(hh, 4, lanes))

    def once():
        _check(hh, L.hbt_accumulate_same_dev(hh, d.data_ptr(), n, 0.0))
        _check(hh, L.hbt_accumulate_mixed_dev(hh, d.data_ptr(), off.ctypes.data, nev, None, None, 0, ids.ctypes.data,
                                              cs.ctypes.data, nmix, 0.0))
        h.synchronize()

    once()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        once()
        ts.append((time.perf_counter() - t0) * 1e3)
    st = h.stage_counters()
    print(json.dumps({"setting": tag, "lanes": lanes, "wall_ms_min": round(min(ts), 3), "wall_ms": [round(t, 3) for t in ts],
                      "accepted_same": int(st[4]) // (reps + 1), "accepted_mixed": int(st[10]) // (reps + 1)}), flush=True)
    h.close()


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--worker":
        return worker(sys.argv[2])
    for s in sys.argv[1:]:
        a, b, lanes = s.split(":")
        env = dict(os.environ, HBT_B200_OCC=a, HBT_B200_OCC4=b, CORUN_LANES=lanes)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", s], env=env, capture_output=True, text=True, timeout=600)
        sys.stdout.write(r.stdout if r.returncode == 0 else json.dumps({"setting": s, "error": r.stderr[-600:]}) + "\n")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
