#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(oracle/_ref/ref_driver, compiled from /root/reference by oracle/Makefile) on the
reference's own unit-test fixtures (/root/reference/unit_tests/test_gzip_reader and
test_reader_files — config C1 of BASELINE.json) through the reference's own reader loop
(src/Analysis.cpp:817-835).

For each case two files are committed:
  <case>.particles.bin  HBTIN001: the filtered particle lists the reference's reader
                        produced (so the case can be replayed where /root/reference is absent)
  <case>.ref.npz        the reference's raw accumulators (HBTOUT01 -> npz)
plus, for the cases in TEXT_CASES, <case>.<output file name>.gz: the reference's own text
output (first K_T bin; the q_inv file; the last K_phi file) for the output-format check.

Run here (needs /root/reference):  python tests/golden/make_golden.py
"""
import gzip
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from hadronic_afterburner_toolkit_b200 import hbtio  # noqa: E402
from hadronic_afterburner_toolkit_b200.params import C1  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

REF = "/root/reference/unit_tests"
HERE = os.path.dirname(os.path.abspath(__file__))

# case name -> (fixture dir, read_in_mode, files to link, params, extra parameters.dat keys)
UNIT = C1.with_(n_KT=6)  # the K_T grid of unit_tests/parameters.dat
CASES = {
    # C1 proper: pi+ pi+, 1 K_T bin
    "c1_urqmd_gz": ("test_gzip_reader", 2, C1, {}),
    "c1_urqmd_gz_rap10": ("test_gzip_reader", 2, C1.with_(HBTrap_min=-10.0, HBTrap_max=10.0), {}),
    "c1_iss_gz": ("test_gzip_reader", 10, C1, {}),
    "c1_iss_gz_rap10": ("test_gzip_reader", 10, C1.with_(HBTrap_min=-10.0, HBTrap_max=10.0), {}),
    # the switches HBT_unittest.cc:50-89 exercises, on the unit-test K_T grid
    "unit_iss_gz_3d": ("test_gzip_reader", 10, UNIT, {}),
    "unit_iss_gz_inv": ("test_gzip_reader", 10, UNIT.with_(invariant_radius_flag=1), {}),
    "unit_iss_gz_az": ("test_gzip_reader", 10, UNIT.with_(azimuthal_flag=1, n_Kphi=4, n_KT=3, qnpts=21), {}),
    "unit_iss_gz_noboost": ("test_gzip_reader", 10, UNIT.with_(long_comoving_boost=0), {}),
    "unit_iss_gz_cap": ("test_gzip_reader", 10, UNIT.with_(needed_number_of_pairs=50.0), {}),
    "unit_iss_gz_realmixed": ("test_gzip_reader", 10, UNIT, {"read_in_real_mixed_events": 1}),
    "unit_urqmd_txt": ("test_reader_files", 1, UNIT.with_(HBTrap_min=-10.0, HBTrap_max=10.0), {}),
    "unit_oscar_kplus": ("test_reader_files", 0, UNIT.with_(particle_monval=321, HBTrap_min=-10.0, HBTrap_max=10.0), {}),
}


from hadronic_afterburner_toolkit_b200.params import C3, C4  # noqa: E402

SYNTH_CASES = (
    ("synth_c3_small", C3.with_(qnpts=21), 2, 5, 200),
    ("synth_c4_small", C4.with_(qnpts=11), 2, 5, 200),
    ("synth_c3_cap", C3.with_(qnpts=21, needed_number_of_pairs=3000.0), 3, 4, 200),
)
TEXT_CASES = ("c1_iss_gz_rap10", "unit_iss_gz_inv", "unit_iss_gz_az")


def main():
    assert O.have_reference(), "build oracle/_ref first (make -C oracle ref)"
    meta = {}
    for name, (fixdir, mode, P, extra) in CASES.items():
        with tempfile.TemporaryDirectory() as td:
            os.makedirs(os.path.join(td, "EOS"))
            os.symlink(os.path.join(O.REF_DIR, "EOS", "pdg.dat"), os.path.join(td, "EOS", "pdg.dat"))
            res = os.path.join(td, "results")
            os.makedirs(res)
            for fn in os.listdir(os.path.join(REF, fixdir)):
                os.symlink(os.path.realpath(os.path.join(REF, fixdir, fn)), os.path.join(res, fn))
            with open(os.path.join(td, "parameters.dat"), "w") as f:
                f.write(P.parameters_dat(read_in_mode=mode, event_buffer_size=100000, **extra))
            out = os.path.join(td, "out.bin")
            pout = os.path.join(HERE, name + ".particles.bin")
            r = subprocess.run([O.REF_DRIVER, "files", "parameters.dat", "results", out, pout], cwd=td,
                               stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
            assert r.returncode == 0, r.stderr.decode()
            acc = hbtio.read_accumulators(out)
            hbtio.save_accumulators_npz(os.path.join(HERE, name + ".ref.npz"), acc)
            dats = sorted(fn for fn in os.listdir(res) if fn.startswith("HBT_correlation_function"))
            if name in TEXT_CASES:  # keep the text output of a few cases only (size)
                keep = [dats[0]] + [d for d in dats if "_inv_" in d][:1] + [d for d in dats if "_Kphi_" in d][-1:]
                for d in sorted(set(keep)):
                    with open(os.path.join(res, d), "rb") as f, \
                            gzip.GzipFile(os.path.join(HERE, f"{name}.{d}.gz"), "wb", mtime=0) as g:
                        g.write(f.read())
            batches = hbtio.read_batches(pout)
            meta[name] = {
                "fixture": fixdir, "read_in_mode": mode, "extra": extra,
                "params": {k: getattr(P, k) for k in P.__dataclass_fields__},
                "events_per_batch": [[len(e) for e in b.same] for b in batches],
                "real_mixed": batches[0].mixed is not None,
                "psi_ref": acc.psi_ref, "pairs_same": acc.pairs_same,
                "npairs_num": [int(x) for x in acc.npairs_num], "npairs_den": [int(x) for x in acc.npairs_den],
                "dat_files": dats,
            }
            print(name, meta[name]["events_per_batch"], meta[name]["npairs_num"], meta[name]["npairs_den"])
    # small synthetic slices of the C3 / C4 shapes through the reference (mem mode)
    from hadronic_afterburner_toolkit_b200 import synth
    from hadronic_afterburner_toolkit_b200.params import C3, C4
    for name, P, ngrp, nev, mult in SYNTH_CASES:
        batches = synth.make_batches(20260003, ngrp, nev, multiplicity=mult)
        hbtio.write_batches(os.path.join(HERE, name + ".particles.bin"), batches)
        acc = O.run_reference(P, batches)
        hbtio.save_accumulators_npz(os.path.join(HERE, name + ".ref.npz"), acc)
        meta[name] = {"fixture": "synthetic", "params": {k: getattr(P, k) for k in P.__dataclass_fields__},
                      "events_per_batch": [[len(e) for e in b.same] for b in batches], "real_mixed": False,
                      "psi_ref": acc.psi_ref, "pairs_same": acc.pairs_same,
                      "npairs_num": [int(x) for x in acc.npairs_num], "npairs_den": [int(x) for x in acc.npairs_den]}
        print(name, meta[name]["npairs_num"], meta[name]["npairs_den"])
    # psi_ref known answers of unit_tests/HBT_unittest.cc:39,43,47 come from the mode-2 fixture
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
