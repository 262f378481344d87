#!/usr/bin/env python
"""Golden vectors for the fast reader's formats beyond the iSS one: synthetic `particle_list.dat` (read_in_mode 2,
gzipped UrQMD text; read_in_mode 1, UrQMD file-13 style text), `particle_list.bin` (read_in_mode 21) and
`OSCAR.DAT` (read_in_mode 0) files are pushed through the UNMODIFIED
reference reader (oracle/_ref/ref_driver files: particleSamples + its filter, the loop of
src/Analysis.cpp:817-835) and the filtered particle lists it produced are committed next to the
inputs:

  tests/golden/urqmd_small.particle_list.dat      mode-2 input (gz)
  tests/golden/urqmd_small.particle_list.bin      mode-21 input
  tests/golden/urqmd_small.f13.dat                mode-1 input
  tests/golden/urqmd_small.OSCAR.DAT              mode-0 input
  tests/golden/urqmd_small.iss.bin                mode-9 input (iSS binary)
  tests/golden/urqmd_small.smash.dat              mode-7 input (gzipped SMASH text)
  tests/golden/urqmd_small.f13_3p3.dat / f13_nohdr.dat / jam.dat   mode-4 / mode-3 / mode-5 inputs (3 events each)
  tests/golden/urqmd_small.smash.bin              mode-8 input (extended SMASH binary, 3 events)
  tests/golden/urqmd_<mode>_<case>.particles.bin  HBTIN001 dumps of the reference reader

Run here (needs /root/reference compiled into oracle/_ref):  python tests/golden/make_golden_readers.py
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from hadronic_afterburner_toolkit_b200 import synth  # noqa: E402
from hadronic_afterburner_toolkit_b200.params import HBTParams  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
# case -> (particle_monval, event_buffer_size, rapidity_shift)
CASES = {"pip_all": (211, 100000, 0.0), "pip_buf250": (211, 250, 0.0), "kp_buf1": (321, 1, 0.0), "pim_shift": (-211, 400, 0.37)}


def main():
    assert O.have_reference(), "build oracle/_ref first (make -C oracle ref)"
    rng = np.random.default_rng(20260421)
    events = [ev for b in synth.make_batches(20260421, 2, 4, multiplicity=90) for ev in b.same]
    events.insert(3, events[0][:0])  # an empty event
    records = synth.urqmd_records(events, rng)
    ftxt, fbin = os.path.join(HERE, "urqmd_small.particle_list.dat"), os.path.join(HERE, "urqmd_small.particle_list.bin")
    synth.write_urqmd_gz(ftxt, records)
    synth.write_urqmd_bin(fbin, records)
    ff13, fosc = os.path.join(HERE, "urqmd_small.f13.dat"), os.path.join(HERE, "urqmd_small.OSCAR.DAT")
    # shorter inputs for the two plain-text formats (they are not compressed in the repository)
    short = records[:5]
    synth.write_urqmd_f13(ff13, short)
    synth.write_oscar(fosc, short)
    fiss, fsm = os.path.join(HERE, "urqmd_small.iss.bin"), os.path.join(HERE, "urqmd_small.smash.dat")
    synth.write_iss_bin(fiss, records)
    synth.write_smash_gz(fsm, short)
    f3p3, fnoh, fjam = (os.path.join(HERE, "urqmd_small." + n) for n in ("f13_3p3.dat", "f13_nohdr.dat", "jam.dat"))
    synth.write_urqmd_f13(f3p3, short[:3], header_lines=14)
    synth.write_urqmd_f13(fnoh, short[:3], header_lines=0)
    synth.write_jam(fjam, short[:3])
    fsmb = os.path.join(HERE, "urqmd_small.smash.bin")
    synth.write_smash_bin(fsmb, short[:3], np.random.default_rng(8))
    meta = {}
    for mode, src, name in ((2, ftxt, "particle_list.dat"), (21, fbin, "particle_list.bin"), (1, ff13, "particle_list.dat"),
                            (0, fosc, "OSCAR.DAT"), (9, fiss, "particle_list.bin"), (7, fsm, "particle_list.dat"),
                            (4, f3p3, "particle_list.dat"), (3, fnoh, "particle_list.dat"), (5, fjam, "particle_list.dat"),
                            (8, fsmb, "particles_binary.bin")):
        for case, (monval, buf, shift) in CASES.items():
            with tempfile.TemporaryDirectory() as td:
                os.makedirs(os.path.join(td, "EOS"))
                os.symlink(os.path.join(O.REF_DIR, "EOS", "pdg.dat"), os.path.join(td, "EOS", "pdg.dat"))
                os.makedirs(os.path.join(td, "results"))
                shutil.copy(src, os.path.join(td, "results", name))
                P = HBTParams(qnpts=11, particle_monval=monval, HBTrap_min=-10.0, HBTrap_max=10.0)
                with open(os.path.join(td, "parameters.dat"), "w") as f:
                    f.write(P.parameters_dat(read_in_mode=mode, event_buffer_size=buf, rapidity_shift=shift))
                pout = os.path.join(HERE, f"urqmd_{mode}_{case}.particles.bin")
                r = subprocess.run([O.REF_DRIVER, "files", "parameters.dat", "results", os.path.join(td, "out.bin"), pout],
                                   cwd=td, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
                assert r.returncode == 0, r.stderr.decode()[-2000:]
                meta[f"urqmd_{mode}_{case}"] = {"read_in_mode": mode, "particle_monval": monval, "event_buffer_size": buf,
                                                "rapidity_shift": shift, "input": os.path.basename(src)}
                print(f"urqmd_{mode}_{case}: {os.path.getsize(pout)} bytes")
    json.dump(meta, open(os.path.join(HERE, "reader_cases.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
