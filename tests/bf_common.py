"""Shared helpers of the BalanceFunction tests: the golden input as the reference's reader hands it
to the operator (events grouped by event_buffer_size, species lists with pT, phi_p, rap_y, rap_eta)."""
import gzip
import json
import os

import numpy as np

from oracle import bf_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = json.load(open(os.path.join(GOLDEN, "bf_cases.json")))


def read_all_species(path):
    """events of the mode-10 text: list of arrays [n, 10] = monval mass t x y z E px py pz"""
    events = []
    with gzip.open(path, "rt") as f:
        while True:
            head = f.readline()
            if not head.strip():
                break
            n = int(head.split()[0])
            events.append(np.array([[float(v) for v in f.readline().split()] for _ in range(n)]).reshape(n, 10))
    return events


def batches_of(events, event_buffer_size):
    """read_in_particle_samples_gzipped's grouping rule (src/particleSamples.cpp:1256-1284)"""
    out, cur, num = [], [], 0
    for ev in events:
        cur.append(ev)
        num += len(ev)
        if num >= event_buffer_size:
            out.append(cur)
            cur, num = [], 0
    if cur:
        out.append(cur)
    return out


def species_lists(batch, alpha, beta):
    """get_balance_function_particle_list_{a,abar,b,bbar}: single-species filters of the batch
    (src/particleSamples.cpp:66-69, 1315-1326) with the kinematics of boostParticles"""
    def one(monval):
        lst = []
        for ev in batch:
            s = ev[ev[:, 0] == monval]
            pT, phi, ry, reta = bf_oracle.kinematics(s[:, 7], s[:, 8], s[:, 9], s[:, 6], s[:, 1])
            lst.append({"pT": pT, "phi": phi, "rap_y": ry, "rap_eta": reta})
        return lst
    return {"a": one(alpha), "abar": one(-alpha), "b": one(beta), "bbar": one(-beta)}


def golden_text(name):
    return {fn: open(os.path.join(GOLDEN, f"{name}.{fn}")).read() for fn in CASES[name]["files"]}
