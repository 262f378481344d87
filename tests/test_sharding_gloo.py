"""N>1 host logic on CPU (gloo, world_size 2): groups dealt round-robin, each rank
fast-forwards the shared RNG stream past the groups it does not own, histograms are summed
with one all-reduce — the result must equal the single-process run bit-for-bit on the integer
accumulators.  The pair loops themselves are done by the CPU oracle here (test infrastructure);
the GPU twin of this test is tests/test_gpu_multi.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

FIELDS = ("num_count", "den_count", "npairs_num", "npairs_den", "num_cos", "sum_qo", "sum_qs", "sum_ql")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hadronic_afterburner_toolkit_b200 import sharding, synth
    from hadronic_afterburner_toolkit_b200.hbt_correlation import Random
    from hadronic_afterburner_toolkit_b200.params import HBTParams
    from oracle import oracle_py as O

    P = HBTParams(qnpts=11, randomSeed=777)
    counts = [(3, 3), (4, 4), (2, 2), (5, 5), (3, 3)]
    batches = [synth.make_batches(90, 1, nev, multiplicity=120, first_group=g)[0] for g, (nev, _) in enumerate(counts)]
    o = O.Oracle(P)
    prod_rng = Random(P.randomSeed)  # the product's host RNG must walk the stream identically
    mine = []
    for g in sharding.walk(rank, world, counts, o):
        o.process_batch(batches[g])
        mine.append(g)
    for g in sharding.walk(rank, world, counts, prod_rng):
        ids, cs, ang = prod_rng.mixed_plan(counts[g][0], counts[g][1], want_angles=True)
        if g == mine[-1]:
            ids_o, ang_o = o.last_plan()  # the oracle's draws for its last own group
            assert np.array_equal(ids, ids_o) and np.array_equal(ang, ang_o)
    assert mine == list(sharding.my_groups(rank, world, len(counts)))
    acc = o.accumulators()
    out = {}
    for k in FIELDS:
        t = torch.from_numpy(np.asarray(getattr(acc, k)).astype(np.float64))
        dist.all_reduce(t)
        out[k] = t.numpy()
    # both ranks end at the same stream position as the single-process run
    after = torch.tensor([o.rand_int_uniform()], dtype=torch.int64)
    gathered = [torch.zeros_like(after) for _ in range(world)]
    dist.all_gather(gathered, after)
    if rank == 0:
        np.savez(os.path.join(outdir, "reduced.npz"), after=np.array([int(x) for x in gathered]), **out)
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    z = np.load(tmp_path / "reduced.npz")
    from hadronic_afterburner_toolkit_b200 import synth
    from hadronic_afterburner_toolkit_b200.params import HBTParams
    from oracle import oracle_py as O

    P = HBTParams(qnpts=11, randomSeed=777)
    counts = [(3, 3), (4, 4), (2, 2), (5, 5), (3, 3)]
    o = O.Oracle(P)
    for g, (nev, _) in enumerate(counts):
        o.process_batch(synth.make_batches(90, 1, nev, multiplicity=120, first_group=g)[0])
    ref = o.accumulators()
    for k in ("num_count", "den_count", "npairs_num", "npairs_den"):
        assert np.array_equal(z[k], np.asarray(getattr(ref, k)).astype(np.float64)), k
    for k in ("num_cos", "sum_qo", "sum_qs", "sum_ql"):
        assert np.allclose(z[k], getattr(ref, k), rtol=1e-12, atol=1e-13), k
    nxt = o.rand_int_uniform()
    assert list(z["after"]) == [nxt, nxt]


def test_round_robin_cover():
    from hadronic_afterburner_toolkit_b200 import sharding

    for world in (1, 2, 3, 8):
        seen = sorted(g for r in range(world) for g in sharding.my_groups(r, world, 200))
        assert seen == list(range(200))
        sizes = [len(sharding.my_groups(r, world, 200)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
