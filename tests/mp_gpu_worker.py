"""Worker of tests/test_gpu_multi.py: one rank per GPU (torchrun), groups sharded round-robin,
library-owned NCCL communicator, one all-reduce of the histograms, rank 0 checks the global
result against the single-process CPU oracle."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hadronic_afterburner_toolkit_b200 import capi, hbtio, sharding, synth  # noqa: E402
from hadronic_afterburner_toolkit_b200.hbt_correlation import HBT_correlation, _check  # noqa: E402
from hadronic_afterburner_toolkit_b200.params import HBTParams  # noqa: E402


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    L = capi.lib()
    P = HBTParams(qnpts=21, randomSeed=4242)
    counts = [(4, 4), (3, 3), (5, 5), (4, 4), (2, 2), (4, 4), (3, 3)]
    batches = [synth.make_batches(91, 1, nev, multiplicity=400, first_group=g)[0] for g, (nev, _) in enumerate(counts)]
    eng = HBT_correlation(P, device=local)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        _check(None, L.hbt_comm_unique_id(buf))
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    dist.broadcast(uid, 0)
    _check(eng._h, L.hbt_comm_init_rank(eng._h, world, rank, ctypes.create_string_buffer(uid.numpy().tobytes(), 128)))
    for g in sharding.walk(rank, world, counts, eng.ran_gen):
        eng.calculate_HBT_correlation_function(batches[g])
    _check(eng._h, L.hbt_allreduce(eng._h))
    acc = eng.accumulators()  # every rank now reads the global sums
    tot = torch.tensor([float(acc.num_count.sum()), float(acc.den_count.sum())], dtype=torch.float64)
    lst = [torch.zeros_like(tot) for _ in range(world)]
    dist.all_gather(lst, tot)
    assert all(torch.equal(x, lst[0]) for x in lst), "ranks disagree on the reduced histograms"
    if rank == 0:
        from oracle import oracle_py as O

        o = O.Oracle(P)
        for b in batches:
            o.process_batch(b)
        hbtio.compare(o.accumulators(), acc, rtol=1e-10, check_stage="cheap")
        # more batches after an all-reduce keep accumulating locally (no double counting)
    eng.calculate_HBT_correlation_function(batches[0]) if rank == 0 else None
    _check(eng._h, L.hbt_allreduce(eng._h))
    acc2 = eng.accumulators()
    if rank == 0:
        from oracle import oracle_py as O

        o = O.Oracle(P)
        for b in batches:
            o.process_batch(b)
        extra = O.Oracle(P)
        # rank 0's stream position after its shard differs from the oracle's; only the counts of
        # the same-event part are deterministic here
        n = sum(len(e) for e in batches[0].same)
        assert int(acc2.stage[0]) == int(acc.stage[0]) + n * (n - 1) // 2
        open(os.environ["HBT_MP_OK"], "w").write("ok %d ranks\n" % world)
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
