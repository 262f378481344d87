#!/usr/bin/env python
"""bench.py — HBT pairs/sec (same+mixed) on N B200s, with the FP64 roofline and the reference's
own CPU implementation timed beside it.

    python bench.py [--gpus N --steps K --warmup W]            our arm (libhbt_b200.so via the C ABI)
    python bench.py --impl reference [...]                      the reference's CPU path (oracle/_ref)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): oversample groups of BASELINE.json's config 5 shape —
`--events-per-group` (100) synthetic central events x 1500 pi+, 41^3 q grid, 4 K_T bins,
same-event + mixed-event pairs; ONE STEP = `--groups-per-gpu` such groups on every GPU (weak
scaling: each rank owns distinct groups, exactly how config 5 shards its 200 groups; no
data-path collective) followed by the single NCCL all-reduce of the histograms.

    --config c2|c3|c4|c4k     the other BASELINE configurations' group shapes (c2: 10 events, same-event only; c3: + mixed;
                              c4 / c4k: 50 events of pi+ / K+, 8 K_T x 8 K_phi bins x 41^3 = 238 MB of histograms)
    --scaling strong          config 5 as written: `--total-groups` (200) groups sharded round-robin over the ranks, ONE
                              all-reduce at the end; a step = the whole job, `value` = its pairs / time-to-result
Every line carries `checksum`: integer sums / a position-weighted hash of the all-reduced num_count and den_count of a
FIXED job (strong: the job itself; weak: 8 groups sharded over the ranks), so the same value must appear at N = 1, 2, 4, 8.

`value`  = pairs of all ranks / max-over-ranks device time of the step, particles already in HBM.
`e2e`    = same metric through the reference-facing host call (hbt_accumulate_batch) with the
           particles in pinned HOST memory: staging copy, H2D, mixed-event plan (RNG draws),
           kernels, all-reduce and the D2H read of the per-K_T pair counters inside the timing.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from hadronic_afterburner_toolkit_b200 import hbtio, synth  # noqa: E402
from hadronic_afterburner_toolkit_b200.params import C2, C3, C4, C5, EVENT_MULTIPLICITY, KAON_MASS, PION_MASS  # noqa: E402

METRIC = "HBT pairs/sec (same+mixed)"
UNIT = "pairs/s"
SEED = 20260005


# group shapes of BASELINE.json's configurations (C2/C3 10 events x 1500, C4 50 x 1500, C5 100 x 1500)
CONFIGS = {
    "c5": dict(P=C5, events=100, mass=PION_MASS, species="pi+", mixed=True, groups=2,
               desc="config-5-shape oversample groups, same+mixed pairs, 41^3 q grid x 4 K_T bins"),
    "c3": dict(P=C3, events=10, mass=PION_MASS, species="pi+", mixed=True, groups=32,
               desc="config-3-shape oversample groups (oversampling 10), same+mixed pairs, 41^3 q grid x 4 K_T bins"),
    "c2": dict(P=C2, events=10, mass=PION_MASS, species="pi+", mixed=False, groups=32,
               desc="config-2-shape oversample groups (oversampling 10), same-event pairs only, 41^3 q grid x 4 K_T bins"),
    "c4": dict(P=C4, events=50, mass=PION_MASS, species="pi+", mixed=True, groups=2,
               desc="config-4-shape oversample groups (oversampling 50), same+mixed pairs, 8 K_T x 8 K_phi bins x 41^3 (238 MB of histograms)"),
    "c4k": dict(P=C4, events=50, mass=KAON_MASS, species="K+", mixed=True, groups=2,
                desc="config-4-shape oversample groups (oversampling 50), same+mixed pairs, 8 K_T x 8 K_phi bins x 41^3 (238 MB of histograms)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c5", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--total-groups", type=int, default=200, help="--scaling strong: groups of the whole job (config 5: 200)")
    ap.add_argument("--groups-per-gpu", type=int, default=None, help="--scaling weak: groups per GPU per step")
    ap.add_argument("--events-per-group", type=int, default=None)
    ap.add_argument("--multiplicity", type=int, default=EVENT_MULTIPLICITY)
    ap.add_argument("--cpu-events", type=int, default=10, help="events in the CPU baseline's bounded sample group")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    c = CONFIGS[a.config]
    if a.groups_per_gpu is None:
        a.groups_per_gpu = c["groups"]
    if a.events_per_group is None:
        a.events_per_group = c["events"]
    return a


def workload_config(a, world):
    c = CONFIGS[a.config]
    P = c["P"]
    if a.scaling == "strong":
        per_step = f"{a.total_groups} groups in all, sharded round-robin over the ranks, one all-reduce at the end (a step = the whole job)"
    else:
        per_step = f"{a.groups_per_gpu} groups/GPU/step"
    return {
        "workload": f"{c['desc']}: {a.events_per_group} events x {a.multiplicity} {c['species']} per group; {per_step}",
        "name": a.config, "scaling": a.scaling,
        "groups_per_gpu_per_step": a.groups_per_gpu if a.scaling == "weak" else None,
        "total_groups": a.total_groups if a.scaling == "strong" else None,
        "events_per_group": a.events_per_group,
        "multiplicity": a.multiplicity, "qnpts": P.qnpts, "n_KT": P.n_KT, "n_Kphi": P.n_Kphi if P.azimuthal_flag else None,
        "needed_number_of_pairs": P.needed_number_of_pairs,
        "sharding": f"event groups over {world} rank(s), one NCCL all-reduce of the histograms per step",
        "l2": "flushed between timed steps (256 MiB write); accumulators stay resident by design",
        "reference_arm_sample": f"the reference arm and cpu_baseline time groups of {a.cpu_events} events x {a.multiplicity} particles "
                                f"(one per host core; {a.cpu_events * a.multiplicity} particles per group against "
                                f"{a.events_per_group * a.multiplicity} here: a full-size group of config 5 takes the reference 15 min "
                                "per core); smaller groups have the smaller working set, which favours the reference",
    }


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation on the host cores
def run_cpu_reference(a, cores, repeats=1):
    """One process per core, each pushing one group of n_events x multiplicity particles through the
    UNMODIFIED reference (oracle/_ref/ref_driver mem).  Returns (pairs/s aggregate, pairs per
    process, kind).  Falls back to the C oracle port if the compiled reference is absent."""
    from oracle import oracle_py as O

    c = CONFIGS[a.config]
    P, n_events, multiplicity, mass = c["P"], a.cpu_events, a.multiplicity, c["mass"]
    nmix = n_events // 2 + 1
    n = n_events * multiplicity
    pairs = n * (n - 1) // 2 + (n_events * multiplicity * nmix * multiplicity if c["mixed"] else 0)
    if O.have_reference():
        with tempfile.TemporaryDirectory() as td:
            fpar = os.path.join(td, "parameters.dat")
            with open(fpar, "w") as f:
                f.write(P.parameters_dat())
            fins = []
            for k in range(cores):
                fin = os.path.join(td, f"in{k}.bin")
                hbtio.write_batches(fin, synth.make_batches(SEED, 1, n_events, mass, multiplicity, first_group=1000 + k))
                fins.append(fin)
            best = 0.0
            for _ in range(repeats):
                t0 = time.perf_counter()
                procs = [subprocess.Popen([O.REF_DRIVER, "mem", fpar, fins[k], os.path.join(td, f"out{k}.bin")]
                                          + ([] if c["mixed"] else ["same_only"]),
                                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for k in range(cores)]
                rcs = [p.wait() for p in procs]
                wall = time.perf_counter() - t0
                assert all(r == 0 for r in rcs), "ref_driver failed"
                # time inside the reference's two loop functions, slowest process
                t_loop = max(hbtio.read_accumulators(os.path.join(td, f"out{k}.bin")).t_total for k in range(cores))
                best = max(best, cores * pairs / t_loop)
            return best, pairs, "reference", wall
    # port: the C restatement, single-threaded per process (run in-process, one core)
    o = O.Oracle(P)
    b = synth.make_batches(SEED, 1, n_events, mass, multiplicity, first_group=1000)[0]
    t0 = time.perf_counter()
    o.process_batch(b, do_mixed=c["mixed"])
    wall = time.perf_counter() - t0
    ts, tm = o.times()
    return pairs / (ts + tm), pairs, "port", wall


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def reference_arm(a, rank, world):
    if rank != 0:
        return
    cores = host_cores()
    vals, walls = [], []
    for i in range(a.warmup + a.steps):
        v, pairs, kind, wall = run_cpu_reference(a, cores)
        if i >= a.warmup:
            vals.append(v)
            walls.append(wall)
    value = float(np.mean(vals))
    sample = (f"per step: {cores} processes x 1 group of {a.cpu_events} events x {a.multiplicity} particles "
              f"({pairs:.3e} pairs each; NOT the {a.events_per_group}-event groups of the GPU arm, see config.reference_arm_sample), "
              "time inside the reference's two pair-loop functions")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * float(np.mean(walls)), "higher_is_better": True, "scaling": a.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(a, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except FileNotFoundError:
            self.p = None
        self.t0 = self.t1 = None

    def window_start(self):
        import datetime
        self.t0 = datetime.datetime.now()

    def window_end(self):
        import datetime
        self.t1 = datetime.datetime.now()

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, power, reasons = [], 0.0, [], set()
        import datetime
        for r in rows:
            try:
                ts = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f")
                if self.t0 and self.t1 and not (self.t0 <= ts <= self.t1):
                    continue  # only samples taken DURING the timed region
                sm.append(float(r[1])); mx = max(mx, float(r[2])); power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.strip().lower().startswith("active"):
                    reasons.add(name)
        load = sm  # every kept sample lies inside the timed region
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": mx or None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_ops(stage, boost=True, az=False):
    """lazily-evaluated FP64 operation count from the stage populations."""
    d = 17 if boost else 2
    e_extra = 3 if az else 0
    s, m = stage[:6].astype(np.float64), stage[6:].astype(np.float64)
    same = 7 * s[0] + 10 * s[1] + 5 * s[2] + d * s[3] + (20 + e_extra) * s[4]
    mixed = 7 * m[0] + 10 * m[1] + 5 * m[2] + d * m[3] + (3 + e_extra) * m[4]
    return same, mixed


_REAL_STDOUT = None


def emit(obj) -> None:
    """The ONE JSON line of the contract, on the process's original stdout."""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def checksum(num_count, den_count):
    """Integer fingerprint of the two count histograms: sums and a position-weighted hash (exact in Python ints)."""
    w = (np.arange(num_count.size, dtype=np.uint64) % np.uint64(65521)) + np.uint64(1)
    M = (1 << 61) - 1
    hn = int(np.sum((num_count.astype(object) * w.astype(object)))) % M
    hd = int(np.sum((den_count.astype(object) * w.astype(object)))) % M
    return {"num_count_sum": int(num_count.sum()), "den_count_sum": int(den_count.sum()), "num_hash": hn, "den_hash": hd}


def main():
    global _REAL_STDOUT
    a = parse()
    # libraries print to stdout too (NCCL's version banner under NCCL_DEBUG=VERSION/WARN): everything but the
    # JSON line goes to stderr
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if a.impl == "reference":
        reference_arm(a, rank, world)
        return

    import torch
    import torch.distributed as dist

    from hadronic_afterburner_toolkit_b200 import capi
    from hadronic_afterburner_toolkit_b200.hbt_correlation import HBT_correlation, Random, _check, gather_rapidity, psi_ref

    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the product has no CPU path"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = capi.lib()
    cfg = CONFIGS[a.config]
    P, mass, do_mixed = cfg["P"], cfg["mass"], cfg["mixed"]
    az = P.azimuthal_flag == 1
    eng = HBT_correlation(P, device=local)
    h = eng._h
    if (os.environ.get("HBT_B200_FUSE", "1") == "0" or os.environ.get("HBT_B200_KERNEL", "2") == "1") and "HBT_B200_LANES" not in os.environ:
        # per-loop timings: one compute stream, so that the two kernels of a group do not overlap the next group's
        _check(h, L.hbt_set_option(h, 4, 1))
    if world > 1:  # NCCL communicator of the library itself; torch only ferries the 128-byte id
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            _check(None, L.hbt_comm_unique_id(buf))
            uid.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        raw = bytes(uid.cpu().numpy().tobytes())
        _check(h, L.hbt_comm_init_rank(h, world, rank, ctypes.create_string_buffer(raw, 128)))

    # ---- synthetic input: this rank's groups, pinned host copies and HBM-resident copies
    nev, mult = a.events_per_group, a.multiplicity
    strong = a.scaling == "strong"
    if strong:
        my_groups = [g for g in range(a.total_groups) if g % world == rank]  # round-robin, as config 5 shards its groups
        all_groups = a.total_groups
    else:
        my_groups = [rank * a.groups_per_gpu + g for g in range(a.groups_per_gpu)]
        all_groups = world * a.groups_per_gpu
    n = nev * mult
    nmix = nev // 2 + 1
    off = np.arange(nev + 1, dtype=np.int64) * mult
    pairs_group = n * (n - 1) // 2 + (nev * mult * nmix * mult if do_mixed else 0)

    def load_group(seed, g):
        arr = synth.make_group(seed, g, nev, mass, mult).reshape(nev * mult, 8)
        flat = np.ascontiguousarray(np.concatenate([gather_rapidity(P, arr[e * mult:(e + 1) * mult]) for e in range(nev)]))
        assert flat.shape[0] == nev * mult  # |y| < 0.45 < HBTrap: the cut keeps everything
        t = torch.from_numpy(flat).pin_memory()
        # Psi_2 of the group (azimuthally sensitive configs), glibc on the host as the reference computes it
        return t, t.cuda(), (psi_ref(arr, 2) if az else 0.0)

    host, dev, psis = [], [], []
    for g in my_groups:
        t, d, ps = load_group(SEED, g)
        host.append(t); dev.append(d); psis.append(ps)

    def make_plans(seed, groups_all, mine):
        """The reference's draws for the whole job in stream order; a rank keeps the plans of its own groups and
        only advances the stream past the others (src/HBT_correlation.cpp:200-215)."""
        rng = Random(seed)
        plans = {}
        for g in groups_all:
            if g in mine:
                plans[g] = rng.mixed_plan(nev, nev)
            else:
                rng.skip_batch(nev, nev)
        return plans

    plans = make_plans(P.randomSeed, range(all_groups) if strong else my_groups, set(my_groups)) if do_mixed else {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def submit_resident(d_t, psi, plan):
        if do_mixed:
            ids, cs = plan
            # one whole-batch call per group (both loops: two kernels next to each other); separate calls with HBT_B200_FUSE=0 or stage counters on
            _check(h, L.hbt_accumulate_batch_dev(h, d_t.data_ptr(), off.ctypes.data, nev, ids.ctypes.data, cs.ctypes.data, nmix, psi))
        else:
            _check(h, L.hbt_accumulate_same_dev(h, d_t.data_ptr(), n, psi))

    def submit_host(h_t, psi, plan):
        if do_mixed:
            ids, cs = plan
            _check(h, L.hbt_accumulate_batch(h, h_t.data_ptr(), off.ctypes.data, nev, None, None, 0,
                                             ids.ctypes.data, cs.ctypes.data, nmix, psi, 1, 1))
        else:
            _check(h, L.hbt_accumulate_batch(h, h_t.data_ptr(), off.ctypes.data, nev, None, None, 0, None, None, 0, psi, 1, 0))

    def step_resident():
        for k, g in enumerate(my_groups):
            submit_resident(dev[k], psis[k], plans.get(g))
        if world > 1:
            _check(h, L.hbt_allreduce(h))

    kcount = np.zeros(2 * P.n_slabs, dtype=np.uint64)
    e2e_rng = Random(P.randomSeed + 1)

    def step_e2e():
        for k, g in enumerate(my_groups):
            plan = e2e_rng.mixed_plan(nev, nev) if do_mixed else None  # fresh draws, as the host loop would make them
            submit_host(host[k], psis[k], plan)
        if world > 1:
            _check(h, L.hbt_allreduce(h))
        # the step's result: per-K_T accepted-pair counters, device -> host
        _check(h, L.hbt_read(h, None, None, None, None, None, None, kcount.ctypes.data, kcount[P.n_slabs:].ctypes.data))

    def timed(step_fn, steps, device_timer, after=None):
        total = 0.0
        for it in range(steps):
            flush.fill_(1)
            barrier()
            if device_timer:
                _check(h, L.hbt_timer_start(h))
                step_fn()
                ms = ctypes.c_double()
                _check(h, L.hbt_timer_stop(h, ctypes.byref(ms)))
                _check(h, L.hbt_synchronize(h))
                total += ms.value * 1e-3
            else:
                t0 = time.perf_counter()
                step_fn()
                if after is not None and it == steps - 1:
                    after()
                _check(h, L.hbt_synchronize(h))
                torch.cuda.synchronize()
                total += time.perf_counter() - t0
        t = torch.tensor([total], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def launches():
        c = ctypes.c_uint64()
        _check(h, L.hbt_get_launch_count(h, ctypes.byref(c)))
        return c.value

    # ---- stage populations of one step (for the algorithmic-ops formula): an instrumented,
    # UNTIMED pass over the same input; the timed production kernels skip tile pairs whose
    # bounding boxes cannot hold an accepted pair and therefore do not see every pair
    _check(h, L.hbt_set_option(h, 1, 1))
    s0 = eng.stage_counters()
    inst_groups = my_groups if not strong else my_groups[:2]  # (strong: two groups, scaled: all groups have the same shape)
    for k, g in enumerate(inst_groups):
        submit_resident(dev[k], psis[k], plans.get(g))
    st_step = eng.stage_counters() - s0
    st_step = (st_step.astype(np.float64) * (len(my_groups) / max(1, len(inst_groups)))).astype(np.uint64)
    _check(h, L.hbt_set_option(h, 1, 0))

    # ---- resident-input measurement (value) ---------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None  # nvidia-smi needs a moment to start: launch it before the warm-up
    timed(step_resident, a.warmup, True)
    st0 = eng.stage_counters()
    tm0, l0 = eng.timers(), launches()
    if sampler:
        sampler.window_start()
    t_value = timed(step_resident, a.steps, True)
    if sampler:
        sampler.window_end()
    clocks = sampler.stop() if sampler else None
    st1 = eng.stage_counters()
    tm1, l1 = eng.timers(), launches()
    pairs_step = all_groups * pairs_group
    pairs_total = pairs_step * a.steps
    value = pairs_total / t_value

    # ---- end-to-end measurement (host buffers through the reference-facing call) ----------
    e2e = None
    if not a.no_e2e:
        nb = P.n_bins
        # result buffers in page-locked memory, like the inputs (what a caller that cares about throughput hands over:
        # config 4's 238 MB of histograms arrive at PCIe speed instead of through the driver's pageable staging)
        pinned = [torch.zeros(nb, dtype=torch.int64).pin_memory()] + [torch.zeros(nb, dtype=torch.float64).pin_memory() for _ in range(4)] \
            + [torch.zeros(nb, dtype=torch.int64).pin_memory()]
        full = [t.numpy() for t in pinned]

        def read_all():
            # the path's result: all six histograms, read once per analysis (here: once per timed region)
            _check(h, L.hbt_read(h, full[0].ctypes.data, full[1].ctypes.data, full[2].ctypes.data, full[3].ctypes.data,
                                 full[4].ctypes.data, full[5].ctypes.data, kcount.ctypes.data, kcount[P.n_slabs:].ctypes.data))

        timed(step_e2e, max(1, a.warmup - 1), False)
        t_e2e = timed(step_e2e, a.steps, False, after=read_all)
        blob = sum(x.nbytes for x in full)
        e2e = {"value": pairs_total / t_e2e, "unit": UNIT,
               "h2d_bytes_per_step": int(len(my_groups) * (n * 64 + (nev * nmix * 48 if do_mixed else 0))),
               "d2h_bytes_per_step": int(kcount.nbytes + blob / a.steps),
               "d2h_note": f"per step the per-K_T pair counters ({kcount.nbytes} B); the six histograms ({blob} B) are read ONCE, "
                           f"inside the timed region, after the last step (amortised over the {a.steps} steps)",
               "ms_per_step": 1e3 * t_e2e / a.steps}

    # ---- roofline of the pair kernels (rank 0's device) -------------------------------------
    peak = ctypes.c_double()
    _check(None, L.hbt_measure_fp64_peak(local, 300.0, ctypes.byref(peak)))
    # stage populations of this rank over the timed steps = steps x (one instrumented step)
    dst = (st_step * np.uint64(a.steps)).astype(np.uint64)
    ops_same, ops_mixed = algorithmic_ops(dst, boost=P.long_comoving_boost == 1, az=az)
    if not do_mixed:
        ops_mixed = 0.0
    fused = do_mixed and os.environ.get("HBT_B200_FUSE", "1") != "0" and os.environ.get("HBT_B200_KERNEL", "2") != "1"
    # a whole batch = the same-event kernel and the v4 mixed-event kernel next to each other on two streams
    # (HBT_B200_SPLIT=0: the one fused v3 kernel of rounds 1-2)
    split = fused and os.environ.get("HBT_B200_SPLIT", "1") != "0" and os.environ.get("HBT_B200_MIXED4", "1") != "0"
    ks = (tm1["same_ms"] - tm0["same_ms"]) * 1e-3
    km = (tm1["mixed_ms"] - tm0["mixed_ms"]) * 1e-3
    ach = (ops_same + ops_mixed) / (ks + km) / 1e12
    traffic = None
    traffic_detail = None
    if a.config == "c5":
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch (one C5-shape group)
            one = fused and not split
            traffic = tj["fused" if one else "same"]["dram_bytes_per_launch"] + (0 if one else tj["mixed"]["dram_bytes_per_launch"])
            traffic_detail = {"kind": "STATIC: a constant read from profiles/traffic.json (one `ncu --set full` capture with a cold L2), "
                                      "NOT measured in this run",
                              "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, one C5-shape group per launch)",
                              "fused": tj["fused"]["dram_bytes_per_launch"], "same": tj["same"]["dram_bytes_per_launch"],
                              "mixed": tj["mixed"]["dram_bytes_per_launch"],
                              "algorithmic_bytes_per_launch": tj["fused"]["algorithmic_bytes_per_launch"], "source": tj["source"]}
        except (OSError, KeyError, ValueError):
            pass
    # what actually binds (profiles/r03_ncu_summary.txt, profiles/r03_controls_corun.txt): the same-event kernel ALONE
    # is held by its reductions into the L2 (5 per accepted pair: 15.3 ms, 11.6 ms without them, and neither fewer
    # instructions nor fewer shared-memory wavefronts nor 20 instead of 18 warps move it); the mixed-event kernel by
    # instruction issue (80 % of the slots).  Next to each other the mixed-event warps fill the same-event warps'
    # reduction stalls: 24.4 ms for a C5 group against 26.4 one after the other (and 27.1 for the fused v3 kernel);
    # with the reductions off both ways take ~23 ms, i.e. what is left is the sum of two issue-bound kernels
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    n_sm = torch.cuda.get_device_properties(local).multi_processor_count
    red_lanes = 5.0 * float(dst[4]) + float(dst[10])
    red_rate = red_lanes / (ks + km) / (sm_mhz * 1e6) / n_sm
    bound_actual = {
        "issue_slots": {"busy_pct_same_event_kernel_alone": 51.9, "busy_pct_mixed_event_kernel_alone": 80.5,
                        "warp_instructions_per_c5_group": 9.29e9 + 10.25e9, "ms_at_full_issue": 16.8,
                        "kind": "STATIC: ncu captures profiles/r03_ncu_summary.txt of the two kernels of a batch, each alone "
                                "(smsp__issue_active, smsp__inst_executed); under ncu they cannot run next to each other"},
        "lsu_data_pipe": {"busy_pct_same_event_kernel_alone": 70.7, "busy_pct_mixed_event_kernel_alone": 62.0,
                          "kind": "STATIC: same captures, l1tex__data_pipe_lsu_wavefronts (same-event: 58 % of the wavefronts are the "
                                  "reductions, 14 per 64-bit RED of 27 lanes with spread addresses)"},
        "reductions": {"red_lanes_per_clk_per_sm": red_rate, "ceiling_all_sms_busy": 0.68, "ceiling_per_sm": 0.94,
                       "frac": red_rate / 0.68, "lanes": "5 per accepted same-event pair + 1 per accepted mixed-event pair",
                       "ceiling_source": "profiles/r02_tma_red_bench.txt (scripts/micro/tma_red_bench.cu: spread REDs from every SM / "
                                         "from a quarter of the SMs)"},
        "controls": "profiles/r03_controls_corun.txt (one C5-shape group): the two kernels next to each other 24.3 ms, one after the "
                    "other 26.4 (same-event 15.3 + mixed-event 11.1); same-event reductions predicated off with all four sums still "
                    "evaluated (HBT_DBG_RED=7): 23.2 / 22.7 (same-event kernel alone 11.6); profiles/r02_controls.txt has the "
                    "controls of the fused v3 kernel (27.25 ms; 25.44 without reductions)",
    }
    roofline = {
        "bound": "fp64", "achieved": ach, "peak": peak.value, "unit": "TFLOP/s", "frac": ach / peak.value,
        "traffic": traffic, "traffic_detail": traffic_detail, "bound_actual": bound_actual,
        "peak_source": "measured on this device: DFMA dependent-chain microbenchmark (hbt_measure_fp64_peak); "
                       "FP64 is not in MEASURED_PEAKS.json",
        "definition": "algorithmic FP64 ops (SURVEY.md 8d: 7nA+10nB+5nC+17nD+20nE same, ...+3nE mixed; stage populations from an "
                      "instrumented, untimed pass over the same input) / CUDA-event time of the production pair kernels "
                      "(incl. the same-event sort + cull kernels) on the launching stream",
        "note": "the bound SURVEY.md 8d names is the FP64 pipe; it is NOT what binds this design: the production prefilter runs in "
                "packed FP32 and mixed-event survivors are binned in FP32 wherever a rigorous error band allows, so the FP64 pipe "
                "itself is 31 % busy in the same-event kernel and 2 % in the mixed-event kernel (ncu) and `frac` is an "
                "algorithmic-throughput fraction; see bound_actual",
        "kernels": ({
            # per group: the same-event kernel and the v4 mixed-event kernel share the SMs on two streams (split), or one
            # fused v3 kernel works through both unit lists interleaved (HBT_B200_SPLIT=0); timed from the first to the
            # last of them on the launching stream (the second stream joins it)
            ("hbt_pairs_v3<same-event> || hbt_pairs_v4_mixed" if split else "hbt_pairs_v3_fused"): {"ms_per_launch": 1e3 * (ks + km) / max(1, tm1["same_launches"] - tm0["same_launches"]),
                                   "pairs_per_s": float(dst[0] + dst[6]) / (ks + km), "tflops": ach, "frac": ach / peak.value,
                                   "ops_per_pair_same": ops_same / float(dst[0]), "ops_per_pair_mixed": ops_mixed / float(dst[6])},
        } if fused else {
            "same": {"ms_per_launch": 1e3 * ks / max(1, tm1["same_launches"] - tm0["same_launches"]),
                     "pairs_per_s": float(dst[0]) / ks, "tflops": ops_same / ks / 1e12, "frac": ops_same / ks / 1e12 / peak.value,
                     "ops_per_pair": ops_same / float(dst[0])},
            **({"mixed": {"ms_per_launch": 1e3 * km / max(1, tm1["mixed_launches"] - tm0["mixed_launches"]),
                          "pairs_per_s": float(dst[6]) / km, "tflops": ops_mixed / km / 1e12, "frac": ops_mixed / km / 1e12 / peak.value,
                          "ops_per_pair": ops_mixed / float(dst[6])}} if do_mixed else {}),
        }),
        "stage_fractions_same": [float(x) / float(dst[0]) for x in dst[:6]],
        "kernel_share_of_step": (ks + km) / (t_value if world == 1 else max(t_value, 1e-12)),
    }

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = host_cores()
        v, pairs, kind, wall = run_cpu_reference(a, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{cores} processes x 1 group of {a.cpu_events} events x {mult} particles ({pairs:.3e} pairs each; the GPU arm's "
                         f"groups have {nev} events), time inside the reference's two pair-loop functions; wall {wall:.1f} s",
               "per_core": v / cores}

    # ---- checksum of a FIXED job: identical at every N (driver-visible multi-GPU parity) ---------------------
    _check(h, L.hbt_reset(h))
    if strong:
        ck_groups, ck_seed = list(range(a.total_groups)), SEED
        for k, g in enumerate(my_groups):
            submit_resident(dev[k], psis[k], plans.get(g))
    else:
        ck_groups, ck_seed = list(range(8)), SEED + 77
        mine = [g for g in ck_groups if g % world == rank]
        ck_plans = make_plans(P.randomSeed, ck_groups, set(mine)) if do_mixed else {}
        keep = []
        for g in mine:
            t, d, ps = load_group(ck_seed, g)
            keep.append(d)
            submit_resident(d, ps, ck_plans.get(g))
    if world > 1:
        _check(h, L.hbt_allreduce(h))
    numc, denc = np.zeros(P.n_bins, dtype=np.uint64), np.zeros(P.n_bins, dtype=np.uint64)
    _check(h, L.hbt_read(h, numc.ctypes.data, None, None, None, None, denc.ctypes.data, None, None))
    ck = checksum(numc, denc)
    ck["job"] = f"{len(ck_groups)} groups (seed {ck_seed}) of {nev} events x {mult} {cfg['species']}, sharded round-robin over {world} rank(s), all-reduced"

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * t_value / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(a, world), "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(l1 - l0), "roofline": roofline, "cpu_baseline": cpu,
            "pairs_per_step": pairs_step, "kernel": os.environ.get("HBT_B200_KERNEL", "default"),
            "deferred_pairs": eng.deferred_pairs(), "checksum": ck,
        }
        emit(out)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
