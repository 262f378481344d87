"""End-to-end drop-in check: the reference's own command-line binary and the SAME binary with
the B200 HBT_correlation class inside (hadronic_afterburner_toolkit_b200/host) read the same
particle_samples.gz through the reference's unchanged reader, with the same parameters.dat and
randomSeed, and must write the same HBT_correlation_function_*.dat files."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT
from hadronic_afterburner_toolkit_b200 import synth
from hadronic_afterburner_toolkit_b200.params import C3, C4, HBTParams

REF_EXE = os.path.join(ROOT, "oracle", "_ref", "hadronic_afterburner_tools.e")
OUR_EXE = os.path.join(ROOT, "hadronic_afterburner_toolkit_b200", "host", "build", "hadronic_afterburner_tools_b200.e")
# the HBT analysis without any reference code: fast gz reader -> GPU -> the same output files
FAST_EXE = os.path.join(ROOT, "hadronic_afterburner_toolkit_b200", "host", "build", "hbt_fast_analysis.e")
PDG = os.path.join(ROOT, "oracle", "_ref", "EOS", "pdg.dat")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (os.path.exists(REF_EXE) and os.path.exists(OUR_EXE)),
                                 reason="compiled reference / drop-in binary not present")]


def run_binary(exe, workdir, params_text, gz_src, env=None, input_name="particle_samples.gz", extra_files=None):
    os.makedirs(os.path.join(workdir, "EOS"))
    os.makedirs(os.path.join(workdir, "results"))
    shutil.copy(PDG, os.path.join(workdir, "EOS", "pdg.dat"))
    shutil.copy(gz_src, os.path.join(workdir, "results", input_name))
    for name, src in (extra_files or {}).items():
        shutil.copy(src, os.path.join(workdir, "results", name))
    with open(os.path.join(workdir, "parameters.dat"), "w") as f:
        f.write(params_text)
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([exe], cwd=workdir, capture_output=True, text=True, env=e)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = os.path.join(workdir, "results")
    return {fn: open(os.path.join(res, fn)).read().splitlines() for fn in sorted(os.listdir(res))
            if fn.startswith("HBT_correlation_function")}, r.stdout


def same_text(want, got):
    assert sorted(want) == sorted(got)
    for fn in want:
        a, b = want[fn], got[fn]
        assert len(a) == len(b), fn
        for la, lb in zip(a, b):
            if la == lb:
                continue
            assert len(la) == len(lb), (fn, la, lb)
            for x, y in zip(la.split(), lb.split()):
                if x != y:  # a <=1e-10 difference straddling the 9th digit
                    assert abs(float(x) - float(y)) <= 2e-8 * max(abs(float(x)), 1e-300), (fn, la, lb)


CASES = {
    "c3_shape": (C3.with_(qnpts=15), 3, 5, 300),
    "c4_shape_az": (C4.with_(qnpts=9, n_KT=4, n_Kphi=4), 2, 4, 300),
    "qinv": (HBTParams(qnpts=21, invariant_radius_flag=1), 2, 4, 200),
    "cap_reached": (C3.with_(qnpts=15, needed_number_of_pairs=2500.0), 4, 4, 300),
    "qinv_cap_reached": (HBTParams(qnpts=21, invariant_radius_flag=1, needed_number_of_pairs=300.0), 3, 4, 300),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_same_files_as_reference_binary(name, tmp_path):
    P, ngrp, nev, mult = CASES[name]
    batches = synth.make_batches(41, ngrp, nev, multiplicity=mult)
    gz = str(tmp_path / "input.gz")
    synth.write_iss_gz(gz, batches)
    text = P.parameters_dat(event_buffer_size=nev * mult)  # groups of exactly nev events
    want, _ = run_binary(REF_EXE, str(tmp_path / "ref"), text, gz)
    got, out = run_binary(OUR_EXE, str(tmp_path / "ours"), text, gz, env={"HBT_B200_DEVICES": "1"})
    assert "HBT pair loops run on 1 GPU context(s)" in out
    same_text(want, got)
    assert len(want) == (P.n_KT - 1) * (P.n_Kphi if P.azimuthal_flag else 1) * (2 if P.invariant_radius_flag else 1)
    # third arm: our own reader and driver (no reference code at all) on the same directory layout
    fast, out = run_binary(FAST_EXE, str(tmp_path / "fast"), text, gz)
    assert "hbt_fast_analysis: %d batches, %d events" % (ngrp, ngrp * nev) in out
    same_text(want, fast)


def test_fast_driver_refuses_what_it_cannot_read(tmp_path):
    P = C3.with_(qnpts=15)
    batches = synth.make_batches(44, 1, 2, multiplicity=50)
    gz = str(tmp_path / "input.gz")
    synth.write_iss_gz(gz, batches)
    for key, value in (("read_in_mode", 6), ("resonance_feed_down_flag", 1), ("particle_monval", 9999),
                       ("read_in_real_mixed_events", 1)):  # (the last: no mixed-event file in the directory)
        text = P.parameters_dat(event_buffer_size=100, **{key: value})
        with pytest.raises(AssertionError, match="hbt_fast_analysis"):
            run_binary(FAST_EXE, str(tmp_path / key), text, gz)


def test_real_mixed_events_same_files(tmp_path):
    """read_in_real_mixed_events = 1: the partner events come from results/particle_samples_mixed_event.gz, one
    batch per batch of the main file (src/Analysis.cpp:824-825); batches of different event counts in the two
    files, so the draws use nev_mixed != nev.  Reference binary, drop-in binary and fast driver: same files."""
    P = C3.with_(qnpts=15)
    main = synth.make_batches(51, 3, 4, multiplicity=300)
    mixed = synth.make_batches(52, 3, 6, multiplicity=200)   # 6 events of 200 per 1200-particle batch
    gz, gz2 = str(tmp_path / "input.gz"), str(tmp_path / "mixed.gz")
    synth.write_iss_gz(gz, main)
    synth.write_iss_gz(gz2, mixed)
    text = P.parameters_dat(event_buffer_size=1200, read_in_real_mixed_events=1)
    extra = {"particle_samples_mixed_event.gz": gz2}
    want, _ = run_binary(REF_EXE, str(tmp_path / "ref"), text, gz, extra_files=extra)
    got, _ = run_binary(OUR_EXE, str(tmp_path / "ours"), text, gz, env={"HBT_B200_DEVICES": "1"}, extra_files=extra)
    same_text(want, got)
    fast, out = run_binary(FAST_EXE, str(tmp_path / "fast"), text, gz, extra_files=extra)
    assert "hbt_fast_analysis: 3 batches, 12 events" in out
    same_text(want, fast)


@pytest.mark.parametrize("mode,input_name,src_name", [(2, "particle_list.dat", "urqmd_small.particle_list.dat"),
                                                      (21, "particle_list.bin", "urqmd_small.particle_list.bin"),
                                                      (1, "particle_list.dat", "urqmd_small.f13.dat"),
                                                      (0, "OSCAR.DAT", "urqmd_small.OSCAR.DAT")])
def test_urqmd_input_formats_same_files(mode, input_name, src_name, tmp_path):
    """read_in_mode 2 (gzipped UrQMD text) and 21 (UrQMD binary): the reference binary, the drop-in
    binary (reference reader, GPU pair loops) and hbt_fast_analysis.e (our reader, no reference code)
    write the same files from the committed synthetic UrQMD-format inputs."""
    src = os.path.join(ROOT, "tests", "golden", src_name)
    P = HBTParams(qnpts=11, KT_min=0.0, KT_max=1.0, n_KT=3, HBTrap_min=-1.0, HBTrap_max=1.0)
    text = P.parameters_dat(read_in_mode=mode, event_buffer_size=400)
    want, _ = run_binary(REF_EXE, str(tmp_path / "ref"), text, src, input_name=input_name)
    got, _ = run_binary(OUR_EXE, str(tmp_path / "ours"), text, src, env={"HBT_B200_DEVICES": "1"}, input_name=input_name)
    same_text(want, got)
    fast, out = run_binary(FAST_EXE, str(tmp_path / "fast"), text, src, input_name=input_name)
    assert ("hbt_fast_analysis: 2 batches, 9 events" if mode in (2, 21) else "hbt_fast_analysis: 1 batches, 5 events") in out
    same_text(want, fast)
    assert any(float(l.split()[3]) != 0.0 for fn in want for l in want[fn])  # the files hold pairs


@pytest.mark.parametrize("alpha,beta,rap_type,buf", [(211, -211, 1, 1300), (9998, -9998, 0, 100000)])
def test_balance_function_operator_same_files(alpha, beta, rap_type, buf, tmp_path):
    """Analysis::BalanceFunctionAnalysis with OUR BalanceFunction class inside the
    reference binary (reference reader incl. the charged-hadron species groups, reference RNG) writes
    the reference's three files, character for character (all histogram entries are integers)."""
    gz = os.path.join(ROOT, "tests", "golden", "bf_input.gz")
    text = HBTParams(randomSeed=77).parameters_dat(
        analyze_HBT=0, analyze_balance_function=1, event_buffer_size=buf, particle_alpha=alpha, particle_beta=beta,
        Bnpts=21, Brap_max=2.0, BpT_min=0.2, BpT_max=3.0, rap_type=rap_type)

    def run(exe, wd):
        run_binary(exe, wd, text, gz)
        res = os.path.join(wd, "results")
        return {fn: open(os.path.join(res, fn)).read() for fn in sorted(os.listdir(res)) if fn.endswith(".dat")}
    want = run(REF_EXE, str(tmp_path / "ref"))
    got = run(OUR_EXE, str(tmp_path / "ours"))
    assert len(want) == 3 and sorted(want) == sorted(got)
    for fn in want:
        assert want[fn] == got[fn], fn


SHARDED = {
    "uncapped": C3.with_(qnpts=15),
    # the ordered pair cap engages in the middle of the run: the group must cut at the reference's pair
    "capped": C3.with_(qnpts=15, needed_number_of_pairs=6000.0),
    "capped_qinv": HBTParams(qnpts=21, invariant_radius_flag=1, needed_number_of_pairs=400.0),
}


@pytest.mark.parametrize("devices", ["0,0", "0,0,0", "all"])
@pytest.mark.parametrize("name", sorted(SHARDED))
def test_groups_sharded_over_several_contexts_same_files(name, devices, tmp_path):
    """Oversample groups dealt round-robin to several engine contexts (HBT_B200_DEVICES): one per GPU of the box
    ("all"; skipped on a single-GPU box), or several on GPU 0 — the same code path, which also runs where the driver
    has one GPU.  With and without the needed_number_of_pairs cap: same files as the reference binary."""
    from hadronic_afterburner_toolkit_b200 import capi

    ndev = capi.lib().hbt_device_count()
    if devices == "all" and ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    nctx = ndev if devices == "all" else devices.count(",") + 1
    P, ngrp, nev, mult = SHARDED[name], 7, 4, 300
    batches = synth.make_batches(43, ngrp, nev, multiplicity=mult)
    gz = str(tmp_path / "input.gz")
    synth.write_iss_gz(gz, batches)
    text = P.parameters_dat(event_buffer_size=nev * mult)
    want, _ = run_binary(REF_EXE, str(tmp_path / "ref"), text, gz)
    got, out = run_binary(OUR_EXE, str(tmp_path / "ours"), text, gz, env={"HBT_B200_DEVICES": devices})
    assert f"HBT pair loops run on {nctx} GPU context(s)" in out
    same_text(want, got)
    fast, _ = run_binary(FAST_EXE, str(tmp_path / "fast"), text, gz, env={"HBT_B200_DEVICES": devices})
    same_text(want, fast)
