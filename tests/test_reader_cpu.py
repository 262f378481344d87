"""The fast gz reader (hbt_reader_*, host code of libhbt_b200.so) against a line-by-line Python
restatement of the reference's reader for read_in_mode=10
(src/particleSamples.cpp:1247-1286 + boostParticles :441-470 + the single-species filter
:672-676 + the HBT gather's rapidity cut, src/HBT_correlation.cpp:255-281).  No GPU needed."""
import gzip
import math

import numpy as np
import pytest

from hadronic_afterburner_toolkit_b200 import capi, hbtio, synth
from hadronic_afterburner_toolkit_b200.params import HBTParams


def reference_reader(path, monval, buffer_size, rap_shift=0.0, cut=None):
    """The reference's loop, literally: returns [(events, all_particles)] per batch."""
    with gzip.open(path, "rb") as f:
        data = f.read()
    pos = 0
    past_eof = False

    def readline():  # gz_readline: bytes up to '\n'; past_eof once a read went beyond the last byte
        nonlocal pos, past_eof
        k = data.find(b"\n", pos)
        if k < 0:
            line, pos, past_eof = data[pos:], len(data), True
        else:
            line, pos = data[pos:k], k + 1
        return line.decode()

    ch, sh = math.cosh(rap_shift), math.sinh(rap_shift)
    lo = math.tanh(cut.HBTrap_min) if cut else None
    hi = math.tanh(cut.HBTrap_max) if cut else None
    out = []
    while True:
        events, num = [], 0
        while num < buffer_size:
            line = readline()
            if past_eof:
                break
            tok = line.split()
            n = int(tok[0]) if tok else 0
            ev = []
            for _ in range(n):
                t = readline().split()
                mv = int(t[0])
                mass, tt, x, y, z, E, px, py, pz = (float(v) for v in t[1:10])
                if mv != monval:
                    continue
                Es = E * ch + pz * sh
                pzs = pz * ch + E * sh
                if cut and not (lo < pzs / Es < hi):
                    continue
                ev.append([px, py, pzs, Es, x, y, z, tt])
            events.append(np.array(ev, dtype=np.float64).reshape(-1, 8))
            num += n
        if not events:
            break
        out.append((events, num))
    return out


def write_mixed_species(path, batches, trailing_newline=True, empty_event_at=None):
    """Events of pi+ with K+ and p lines interleaved (they count for event_buffer_size)."""
    rng = np.random.default_rng(7)
    lines = []
    k = 0
    for b in batches:
        for ev in b.same:
            if empty_event_at is not None and k == empty_event_at:
                lines.append("0")
            k += 1
            rows = []
            for p in ev:
                rows.append("211 %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g"
                            % (0.138, p[7], p[4], p[5], p[6], p[3], p[0], p[1], p[2]))
                if rng.random() < 0.3:
                    q = p * rng.uniform(0.5, 1.5)
                    rows.append("%d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g"
                                % (rng.choice([321, 2212, -211]), 0.494, q[7], q[4], q[5], q[6], q[3], q[0], q[1], q[2]))
            lines.append(str(len(rows)))
            lines.extend(rows)
    text = "\n".join(lines) + ("\n" if trailing_newline else "")
    with gzip.open(path, "wt", compresslevel=1) as f:
        f.write(text)


def collect(path, monval, buffer_size, **kw):
    r = hbtio.FastReader(str(path), monval, buffer_size, **kw)
    got = []
    for b in r:
        got.append((b.same, r.all_particles))
    r.close()
    return got


def same(a, b):
    assert len(a) == len(b)
    for (ea, na), (eb, nb) in zip(a, b):
        assert na == nb and len(ea) == len(eb)
        for x, y in zip(ea, eb):
            assert x.shape == y.shape and np.array_equal(x, y)  # the same doubles, bit for bit


@pytest.mark.parametrize("buffer_size", [1, 250, 1000, 10 ** 9])
@pytest.mark.parametrize("trailing_newline", [True, False])
def test_grouping_filter_and_values(tmp_path, buffer_size, trailing_newline):
    batches = synth.make_batches(20260020, 2, 7, multiplicity=120)
    path = tmp_path / "particle_samples.gz"
    write_mixed_species(path, batches, trailing_newline=trailing_newline, empty_event_at=5)
    same(reference_reader(path, 211, buffer_size), collect(path, 211, buffer_size))
    same(reference_reader(path, 321, buffer_size), collect(path, 321, buffer_size))  # sparse species, empty events


def test_rapidity_shift_and_cut(tmp_path):
    batches = synth.make_batches(20260021, 1, 5, multiplicity=200)
    path = tmp_path / "particle_samples.gz"
    write_mixed_species(path, batches)
    cut = HBTParams(HBTrap_min=-0.3, HBTrap_max=0.2)
    ref = reference_reader(path, 211, 400, rap_shift=0.37, cut=cut)
    got = collect(path, 211, 400, rap_shift=0.37, rapidity_cut=cut)
    same(ref, got)
    assert sum(len(e) for evs, _ in got for e in evs) < sum(len(e) for e in batches[0].same)  # the cut did something


def test_same_batches_as_the_writer(tmp_path):
    """synth.write_iss_gz -> FastReader returns the batches it was given (oversample groups of
    exactly `oversample` events: event_buffer_size = oversample x multiplicity)."""
    batches = synth.make_batches(20260022, 3, 4, multiplicity=150)
    path = tmp_path / "particle_samples.gz"
    synth.write_iss_gz(str(path), batches)
    got = collect(path, 211, 4 * 150)
    assert len(got) == 3
    for (evs, n), b in zip(got, batches):
        assert n == 600 and len(evs) == 4
        for x, y in zip(evs, b.same):
            assert np.array_equal(x, y)


def test_errors(tmp_path):
    import ctypes

    L = capi.lib()
    h = ctypes.c_void_p()
    assert L.hbt_reader_open(str(tmp_path / "missing.gz").encode(), 10, 211, 100, 0.0, None, ctypes.byref(h)) == -1
    path = tmp_path / "p.gz"
    with gzip.open(path, "wt") as f:
        f.write("2\n211 0.138 1 0 0 0 1 0.1 0.1 0.1\n")  # the event announces 2 particles, holds 1
    assert L.hbt_reader_open(str(path).encode(), 6, 211, 100, 0.0, None, ctypes.byref(h)) == -1  # not a read_in_mode of the reference
    assert L.hbt_reader_open(str(path).encode(), 8, 211, 100, 0.0, None, ctypes.byref(h)) == -1  # not an extended SMASH binary
    assert L.hbt_reader_open(str(path).encode(), 10, 9999, 100, 0.0, None, ctypes.byref(h)) == -1  # species groups need pdg.dat
    r = hbtio.FastReader(str(path), 211, 100)
    with pytest.raises(capi.HBTError):
        next(r)
    r.close()


# ---- UrQMD and OSCAR formats (read_in_mode 2, 21, 1, 0) against the reference reader's own output ----
import json
import os

from conftest import GOLDEN

READER_CASES = json.load(open(os.path.join(GOLDEN, "reader_cases.json")))


@pytest.mark.parametrize("name", sorted(READER_CASES))
def test_urqmd_modes_match_the_reference_reader(name):
    """tests/golden/make_golden_readers.py pushed the committed particle_list.dat (gzipped UrQMD
    text) / particle_list.bin (UrQMD binary) / f13.dat (UrQMD file-13 text) / OSCAR.DAT through the unmodified reference reader and dumped
    the filtered particle lists per batch; hbt_reader must return the same batches, events and
    doubles (id map, species filter, unknown ids and other species counted for event_buffer_size,
    an empty event, float32 -> double, the rapidity shift)."""
    c = READER_CASES[name]
    want = hbtio.read_batches(os.path.join(GOLDEN, name + ".particles.bin"))
    # when the file ends exactly on a batch boundary the reference's loop makes one more read that
    # returns no event (src/Analysis.cpp:821: an analysis call on zero events, a no-op); hbt_reader
    # reports the end of the file instead
    want = [b for b in want if len(b.same) > 0]
    got = collect(os.path.join(GOLDEN, c["input"]), c["particle_monval"], c["event_buffer_size"],
                  rap_shift=c["rapidity_shift"], read_in_mode=c["read_in_mode"])
    assert len(got) == len(want) > 0
    for (evs, _), b in zip(got, want):
        assert len(evs) == len(b.same)
        for x, y in zip(evs, b.same):
            assert x.shape == y.shape and np.array_equal(x, y)


def test_urqmd_text_and_binary_agree_up_to_float32(tmp_path):
    """The two writers of synth.py describe the same events; the binary format stores float32."""
    rng = np.random.default_rng(5)
    events = [ev for b in synth.make_batches(20260023, 1, 3, multiplicity=50) for ev in b.same]
    rec = synth.urqmd_records(events, rng)
    synth.write_urqmd_gz(str(tmp_path / "particle_list.dat"), rec, trailing_newline=False)
    synth.write_urqmd_bin(str(tmp_path / "particle_list.bin"), rec)
    a = collect(tmp_path / "particle_list.dat", 211, 10 ** 9, read_in_mode=2)
    b = collect(tmp_path / "particle_list.bin", 211, 10 ** 9, read_in_mode=21)
    assert len(a) == len(b) == 1 and a[0][1] == b[0][1] == sum(len(r) for r in rec)
    for x, y, ev in zip(a[0][0], b[0][0], events):
        assert np.array_equal(x, ev)  # %.17g text: the generator's doubles
        assert np.array_equal(y, x.astype(np.float32).astype(np.float64))


def test_urqmd_truncated_files_are_errors(tmp_path):
    rng = np.random.default_rng(6)
    events = [ev for b in synth.make_batches(20260024, 1, 2, multiplicity=20) for ev in b.same]
    rec = synth.urqmd_records(events, rng)
    synth.write_urqmd_bin(str(tmp_path / "full.bin"), rec)
    data = open(tmp_path / "full.bin", "rb").read()
    open(tmp_path / "cut.bin", "wb").write(data[:len(data) - 17])
    r = hbtio.FastReader(str(tmp_path / "cut.bin"), 211, 10 ** 9, read_in_mode=21)
    with pytest.raises(capi.HBTError):
        next(r)
    r.close()
    with gzip.open(tmp_path / "cut.dat", "wt") as f:
        f.write("3 \nskipped\n101 2 1 1 6 99 0.138 1 0 0 0 1 0.1 0.1 0.1\n")
    r = hbtio.FastReader(str(tmp_path / "cut.dat"), 211, 10 ** 9, read_in_mode=2)
    with pytest.raises(capi.HBTError):
        next(r)
    r.close()


REF_FIXTURES = "/root/reference/unit_tests/test_reader_files"


@pytest.mark.skipif(not os.path.isdir(REF_FIXTURES), reason="the reference's own fixtures are only in the build container")
@pytest.mark.parametrize("name,mode,fn", [("unit_urqmd_txt", 1, "particle_list.dat"), ("unit_oscar_kplus", 0, "OSCAR.DAT")])
def test_reference_unit_test_fixtures(name, mode, fn):
    """The reference's own reader fixtures (unit_tests/test_reader_files: two UrQMD events as file-13 text and as
    OSCAR1997A) through hbt_reader, against the particle lists the reference reader produced from them
    (tests/golden/<name>.particles.bin, made by tests/golden/make_golden.py)."""
    import json as _json
    meta = _json.load(open(os.path.join(GOLDEN, "cases.json")))[name]
    want = [b for b in hbtio.read_batches(os.path.join(GOLDEN, name + ".particles.bin")) if len(b.same) > 0]
    got = collect(os.path.join(REF_FIXTURES, fn), meta["params"]["particle_monval"], 100000, read_in_mode=mode)
    assert len(got) == len(want) > 0
    for (evs, _), b in zip(got, want):
        assert len(evs) == len(b.same)
        for x, y in zip(evs, b.same):
            assert x.shape == y.shape and np.array_equal(x, y)
