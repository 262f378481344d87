"""Downstream consumers of the frozen output format: our event
averaging and radius fit against golden vectors produced by the reference's own scripts
(tests/golden/make_golden_ebe.py ran /root/reference/ebe_scripts/*.py in the build container)."""
import gzip
import os
import sys
import types

import numpy as np
import pytest

from conftest import GOLDEN
from hadronic_afterburner_toolkit_b200.ebe_scripts import average_event_HBT_correlation_function as avg
from hadronic_afterburner_toolkit_b200.ebe_scripts import fit_HBT_radii as fit


def _groups(npz, prefix):
    return {k[len(prefix):]: npz[k] for k in npz.files if k.startswith(prefix)}


def test_event_average_matches_the_reference_script(tmp_path):
    g = np.load(os.path.join(GOLDEN, "ebe_avg.npz"))
    want = _groups(g, "out/")
    for key, table in _groups(g, "in/").items():
        ev, name = key.split("/")
        d = tmp_path / "work" / f"UrQMD_{ev}" / "UrQMD_results"
        d.mkdir(parents=True, exist_ok=True)
        np.savetxt(d / name, table, fmt="%18.8e", delimiter="    ")
    got = avg.average_event_folders(str(tmp_path / "work"), str(tmp_path / "avg"))
    assert sorted(got) == sorted(want)
    for name in want:
        # `want` was read back from the reference script's text output (11 significant digits)
        np.testing.assert_allclose(got[name], want[name], rtol=1e-10, atol=0)
        # the file on disk: same layout as the reference's savetxt ('%.10e', two blanks)
        ref_text = bytes(g["text/" + name]).decode().splitlines()
        our_text = open(tmp_path / "avg" / name).read().splitlines()
        assert len(our_text) == len(ref_text)
        assert [len(l.split("  ")) for l in our_text] == [len(l.split("  ")) for l in ref_text]
        # (the folder order of glob() is the file system's, so a sum may differ in its last bit
        # and a printed digit may flip: compare the numbers, not the characters)
        np.testing.assert_allclose(np.loadtxt(tmp_path / "avg" / name), want[name], rtol=2e-10)
    assert avg.main(["prog"]) == 1  # usage
    with pytest.raises(FileNotFoundError):
        avg.average_event_folders(str(tmp_path / "nothing"), str(tmp_path / "avg2"))
    with pytest.raises(ValueError):
        avg.average_tables([np.zeros((3, 5)), np.zeros((4, 5))])


def test_radius_fit_matches_the_reference_script():
    g = np.load(os.path.join(GOLDEN, "ebe_fit.npz"))
    want = _groups(g, "out/")
    assert bytes(g["header"]).decode() == fit.HEADER
    for cut, table in _groups(g, "in/").items():
        got = fit.fit_table(table)
        assert got.shape == want[cut].shape == (5, 13)
        np.testing.assert_allclose(got, want[cut], rtol=1e-8, atol=1e-10)
    assert fit.grid_points(41 ** 3) == 41 and fit.grid_points(31 ** 3) == 31
    with pytest.raises(ValueError):
        fit.grid_points(1000 + 1)


def test_database_flow_with_a_stand_in_for_h5py(monkeypatch):
    """fit_database is the reference's in-place HDF5 flow; h5py is not installed here, so the
    flow is exercised against a dict-backed stand-in with the h5py calls the flow makes."""
    g = np.load(os.path.join(GOLDEN, "ebe_fit.npz"))

    class Attrs(dict):
        def create(self, k, v):
            self[k] = v

    class Dataset:
        def __init__(self, data):
            self.data, self.attrs = np.array(data), Attrs()

    class Group(dict):
        def get(self, k):
            v = self[k]
            return v.data if isinstance(v, Dataset) else v

        def create_dataset(self, name, data=None, **kw):
            assert kw == {"compression": "gzip", "compression_opts": 9}
            self[name] = Dataset(data)
            return self[name]

    class File(dict):
        closed = False

        def close(self):
            self.closed = True

    db = File(event_0=Group({f"HBT_correlation_function_KT_{c}.dat": t for c, t in _groups(g, "in/").items()}))
    db["event_0"]["HBT_radii_KT_0_0.2.dat"] = Dataset(np.zeros((1, 1)))  # a stale result is replaced
    fake = types.ModuleType("h5py")
    fake.File = lambda name: db
    monkeypatch.setitem(sys.modules, "h5py", fake)
    fit.fit_database("database.h5", verbose=False)
    assert db.closed
    for cut, want in _groups(g, "out/").items():
        ds = db["event_0"][f"HBT_radii_KT_{cut}.dat"]
        np.testing.assert_allclose(ds.data, want, rtol=1e-8, atol=1e-10)
        assert bytes(ds.attrs["header"]).decode() == fit.HEADER


def test_fit_on_the_reference_binarys_own_output_file(tmp_path):
    """One of the reference binary's files (golden, 31^3 rows x 5 columns): the adapter builds the
    correlation + error columns and the fit runs on the bins that hold pairs."""
    src = os.path.join(GOLDEN, "c1_iss_gz_rap10.HBT_correlation_function_KT_0_1.dat.gz")
    table5 = np.loadtxt(gzip.open(src))
    assert table5.shape == (31 ** 3, 5)
    t8, ok = fit.correlation_table(table5)
    assert t8.shape == (31 ** 3, 8) and ok.sum() > 1000
    np.testing.assert_array_equal(t8[ok, 6], table5[ok, 3] / table5[ok, 4])
    assert np.all(t8[~ok, 6] == 0.0)
    dst = tmp_path / "HBT_correlation_function_KT_0_1.dat"
    np.savetxt(dst, table5, fmt="%18.8e", delimiter="    ")
    radii = fit.fit_dat_file(str(dst), q_cut_max_list=(0.1, 0.15))
    assert radii.shape == (2, 13) and np.all(np.isfinite(radii))
    saved = np.loadtxt(tmp_path / "HBT_radii_KT_0_1.dat")
    np.testing.assert_allclose(saved, radii, rtol=1e-9)
    assert open(tmp_path / "HBT_radii_KT_0_1.dat").readline().startswith("# q_cut[GeV]  lambda")
    with pytest.raises(ValueError):
        fit.correlation_table(np.zeros((8, 4)))
