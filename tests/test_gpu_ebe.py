"""End-to-end physics check through the frozen file format: the
correlation-function files written from the GPU accumulators and from the CPU oracle's
accumulators give the same event average and the same fitted HBT radii."""
import os

import numpy as np
import pytest

from hadronic_afterburner_toolkit_b200 import synth
from hadronic_afterburner_toolkit_b200.ebe_scripts import average_event_HBT_correlation_function as avg
from hadronic_afterburner_toolkit_b200.ebe_scripts import fit_HBT_radii as fit
from hadronic_afterburner_toolkit_b200.hbt_correlation import HBT_correlation, write_correlation_function
from hadronic_afterburner_toolkit_b200.params import C3
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu


def test_radii_from_gpu_files_equal_radii_from_cpu_files(tmp_path):
    P = C3.with_(qnpts=21)
    roots = {"gpu": tmp_path / "gpu", "cpu": tmp_path / "cpu"}
    for ev in range(2):  # two "events" (independent analyses), each one oversample group
        batches = synth.make_batches(20260030 + ev, 1, 8, multiplicity=1000)
        h = HBT_correlation(P)
        o = O.Oracle(P)
        for b in batches:
            h.calculate_HBT_correlation_function(b)
            o.process_batch(b)
        for kind, acc in (("gpu", h.accumulators()), ("cpu", o.accumulators())):
            d = roots[kind] / f"UrQMD_{ev}" / "UrQMD_results"
            d.mkdir(parents=True)
            write_correlation_function(str(d), P, acc)
        h.close()
    tables = {k: avg.average_event_folders(str(roots[k]), str(roots[k] / "avg")) for k in roots}
    assert sorted(tables["gpu"]) == sorted(tables["cpu"]) and len(tables["gpu"]) == P.n_KT - 1
    for name in tables["cpu"]:
        # files carry 9 significant digits; sums agree to 1e-10, so the printed numbers agree to a last-digit flip
        np.testing.assert_allclose(tables["gpu"][name], tables["cpu"][name], rtol=2e-8, atol=1e-12)
        rg = fit.fit_dat_file(os.path.join(str(roots["gpu"] / "avg"), name), q_cut_max_list=(0.1, 0.15))
        rc = fit.fit_dat_file(os.path.join(str(roots["cpu"] / "avg"), name), q_cut_max_list=(0.1, 0.15))
        np.testing.assert_allclose(rg, rc, rtol=1e-5, atol=1e-7)
        assert np.all(np.isfinite(rg))
        # the synthetic source is a 4 fm Gaussian in x and y: the sideward radius must come out near it
        assert 2.5 < abs(rg[1, 5]) < 5.5, rg[1]
