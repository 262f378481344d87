import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: longer-running parity case")


@pytest.fixture(scope="session")
def golden_cases():
    with open(os.path.join(GOLDEN, "cases.json")) as f:
        return json.load(f)


def load_golden(name):
    from hadronic_afterburner_toolkit_b200 import hbtio
    from hadronic_afterburner_toolkit_b200.params import HBTParams

    with open(os.path.join(GOLDEN, "cases.json")) as f:
        meta = json.load(f)[name]
    P = HBTParams(**meta["params"])
    batches = hbtio.read_batches(os.path.join(GOLDEN, name + ".particles.bin"))
    ref = hbtio.load_accumulators_npz(os.path.join(GOLDEN, name + ".ref.npz"))
    return P, batches, ref, meta


GOLDEN_NAMES = sorted(json.load(open(os.path.join(GOLDEN, "cases.json"))).keys())
