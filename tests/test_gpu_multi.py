"""Multi-GPU path on real devices: torchrun, one rank per GPU, NCCL all-reduce inside
libhbt_b200.so.  Skipped on a single-GPU box (the gloo twin tests/test_sharding_gloo.py covers
the host logic everywhere)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_sharded_groups_nccl_allreduce(tmp_path):
    from hadronic_afterburner_toolkit_b200 import capi

    ndev = capi.lib().hbt_device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(ndev, 4)
    ok = tmp_path / "ok.txt"
    env = dict(os.environ, HBT_MP_OK=str(ok))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "tests", "mp_gpu_worker.py")], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert ok.read_text().startswith("ok")
