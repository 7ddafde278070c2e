"""NVLink peer-memory gradient exchange (lrx_p2p_*): needs >= 2 GPUs on the box, skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_p2p_exchange_matches_nccl_on_two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(here, "p2p_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "P2P_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
