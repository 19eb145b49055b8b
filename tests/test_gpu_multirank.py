"""Multi-GPU uniq through the C ABI's peer group (ck_peer_export / ck_peer_attach): needs >= 2 GPUs on the box (one process
per GPU -- ranks that wait for one another must not share a GPU), skipped otherwise.  Run it with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4])
def test_peer_group_uniq_matches_single_table_and_oracle(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29510 + world), os.path.join(ROOT, "tests", "multirank_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "MULTIRANK OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
