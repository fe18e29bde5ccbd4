"""GPU: the multi-GPU index build (NCCL all-to-all) equals the oracle's single-process map.  Needs >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

import pgr_tk_b200 as pg

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world", [1, 2])
def test_distributed_index_build(world):
    if pg.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(HERE, "dist_index_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIST_INDEX_CHECK OK" in r.stdout
