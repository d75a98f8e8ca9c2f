"""Bit-identity of the i-sharded multi-GPU step against the single-GPU step (needs >= 2 GPUs)."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import REPO_ROOT

pytestmark = pytest.mark.gpu


def test_sharded_equals_unsharded_bitwise(mapc, gpu):
    if gpu < 2:
        pytest.skip("needs at least 2 visible GPUs (run with gpurun --gpus 2)")
    world = 8 if gpu >= 8 else (4 if gpu >= 4 else 2)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(REPO_ROOT, "tests", "mgpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO_ROOT)
    sys.stdout.write(res.stdout[-4000:])
    sys.stderr.write(res.stderr[-4000:])
    assert res.returncode == 0
    assert "bit-identical=True" in res.stdout and "bit-identical=False" not in res.stdout
