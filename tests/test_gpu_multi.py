"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the NVLink peer-memory exchange of the sharded
search against the NCCL all-gather path and the oracle.  Runs tests/multi_gpu_worker.py under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_search_peer_exchange(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29631 + world), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"multi_worker_w{world}.log"), "w") as f:  # the full worker output, for debugging
        f.write(r.stdout + "\n---- stderr ----\n" + r.stderr)
    errs = [l for l in r.stderr.splitlines() if "Error" in l or "assert" in l.lower()]
    assert r.returncode == 0 and "multi-gpu worker OK" in r.stdout, "\n".join(errs[-12:])
