import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` on the GPU box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, f"{name}.npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def csr_to_lists(ids, off):
    return [ids[off[i]:off[i + 1]] for i in range(len(off) - 1)]


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(scope="session")
def cuda_device():
    """GPU tests must run on the real thing: fail (not skip) when CUDA or librir.so is missing."""
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import research_image_retrieval_b200 as rir

    lib = rir.load()
    assert lib.rir_device_check() == 0, lib.rir_last_error()
    return torch.device("cuda", 0)
