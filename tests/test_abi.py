"""CPU: librir.so loads and exports every symbol include/rir.h declares; the product path fails loudly without a GPU."""
import os
import re

import pytest
import torch

import research_image_retrieval_b200 as rir
from research_image_retrieval_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rir.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rir_[a-z0-9_]+)\s*\(", text)))


def test_library_built_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run `python -m research_image_retrieval_b200.build`"
    lib = rir.load()
    assert lib.rir_version() == 100


def test_every_declared_symbol_is_exported_and_typed():
    lib = rir.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/rir.h but not exported by librir.so"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert sorted(_lib.SIGNATURES) == declared


def test_constants_match_header():
    text = open(os.path.join(ROOT, "include", "rir.h")).read()
    for name in ["RIR_OK", "RIR_E_ARG", "RIR_E_ARCH", "RIR_E_CUDA", "RIR_E_WORKSPACE", "RIR_F32", "RIR_BF16",
                 "RIR_FP8E4M3", "RIR_POOL_GEM", "RIR_POOL_MAX", "RIR_POOL_AVG", "RIR_PATH_AUTO", "RIR_PATH_STREAM",
                 "RIR_PATH_MMA", "RIR_PATH_EXACT", "RIR_MAP_OK", "RIR_MAP_EMPTY_OK", "RIR_MAP_NO_POS_RETRIEVED"]:
        m = re.search(rf"#define {name} \(?(-?\d+)\)?", text)
        assert m, name
        assert int(m.group(1)) == getattr(_lib, name), name


def test_workspace_query_is_pure_host_code():
    lib = rir.load()
    # 1M x 2048 bf16, top-100: sample + candidates for up to 4096 queries per pass
    w1 = lib.rir_sim_topk_workspace(1, 1007323, 2048, 100, _lib.RIR_BF16)
    w70 = lib.rir_sim_topk_workspace(70, 1007323, 2048, 100, _lib.RIR_BF16)
    w1k = lib.rir_sim_topk_workspace(1024, 1007323, 2048, 100, _lib.RIR_BF16)
    assert 0 < w1 <= w70 < w1k < 1 << 30
    assert lib.rir_sim_topk_workspace(70, 4993, 2048, 4993, _lib.RIR_F32) > 70 * 4993 * 8
    assert lib.rir_sim_topk_workspace(1, 1007323, 2048, 9000, _lib.RIR_BF16) == 0  # k too large for a filtered scan
    assert lib.rir_sim_topk_workspace(1, 100, 64, 10, 99) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = rir.load()
    assert lib.rir_device_check() == _lib.RIR_E_ARCH
    assert b"fallback" in lib.rir_last_error()
    x = torch.rand(2, 4, 3, 3)
    with pytest.raises(TypeError):
        rir.gem_pool(x)
    with pytest.raises(TypeError):
        rir.l2n(torch.rand(3, 4))
    with pytest.raises(TypeError):
        rir.sim_topk(torch.rand(2, 8).bfloat16(), torch.rand(9, 8).bfloat16(), 3)
    with pytest.raises(TypeError):
        rir.pack_descriptors(torch.rand(4, 8))
    # even with device pointers faked, the library itself refuses to launch without an sm_100 device
    assert lib.rir_l2_normalize(1, 1, 1, 1e-12, 1, None) == _lib.RIR_E_ARCH
    assert lib.rir_merge_topk(1, 1, 1, 1, 1, 1, 1, None, 0, None) == _lib.RIR_E_ARCH


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "research_image_retrieval_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
