"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol the header
declares, and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from hydra_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        capi.build()
    return capi.load()


def _declared():
    txt = open(os.path.join(ROOT, "include", "hydra_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_all_exported(lib):
    from hydra_b200 import capi
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hydra_b200.h but not exported"
    assert sorted(capi.EXPORTS) == names


def test_abi_version(lib):
    assert lib.hb_abi_version() == 2


def test_struct_sizes_match_header(lib):
    from hydra_b200 import capi
    assert C.sizeof(capi.HbConfig) == lib.hb_sizeof_config()
    assert C.sizeof(capi.HbBrrTape) == lib.hb_sizeof_brr_tape() == 12 * 8
    assert C.sizeof(capi.HbFhConfig) == 5 * 8
    assert C.sizeof(capi.HbBrrIterOut) == lib.hb_sizeof_iter_out()


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import hydra_b200
    with pytest.raises(hydra_b200.HydraError, match="no CUDA device|CUDA"):
        hydra_b200.GenotypeStore(100, 10)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hydra_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "hydra_oracle" not in src, f
