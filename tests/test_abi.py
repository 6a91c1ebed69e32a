"""The C-ABI: both libraries export every symbol include/rpbmd.h declares; the product library has no CPU fallback."""
import ctypes
import os
import re

import pytest

from reactive_pb_nn_md_b200 import _binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "rpbmd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rpb_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_binding.ABI_SYMBOLS)


@pytest.mark.parametrize("path", [_binding.CUDA_LIB_PATH, os.path.join(ROOT, "oracle", "librpbmd_oracle.so")])
def test_library_exports_every_declared_symbol(path, oracle_lib):
    assert os.path.exists(path), "build first: python __graft_entry__.py"
    dll = ctypes.CDLL(path)
    for name in header_symbols():
        assert hasattr(dll, name), (path, name)


def test_backends_named(oracle_lib):
    assert oracle_lib.backend == "oracle-cpu"
    lib = _binding.Library(_binding.CUDA_LIB_PATH)          # loading needs no GPU
    assert lib.backend == "cuda-sm100a"
    assert lib.dll.rpb_timer_count() > 10
    names = [lib.dll.rpb_timer_name(i).decode() for i in range(lib.dll.rpb_timer_count())]
    assert "pme_spread" in names and "evb_jacobi" in names


def test_cuda_library_fails_loudly_without_gpu():
    """No CPU fallback: without a device rpb_create must return RPB_ERR_CUDA, never compute anything."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    from reactive_pb_nn_md_b200 import engine
    from tests.util import small_params, water_system
    lib = _binding.Library(_binding.CUDA_LIB_PATH)
    with pytest.raises(_binding.RpbError) as ei:
        engine.Simulation(water_system(10), small_params(), library=lib)
    assert ei.value.code == -2


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "reactive_pb_nn_md_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "librpbmd_oracle" not in text and "oracle/" not in text.replace("tests/ additionally load the CPU oracle", ""), f
