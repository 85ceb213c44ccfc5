"""The C-ABI library loads and exports every symbol include/quadsim_abi.h declares (no compute calls)."""
import ctypes
import os
import re

from uav_reinforcement_learning_control_b200 import build, config, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "quadsim_abi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qs_[a-z_]+)\s*\(", text)))


def test_header_and_binding_agree():
    decl = _declared()
    assert len(decl) >= 15
    assert sorted(engine.exported_symbols()) == decl


def test_library_exports_every_declared_symbol():
    path = build.build_library()
    lib = ctypes.CDLL(path)
    for sym in _declared():
        assert hasattr(lib, sym), sym
    lib.qs_abi_version.restype = ctypes.c_int
    assert lib.qs_abi_version() == 1
    assert lib.qs_params_size() == ctypes.sizeof(config.QsParams)
    lib.qs_last_error_string.restype = ctypes.c_char_p
    assert lib.qs_last_error_string() is not None


def test_no_cpu_fallback_without_gpu():
    """On a box without CUDA the product path must fail loudly, never route through the oracle."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    with pytest.raises(engine.QuadSimError):
        engine.Engine(config.EnvConfig.north_star(), 16)


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "uav_reinforcement_learning_control_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\.{0,2}oracle\b", src, flags=re.M), f"{f} imports oracle"
                assert not re.search(r"#include[^\n]*oracle", src), f"{f} includes oracle sources"
                assert "cpu_ref" not in src and "libcpu_ref" not in src, f"{f} references the CPU port"
