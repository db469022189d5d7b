"""CPU tier: the C-ABI library loads, exports every symbol the header declares,
and refuses to run without a GPU (no silent CPU fallback)."""
import os
import re

import pytest

from pyfvm_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "fvm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fvm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = engine.load_library()
    names = _declared()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/fvm_b200.h but not exported"
    assert sorted(engine.SYMBOLS) == names, "engine.SYMBOLS must mirror the header"
    assert lib.fvm_version() >= 100


def test_kernel_source_embeds_template():
    lib = engine.load_library()
    import ctypes

    need = ctypes.c_size_t()
    lib.fvm_kernel_source(b"// functors here\n", None, 0, ctypes.byref(need))
    buf = ctypes.create_string_buffer(need.value)
    lib.fvm_kernel_source(b"// functors here\n", buf, need.value, ctypes.byref(need))
    src = buf.value.decode()
    assert "struct FvmTileArgs" in src and "cp.async.bulk" in src and "// functors here" in src


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.FvmError):
        engine.Engine(0)
