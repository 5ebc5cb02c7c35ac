"""CPU tier: the product shared library loads and exports every symbol include/slicq.h declares."""
import os
import re

import pytest

from xumx_slicq_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "slicq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(slicq_[a-z_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_cabi.EXPORTS)


def test_library_exports_all_symbols():
    if not os.path.exists(_cabi.library_path()):
        import __graft_entry__
        __graft_entry__.build()
    lib = _cabi.load()
    for s in declared_symbols():
        assert hasattr(lib, s), s
    assert lib.slicq_abi_version() == 1


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(ImportError):
        _cabi.load(str(tmp_path / "libslicq.so"))


def test_cpu_tensors_are_rejected_by_the_product_backend():
    """No CPU fallback: the product backend refuses CPU tensors."""
    import torch
    from xumx_slicq_b200 import NSGTBase
    base = NSGTBase("bark", 262, 32.9, device="cpu")
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        base.nsgt.forward((torch.zeros(1, 1000),))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "xumx_slicq_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) and f != "dft_codelets.cuh":
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), os.path.join(dp, f)
