"""CPU test: the C-ABI shared library loads without a GPU and exports every symbol include/mvd_b200.h declares;
argument validation fails loudly (no compute is attempted here)."""
import ctypes
import os

import pytest


def test_library_exports_every_declared_symbol():
    from mvd_b200 import _lib

    lib = _lib.lib()
    declared = _lib.declared_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    unbound = [s for s in declared if s not in _lib._SIGNATURES]
    assert not unbound, f"declared in the header but not bound in _lib.py: {unbound}"
    assert lib.mvd_abi_version() == 1


def test_invalid_arguments_are_hard_errors():
    from mvd_b200 import _lib

    lib = _lib.lib()
    # K not a multiple of 64 -> MVD_ERR_INVALID before anything touches a device
    rc = lib.mvd_linear_bf16(None, 0, 100, None, 0, 0, None, 0, None, None, 0, 0, None, 0, None, 0, 8, 64, 0, 0, None)
    assert rc == -1
    assert b"multiple of 64" in lib.mvd_last_error()
    rc = lib.mvd_attention_bf16(None, 0, 0, None, 0, 0, None, 0, 0, None, 0, 0, 0, 5, 64, 64, 0.125, None)
    assert rc == -1 and b"empty problem" in lib.mvd_last_error()
    rc = lib.mvd_small_linear_f32(None, 0, None, None, None, 0, 17, 8, 8, 0, 0, None)
    assert rc == -1


def test_product_refuses_cpu_tensors():
    import torch
    from mvd_b200 import ops

    with pytest.raises(ValueError, match="CUDA"):
        ops.linear(torch.zeros(8, 64, dtype=torch.bfloat16), torch.zeros(64, 64, dtype=torch.bfloat16))
    with pytest.raises(ValueError, match="CUDA"):
        ops.groupnorm(torch.zeros(1, 4, 64, dtype=torch.bfloat16), torch.ones(64, dtype=torch.bfloat16),
                      torch.zeros(64, dtype=torch.bfloat16))


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mvd_b200")
    for fn in os.listdir(root):
        if fn.endswith(".py"):
            src = open(os.path.join(root, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
