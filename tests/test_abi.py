"""The C-ABI library loads and exports every symbol include/equss_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "equss_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(equss_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    import equss_b200
    assert sorted(equss_b200._native.EXPORTS) == _declared_symbols()


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    import equss_b200
    lib = ctypes.CDLL(equss_b200._native.lib_path())
    for sym in _declared_symbols():
        assert hasattr(lib, sym), f"{sym} is declared in include/equss_b200.h but not exported"


def test_version_and_error_string():
    import equss_b200
    L = equss_b200._native.load()
    assert L.equss_version() >= 100
    assert isinstance(L.equss_last_error_string(), bytes)
    assert L.equss_probe_cpad(27) == 28 and L.equss_probe_cpad(54) == 56


def test_no_cpu_fallback():
    """CPU tensors are rejected loudly instead of being computed some other way."""
    import torch
    import equss_b200
    from equss_b200 import ops
    with pytest.raises(equss_b200._native.EqussNativeError):
        ops.pq_assign(torch.randn(8, 16), torch.randn(2, 4, 8))
    with pytest.raises(equss_b200._native.EqussNativeError):
        ops.confusion_update(torch.zeros(4, dtype=torch.long), torch.zeros(4, dtype=torch.long), 3,
                             torch.zeros(3, 3, dtype=torch.long))


def test_argument_errors_mirror_reference():
    """Bad `normalize` / indivisible embed_dim raise ValueError like model/quantizer.py:455,574-575."""
    import torch
    from equss_b200 import _native as N
    with pytest.raises(ValueError, match="divisible"):
        N.zdesc_for(torch.zeros(4, 10), 3)
