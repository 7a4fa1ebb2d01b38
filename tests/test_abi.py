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


def test_host_side_shape_planning():
    """The `*_supported` predicates and the workspace planners are host code (no device call): they answer on a box
    without a GPU, follow the support matrix include/equss_b200.h documents, and size the workspaces the kernels need --
    in particular no N x N (or shard x N) similarity matrix for the kNN path."""
    import torch
    import equss_b200
    from equss_b200 import _native as N
    L = N.load()
    # fused assign + gather: l2 rows, d in {16, 32}, K <= 256, flat and NCHW (equss_b200.h: equss_pq_assign_gather)
    flat, _, _ = N.zdesc_for(torch.zeros(64, 1024), 64)
    nchw, _, _ = N.zdesc_for(torch.zeros(2, 1024, 4, 8), 64)
    sup = lambda zd, M, K, d, mode: L.equss_pq_assign_gather_supported(ctypes.byref(zd), M, K, d, mode)
    assert sup(flat, 64, 256, 16, 1) == 1 and sup(nchw, 64, 256, 16, 1) == 1 and sup(flat, 32, 256, 32, 1) == 1
    assert sup(flat, 16, 512, 64, 1) == 0                      # C4 shape: two passes (assign, then gather)
    assert sup(flat, 64, 256, 16, 2) == 0 and sup(flat, 64, 256, 16, 0) == 0      # z_norm / none rows: not fused
    # assign workspace: the exact SIMT scan needs none; the tensor-core kernels need operand images + the re-score list
    ws = [L.equss_pq_assign_workspace_bytes(51200, 64, 256, 16, a) for a in (0, 1, 2, 3)]
    assert ws[1] == 0 and ws[0] > 0 and ws[2] > 0 and ws[3] > 0
    assert ws[2] < 4 * 51200 * 64 * 256 // 16                  # far below a distance matrix (3.4 GB at this shape)
    # probe logits on the tensor cores: D % 32 == 0, C_pad <= 64, h * w % 32 == 0
    tc = L.equss_probe_logits_tc_supported
    assert tc(1024, 40, 40, 56) == 1 and tc(1000, 40, 40, 56) == 0 and tc(1024, 40, 40, 68) == 0 and tc(1024, 5, 5, 56) == 0
    assert L.equss_probe_image_bytes(1024, 56) == 2 * 4 * 1024 * 64       # hi + lo tf32 pieces of [D][64] (512 KB)
    # kNN: workspace = fp16 copies + per-split survivor lists, O(rows) -- against 1.25 GB for a 6250 x 50000 fp32 shard
    kws = L.equss_knn_workspace_bytes(6250, 50000, 768, 8)
    assert 2 * (6250 + 50000) * 768 <= kws < 6250 * 50000 * 4 // 4
    assert L.equss_knn_workspace_bytes(100, 100, 100, 8) > 0
    # soft-assignment statistics kernel, EMA tail scratch, expansion-head GEMM
    soft = L.equss_pq_soft_stats_supported
    assert soft(256, 16) == 1 and soft(37, 16) == 1 and soft(2048, 16) == 0
    assert L.equss_pq_train_tail_scratch_floats(64) >= 10 * 64
    hg = L.equss_head_gemm_supported
    assert hg(384, 0, 1600, 1) == 1 and hg(384, 384, 1600, 1) == 1 and hg(768, 768, 784, 1) == 1 and hg(100, 0, 1600, 1) == 0
    assert L.equss_last_error_string() == b"ok" or isinstance(L.equss_last_error_string(), bytes)


def test_product_path_never_touches_the_oracle():
    """oracle/ is the checker: only tests/, smoke() and bench.py's CPU legs may import it.  The package (host mirrors
    and CUDA sources) must not mention it, and it must fail loudly -- not fall back -- without the native library."""
    pkg = os.path.join(ROOT, "expand-and-quantize-for-unsupervised-semantic-segmentation_b200")
    for base, _, files in os.walk(pkg):
        if "_obj" in base or "__pycache__" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "equss_oracle" not in text and "oracle/" not in text and "import oracle" not in text, f
                assert "kernel_standins" not in text, f          # the CPU suite's torch stand-ins live in tests/ only
    import equss_b200
    assert issubclass(equss_b200._native.EqussNativeError, RuntimeError)
