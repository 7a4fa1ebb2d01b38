"""ctypes binding of libequss_b200.so (C-ABI declared in include/equss_b200.h).

PyTorch is used only for device memory and streams: every wrapper below takes CUDA tensors, passes
their raw ``data_ptr()`` and the current stream to the C entry point, and raises on any error code.
There is no CPU or eager-PyTorch fallback: if the library is missing, cannot be loaded, or no sm_100
device is visible, calls fail loudly with :class:`EqussNativeError`.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

__all__ = [
    "EqussNativeError", "lib", "lib_path", "load", "NORM_MODES", "ZDesc", "zdesc_for",
    "ASSIGN_AUTO", "ASSIGN_SIMT", "ASSIGN_TCGEN05", "ASSIGN_TCGEN05_TF32", "EXPORTS",
]

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libequss_b200.so"

NORM_MODES = {"none": 0, None: 0, "l2": 1, "z_norm": 2, "z_trainable": 3, "affine": 3}
ASSIGN_AUTO, ASSIGN_SIMT, ASSIGN_TCGEN05, ASSIGN_TCGEN05_TF32 = 0, 1, 2, 3
LAYOUT_FLAT, LAYOUT_NCHW = 0, 1

# every symbol include/equss_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "equss_last_error_string", "equss_version", "equss_device_check", "equss_launch_count",
    "equss_pq_assign_workspace_bytes", "equss_pq_assign", "equss_pq_cnorm2", "equss_pq_gather_loss",
    "equss_pq_gather_loss_bwd", "equss_pq_accumulate", "equss_ema_update", "equss_pq_distance_prob",
    "equss_probe_cpad", "equss_probe_logits", "equss_probe_argmax_schedule", "equss_probe_argmax_confusion", "equss_confusion_update",
    "equss_usage_percentiles", "equss_pq_assign_gather_supported", "equss_pq_assign_gather",
    "equss_probe_image_bytes", "equss_probe_build_image", "equss_probe_logits_tc_supported", "equss_probe_logits_tc",
    "equss_knn_workspace_bytes", "equss_knn_topk",
    "equss_head_gemm_supported", "equss_head_gemm",
    "equss_pq_soft_stats_supported", "equss_pq_soft_stats", "equss_channel_moments",
    "equss_pq_train_tail_scratch_floats", "equss_pq_train_tail", "equss_pq_prepare_codebook",
    "equss_pq_train_tail_peers", "equss_token_gram", "equss_probe_losses_supported", "equss_probe_losses", "equss_stego_feature_corr",
]


class EqussNativeError(RuntimeError):
    pass


class ZDesc(C.Structure):
    _fields_ = [("n_pixels", C.c_int64), ("hw", C.c_int64), ("stride_b", C.c_int64),
                ("stride_s", C.c_int64), ("stride_c", C.c_int64), ("dim", C.c_int32), ("layout", C.c_int32)]


_lib: Optional[C.CDLL] = None


def lib_path() -> str:
    return os.path.join(_PKG_DIR, _LIB_NAME)


def load() -> C.CDLL:
    """Load the shared library (building is the job of ``__graft_entry__.build()`` / ``build.py``)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise EqussNativeError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). equss_b200 has no CPU or PyTorch fallback.")
    try:
        L = C.CDLL(path)
    except OSError as e:  # pragma: no cover
        raise EqussNativeError(f"cannot load {path}: {e}") from e
    _declare(L)
    _lib = L
    return L


def lib() -> C.CDLL:
    return load()


def _declare(L: C.CDLL) -> None:
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
    zp = C.POINTER(ZDesc)
    L.equss_last_error_string.restype = C.c_char_p
    L.equss_last_error_string.argtypes = []
    L.equss_version.restype = i32
    L.equss_device_check.restype = i32
    L.equss_device_check.argtypes = [i32]
    L.equss_launch_count.restype = i64
    L.equss_pq_assign_workspace_bytes.restype = i64
    L.equss_pq_assign_workspace_bytes.argtypes = [i64, i32, i32, i32, i32]
    L.equss_pq_assign.restype = i32
    L.equss_pq_assign.argtypes = [vp, zp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, i64, i32, vp]
    L.equss_pq_cnorm2.restype = i32
    L.equss_pq_cnorm2.argtypes = [vp, i32, i32, i32, vp, vp]
    L.equss_pq_gather_loss.restype = i32
    L.equss_pq_gather_loss.argtypes = [vp, zp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    L.equss_pq_gather_loss_bwd.restype = i32
    L.equss_pq_gather_loss_bwd.argtypes = [vp, zp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.equss_pq_accumulate.restype = i32
    L.equss_pq_accumulate.argtypes = [vp, zp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]
    L.equss_ema_update.restype = i32
    L.equss_ema_update.argtypes = [vp, i32, i32, i32, f64, f64, vp, vp, vp, vp, vp, vp]
    L.equss_pq_distance_prob.restype = i32
    L.equss_pq_distance_prob.argtypes = [vp, zp, vp, vp, i32, i32, i32, i32, vp, vp, f32, vp, vp]
    L.equss_probe_cpad.restype = i32
    L.equss_probe_cpad.argtypes = [i32]
    L.equss_probe_logits.restype = i32
    L.equss_probe_logits.argtypes = [vp, i32, i32, i32, i32, vp, vp, i32, vp, vp]
    L.equss_probe_argmax_schedule.restype = i32
    L.equss_probe_argmax_schedule.argtypes = [i32, i32, i32, C.POINTER(C.c_int32)]
    L.equss_usage_percentiles.restype = i32
    L.equss_usage_percentiles.argtypes = [vp, i64, i64, i32, i32, vp, vp]
    L.equss_pq_assign_gather_supported.restype = i32
    L.equss_pq_assign_gather_supported.argtypes = [zp, i32, i32, i32, i32]
    L.equss_pq_assign_gather.restype = i32
    L.equss_pq_assign_gather.argtypes = [vp, zp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, i64, vp]
    L.equss_probe_image_bytes.restype = i64
    L.equss_probe_image_bytes.argtypes = [i32, i32]
    L.equss_probe_build_image.restype = i32
    L.equss_probe_build_image.argtypes = [vp, i32, i32, vp, vp]
    L.equss_probe_logits_tc_supported.restype = i32
    L.equss_probe_logits_tc_supported.argtypes = [i32, i32, i32, i32]
    L.equss_probe_logits_tc.restype = i32
    L.equss_probe_logits_tc.argtypes = [vp, i32, i32, i32, i32, vp, vp, i32, vp, vp]
    L.equss_head_gemm_supported.restype = i32
    L.equss_head_gemm_supported.argtypes = [i32, i32, i32, i32]
    L.equss_head_gemm.restype = i32
    L.equss_head_gemm.argtypes = [vp, i32, i32, vp, i32, i32, i32, vp, vp, i32, i32, vp, i64, vp]
    L.equss_pq_soft_stats_supported.restype = i32
    L.equss_pq_soft_stats_supported.argtypes = [i32, i32]
    L.equss_pq_soft_stats.restype = i32
    L.equss_pq_soft_stats.argtypes = [vp, zp, vp, vp, i32, i32, i32, i32, vp, vp, f32, vp, vp, vp]
    L.equss_pq_train_tail_scratch_floats.restype = i32
    L.equss_pq_train_tail_scratch_floats.argtypes = [i32]
    L.equss_pq_train_tail.restype = i32
    L.equss_pq_train_tail.argtypes = [vp, i32, i32, i32, f64, f64, vp, vp, vp, vp, vp, i64, f64, vp, vp, vp]
    L.equss_pq_train_tail_peers.restype = i32
    L.equss_pq_train_tail_peers.argtypes = [vp, i32, vp, i32, i32, i32, f64, f64, vp, vp, vp, vp, vp, i64, f64, vp, vp, vp, vp]
    L.equss_pq_prepare_codebook.restype = i32
    L.equss_pq_prepare_codebook.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp]
    L.equss_token_gram.restype = i32
    L.equss_token_gram.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    L.equss_probe_losses_supported.restype = i32
    L.equss_probe_losses_supported.argtypes = [i32] * 9
    L.equss_probe_losses.restype = i32
    L.equss_probe_losses.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp]
    L.equss_stego_feature_corr.restype = i32
    L.equss_stego_feature_corr.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp]
    L.equss_channel_moments.restype = i32
    L.equss_channel_moments.argtypes = [vp, zp, vp, vp]
    L.equss_probe_argmax_confusion.restype = i32
    L.equss_probe_argmax_confusion.argtypes = [vp, i32, i32, i32, i32, vp, i32, i32, i32, i32,
                                               C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                               C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int32), vp]
    L.equss_confusion_update.restype = i32
    L.equss_confusion_update.argtypes = [vp, vp, i64, i32, i32, vp, vp]
    L.equss_knn_workspace_bytes.restype = i64
    L.equss_knn_workspace_bytes.argtypes = [i64, i64, i32, i32]
    L.equss_knn_topk.restype = i32
    L.equss_knn_topk.argtypes = [vp, i64, vp, i64, i32, i32, vp, vp, vp, i64, vp]


# ---------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------
def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().equss_last_error_string().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise EqussNativeError(f"{what} failed with code {rc}: {msg}")


def require_cuda(*tensors: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise EqussNativeError(
                "equss_b200 kernels need CUDA tensors on a B200 (sm_100a); got a "
                f"{t.device} tensor. There is no CPU fallback.")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise EqussNativeError(f"tensors on different devices: {t.device} vs {dev}")
    if dev is None:
        raise EqussNativeError("no tensor given")
    # kernels are launched on the CURRENT device's context (stream handle, SM count, function attributes): a tensor
    # on another device would be launched against the wrong context
    if dev.index is not None and dev.index != torch.cuda.current_device():
        raise EqussNativeError(
            f"tensor lives on {dev} but the current CUDA device is cuda:{torch.cuda.current_device()}; call "
            "torch.cuda.set_device() (one process per GPU) or wrap the call in `with torch.cuda.device(...)`")
    return dev


_checked_devices = set()


def ensure_device(dev: torch.device) -> None:
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx in _checked_devices:
        return
    check(lib().equss_device_check(idx), "equss_device_check")
    _checked_devices.add(idx)


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32 + contiguous view/copy (host plumbing)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def is_channels_last(z: torch.Tensor) -> bool:
    """4-D tensor whose memory is NHWC-dense (and not also NCHW-dense)."""
    return z.dim() == 4 and not z.is_contiguous() and z.is_contiguous(memory_format=torch.channels_last)


def f32_dense(t: torch.Tensor, like: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 tensor in a layout the kernels take without a copy where possible: NCHW- or NHWC-dense 4-D, dense
    otherwise.  With ``like`` the result uses the memory format of ``like`` (gradients of a channels-last z)."""
    if t.dtype != torch.float32:
        t = t.float()
    if like is not None and is_channels_last(like):
        return t.contiguous(memory_format=torch.channels_last)
    if like is None and is_channels_last(t):
        return t
    return t.contiguous()


def zdesc_for(z: torch.Tensor, M: int) -> "tuple[ZDesc, int, str]":
    """Describe an activation tensor without copying it.

    2-D (n, D) contiguous          -> flat layout            (model/quantizer.py:396)
    4-D (B, D, h, w) contiguous    -> NCHW layout            (model/quantizer.py:112 consumes this after a permute)
    Returns (descriptor, d, layout_name).
    """
    if z.dim() == 2:
        n, D = z.shape
        if D % M != 0:
            raise ValueError(f"Embed dim {D} should be divisible by #PQ {M}.")
        zd = ZDesc(n, max(n, 1), max(n, 1) * D, D, 1, D, LAYOUT_FLAT)
        return zd, D // M, "flat"
    if z.dim() == 4:
        B, D, h, w = z.shape
        if D % M != 0:
            raise ValueError(f"Embed dim {D} should be divisible by #PQ {M}.")
        hw = h * w
        if is_channels_last(z):        # (B, D, h, w) view of NHWC memory, e.g. the expansion head's output: flat rows
            zd = ZDesc(B * hw, max(B * hw, 1), max(B * hw, 1) * D, D, 1, D, LAYOUT_FLAT)
            return zd, D // M, "flat"
        zd = ZDesc(B * hw, max(hw, 1), D * hw, 1, hw, D, LAYOUT_NCHW)
        return zd, D // M, "nchw"
    raise ValueError(f"expected a (n, D) or (B, D, h, w) tensor, got shape {tuple(z.shape)}")


def as_voidp_array(ptrs: Sequence[Optional[int]]):
    arr = (C.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


def as_i32_array(vals: Sequence[int]):
    arr = (C.c_int32 * len(vals))()
    for i, v in enumerate(vals):
        arr[i] = int(v)
    return arr
