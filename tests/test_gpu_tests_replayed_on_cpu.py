"""The module-level ``-m gpu`` tests, re-run on CPU tensors with the kernel entry points replaced by their torch
definitions (tests/kernel_standins.py).  Same test bodies, same fixtures, same tolerances: only ``DEV`` is "cpu" and the
kernels are stand-ins, so what this covers is (i) the host logic of every nn.Module mirror the GPU tests drive and
(ii) the GPU tests themselves (they run here before they ever reach a device).  Tests that call a kernel entry point
directly -- there is nothing to stand in for -- stay GPU-only."""
import importlib

import pytest

import kernel_standins

_SOURCES = ("test_gpu_modules", "test_gpu_variants", "test_gpu_flags", "test_gpu_flags_dropout")   # module-level tests
_mods = {name: importlib.import_module(name) for name in _SOURCES}
_GPU_ONLY = {
    "test_gpu_modules.test_knn_module_and_npz_contract",        # kNN kernel
    "test_gpu_variants.test_stego_loss_matches_reference",      # feature-correlation kernel
    "test_gpu_variants.test_channel_moments_and_soft_stats_kernels",   # compares kernels with torch: nothing left to check
    "test_gpu_flags.test_ema_gumbel_matches_reference",         # these two call monkeypatch.undo(), which would also
    "test_gpu_flags.test_param_gumbel_matches_reference",       # remove the stand-ins (the draw itself: test_host_paths_cpu)
    "test_gpu_flags_dropout.test_keep_mask_draw_consumes_the_device_generator_like_the_reference",   # CUDA generator
}


@pytest.fixture(autouse=True)
def _cpu_with_standins(monkeypatch):
    kernel_standins.install(monkeypatch)
    for m in _mods.values():
        monkeypatch.setattr(m, "DEV", "cpu")


for _name, _m in _mods.items():
    for _attr in dir(_m):
        if _attr.startswith("test_") and f"{_name}.{_attr}" not in _GPU_ONLY:
            globals()[f"{_attr}__{_name[len('test_gpu_'):]}"] = getattr(_m, _attr)
