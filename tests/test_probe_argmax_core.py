"""CPU check of the tournament argmax used by probe_argmax_rows_t_kernel (csrc/probe_argmax_core.cuh).

The header compiles as plain C++; tests/native/probe_argmax_core_check.cpp runs it against the sequential
first-maximal-index loop (the definition: torch.argmax of the interpolated logits, model/evaluator.py:71,106) on
random, heavily tied, NaN / inf / signed-zero, all-equal and padded inputs for every channel count the kernel is
instantiated for.  The GPU parity tests (tests/test_gpu_eval.py) cover the kernel itself."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "expand-and-quantize-for-unsupervised-semantic-segmentation_b200", "csrc")


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_tournament_argmax_equals_sequential_loop(tmp_path):
    exe = str(tmp_path / "pacheck")
    src = os.path.join(ROOT, "tests", "native", "probe_argmax_core_check.cpp")
    # -ffp-contract=off: the host emulation of mul.rn / fma.rn must not be re-associated
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-I", CSRC, src, "-o", exe], check=True)
    out = subprocess.run([exe, "100000"], check=True, capture_output=True, text=True, timeout=120).stdout
    assert out.startswith("ok "), out
    assert int(out.split()[1]) == 8 * 100000
