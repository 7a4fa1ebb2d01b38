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


def _schedule(n_items, ctas, rows):
    import ctypes
    import __graft_entry__
    __graft_entry__.build()
    import equss_b200
    L = equss_b200._native.load()
    out = (ctypes.c_int32 * 3)()
    rc = L.equss_probe_argmax_schedule(n_items, ctas, rows, out)
    return rc, tuple(out)


def test_probe_schedule_covers_every_row_once():
    """Host logic of the persistent probe kernel (pure host code in the C-ABI library, no device call): decoding
    every slot the way the kernel does must visit every (item, row) exactly once, the split remainder must fit in one
    round, and the cocostuff27 shape (32 images x 41 row blocks on 296 CTAs) gets four full rounds + one half round."""
    import random
    rc, (n_full, n_sched, parts) = _schedule(32 * 41, 296, 8)
    assert rc == 0 and (n_full, n_sched, parts) == (1184, 1184 + 128 * 2, 2)
    rng = random.Random(0)
    cases = [(1, 1, 8), (5, 5, 8), (7, 3, 1), (296, 296, 8), (297, 296, 8), (1000, 296, 6), (3 * 9, 10, 3)]
    cases += [(rng.randint(1, 5000), rng.randint(1, 600), rng.choice([1, 2, 3, 4, 5, 6, 7, 8])) for _ in range(300)]
    for n_items, ctas, rows in cases:
        ctas = min(ctas, n_items)
        rc, (n_full, n_sched, parts) = _schedule(n_items, ctas, rows)
        assert rc == 0
        rem = n_items - n_full
        assert n_full % ctas == 0 and 0 <= rem < ctas and n_sched == n_full + rem * parts
        assert rows % parts == 0 and rem * parts <= ctas and (parts == 1 or rem > 0)
        # the next power of two would not fit (or does not divide the rows)
        assert rem == 0 or rows % (2 * parts) != 0 or rem * parts * 2 > ctas
        seen = {}
        for slot in range(n_sched):                      # the kernel's decode (probe_argmax_rows_t_kernel)
            item, r_lo, r_hi = slot, 0, rows
            if slot >= n_full:
                t = slot - n_full
                item = n_full + t // parts
                r_lo = (t % parts) * (rows // parts)
                r_hi = r_lo + rows // parts
            for r in range(r_lo, r_hi):
                seen[(item, r)] = seen.get((item, r), 0) + 1
        assert len(seen) == n_items * rows and set(seen.values()) == {1}
    assert _schedule(4, 8, 8)[0] != 0 and _schedule(0, 1, 8)[0] != 0      # ctas > n_items / empty: rejected
