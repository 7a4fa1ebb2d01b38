"""Host check of the split-tf32 contraction used by the probe-logits, expansion-head, split-tf32 assign and split-tf32
kNN kernels (csrc/probe_logits_tc.cu, head_gemm_tc.cu, pq_assign_tc_kernel.cuh, knn_tc.cu):

    x . w  ~=  hi(x).hi(w) + lo(x).hi(w) + hi(x).lo(w),     hi = the fp32 word with its low 13 mantissa bits cleared
                                                            (what the tensor core reads of an fp32 operand),
                                                            lo = x - hi (exact in fp32), itself read as tf32

The dropped lo.lo term and the truncation of lo are each ~2^-21 relative per element, so the three-product sum
carries fp32-level accuracy (3-6e-6 of the output scale for the head, 4e-7 for the probe logits)
where one tf32 product is ~1e-3.  Operands are emulated bit-exactly in numpy; accumulation in float64 isolates the
operand-side error.  No device code runs here."""
import numpy as np
import pytest


def _tf32(x: np.ndarray) -> np.ndarray:
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


@pytest.mark.parametrize("D,C", [(1024, 54), (384, 384), (16, 256)])
def test_three_tf32_products_reach_fp32_accuracy(D, C):
    rng = np.random.default_rng(D + C)
    x = (rng.standard_normal((512, D)) * np.exp(rng.standard_normal((512, 1)))).astype(np.float32)
    w = (rng.standard_normal((C, D)) * 0.05).astype(np.float32)
    xh, wh = _tf32(x), _tf32(w)
    xl, wl = x - xh, w - wh                                   # exact: the low 13 bits as an fp32 number
    assert np.array_equal(xh.astype(np.float64) + xl.astype(np.float64), x.astype(np.float64))
    xl_t, wl_t = _tf32(xl), _tf32(wl)                         # the tensor core drops the low bits of lo as well
    f = np.float64
    exact = x.astype(f) @ w.astype(f).T
    one = xh.astype(f) @ wh.astype(f).T
    three = one + xl_t.astype(f) @ wh.astype(f).T + xh.astype(f) @ wl_t.astype(f).T
    scale = np.linalg.norm(x.astype(f), axis=1)[:, None] * np.linalg.norm(w.astype(f), axis=1)[None, :]
    e1 = float((np.abs(one - exact) / scale).max())
    e3 = float((np.abs(three - exact) / scale).max())
    assert e3 <= 2.0 ** -20, e3                               # Cauchy-Schwarz bound: 3 terms of <= 2^-21..2^-22 each
    assert e3 < 1e-3 * e1 or e1 < 1e-6                        # three orders of magnitude better than one tf32 product
    # the fp32 reference itself (sequential fp32 accumulation) is no closer to float64 than the split is
    ref32 = (x @ w.T).astype(f)
    assert e3 <= 4 * float((np.abs(ref32 - exact) / scale).max()) + 2.0 ** -22
