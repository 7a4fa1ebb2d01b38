"""Host check of the fp16-split assignment's decision rule (csrc/pq_assign_h.cu build_image_kernel).

The tensor-core kernel scores code k of a row by  x1.y1 + x2.y1 + x1.y2 + (b1 + b2 + b3)  with x = z_norm = x1 + x2 + ..,
y = beta * c = y1 + y2 + .., b = -beta |c|^2 / 2 (fp16 pieces, beta a power of two), accumulated in fp32.  A row whose
best and second-best scores differ by more than

    tol = 2^-16 * R + 2^-23 * (2 sqrt(d) + 1),      R = 1.0001 * beta * max|c| + 0.5 * beta * max|c|^2

keeps the winner; every other row is re-scored exactly.  The rule is sound if every score is within tol / 2 of the true
beta * (<x, c> - |c|^2 / 2).  Here the operand pieces are built with numpy's round-to-nearest float16 cast (what
``__floats2half2_rn`` does), the score error is measured against float64, and the whole rule -- approximate scores,
flag, exact re-score of the flagged rows -- is run in numpy and compared with the reference's fp32 argmin.  No device
code runs here; the kernels are compared with the oracle index by index in tests/test_gpu_pq.py / test_gpu_fullsize.py."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def _h(x):
    return x.astype(np.float16).astype(np.float32)


def _split_scores(zn, cbn):
    """(approximate scores fp32 [n, K], tol, beta, R) as build_image_kernel + the MMAs define them."""
    d = cbn.shape[1]
    cn2 = (cbn * cbn).sum(axis=1, dtype=np.float32)
    cmax2 = np.float32(cn2.max())
    cmax = np.float32(np.sqrt(cmax2))
    e = int(np.frexp(cmax)[1]) if 0 < cmax < np.inf else 0
    e = max(-12, min(12, e))
    beta = np.float32(np.ldexp(1.0, -e))
    y = cbn * beta
    y1 = _h(y); y2 = _h(y - y1)
    b = np.float32(-0.5) * beta * cn2
    b1 = _h(b); b2 = _h(b - b1); b3 = _h((b - b1) - b2)
    x1 = _h(zn); x2 = _h(zn - x1)
    acc = x1 @ y1.T                      # fp32 accumulation (the order differs from the tensor core's; so does its rounding)
    acc = acc + x2 @ y1.T
    acc = acc + x1 @ y2.T
    acc = acc + (b1 + b2 + b3)[None, :]
    R = np.float32(1.0001) * beta * cmax + np.float32(0.5) * beta * cmax2
    tol = np.float32(2.0 ** -16) * R + np.float32(2.0 ** -23) * np.float32(2 * np.sqrt(d) + 1)
    return acc.astype(np.float32), float(tol), float(beta), float(R)


def _reference_argmin(z, cb):
    zn, cn = F.normalize(torch.from_numpy(z), dim=1), F.normalize(torch.from_numpy(cb), dim=1)
    dist = (zn ** 2).sum(1, keepdim=True) + (cn ** 2).sum(1) - 2 * zn @ cn.t()           # model/quantizer.py:457-461
    return zn.numpy(), cn.numpy(), torch.argmin(dist, dim=1).numpy()


@pytest.mark.parametrize("d,K", [(16, 256), (32, 256), (64, 512), (16, 37)])
def test_split_scores_decide_like_the_reference(d, K):
    rng = np.random.default_rng(d * 1000 + K)
    n = 6000
    cb = rng.standard_normal((K, d)).astype(np.float32) * np.float32(rng.choice([1e-3, 1.0, 40.0]))
    z = rng.standard_normal((n, d)).astype(np.float32)
    # a third of the rows sit (almost) on the bisector of two codes: the near-ties the flag exists for
    a, b = rng.integers(0, K, n // 3), rng.integers(0, K, n // 3)
    cn_t = cb / np.linalg.norm(cb, axis=1, keepdims=True)
    z[: n // 3] = (cn_t[a] + cn_t[b]) * 0.5 + 1e-6 * rng.standard_normal((n // 3, d)).astype(np.float32)
    z[n // 3] = 0.0                                                       # a zero row: every code ties
    zn, cn, ref_idx = _reference_argmin(z, cb)
    score, tol, beta, R = _split_scores(zn, cn)
    true = beta * (zn.astype(np.float64) @ cn.astype(np.float64).T - 0.5 * (cn.astype(np.float64) ** 2).sum(1)[None, :])
    err = np.abs(score.astype(np.float64) - true).max()
    assert err <= 0.25 * tol, (err, tol)            # measured ~0.02-0.1 tol: the rule has a wide margin in practice
    assert np.abs(true).max() <= R
    order = np.argsort(-score, axis=1, kind="stable")
    best, second = order[:, 0], order[:, 1]
    gap = score[np.arange(n), best] - score[np.arange(n), second]
    certain = gap > tol
    assert certain.mean() > 0.6                                           # the exact path is the exception
    assert (~certain[: n // 3]).mean() > 0.5 and not certain[n // 3]      # the planted near-ties are caught
    # rows the kernel keeps without a re-score carry the reference's index; flagged rows take the exact path
    assert np.array_equal(best[certain], ref_idx[certain])
    # and wherever the approximate winner is NOT the reference's, the row was flagged
    assert not (certain & (best != ref_idx)).any()
