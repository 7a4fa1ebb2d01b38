"""Host check of the kNN screening margin (csrc/knn_h.cu): the fp16 screening product may only DROP a database
row whose approximate similarity is more than ``eps`` below the approximate k-th best, with

    eps = max|q| * max|d| * (2 * 2^-10 + 1e-5)          (norms of the power-of-two scaled rows)

The kernel's claim is |S~ - S| <= 2^-10 * max|q| * max|d| for operands rounded to fp16 after the power-of-two scaling
(relative error 2^-11 per normal element, 2^-25 absolute per subnormal one), hence every member of the exact top-k survives
the cut and the fp32 rescore decides among the survivors.  numpy's float16 cast rounds to nearest even like
``__floats2half2_rn``; the accumulation is emulated in fp32.  No device code runs here."""
import numpy as np
import pytest


def _pow2_scale(absmax: float) -> float:
    """knn_h.cu pow2_scale: power of two s with absmax * s in [1/2, 1)."""
    if not absmax > 0:
        return 1.0
    _, e = np.frexp(np.float32(absmax))
    return float(np.ldexp(1.0, -int(e)))


def _cases():
    rng = np.random.default_rng(0)
    F, n = 768, 1200
    x = rng.standard_normal((n, F))
    unit = x / np.linalg.norm(x, axis=1, keepdims=True)
    heavy = rng.standard_cauchy((n, F))
    heavy /= np.linalg.norm(heavy, axis=1, keepdims=True)
    sub = unit.copy()
    sub[:, :700] *= 1e-5                    # most elements become fp16 subnormals after the scaling
    sub[0, 0] = 1.0
    wild = rng.standard_normal((n, F)) * np.exp(3 * rng.standard_normal((n, 1)))     # any scale (not unit norm)
    dup = unit.copy()
    dup[1::2] = dup[::2] + 1e-4 * rng.standard_normal((n // 2, F))                  # near-duplicate rows: tight top-k
    return {"unit": unit, "heavy_tailed": heavy, "subnormal": sub, "any_scale": wild, "near_duplicates": dup}


@pytest.mark.parametrize("name", ["unit", "heavy_tailed", "subnormal", "any_scale", "near_duplicates"])
def test_fp16_screen_keeps_the_exact_top_k(name):
    x = _cases()[name].astype(np.float32)
    s = np.float32(_pow2_scale(float(np.abs(x).max())))
    xs = x * s                                                   # exact: a power of two
    assert 0.5 <= float(np.abs(xs).max()) <= 1.0
    xh = xs.astype(np.float16)
    exact = xs.astype(np.float64) @ xs.astype(np.float64).T      # what the fp32 rescore resolves (scaled domain)
    approx = xh.astype(np.float32) @ xh.astype(np.float32).T     # fp16 operands, fp32 accumulate
    nrm = float(np.linalg.norm(xs.astype(np.float64), axis=1).max()) ** 2
    err = float(np.abs(approx.astype(np.float64) - exact).max())
    assert err <= 2.0 ** -10 * nrm, (err, nrm)                   # the bound of knn_h.cu (operand rounding + accumulation)
    eps = nrm * (2 * 2.0 ** -10 + 1e-5)
    for k in (8, 30):
        kth_approx = np.partition(approx, -k, axis=1)[:, -k]
        survivors = approx >= (kth_approx - np.float32(eps))[:, None]
        order = np.argsort(-exact, axis=1, kind="stable")[:, :k]          # larger first, lower index on ties
        kept = np.take_along_axis(survivors, order, axis=1)
        assert kept.all(), f"{name}: {int((~kept).sum())} true top-{k} members cut by the screen"
        # and the screen is selective: the rescore looks at a small fraction of the database
        if name in ("unit", "heavy_tailed"):
            assert survivors.sum(axis=1).mean() < 0.2 * x.shape[0]
