"""equss_b200 -- B200-native (sm_100a) implementation of the EQUSS product-quantization hot path.

Scope: PQ distance+argmin, gather+losses, EMA codebook update, cluster-probe
argmax + confusion histogram, global-feature kNN, and (first "next" row) the expansion head in front of the quantiser.
Host code is Python/PyTorch and mirrors the
reference's module names and signatures; all arithmetic runs in hand-written CUDA kernels behind the
C-ABI declared in ``include/equss_b200.h`` (``libequss_b200.so``).  There is no CPU fallback.
"""
from . import _native  # noqa: F401
from . import ops  # noqa: F401

__version__ = "0.1.0"


def _lazy(name):
    import importlib
    return importlib.import_module(f"{__name__}.{name}")


def __getattr__(name):
    if name in ("quantizer", "quantizer_v2", "codebooks", "evaluator", "metric", "knn", "head", "dist_utils", "build", "losses"):
        return _lazy(name)
    raise AttributeError(name)
