"""Shared host logic of the PQ head: the fused multi-subspace forward built from the C-ABI kernels, its
autograd wrapper, and the small batched statistics the reference computes per subspace in Python.

All M subspaces are processed by ONE kernel per stage (assign -> gather/loss -> accumulate -> all-reduce
-> EMA update) instead of the reference's M sequential Python iterations of ~40 launches each
(model/quantizer.py:595-604).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from . import ops
from .dist_utils import all_reduce_packed_

__all__ = ["normalize_codebook", "pq_quantize", "percentile_stats", "flat_pixels", "PQGatherLoss"]


def normalize_codebook(codebook: torch.Tensor, mode: Optional[str], *, ema_style: bool = True,
                       z_mean: Optional[torch.Tensor] = None, z_std: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Codebook-side normalisation on a stacked [M, K, d] codebook (model/quantizer.py:419-455).  This is
    M*K*d elements of differentiable host plumbing; the N*D activation side runs inside the kernels."""
    if mode == "l2":
        return F.normalize(codebook, dim=2)                                   # :421
    if mode == "z_norm":
        s, m = torch.std_mean(codebook, dim=2, keepdim=True)                  # :426
        return (codebook - m) / (s + 1e-5)
    if mode == "z_trainable":
        if ema_style:
            s, m = torch.std_mean(codebook, dim=1, keepdim=True)              # :449 (over the K codes)
            return (codebook - m) / (s + 1e-5)
        return (codebook - z_mean.unsqueeze(1)) / (z_std.unsqueeze(1) + 1e-5)  # :133
    if mode == "none" or mode is None:
        if mode is None:
            raise ValueError(f"Unsupported normalize type {mode}")            # :455
        return codebook
    raise ValueError(f"Unsupported normalize type {mode}")


class PQGatherLoss(torch.autograd.Function):
    """K3 with gradients.  Returns (out, mse_commit[M], mse_codebook[M]); the two MSE vectors have equal
    values but route their gradients like the reference's two mse_loss calls (model/quantizer.py:175-176):
    commitment -> activations (through the normalisation), codebook -> gathered rows."""

    @staticmethod
    def forward(ctx, z, gather_src, idx, normalize, norm_a, norm_b, pre_out=None, pre_sqerr=None):
        # pre_out / pre_sqerr: K3 results already produced by the fused assign+gather kernel
        if pre_out is not None:
            out, sqerr = pre_out, pre_sqerr
        else:
            out, sqerr, _ = ops.pq_gather_loss(z, gather_src, idx, normalize, norm_a, norm_b)
        M, K, d = gather_src.shape
        n = idx.shape[1]
        mse = (sqerr / max(n * d, 1)).to(torch.float32)
        ctx.save_for_backward(z, gather_src, idx, norm_a if norm_a is not None else z.new_empty(0),
                              norm_b if norm_b is not None else z.new_empty(0))
        ctx.normalize = normalize
        ctx.scale = 2.0 / max(n * d, 1)
        ctx.mark_non_differentiable(idx)
        return out, mse, mse.clone()

    @staticmethod
    def backward(ctx, g_out, g_commit, g_cb):
        z, src, idx, na, nb = ctx.saved_tensors
        na = na if na.numel() else None
        nb = nb if nb.numel() else None
        need_z, need_src = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        M = src.shape[0]
        coef = (g_commit.float() * ctx.scale) if g_commit is not None else torch.zeros(M, device=z.device)
        cb_coef = (g_cb.float() * ctx.scale) if (need_src and g_cb is not None) else None
        gz, gsrc = ops.pq_gather_loss_bwd(z, src, idx, ctx.normalize, g_out if need_z else None, coef, na, nb,
                                          want_grad_z=need_z, cb_coef=cb_coef)
        if gz is not None and gz.dtype != z.dtype:
            gz = gz.to(z.dtype)
        return gz, gsrc, None, None, None, None, None, None


def pq_quantize(z: torch.Tensor, codebook_norm: torch.Tensor, gather_src: torch.Tensor, normalize: Optional[str],
                norm_a=None, norm_b=None, *, want_prob: bool = True, temperature: float = 1.0, algo: int = 0
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """assign -> (soft assignment) -> gather/loss for all M subspaces at once.
    Returns (idx int32 [M,N], out like z, mse_commit [M], mse_codebook [M], prob [N, M*K] or None)."""
    z32 = z if z.dtype == torch.float32 else z.float()
    cbn = codebook_norm.detach()
    cn2 = ops.pq_cnorm2(cbn)
    pre_out = pre_sq = None
    if algo == 0:
        # one fused kernel (K1 + K3, activations read once) where the shape allows it, else K1 then K3
        idx, pre_out, pre_sq = ops.pq_assign_gather(z32, cbn, gather_src.detach(), cn2, normalize, norm_a, norm_b)
    else:
        idx = ops.pq_assign(z32, cbn, cn2, normalize, norm_a, norm_b, algo=algo)
    prob = ops.pq_distance_prob(z32, cbn, cn2, normalize, norm_a, norm_b, temperature) if want_prob else None
    out, mse_commit, mse_cb = PQGatherLoss.apply(z32, gather_src, idx, normalize, norm_a, norm_b, pre_out, pre_sq)
    return idx, out, mse_commit, mse_cb, prob


@torch.no_grad()
def ema_statistics(z: torch.Tensor, idx: torch.Tensor, K: int, *, use_norm: bool = False,
                   normalize: Optional[str] = None, norm_a=None, norm_b=None) -> torch.Tensor:
    """K4 + K5: packed per-code sums/counts [M, K, d+1], all-reduced over the data-parallel group."""
    packed = ops.pq_accumulate(z.detach().float(), idx, K, use_norm=use_norm, normalize=normalize,
                               norm_a=norm_a, norm_b=norm_b)
    return all_reduce_packed_(packed)


@torch.no_grad()
def percentile_stats(count: torch.Tensor, prefix: str) -> Dict[str, torch.Tensor]:
    """Batched get_histogram_count (model/quantizer.py:15-30) for counts [M, K]: per subspace the first
    rank whose cumulative sorted usage reaches 10/50/90 %, divided by K; returned as the MEAN over
    subspaces (what ProductQuantizerWrapper.forward reports, :607-608) in 0-dim tensors -- one kernel launch
    and no host synchronisation (the reference performs ~6K tensor->bool syncs per subspace here)."""
    mean = ops.usage_percentiles(count).mean(dim=0)
    return {f"{prefix}-p10": mean[0], f"{prefix}-p50": mean[1], f"{prefix}-p90": mean[2]}


def flat_pixels(z: torch.Tensor) -> int:
    return z.shape[0] if z.dim() == 2 else z.shape[0] * z.shape[2] * z.shape[3]
