"""Shared host logic of the PQ head: the fused multi-subspace forward built from the C-ABI kernels, its
autograd wrappers, and the small batched statistics the reference computes per subspace in Python.

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

__all__ = ["normalize_codebook", "pq_quantize", "percentile_stats", "flat_pixels", "PQGatherLoss", "DistanceProb",
           "soft_assignment_stats"]


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


def _rows(z: torch.Tensor, M: int) -> torch.Tensor:
    """(N, M, d) view of a flat (N, D) or NCHW (B, D, h, w) activation (host-side autograd plumbing only)."""
    if z.dim() == 2:
        return z.reshape(z.shape[0], M, -1)
    B, D, h, w = z.shape
    return z.permute(0, 2, 3, 1).reshape(B * h * w, M, D // M)


def _normalize_rows(zr: torch.Tensor, mode: Optional[str], a: Optional[torch.Tensor], b: Optional[torch.Tensor]) -> torch.Tensor:
    """Differentiable z-side normalisation of (N, M, d) rows -- used only to route gradients in backward passes."""
    if mode == "l2":
        return F.normalize(zr, dim=-1)
    if mode == "z_norm":
        s, m = torch.std_mean(zr, dim=-1, keepdim=True)
        return (zr - m) / (s + 1e-5)
    if mode == "z_trainable":
        M, d = zr.shape[1], zr.shape[2]
        return (zr - a.reshape(M, d)) / b.reshape(M, d)
    return zr


def _affine_param_grads(gz: torch.Tensor, z: torch.Tensor, M: int, a: torch.Tensor, b: torch.Tensor, need_a: bool, need_b: bool):
    """Gradients of z_norm = (z - a) / b w.r.t. the per-channel vectors, from grad_z = g_znorm / b:
    g_a = -sum_n grad_z,  g_b = -sum_n grad_z * (z - a) / b."""
    gr, zr = _rows(gz, M), _rows(z, M)
    ga = (-gr.sum(dim=0)).reshape(-1) if need_a else None
    gb = (-(gr * (zr - a.reshape(1, M, -1))).sum(dim=0) / b.reshape(M, -1)).reshape(-1) if need_b else None
    return ga, gb


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
        need_a, need_b = ctx.needs_input_grad[4], ctx.needs_input_grad[5]
        M = src.shape[0]
        coef = (g_commit.float() * ctx.scale) if g_commit is not None else torch.zeros(M, device=z.device)
        cb_coef = (g_cb.float() * ctx.scale) if (need_src and g_cb is not None) else None
        want_gz = need_z or need_a or need_b
        gz, gsrc = ops.pq_gather_loss_bwd(z, src, idx, ctx.normalize, g_out if want_gz else None, coef, na, nb,
                                          want_grad_z=want_gz, cb_coef=cb_coef)
        ga = gb = None
        if (need_a or need_b) and na is not None:
            ga, gb = _affine_param_grads(gz, z, M, na, nb, need_a, need_b)
        if gz is not None and gz.dtype != z.dtype:
            gz = gz.to(z.dtype)
        return (gz if need_z else None), gsrc, None, None, ga, gb, None, None


class DistanceProb(torch.autograd.Function):
    """K2 with gradients: ``softmax(-distance / T)`` [N, M*K] (model/quantizer.py:468; dino_new_vq.py:398).  The
    reference's soft assignment is differentiable w.r.t. the activations and (learned codebooks) the codebook; its
    trainers back-propagate the JSD / entropy / contrastive terms through it (model/dino_new_vq.py:447-450,
    dino_contra.py:253-257, dino_vae.py:220-224).  Forward is the kernel; backward evaluates the softmax and
    distance Jacobians as batched contractions over the saved probabilities and routes the result through the
    row normalisation with autograd."""

    @staticmethod
    def forward(ctx, z, codebook_norm, cnorm2, normalize, norm_a, norm_b, temperature):
        prob = ops.pq_distance_prob(z, codebook_norm, cnorm2, normalize, norm_a, norm_b, temperature)
        empty = z.new_empty(0)
        ctx.save_for_backward(z, codebook_norm, norm_a if norm_a is not None else empty,
                              norm_b if norm_b is not None else empty, prob)
        ctx.normalize, ctx.temperature = normalize, float(temperature)
        return prob

    @staticmethod
    def backward(ctx, g):
        z, cbn, na, nb, prob = ctx.saved_tensors
        need_z, need_c, need_a, need_b = (ctx.needs_input_grad[i] for i in (0, 1, 4, 5))
        M, K, d = cbn.shape
        N = prob.shape[0]
        p = prob.view(N, M, K)
        g = g.reshape(N, M, K).float()
        gd = p * ((p * g).sum(dim=-1, keepdim=True) - g) / ctx.temperature        # dL/d distance
        c = cbn.detach().float()
        with torch.enable_grad():
            zl = z.detach().requires_grad_(need_z)
            a = na.detach().requires_grad_(need_a) if na.numel() else None
            b = nb.detach().requires_grad_(need_b) if nb.numel() else None
            zn = _normalize_rows(_rows(zl.float(), M), ctx.normalize, a, b)
        znd = zn.detach()
        gz = ga = gb = gc = None
        wanted = [t for t, need in ((zl, need_z), (a, need_a), (b, need_b)) if need and t is not None]
        if wanted:
            g_zn = 2.0 * (znd * gd.sum(dim=-1, keepdim=True) - torch.einsum("nmk,mkd->nmd", gd, c))
            if zn.requires_grad:
                got = list(torch.autograd.grad(zn, wanted, g_zn))
                if need_z:
                    gz = got.pop(0).to(z.dtype)
                if need_a and a is not None:
                    ga = got.pop(0)
                if need_b and b is not None:
                    gb = got.pop(0)
        if need_c:
            gc = 2.0 * (c * gd.sum(dim=0).unsqueeze(-1) - torch.einsum("nmk,nmd->mkd", gd, znd))
        return gz, gc, None, None, ga, gb, None


def _wants_grad(*ts) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


def pq_quantize(z: torch.Tensor, codebook_norm: torch.Tensor, gather_src: torch.Tensor, normalize: Optional[str],
                norm_a=None, norm_b=None, *, want_prob: bool = True, temperature: float = 1.0, algo: int = 0,
                cnorm2: Optional[torch.Tensor] = None, idx: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """assign -> (soft assignment) -> gather/loss for all M subspaces at once.
    ``idx``: externally chosen indices (the Gumbel draw of the use_gumbel flag) -- the assignment kernel is skipped.
    Returns (idx int32 [M,N], out like z, mse_commit [M], mse_codebook [M], prob [N, M*K] or None)."""
    z32 = z if z.dtype == torch.float32 else z.float()
    cbn = codebook_norm.detach()
    cn2 = cnorm2 if cnorm2 is not None else ops.pq_cnorm2(cbn)
    pre_out = pre_sq = None
    if idx is not None:
        idx = idx.to(torch.int32).contiguous()
    elif algo == 0:
        # one fused kernel (K1 + K3, activations read once) where the shape allows it, else K1 then K3
        idx, pre_out, pre_sq = ops.pq_assign_gather(z32, cbn, gather_src.detach(), cn2, normalize, norm_a, norm_b)
    else:
        idx = ops.pq_assign(z32, cbn, cn2, normalize, norm_a, norm_b, algo=algo)
    prob = None
    if want_prob:
        if _wants_grad(z32, codebook_norm, norm_a, norm_b):
            prob = DistanceProb.apply(z32, codebook_norm, cn2, normalize, norm_a, norm_b, temperature)
        else:
            prob = ops.pq_distance_prob(z32, cbn, cn2, normalize, norm_a, norm_b, temperature)
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


def soft_assignment_stats(prob: torch.Tensor, M: int, K: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """The two ``distance_prob`` consumers of model/dino_new_vq.py:447-450, per subspace then averaged over the M
    subspaces (the wrapper's mean, :724-725), evaluated on a materialised, differentiable soft assignment:

      jsd     = JSDLoss(p1, p2)      model/loss.py:508-525   (p1 | p2 = the two halves of the batch)
      entropy = EntropyLoss(p1, p2)  model/loss.py:490-505   (minus the entropy of mean_n p1)

    Used when a gradient is required; the no-grad path runs the fused kernel (ops.pq_soft_stats) and never
    materialises the N x K*M tensor."""
    n = prob.shape[0]
    pv = prob.view(n, M, K)
    if n % 2 != 0:
        raise ValueError("JSD needs an even number of rows: the batch holds two views (model/dino_new_vq.py:447)")
    p1, p2 = pv[: n // 2], pv[n // 2:]
    e = 1e-6
    lm = ((p1 + p2 + e) * 0.5).log()
    t1, t2 = (p1 + e).log(), (p2 + e).log()
    # KLDivLoss(batchmean, log_target=True)(m, t) = sum exp(t) * (t - m) / rows
    kl = ((p1 + e) * (t1 - lm)).sum(dim=(0, 2)) + ((p2 + e) * (t2 - lm)).sum(dim=(0, 2))
    jsd = (0.5 * kl / (n // 2)).mean()
    avg = p1.mean(dim=0)                                          # [M, K]
    ent = (avg * torch.log(avg + 1e-8)).sum(dim=-1).mean()        # = -(entropy)
    return jsd, ent


def flat_pixels(z: torch.Tensor) -> int:
    return z.shape[0] if z.dim() == 2 else z.shape[0] * z.shape[2] * z.shape[3]
