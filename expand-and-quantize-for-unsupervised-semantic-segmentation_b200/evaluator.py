"""Drop-in mirror of the reference's ``model/evaluator.py`` (UnSegEvaluator, ClusterLookup).

The reference bilinearly upsamples the (B, D, h, w) features to label resolution and runs both probes
there (model/evaluator.py:53-54,67-70,95-106).  Here both probes run at TOKEN resolution in one kernel
(K8 step 1) and a second kernel interpolates the 27+27 logits per label pixel, takes both argmaxes and
(optionally) feeds the confusion histograms (K8 step 2 + K9), see csrc/eval_probe.cu.

The two training losses are "next" rows: they are evaluated from the low-resolution
logits with differentiable PyTorch ops (interpolation of logits instead of features; the per-pixel
feature norm from 2x2 Gram terms), never materialising the upsampled feature map.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa

from . import ops

__all__ = ["UnSegEvaluator", "ClusterLookup"]


class ClusterLookup(nn.Module):
    """Cosine cluster probe, model/evaluator.py:85-111.  ``UnSegEvaluator`` runs the ``alpha=None`` case through the
    probe kernels at token resolution; this module-level forward is the plain-torch definition kept for the callers
    that use the class on its own (CRF branch, soft assignment with a temperature) -- same parameter, same returns."""

    def __init__(self, dim: int, n_classes: int):
        super().__init__()
        self.n_classes = n_classes
        self.dim = dim
        self.clusters = torch.nn.Parameter(torch.randn(n_classes, dim))

    def forward(self, x: torch.Tensor, alpha: Optional[float] = 2.0, log_probs: bool = False):
        """x: (b, dim, h, w).  Returns (loss, assignment (b, n_classes, h, w)): the assignment is the one-hot of the
        most similar centre when ``alpha`` is None, else softmax(alpha * cosine); the loss is minus the mean
        assignment-weighted cosine (:106); ``log_probs`` swaps the assignment for log_softmax(alpha * cosine) (:108-109)."""
        centres = F.normalize(self.clusters, dim=1)
        cosine = torch.einsum("bchw,nc->bnhw", F.normalize(x, dim=1), centres)
        if alpha is None:
            assignment = torch.zeros_like(cosine).scatter_(1, cosine.argmax(dim=1, keepdim=True), 1.0)
        else:
            assignment = torch.softmax(cosine * alpha, dim=1)
        loss = -(assignment * cosine).sum(dim=1).mean()
        if log_probs:
            return loss, torch.log_softmax(cosine * alpha, dim=1)
        return loss, assignment


def _taps(out_size: int, in_size: int, device):
    """Source indices / weights of torch's bilinear upsampling, align_corners=False."""
    scale = float(in_size) / float(out_size)
    dst = torch.arange(out_size, dtype=torch.float32, device=device)
    src = torch.clamp((dst + 0.5) * torch.tensor(scale, dtype=torch.float32, device=device) - 0.5, min=0.0)
    i0 = src.floor().long().clamp_(max=in_size - 1)
    i1 = torch.clamp(i0 + 1, max=in_size - 1)
    l1 = src - i0.float()
    return i0, i1, 1.0 - l1, l1


def _upsampled_feature_norm(x: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """|| bilinear_upsample(x)[b, :, Y, X] ||_2 for every label pixel, shape (B, H, W), from the 2x2 Gram
    terms of the token grid: |sum_t w_t x_t|^2 = sum_{t,t'} w_t w_t' <x_t, x_t'>.  Differentiable w.r.t. x (the
    reference's F.normalize of the upsampled features is, model/evaluator.py:96); the squared norm is floored at
    (1e-12)^2, F.normalize's own floor, so a zero feature vector has a finite (zero) gradient."""
    B, D, h, w = x.shape
    y0, y1, wy0, wy1 = _taps(H, h, x.device)
    x0, x1, wx0, wx1 = _taps(W, w, x.device)
    g_self = (x * x).sum(1)                                             # (B,h,w)
    g_h = F.pad((x[..., :, :-1] * x[..., :, 1:]).sum(1), (0, 1))        # <x[y,x], x[y,x+1]>
    g_v = F.pad((x[..., :-1, :] * x[..., 1:, :]).sum(1), (0, 0, 0, 1))  # <x[y,x], x[y+1,x]>
    g_d = F.pad((x[..., :-1, :-1] * x[..., 1:, 1:]).sum(1), (0, 1, 0, 1))   # <x[y,x], x[y+1,x+1]>
    g_a = F.pad((x[..., :-1, 1:] * x[..., 1:, :-1]).sum(1), (1, 0, 0, 1))   # <x[y,x], x[y+1,x-1]> stored at [y, x]
    Y0, X0 = y0[:, None], x0[None, :]
    Y1, X1 = y1[:, None], x1[None, :]
    sx, sy = (X1 != X0), (Y1 != Y0)                                     # taps distinct (not clamped at the border)?

    def at(g, yy, xx):
        return g[:, yy, xx]
    a, b_, c, e = wy0[:, None] * wx0[None, :], wy0[:, None] * wx1[None, :], wy1[:, None] * wx0[None, :], wy1[:, None] * wx1[None, :]
    s00, s01, s10, s11 = at(g_self, Y0, X0), at(g_self, Y0, X1), at(g_self, Y1, X0), at(g_self, Y1, X1)
    h0 = torch.where(sx, at(g_h, Y0, X0), s00)                          # <(y0,x0),(y0,x1)>
    h1 = torch.where(sx, at(g_h, Y1, X0), s10)                          # <(y1,x0),(y1,x1)>
    v0 = torch.where(sy, at(g_v, Y0, X0), s00)                          # <(y0,x0),(y1,x0)>
    v1 = torch.where(sy, at(g_v, Y0, X1), s01)                          # <(y0,x1),(y1,x1)>
    dd = torch.where(sx & sy, at(g_d, Y0, X0), torch.where(sx, h0, torch.where(sy, v0, s00)))   # <(y0,x0),(y1,x1)>
    aa = torch.where(sx & sy, at(g_a, Y0, X1), torch.where(sx, h0, torch.where(sy, v1, s01)))   # <(y0,x1),(y1,x0)>
    n2 = (a * a * s00 + b_ * b_ * s01 + c * c * s10 + e * e * s11 +
          2 * (a * b_ * h0 + c * e * h1 + a * c * v0 + b_ * e * v1 + a * e * dd + b_ * c * aa))
    return n2.clamp_min(1e-24).sqrt()


class _ProbeLosses(torch.autograd.Function):
    """Both evaluator losses with gradients for the probe parameters (through ``wmat`` / ``bias``, which the caller
    assembles differentiably from the cluster centres and the linear probe)."""

    @staticmethod
    def forward(ctx, out, wmat, bias, label, Cc, Cp, C):
        B, D, h, w = out.shape
        out32 = out.detach().float().contiguous()
        logits = ops.probe_logits(out32, wmat.detach(), bias.detach())
        gram = ops.token_gram(out32)
        need = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        sums, n_valid, g = ops.probe_losses(logits, gram, B, h, w, Cp + C, label, C, (0, Cc), (Cp, C), want_grad=need)
        P = label.numel()
        nv = n_valid.to(torch.float64)
        linear_loss = (sums[0] / nv[0]).float()              # 0/0 = nan when no pixel is labelled, like F.cross_entropy
        cluster_loss = (-sums[1] / P).float()
        if need:
            ctx.save_for_backward(out32, g, nv)
        ctx.meta = (Cc, Cp, C, P, wmat.shape[0])
        return linear_loss, cluster_loss

    @staticmethod
    def backward(ctx, g_lin, g_clu):
        out32, g, nv = ctx.saved_tensors
        Cc, Cp, C, P, rows = ctx.meta
        B, D, h, w = out32.shape
        scale = torch.zeros(g.shape[1], device=g.device)
        scale[:Cc] = -g_clu.float() / P
        scale[Cp:Cp + C] = g_lin.float() / nv[0].float()
        gs = g * scale                                                           # [N, C_pad]
        grad_w = torch.einsum("bnc,bdn->cd", gs.view(B, h * w, -1), out32.view(B, D, h * w))[:rows]
        grad_b = gs.sum(dim=0)[:rows]
        return None, grad_w, grad_b, None, None, None, None


class UnSegEvaluator(nn.Module):
    """model/evaluator.py:11-82."""

    def __init__(self, embed_dim: int, num_classes: int, extra_classes: int = 0, num_pq: int = 1) -> None:
        super().__init__()
        self.num_classes = num_classes
        self.linear_probe = nn.Conv2d(embed_dim, num_classes, kernel_size=1, stride=1)
        self.cluster_probe = ClusterLookup(embed_dim, num_classes + extra_classes)
        self.linear_loss = nn.CrossEntropyLoss()
        self.compute_losses = True   # set False to skip the two (training-only) loss scalars
        self._probe_cache = None

    def _packed_probe_weights(self, dev):
        """Both probes as ONE operand for K8: rows [0, Cc) the L2-normalised cluster centres, rows [Cp, Cp + C) the linear
        probe, each head starting at a multiple of four channels (the second kernel reads four logits per 16-byte load;
        padding rows are zero and never compete in the argmax), packed into the kernels' K-major / tensor-core images.
        In eval() mode the packed operand is cached and rebuilt when a parameter's storage or version counter changes;
        in train() mode (the probes are being optimised) it is rebuilt on every call.  In-place updates through ``.data``
        do not bump version counters: call :meth:`invalidate_cache` after such an update."""
        ps = (self.cluster_probe.clusters, self.linear_probe.weight, self.linear_probe.bias)
        key = tuple((p.data_ptr(), p._version) for p in ps) + (str(dev),)
        if not self.training and self._probe_cache is not None and self._probe_cache[0] == key:
            return self._probe_cache[1], self._probe_cache[2]
        Cc, C, D = self.cluster_probe.n_classes, self.num_classes, self.cluster_probe.dim
        Cp = (Cc + 3) // 4 * 4
        wmat = torch.zeros(Cp + C, D, device=dev)
        wmat[:Cc] = F.normalize(ps[0].detach().float(), dim=1)
        wmat[Cp:] = ps[1].detach().float().view(C, D)
        bias = torch.zeros(Cp + C, device=dev)
        bias[Cp:] = ps[2].detach().float()
        pack = ops.probe_pack(wmat)
        self._probe_cache = (key, pack, bias)
        return pack, bias

    def invalidate_cache(self) -> None:
        self._probe_cache = None

    # -- K8: predictions (and optionally K9 confusion matrices) -----------------------------------
    @torch.no_grad()
    def predict(self, out: torch.Tensor, label: torch.Tensor, cluster_confusion: Optional[torch.Tensor] = None,
                linear_confusion: Optional[torch.Tensor] = None, want_preds: bool = True):
        """Fused probe: returns (linear_preds, cluster_preds) at label resolution (None if not wanted) and
        accumulates the int64 confusion buffers in place when given (rows = prediction, cols = label)."""
        B, D, h, w = out.shape
        Cc = self.cluster_probe.n_classes
        C = self.num_classes
        Cp = (Cc + 3) // 4 * 4
        wmat, bias = self._packed_probe_weights(out.device)
        logits = ops.probe_logits(out, wmat, bias)
        confs = None
        if cluster_confusion is not None or linear_confusion is not None:
            confs = [cluster_confusion, linear_confusion]
        preds = ops.probe_argmax_confusion(logits, B, h, w, Cp + C, label, C, [(0, Cc), (Cp, C)],
                                           want_preds=want_preds, confusions=confs)
        return preds[1], preds[0]

    def _losses(self, out: torch.Tensor, label: torch.Tensor, cluster_preds: torch.Tensor):
        """(linear_loss, cluster_loss) of model/evaluator.py:65-80,106.  Kernel path (K8b): token logits -> one pass over
        the label pixels (interpolation + masked cross-entropy + cosine of the winning cluster, and the transposed
        interpolation of the logit gradients); the probe-parameter gradients are one [C_pad x N] x [N x D] contraction.
        Features that require a gradient themselves (no reference caller: the wrappers pass ``out.detach()``) take the
        differentiable PyTorch formulation below."""
        B, D, h, w = out.shape
        H, W = label.shape[-2:]
        C, Cc = self.num_classes, self.cluster_probe.n_classes
        Cp = (Cc + 3) // 4 * 4
        if out.requires_grad or not ops.N.lib().equss_probe_losses_supported(h, w, H, W, Cp + C, Cc, C, 0, Cp):
            return self._losses_torch(out, label, cluster_preds)
        dev = out.device
        wmat = torch.cat([F.normalize(self.cluster_probe.clusters.float(), dim=1), torch.zeros(Cp - Cc, D, device=dev),
                          self.linear_probe.weight.float().view(C, D)], dim=0)
        bias = torch.cat([torch.zeros(Cp, device=dev), self.linear_probe.bias.float()])
        return _ProbeLosses.apply(out, wmat, bias, label, Cc, Cp, C)

    def _losses_torch(self, out: torch.Tensor, label: torch.Tensor, cluster_preds: torch.Tensor):
        B, D, h, w = out.shape
        H, W = label.shape[-2:]
        C = self.num_classes
        # fp32 contraction (cuDNN convolutions default to TF32, which is not within the 1e-5 parity bar)
        lin_low = torch.einsum("bchw,nc->bnhw", out, self.linear_probe.weight.view(C, D)) + \
            self.linear_probe.bias.view(1, C, 1, 1)                                         # :67 at token resolution
        lin_up = F.interpolate(lin_low, (H, W), mode="bilinear", align_corners=False) if (h, w) != (H, W) else lin_low
        label_flat = label.reshape(-1)
        mask = torch.logical_and(label_flat >= 0, label_flat < self.num_classes)           # :73
        logit_flat = lin_up.permute(0, 2, 3, 1).reshape(-1, self.num_classes)
        linear_loss = self.linear_loss(logit_flat[mask], label_flat[mask]).mean()          # :80
        nc = F.normalize(self.cluster_probe.clusters, dim=1)
        inner_low = torch.einsum("bchw,nc->bnhw", out, nc)
        inner_up = F.interpolate(inner_low, (H, W), mode="bilinear", align_corners=False) if (h, w) != (H, W) else inner_low
        norm = _upsampled_feature_norm(out, H, W).clamp_min(1e-12)
        picked = inner_up.gather(1, cluster_preds.unsqueeze(1)).squeeze(1) / norm
        cluster_loss = -picked.mean()                                                      # :106
        return linear_loss, cluster_loss

    def forward(self, out: torch.Tensor, img: torch.Tensor, label: Optional[torch.Tensor] = None,
                is_crf: bool = False) -> Tuple[torch.Tensor, ...]:
        if is_crf:
            # final-eval CRF branch (CPU pydensecrf, SURVEY out of scope): reference formulation
            from utils.crf_utils import batched_crf  # provided by the reference checkout
            if out.shape[-2:] != label.shape[-2:]:
                out = F.interpolate(out, label.shape[-2:], mode="bilinear", align_corners=False)
            linear_log_prob = torch.log_softmax(self.linear_probe(out), dim=1)
            cluster_loss, cluster_log_prob = self.cluster_probe(out, 2, log_probs=True)
            linear_preds = batched_crf(img, linear_log_prob).argmax(1)
            cluster_preds = batched_crf(img, cluster_log_prob).argmax(1)
            return torch.zeros_like(cluster_loss), linear_preds, cluster_loss, cluster_preds
        assert label is not None
        out32 = out.float()
        linear_preds, cluster_preds = self.predict(out32, label)
        if self.compute_losses:
            linear_loss, cluster_loss = self._losses(out32, label, cluster_preds)
        else:
            linear_loss = cluster_loss = torch.zeros((), device=out.device)
        return linear_loss, linear_preds, cluster_loss, cluster_preds
