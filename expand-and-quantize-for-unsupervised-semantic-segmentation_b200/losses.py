"""STEGO correspondence loss next to the PQ head in the training step: mirror of ``STEGOLoss`` and its
helpers in the reference's ``model/loss.py:647-739`` (same constructor, ``helper`` / ``forward`` signatures and cfg keys).

The feature half of ``helper`` -- cosine correlation of the sampled (frozen) backbone features, row centring and
re-centring, six eager kernels and three full-tensor reductions in the reference -- is one kernel
(``equss_stego_feature_corr``).  The code half stays differentiable PyTorch on the small sampled code maps: it carries
the gradient to the expansion head, and its contraction is over the expanded code dimension only.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

__all__ = ["STEGOLoss", "tensor_correlation", "norm", "sample", "super_perm"]


def tensor_correlation(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """(n, c, h, w) x (n, c, i, j) -> (n, h, w, i, j)   (model/loss.py:647-648)."""
    return torch.einsum("nchw,ncij->nhwij", a, b)


def norm(t: torch.Tensor) -> torch.Tensor:
    return F.normalize(t, dim=1, eps=1e-10)                                    # :651-652


def sample(t: torch.Tensor, coords: torch.Tensor) -> torch.Tensor:
    """Bilinear samples of ``t`` at ``coords`` in [-1, 1] (n, S, S, 2)   (:655-656)."""
    return F.grid_sample(t, coords.permute(0, 2, 1, 3), padding_mode="border", align_corners=True)


def super_perm(size: int, device: torch.device) -> torch.Tensor:
    """A permutation of range(size) without fixed points, drawn like the reference (:659-663)."""
    perm = torch.randperm(size, device=device, dtype=torch.long)
    perm[perm == torch.arange(size, device=device)] += 1
    return perm % size


class STEGOLoss(nn.Module):
    """model/loss.py:666-739.  cfg keys: pointwise, zero_clamp, stabilize, feature_samples, neg_samples,
    {pos_intra,pos_inter,neg_inter}_{shift,weight}."""

    def __init__(self, cfg: dict):
        super().__init__()
        self.cfg = cfg

    def standard_scale(self, t):
        centred = t - t.mean()
        return centred / centred.std()

    def helper(self, f1, f2, c1, c2, shift):
        fd = ops.stego_feature_corr(f1, f2, pointwise=bool(self.cfg["pointwise"]))              # :679-687, one kernel
        cd = tensor_correlation(norm(c1), norm(c2))
        floor = 0.0 if self.cfg["zero_clamp"] else -9999.0
        bounded = cd.clamp(floor, 0.8) if self.cfg["stabilize"] else cd.clamp(floor)
        return -bounded * (fd - shift), cd

    def forward(self, orig_feats: torch.Tensor, orig_feats_pos: torch.Tensor, orig_code: torch.Tensor,
                orig_code_pos: torch.Tensor):
        cfg = self.cfg
        S = cfg["feature_samples"]
        shape = [orig_feats.shape[0], S, S, 2]
        coords1 = torch.rand(shape, device=orig_feats.device) * 2 - 1
        coords2 = torch.rand(shape, device=orig_feats.device) * 2 - 1
        feats, code = sample(orig_feats, coords1), sample(orig_code, coords1)
        feats_pos, code_pos = sample(orig_feats_pos, coords2), sample(orig_code_pos, coords2)
        intra, _ = self.helper(feats, feats, code, code, cfg["pos_intra_shift"])
        inter, _ = self.helper(feats, feats_pos, code, code_pos, cfg["pos_inter_shift"])
        negatives = []
        for _ in range(cfg["neg_samples"]):
            perm = super_perm(orig_feats.shape[0], orig_feats.device)
            neg, _ = self.helper(feats, sample(orig_feats[perm], coords2), code, sample(orig_code[perm], coords2),
                                 cfg["neg_inter_shift"])
            negatives.append(neg)
        return (cfg["pos_intra_weight"] * intra.mean() + cfg["pos_inter_weight"] * inter.mean() +
                cfg["neg_inter_weight"] * torch.cat(negatives, dim=0).mean())
