"""Expansion head in front of the product quantiser.

Mirror of ``SegmentationHead`` (model/blocks/module.py:20-44) and of the inline ``cluster1`` / ``cluster2`` pair of
the PQGO / NewVQ models (model/dino_pqgo.py:104-112,127-128; model/dino_new_vq.py same lines):

    code = cluster1(x) + cluster2(x)
    cluster1 = Conv2d(C, D, 1)                       cluster2 = Conv2d(C, C, 1) -> ReLU -> Conv2d(C, D, 1)

The parameters are the reference's own ``nn.Conv2d`` modules under the same attribute names, so checkpoints load
unchanged.  Without autograd (evaluation, feature extraction) the forward runs as two launches of the tcgen05
split-tf32 GEMM ``equss_head_gemm``:

    h    = relu(W2 x + b2)                 x read in place from NCHW
    code = [W1 | W3] [x ; h] + (b1 + b3)   both branches as ONE contraction, written once

and returns ``code`` as a (B, D, h, w) tensor whose memory is NHWC -- the flat (pixel, channel) rows the PQ kernels
consume without the reference's permute + contiguous copy (model/dino_pqgo.py:583-584).  When a gradient is
required the same modules run through PyTorch's convolutions (the caller's framework owns training of the head; the
fused backward is not part of this row).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import ops

__all__ = ["SegmentationHead", "expansion_head", "make_clusterer", "make_nonlinear_clusterer"]


def make_clusterer(in_channels: int, hidden_dim: int) -> nn.Sequential:
    """model/dino_pqgo.py:104-106, model/blocks/module.py:27-29."""
    return nn.Sequential(nn.Conv2d(in_channels, hidden_dim, (1, 1)))


def make_nonlinear_clusterer(in_channels: int, hidden_dim: int) -> nn.Sequential:
    """model/dino_pqgo.py:108-112, model/blocks/module.py:31-35."""
    return nn.Sequential(nn.Conv2d(in_channels, in_channels, (1, 1)), nn.ReLU(),
                         nn.Conv2d(in_channels, hidden_dim, (1, 1)))


class _PackedHead:
    """[W1 | W3] and b1 + b3, rebuilt when a parameter changes (version counters) or moves."""

    def __init__(self):
        self.key = None
        self.w13 = self.b13 = None

    def get(self, c1: nn.Conv2d, c3: nn.Conv2d, reuse: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """``reuse=False`` (modules in train() mode: EMA-teacher style ``p.data`` updates do not bump version counters)
        rebuilds unconditionally -- a D x 2C concatenation, cheap next to the GEMM."""
        ps = [c1.weight, c1.bias, c3.weight, c3.bias]
        key = tuple((p.data_ptr(), p._version, p.device) if p is not None else None for p in ps)
        if key != self.key or not reuse:
            with torch.no_grad():
                D = c1.weight.shape[0]
                self.w13 = torch.cat([c1.weight.reshape(D, -1), c3.weight.reshape(D, -1)], dim=1).float().contiguous()
                b = None
                if c1.bias is not None or c3.bias is not None:
                    b = torch.zeros(D, dtype=torch.float32, device=c1.weight.device)
                    if c1.bias is not None:
                        b = b + c1.bias.float()
                    if c3.bias is not None:
                        b = b + c3.bias.float()
                self.b13 = b
            self.key = key
        return self.w13, self.b13


def _check_pair(cluster1: nn.Sequential, cluster2: nn.Sequential) -> Tuple[nn.Conv2d, nn.Conv2d, nn.Conv2d]:
    ok = (len(cluster1) == 1 and len(cluster2) == 3 and isinstance(cluster1[0], nn.Conv2d)
          and isinstance(cluster2[0], nn.Conv2d) and isinstance(cluster2[1], nn.ReLU) and isinstance(cluster2[2], nn.Conv2d))
    if not ok:
        raise ValueError("expansion_head expects cluster1 = [Conv2d] and cluster2 = [Conv2d, ReLU, Conv2d] "
                         "(model/dino_pqgo.py:104-112)")
    c1, c2, c3 = cluster1[0], cluster2[0], cluster2[2]
    for c in (c1, c2, c3):
        if c.kernel_size != (1, 1) or c.stride != (1, 1) or c.padding != (0, 0) or c.groups != 1:
            raise ValueError("expansion_head: only 1x1, stride-1, ungrouped convolutions")
    return c1, c2, c3


def expansion_head(x: torch.Tensor, cluster1: nn.Sequential, cluster2: nn.Sequential,
                   packed: Optional[_PackedHead] = None) -> torch.Tensor:
    """``cluster1(x) + cluster2(x)`` (model/dino_pqgo.py:127-128) for NCHW ``x``; (B, D, h, w) result, NHWC memory
    on the kernel path."""
    if x.dim() != 4:
        raise ValueError(f"expansion_head expects (B, C, h, w) features, got shape {tuple(x.shape)}")
    c1, c2, c3 = _check_pair(cluster1, cluster2)
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in
                                                                     list(cluster1.parameters()) + list(cluster2.parameters())))
    if needs_grad:
        code = cluster1(x)
        code = code + cluster2(x)
        return code
    B, C, h, w = x.shape
    D = c1.weight.shape[0]
    w13, b13 = (packed or _PackedHead()).get(c1, c3, reuse=not (cluster1.training or cluster2.training))
    hidden = ops.head_gemm(x, c2.weight, c2.bias, relu=True)                  # (n, C) flat
    code = ops.head_gemm(x, w13, b13, a2=hidden)                              # (n, D) flat
    return code.view(B, h, w, D).permute(0, 3, 1, 2)


class SegmentationHead(nn.Module):
    """model/blocks/module.py:20-44 (same constructor, attributes and state_dict keys)."""

    def __init__(self, input_dim: int, hidden_dim: int):
        super().__init__()
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.cluster1 = self.make_clusterer()
        self.cluster2 = self.make_nonlinear_clusterer()
        self._packed = _PackedHead()

    def make_clusterer(self) -> nn.Sequential:
        return make_clusterer(self.input_dim, self.hidden_dim)

    def make_nonlinear_clusterer(self) -> nn.Sequential:
        return make_nonlinear_clusterer(self.input_dim, self.hidden_dim)

    def invalidate_cache(self) -> None:
        """Call after changing the convolution weights through ``.data`` while in eval() mode."""
        self._packed.key = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return expansion_head(x, self.cluster1, self.cluster2, self._packed)
