"""Drop-in mirror of the reference's ``model/quantizer_v2.py`` (live classes, :30-308 and :542-598).

``VectorQuantizer`` is the same learned-codebook quantiser as in ``quantizer.py``.  ``EMAVectorQuantizer``
here is the simplified EMA variant with its reference quirks kept on purpose:
always l2; the output rows are gathered from the NORMALISED INPUT ROWS ``z_norm[idx]`` rather than from
the codebook (quantizer_v2.py:274); EMA sums use ``z_norm``; the all-reduce results are discarded
(:283-284), so the statistics are per rank; no straight-through estimator.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa

from . import _pq_core as core
from . import ops
from .quantizer import VectorQuantizer as _VQ, get_histogram_count  # noqa: F401

__all__ = ["VectorQuantizer", "EMAVectorQuantizer", "ProductQuantizerWrapper"]


class VectorQuantizer(_VQ):
    """model/quantizer_v2.py:30-196 (identical arithmetic to quantizer.VectorQuantizer)."""

    def __init__(self, num_codebook: int, embed_dim: int, beta: float = 0.25, normalize: Optional[str] = None,
                 use_restart: bool = False, use_gumbel: bool = False, use_split: bool = False,
                 use_weighted_sum: bool = False, update_norm: bool = True, **_ignored) -> None:
        super().__init__(num_codebook, embed_dim, beta=beta, normalize=normalize, use_restart=use_restart,
                         use_gumbel=use_gumbel, use_split=use_split, use_weighted_sum=use_weighted_sum,
                         update_norm=update_norm)


class EMAVectorQuantizer(nn.Module):
    """model/quantizer_v2.py:198-308."""

    def __init__(self, n_codes: int, embedding_dim: int, beta: float = 0.25, normalize: Optional[str] = None,
                 decay: float = 0.99, eps: float = 1e-5, use_restart: bool = False, use_gumbel: bool = False,
                 use_split: bool = False, use_weighted_sum: bool = False, update_norm: bool = True, **_ignored) -> None:
        super().__init__()
        weight = torch.randn(n_codes, embedding_dim)
        nn.init.uniform_(weight, -1.0 / n_codes, 1.0 / n_codes)
        self.register_buffer("embeddings", weight)
        self.register_buffer("N", torch.zeros(n_codes))
        self.register_buffer("z_avg", self.embeddings.data.clone())
        self.n_codes = n_codes
        self.embedding_dim = embedding_dim
        self._need_init = False
        self.decay, self.eps, self.beta = decay, eps, beta

    def forward(self, z: torch.Tensor):
        return _v2_group_forward([self], z)


def _v2_group_forward(mods, z: torch.Tensor, want_prob: bool = True):
    q0 = mods[0]
    M, K = len(mods), q0.n_codes
    B, D, h, w = z.shape
    d = D // M
    n = B * h * w
    emb = torch.stack([q.embeddings for q in mods])
    cbn = F.normalize(emb, dim=2)                                               # :261
    z32 = z.float()
    cn2 = ops.pq_cnorm2(cbn)
    idx = ops.pq_assign(z32, cbn, cn2, "l2")                                    # :262-270
    prob = ops.pq_distance_prob(z32, cbn, cn2, "l2") if want_prob else None     # :267
    # quirk (:274): rows of z_norm indexed by the code ids, i.e. the first K pixels act as the "codebook"
    if n < K:
        raise IndexError("index out of range in self")   # what F.embedding raises in the reference
    z_first = z[:, :, :, :].permute(0, 2, 3, 1).reshape(n, D)[:K]               # (K, D) view of the first K pixels
    src = torch.stack([F.normalize(z_first[:, i * d:(i + 1) * d], dim=1) for i in range(M)])   # [M, K, d], differentiable
    out, mse_commit, _ = core.PQGatherLoss.apply(z32.detach(), src.detach(), idx, "l2", None, None)
    gathered = torch.gather(src, 1, idx.long().unsqueeze(-1).expand(M, n, d))   # [M, n, d] keeps the reference's graph
    q = gathered.permute(1, 0, 2).reshape(B, h, w, D).permute(0, 3, 1, 2).contiguous()
    output: Dict[str, torch.Tensor] = {}
    if q0.training:
        with torch.no_grad():
            packed = ops.pq_accumulate(z32, idx, K, use_norm=True, normalize="l2")          # :279-281 (sums of z_norm)
            Ns = torch.stack([m_.N for m_ in mods]).contiguous()
            zavg = torch.stack([m_.z_avg for m_ in mods]).contiguous()
            embs = emb.detach().clone().contiguous()
            ops.ema_update(packed, q0.decay, q0.eps, Ns, zavg, embs)                        # :286-292
            for i, m_ in enumerate(mods):
                m_.N.copy_(Ns[i]); m_.z_avg.copy_(zavg[i]); m_.embeddings.copy_(embs[i])
    commitment = mse_commit.mean()
    output["commitment-loss"] = commitment
    output["loss"] = q0.beta * commitment
    output["codebook-sum"] = torch.sum(torch.abs(torch.stack([m_.embeddings for m_ in mods]))) / M
    return q, output, prob


class ProductQuantizerWrapper(nn.Module):
    """model/quantizer_v2.py:542-598."""

    def __init__(self, num_pq: int, num_codebook: int, embed_dim: int, beta: float = 0.25,
                 normalize: Optional[str] = None, decay: float = 0.99, eps: float = 1e-5, use_restart: bool = False,
                 use_gumbel: bool = False, use_split: bool = False, use_weighted_sum: bool = False,
                 update_norm: bool = True, quantizer_cls=EMAVectorQuantizer) -> None:
        super().__init__()
        if embed_dim % num_pq != 0:
            raise ValueError(f"Embed dim {embed_dim} should be divisible by #PQ {num_pq}.")
        self.num_pq = num_pq
        self.pq_dim = embed_dim // num_pq
        self.materialize_prob = True
        self.quantizers = nn.ModuleList([
            quantizer_cls(num_codebook, self.pq_dim, beta=beta, normalize=normalize, decay=decay, eps=eps,
                          use_restart=use_restart, use_gumbel=use_gumbel, use_split=use_split,
                          use_weighted_sum=use_weighted_sum, update_norm=update_norm)
            for _ in range(self.num_pq)
        ])

    def forward(self, z: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, torch.Tensor], torch.Tensor]:
        qs = list(self.quantizers)
        if all(isinstance(q, EMAVectorQuantizer) for q in qs):
            return _v2_group_forward(qs, z, want_prob=self.materialize_prob)
        from .quantizer import _param_group_forward
        return _param_group_forward(qs, z, want_prob=self.materialize_prob)
