"""Drop-in mirrors of the PQ classes the reference's trainers actually instantiate:
the inline ``Codebook`` / ``EMACodebook`` / ``ProductQuantizerWrapper`` copies in ``model/dino_pqgo.py``
(V5, the ``train.py`` path), ``model/dino_new_vq.py`` (V4, the ``train_vq.py`` path) and
``model/dino_pqgo_cls.py`` (V6).  They differ from ``model/quantizer.py`` in: NCHW input, the quantised
rows come from the RAW codebook, the soft assignment is divided by ``jsd_ts``, the loss keys, and the
return tuples.  Host-RNG research flags (pq_dropout, gumbel, weighted-sum, k-means / restart init) are
not part of the accelerated path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa

from . import _pq_core as core
from . import ops
from .dist_utils import all_reduce_tensor
from .quantizer import EmbeddingEMA

__all__ = ["Codebook", "EMACodebook", "PQGOProductQuantizerWrapper", "NewVQProductQuantizerWrapper",
           "PQGOClsProductQuantizerWrapper", "JSDLoss", "EntropyLoss"]


class EntropyLoss(nn.Module):
    """model/loss.py:490-505."""

    def forward(self, p: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
        avg_p = p.mean(0)
        return -torch.sum(-avg_p * torch.log(avg_p + 1e-8), dim=-1)


class JSDLoss(nn.Module):
    """model/loss.py:508-525."""

    def __init__(self, reduction="batchmean"):
        super().__init__()
        self.kl = nn.KLDivLoss(reduction=reduction, log_target=True)

    def forward(self, p: torch.Tensor, q: torch.Tensor):
        m = (0.5 * (p + q).add(1e-6)).log()
        return 0.5 * (self.kl(m, p.add(1e-6).log()) + self.kl(m, q.add(1e-6).log()))


def _unsupported(**flags) -> None:
    bad = [k for k, v in flags.items() if v]
    if bad:
        raise NotImplementedError(f"{', '.join(bad)} is outside the accelerated PQ path (SURVEY.md 7.5)")


class Codebook(nn.Module):
    """Learned codebook, model/dino_pqgo.py:460-705 (same class in dino_new_vq.py:462-671 and
    dino_pqgo_cls.py:191-405 modulo the forward arity).  ``forward(z, z_pos=None)``."""

    def __init__(self, num_codebook_vectors: int, latent_dim: int, beta=0.25, book=1.0, normalize: str = "none",
                 use_restart: bool = False, use_split: bool = False, use_weighted_sum: bool = False,
                 use_gumbel: bool = False, need_initialized: str = "none", pq_dropout: float = 0.0, jsd_ts: float = 1.0,
                 num_query: int = 3, num_pos: int = 10):
        super().__init__()
        _unsupported(use_weighted_sum=use_weighted_sum, use_gumbel=use_gumbel, pq_dropout=pq_dropout > 0.0,
                     use_split=use_split, need_initialized=need_initialized not in ("none", "uni", "normal"))
        self.latent_dim, self.beta, self.book = latent_dim, beta, book
        self.num_codebook_vectors = num_codebook_vectors
        self.embedding = nn.Embedding(num_codebook_vectors, latent_dim)
        self.embedding.weight.data.uniform_(-1.0 / num_codebook_vectors, 1.0 / num_codebook_vectors)
        self.vq_count = torch.zeros(self.num_codebook_vectors)
        self.normalize = normalize
        if normalize == "z_trainable":
            self.z_mean = nn.Parameter(torch.zeros(self.latent_dim))
            self.z_log_var = nn.Parameter(torch.zeros(self.latent_dim))
        self.use_restart = use_restart
        self.need_initialized = need_initialized
        self.jsd_ts = jsd_ts

    def forward(self, z: torch.Tensor, z_pos: Optional[torch.Tensor] = None):
        q, out, prob, idx = _codebook_group_forward([self], z)
        return q, out, prob[0], idx[0]


def _codebook_group_forward(mods: List[Codebook], z: torch.Tensor, want_prob: bool = True):
    """dino_pqgo.Codebook.forward for M subspaces at once.  Returns (z_q NCHW, outputs, [prob_i (B,h,w,K)],
    [idx_i (B,h,w)])."""
    q0 = mods[0]
    M, K, mode = len(mods), q0.num_codebook_vectors, q0.normalize
    B, D, h, w = z.shape
    d = D // M
    training = q0.training
    if q0.need_initialized != "none" and training:
        for q in mods:
            if q.need_initialized == "uni":
                nn.init.xavier_uniform_(q.embedding.weight)
            elif q.need_initialized == "normal":
                nn.init.xavier_normal_(q.embedding.weight)
            q.need_initialized = "none"
    codebook = torch.stack([q.embedding.weight for q in mods])
    norm_a = norm_b = None
    if mode == "z_trainable":
        norm_a = torch.cat([q.z_mean for q in mods])
        norm_b = torch.cat([q.z_log_var for q in mods]).exp().sqrt() + 1e-5
    cbn = core.normalize_codebook(codebook, mode, ema_style=True)                 # dino_pqgo.py:613-641
    idx, out, mse_commit, mse_cb, prob = core.pq_quantize(z, cbn, codebook, mode, norm_a, norm_b, want_prob=want_prob,
                                                          temperature=q0.jsd_ts)   # :646-665 (raw embedding gathered)
    output: Dict[str, torch.Tensor] = {}
    if training:
        with torch.no_grad():
            packed = ops.pq_accumulate(z.detach().float(), idx, K)
            count = all_reduce_tensor(packed[:, :, d].contiguous(), op="sum")         # :672-673
            for i, q in enumerate(mods):
                q.vq_count = q.vq_count.to(z.device) + count[i]                       # :675
            output["codebook-usage"] = ((K - (count == 0).sum(dim=1).float()) / K).mean()   # :681-682
    output["vq-loss"] = (q0.book * mse_cb + q0.beta * mse_commit).mean()              # :685-687
    n = B * h * w
    idx64 = idx.long()
    probs = [prob.view(B, h, w, M, K)[:, :, :, i, :] for i in range(M)] if prob is not None else [None] * M
    return out, output, probs, [idx64[i].view(B, h, w) for i in range(M)]


class EMACodebook(nn.Module):
    """EMA codebook, model/dino_new_vq.py:241-459.  ``forward(z, i, it)``; decay/eps are hard-coded to
    0.99 / 1e-5 in the reference (:267-268)."""

    def __init__(self, num_codebook_vectors: int, latent_dim: int, beta=0.25, normalize: str = "none",
                 use_restart: bool = False, use_weighted_sum: bool = False, need_initialized: str = "none",
                 pq_dropout: float = 0.0, jsd_ts: float = 1.0, **_ignored):
        super().__init__()
        _unsupported(use_weighted_sum=use_weighted_sum, pq_dropout=pq_dropout > 0.0, use_restart=use_restart,
                     need_initialized=need_initialized not in ("none",))
        self.latent_dim, self.beta = latent_dim, beta
        self.num_codebook_vectors = num_codebook_vectors
        self.codebook = EmbeddingEMA(num_codebook_vectors, latent_dim, decay=0.99, eps=1.0e-5)
        self.register_buffer("vq_count", torch.zeros(num_codebook_vectors), persistent=False)
        self.normalize = normalize
        if normalize == "z_trainable":
            self.z_mean = nn.Parameter(torch.zeros(self.latent_dim))
            self.z_log_var = nn.Parameter(torch.zeros(self.latent_dim))
        self.need_initialized = need_initialized
        self.jsd_loss, self.entropy_loss = JSDLoss(), EntropyLoss()
        self.jsd_ts = jsd_ts

    def forward(self, z: torch.Tensor, i: int = 0, it: int = 0):
        q, out, prob = _ema_codebook_group_forward([self], z)
        return q, out, prob


def _ema_codebook_group_forward(mods: List[EMACodebook], z: torch.Tensor, want_prob: bool = True):
    q0 = mods[0]
    M, K, mode = len(mods), q0.num_codebook_vectors, q0.normalize
    B, D, h, w = z.shape
    d = D // M
    weight = torch.stack([q.codebook.weight for q in mods])
    norm_a = norm_b = None
    if mode == "z_trainable":
        norm_a = torch.cat([q.z_mean for q in mods])
        norm_b = torch.cat([q.z_log_var for q in mods]).exp().sqrt() + 1e-5
    with torch.no_grad():
        cbn = core.normalize_codebook(weight, mode, ema_style=True)
        src = weight.clone()                                                        # raw codebook gathered (:403)
    idx, out, mse_commit, _, prob = core.pq_quantize(z, cbn, src, mode, norm_a, norm_b, want_prob=True,
                                                     temperature=q0.jsd_ts)
    output: Dict[str, torch.Tensor] = {}
    if q0.training:
        with torch.no_grad():
            packed = core.ema_statistics(z, idx, K)                                 # :408-413 (raw z sums)
            exact = torch.stack([q.vq_count.to(z.device) for q in mods]).contiguous()
            vqc = torch.stack([q.codebook.vq_count for q in mods]).contiguous()
            wavg = torch.stack([q.codebook.weight_avg for q in mods]).contiguous()
            wnew = weight.detach().clone().contiguous()
            unused = ops.ema_update(packed, 0.99, 1.0e-5, vqc, wavg, wnew, exact)    # :415-422
            for i, q in enumerate(mods):
                q.vq_count = exact[i]; q.codebook.vq_count.copy_(vqc[i])
                q.codebook.weight_avg.copy_(wavg[i]); q.codebook.weight.copy_(wnew[i])
            output["codebook-usage"] = ((K - unused.float()) / K).mean()            # :431-432
    output["vq-loss"] = q0.beta * mse_commit.mean()                                 # :435-436
    output["codebook-sum"] = torch.sum(torch.abs(torch.stack([q.codebook.weight for q in mods]))) / M
    # JSD / entropy between the two halves of the batch (:447-450), per subspace then averaged
    n = B * h * w
    pv = prob.view(n, M, K)
    p1, p2 = torch.chunk(pv, chunks=2, dim=0)
    jsd = torch.stack([q0.jsd_loss(p1[:, i], p2[:, i]) for i in range(M)]).mean()
    ent = torch.stack([q0.entropy_loss(p1[:, i], p2[:, i]) for i in range(M)]).mean()
    output["jsd"], output["entropy"] = jsd, ent
    return out, output, prob


class _WrapperBase(nn.Module):
    def __init__(self, num_pq: int, embed_dim: int):
        super().__init__()
        if embed_dim % num_pq != 0:
            raise ValueError(f"Embed dim {embed_dim} should be divisible by #PQ {num_pq}.")
        self.num_pq = num_pq
        self.pq_dim = embed_dim // num_pq
        self.materialize_prob = True


class PQGOProductQuantizerWrapper(_WrapperBase):
    """model/dino_pqgo.py:708-776: ``forward(z, z_pos=None, it=-1)`` ->
    (z_q, (z_split, [z_q_i], [idx_i (B,h,w)]), outputs, distance_prob (B,h,w,K*M))."""

    def __init__(self, num_pq: int, num_codebook: int, embed_dim: int, beta: float = 0.25, book: float = 1.0,
                 normalize: Optional[str] = None, decay: float = 0.99, eps: float = 1e-5, use_restart: bool = False,
                 use_split: bool = False, use_gumbel: bool = False, use_weighted_sum: bool = False,
                 update_norm: bool = True, need_initialized: str = "none", pq_dropout: float = 0.0,
                 jsd_ts: float = 1.0, num_query: int = 3, num_pos: int = 10, quantizer_cls=Codebook) -> None:
        super().__init__(num_pq, embed_dim)
        self.quantizers = nn.ModuleList([
            quantizer_cls(num_codebook, self.pq_dim, beta=beta, book=book, normalize=normalize, use_restart=use_restart,
                          use_split=use_split, use_weighted_sum=use_weighted_sum, need_initialized=need_initialized,
                          pq_dropout=pq_dropout, jsd_ts=jsd_ts, num_query=num_query, num_pos=num_pos)
            for _ in range(self.num_pq)
        ])

    def forward(self, z: torch.Tensor, z_pos: torch.Tensor = None, it: int = -1):
        # the reference also pushes z_pos through a second distance+softmax whose result is never used
        # (dino_pqgo.py:650-656,700); that dead work is not reproduced.
        z_q, outputs, probs, idxs = _codebook_group_forward(list(self.quantizers), z, want_prob=self.materialize_prob)
        z_split = torch.chunk(z, chunks=self.num_pq, dim=1)
        z_quantized = list(torch.chunk(z_q, chunks=self.num_pq, dim=1))
        prob = torch.cat(probs, dim=-1) if probs[0] is not None else None
        return z_q, (z_split, z_quantized, idxs), outputs, prob


class NewVQProductQuantizerWrapper(_WrapperBase):
    """model/dino_new_vq.py:674-732: ``forward(z, it)`` -> (z_q, outputs, distance_prob (n, K*M))."""

    def __init__(self, num_pq: int, num_codebook: int, embed_dim: int, beta: float = 0.25,
                 normalize: Optional[str] = None, decay: float = 0.99, eps: float = 1e-5, use_restart: bool = False,
                 use_gumbel: bool = False, use_split: bool = False, use_weighted_sum: bool = False,
                 update_norm: bool = True, need_initialized: str = "none", pq_dropout: float = 0.0,
                 jsd_ts: float = 1.0, quantizer_cls=EMACodebook) -> None:
        super().__init__(num_pq, embed_dim)
        self.quantizers = nn.ModuleList([
            quantizer_cls(num_codebook, self.pq_dim, beta=beta, normalize=normalize, use_restart=use_restart,
                          use_weighted_sum=use_weighted_sum, need_initialized=need_initialized, pq_dropout=pq_dropout,
                          jsd_ts=jsd_ts)
            for _ in range(self.num_pq)
        ])

    def forward(self, z: torch.Tensor, it: int = 0):
        qs = list(self.quantizers)
        if all(isinstance(q, EMACodebook) for q in qs):
            return _ema_codebook_group_forward(qs, z)
        z_q, outputs, probs, _ = _codebook_group_forward(qs, z)
        B, D, h, w = z.shape
        K = qs[0].num_codebook_vectors
        return z_q, outputs, torch.cat([p.reshape(B * h * w, K) for p in probs], dim=-1)


class PQGOClsProductQuantizerWrapper(PQGOProductQuantizerWrapper):
    """model/dino_pqgo_cls.py:408-471: ``forward(z, it=-1)`` -> (z_q, outputs, distance_prob); the
    per-subspace quantiser returns flat (n,) pseudo-label indices."""

    def forward(self, z: torch.Tensor, it: int = -1):
        z_q, outputs, probs, idxs = _codebook_group_forward(list(self.quantizers), z, want_prob=self.materialize_prob)
        prob = torch.cat(probs, dim=-1) if probs[0] is not None else None
        self.pseudo_labels = [i.reshape(-1) for i in idxs]
        return z_q, outputs, prob
