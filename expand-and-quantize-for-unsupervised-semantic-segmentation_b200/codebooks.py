"""Drop-in mirrors of the PQ classes the reference's trainers actually instantiate:
the inline ``Codebook`` / ``EMACodebook`` / ``ProductQuantizerWrapper`` copies in ``model/dino_pqgo.py``
(V5, the ``train.py`` path), ``model/dino_new_vq.py`` (V4, the ``train_vq.py`` path) and
``model/dino_pqgo_cls.py`` (V6).  They differ from ``model/quantizer.py`` in: NCHW input, the quantised
rows come from the RAW codebook, the soft assignment is divided by ``jsd_ts``, the loss keys, and the
return tuples.  One ``Codebook`` class serves the three files; ``variant`` selects the file it stands for:

  variant     forward(...)          returns                                  vq-loss                  restart rows
  "pqgo"      (z, z_pos)            q, out, prob (b,h,w,K), idx (b,h,w)     book*cb + beta*commit    z_norm
  "new_vq"    (z, i, it)            q, out, prob (n,K)  (+ jsd, entropy)    cb + beta*commit         raw z
  "pqgo_cls"  (z)                   q, out, prob (b,h,w,K), idx (n,)        cb + beta*commit         z_norm

``use_weighted_sum`` runs as host PyTorch on the kernel's soft assignment (``_weighted_sum``).  ``pq_dropout`` (pqgo /
new_vq; the reference masks the codebook with ``torch.cuda.FloatTensor`` noise on every call and then indexes the FULL
codebook with indices into the masked one) draws and assigns in host PyTorch (``_host_paths.dropout_assign``: the soft
assignment has a different width per subspace) and feeds the drawn indices to the gather / scatter-add / EMA kernels.
The inline classes' ``use_gumbel`` is only admitted next to the weighted sum,
whose branch takes precedence (dino_pqgo.py:502-503,658-663), exactly as in the reference.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _host_paths as hp
from . import _pq_core as core
from . import ops
from ._host_paths import draw_restart
from .dist_utils import all_reduce_tensor
from .quantizer import EmbeddingEMA

__all__ = ["Codebook", "EMACodebook", "PQGOProductQuantizerWrapper", "NewVQProductQuantizerWrapper",
           "PQGOClsProductQuantizerWrapper", "JSDLoss", "EntropyLoss"]


class EntropyLoss(nn.Module):
    """model/loss.py:490-505: the NEGATIVE entropy of the batch-averaged assignment, sum_k a_k log(a_k + 1e-8) with
    a = mean over rows of ``p``; the second argument is ignored, as in the reference."""

    def forward(self, p: torch.Tensor, q: Optional[torch.Tensor] = None) -> torch.Tensor:
        a = p.mean(dim=0)
        return torch.sum(a * torch.log(a + 1e-8), dim=-1)


class JSDLoss(nn.Module):
    """model/loss.py:508-525: Jensen-Shannon divergence of two row-stochastic matrices with the reference's 1e-6
    smoothing, averaged over rows ("batchmean")."""

    def __init__(self, reduction: str = "batchmean"):
        super().__init__()
        if reduction != "batchmean":
            raise ValueError("JSDLoss mirrors the reference's batchmean reduction only")

    def forward(self, p: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
        lo, hi = torch.aminmax(torch.stack([p.detach().amin(), p.detach().amax(), q.detach().amin(), q.detach().amax()]))
        if lo < 0.0 or hi > 1.0:
            raise ValueError(f"min, max of the inputs : {float(lo)}, {float(hi)}")        # loss.py:519-521
        ps, qs = p + 1e-6, q + 1e-6
        log_mix = torch.log(0.5 * (p + q + 1e-6))
        kl = (ps * (torch.log(ps) - log_mix)).sum() + (qs * (torch.log(qs) - log_mix)).sum()
        return 0.5 * kl / p.shape[0]


def _unsupported(**flags) -> None:
    bad = [k for k, v in flags.items() if v]
    if bad:
        raise NotImplementedError(f"{', '.join(bad)} is outside the accelerated PQ path (SURVEY.md 7.5)")


class Codebook(nn.Module):
    """Learned codebook, model/dino_pqgo.py:460-705, dino_new_vq.py:462-671, dino_pqgo_cls.py:191-405.  Defaults are the
    reference's (``need_initialized="kmeans"``: a directly constructed codebook re-initialises itself from the first
    training batch; the wrappers pass their own default, "none")."""

    def __init__(self, num_codebook_vectors: int, latent_dim: int, beta=0.25, book=1.0, normalize: str = "none",
                 use_restart: bool = False, use_split: bool = False, use_weighted_sum: bool = False,
                 use_gumbel: bool = False, need_initialized: str = "kmeans", pq_dropout: float = 0.0, jsd_ts: float = 1.0,
                 num_query: int = 3, num_pos: int = 10, variant: str = "pqgo"):
        super().__init__()
        _unsupported(pq_dropout=pq_dropout > 0.0 and variant == "pqgo_cls",       # dino_pqgo_cls.py has no such flag
                     need_initialized=need_initialized not in ("none", "uni", "normal", "rand", "kmeans"))
        self.use_split = use_split       # stored and never read, as in the reference (dino_pqgo.py:510; dino_new_vq.py:640 is a comment)
        self.pq_dropout = pq_dropout
        self.use_weighted_sum = use_weighted_sum
        if use_weighted_sum:
            assert normalize == "none", "Weight_sum should be unnormalized"          # dino_pqgo.py:499-500
        if use_gumbel:      # :502-503 -- the reference only admits it next to the weighted sum, whose branch then wins (:658-663)
            assert use_weighted_sum, "Weight_sum and Gumbel should not be both true"
        if variant not in ("pqgo", "new_vq", "pqgo_cls"):
            raise ValueError(f"unknown Codebook variant {variant}")
        self.variant = variant
        self.latent_dim, self.beta, self.book = latent_dim, beta, book
        self.num_codebook_vectors = num_codebook_vectors
        self.embedding = nn.Embedding(num_codebook_vectors, latent_dim)
        self.embedding.weight.data.uniform_(-1.0 / num_codebook_vectors, 1.0 / num_codebook_vectors)
        self.vq_count = torch.zeros(self.num_codebook_vectors)
        self.update_indices = self.update_candidates = None
        self.normalize = normalize
        if normalize == "z_trainable":
            self.z_mean = nn.Parameter(torch.zeros(self.latent_dim))
            self.z_log_var = nn.Parameter(torch.zeros(self.latent_dim))
        self.use_restart = use_restart
        self.need_initialized = need_initialized
        self.jsd_ts = jsd_ts

    @torch.no_grad()
    def prepare_restart(self, vq_current_count: torch.Tensor, z_flat: torch.Tensor) -> None:
        self.update_indices, self.update_candidates = draw_restart(vq_current_count, z_flat)

    @torch.no_grad()
    def restart(self) -> None:
        """dino_pqgo.py:572-577: overwrite the drawn dead codes, clear the exact counter."""
        if self.update_indices is None or self.update_candidates is None:
            return
        self.embedding.weight.data[self.update_indices] = self.update_candidates.float()
        self.vq_count.fill_(0)
        self.update_indices = self.update_candidates = None

    def forward(self, z: torch.Tensor, *args, **kwargs):
        """Arity follows the variant (see module docstring); extra arguments (z_pos / i / it) are not used by the
        arithmetic -- the reference's second pass over ``z_pos`` (dino_pqgo.py:650-656,700) produces nothing that
        leaves the function."""
        q, out, probs, idxs = _codebook_group_forward([self], z)
        if self.variant == "new_vq":
            B, D, h, w = z.shape
            return q, out, probs[0].reshape(B * h * w, -1)
        if self.variant == "pqgo_cls":
            return q, out, probs[0], idxs[0].reshape(-1)
        return q, out, probs[0], idxs[0]


def _init_codebooks(mods: List[Codebook], z: torch.Tensor, d: int) -> None:
    """First-training-call initialisation (dino_pqgo.py:589-609): "rand" draws K rows of z, "kmeans" runs scikit-learn
    k-means on the rows (host, once), "uni" / "normal" re-draw the embedding with Xavier."""
    zf = None
    for i, q in enumerate(mods):
        if q.need_initialized in ("rand", "kmeans") and zf is None:
            zf = z.detach().permute(0, 2, 3, 1).reshape(-1, z.shape[1])
        if q.need_initialized == "rand":
            q.vq_count = q.vq_count.to(z.device)
            q.prepare_restart(torch.zeros(q.num_codebook_vectors, dtype=torch.long, device=z.device), zf[:, i * d:(i + 1) * d])
            q.restart()
        elif q.need_initialized == "kmeans":
            q.embedding.weight.data.copy_(hp.kmeans_centroids(zf[:, i * d:(i + 1) * d], q.num_codebook_vectors))
        elif q.need_initialized == "uni":
            nn.init.xavier_uniform_(q.embedding.weight)
        elif q.need_initialized == "normal":
            nn.init.xavier_normal_(q.embedding.weight)
        q.need_initialized = "none"


def _codebook_group_forward(mods: List[Codebook], z: torch.Tensor, want_prob: bool = True):
    """Codebook.forward of the three inline variants for M subspaces at once.  Returns (z_q NCHW, outputs,
    [prob_i (B,h,w,K)] or [None], [idx_i (B,h,w)])."""
    q0 = mods[0]
    M, K, mode, variant = len(mods), q0.num_codebook_vectors, q0.normalize, q0.variant
    B, D, h, w = z.shape
    d = D // M
    training = q0.training
    if q0.need_initialized != "none" and training:
        _init_codebooks(mods, z, d)
    codebook = torch.stack([q.embedding.weight for q in mods])
    norm_a = norm_b = None
    if mode == "z_trainable":
        norm_a = torch.cat([q.z_mean for q in mods])
        norm_b = torch.cat([q.z_log_var for q in mods]).exp().sqrt() + 1e-5
    cbn = core.normalize_codebook(codebook, mode, ema_style=True)                 # dino_pqgo.py:613-641
    need_soft_stats = variant == "new_vq"
    grad_path = core._wants_grad(z, codebook, norm_a, norm_b)
    make_prob = want_prob or (need_soft_stats and grad_path) or q0.use_weighted_sum
    drop = None
    if q0.pq_dropout > 0.0:                                                          # dino_pqgo.py:641-644, every call
        idx, out, mse_commit, mse_cb, drop = _dropout_quantize(z, cbn, codebook, mode, norm_a, norm_b, M, q0)
        prob = None
    else:
        idx, out, mse_commit, mse_cb, prob = core.pq_quantize(z, cbn, codebook, mode, norm_a, norm_b, want_prob=make_prob,
                                                              temperature=q0.jsd_ts)   # :646-665 (raw embedding gathered)
        if q0.use_weighted_sum:                                                      # dino_pqgo.py:661-662,691
            out, mse_commit, mse_cb = _weighted_sum(z, prob, cbn, mode, norm_a, norm_b, M, K)
    output: Dict[str, torch.Tensor] = {}
    if training:
        with torch.no_grad():
            packed = ops.pq_accumulate(z.detach().float(), idx, K)
            count = all_reduce_tensor(packed[:, :, d].contiguous(), op="sum")         # :672-673
            for i, q in enumerate(mods):
                q.vq_count = q.vq_count.to(z.device) + count[i]                       # :675
            if q0.use_restart:                                                        # :677-679
                rows = z.detach().float().permute(0, 2, 3, 1).reshape(B * h * w, M, d)
                if variant != "new_vq":                                               # pqgo / pqgo_cls draw from z_norm
                    rows = core._normalize_rows(rows, mode, norm_a.detach() if norm_a is not None else None,
                                                norm_b.detach() if norm_b is not None else None)
                for i, q in enumerate(mods):
                    q.prepare_restart(count[i], rows[:, i])
                    q.restart()
            kept = drop.kept if drop is not None else K                               # codes the argmin could see
            output["codebook-usage"] = ((kept - (count == 0).sum(dim=1).float()) / kept).mean()   # :681-682
    book = q0.book if variant == "pqgo" else 1.0
    output["vq-loss"] = (book * mse_cb + q0.beta * mse_commit).mean()                 # :685-687
    if need_soft_stats:                                                               # dino_new_vq.py:666-668
        output["jsd"], output["entropy"] = (drop.soft_stats() if drop is not None else
                                            _soft_stats(z, cbn, mode, norm_a, norm_b, q0.jsd_ts, prob, M, K))
    idx64 = idx.long()
    if drop is not None:
        probs = [p.view(B, h, w, -1) for p in drop.probs] if want_prob else [None] * M
        return out, output, probs, [idx64[i].view(B, h, w) for i in range(M)]
    probs = [prob.view(B, h, w, M, K)[:, :, :, i, :] for i in range(M)] if (prob is not None and want_prob) else [None] * M
    return out, output, probs, [idx64[i].view(B, h, w) for i in range(M)]


def _weighted_sum(z, prob, cbn, mode, norm_a, norm_b, M: int, K: int):
    """``use_weighted_sum`` (dino_new_vq.py:400-401,616-617; dino_pqgo.py:406-407,661-662): the quantised row is the
    soft-assignment-weighted sum of the codes, ``softmax(-d / T) @ codebook_norm``, returned WITHOUT the straight-through
    estimator -- gradients reach z (and a learned codebook) through the soft assignment.  Host PyTorch on the
    materialised probabilities.  Returns (z_q NCHW, mse(z_norm, z_q.detach()) [M],
    mse(z_q, z_norm.detach()) [M])."""
    B, D, h, w = z.shape
    n = B * h * w
    zr = core._normalize_rows(core._rows(z.float(), M), mode, norm_a, norm_b)         # (n, M, d), differentiable
    zq = torch.einsum("nmk,mkd->nmd", prob.view(n, M, K), cbn)
    commit = ((zr - zq.detach()) ** 2).mean(dim=(0, 2))
    cb_loss = ((zq - zr.detach()) ** 2).mean(dim=(0, 2))
    return zq.reshape(B, h, w, D).permute(0, 3, 1, 2).contiguous(), commit, cb_loss


class _Dropped:
    """What one ``pq_dropout`` assignment leaves behind: per-subspace soft assignments of different widths, the keep
    masks, and the number of kept codes as an [M] tensor."""

    def __init__(self, probs: List[torch.Tensor], keeps: List[torch.Tensor]):
        self.probs, self.keeps = probs, keeps
        self.kept = torch.stack([k.sum() for k in keeps]).float()

    def soft_stats(self):
        """jsd / entropy per subspace on the kept codes, then the wrapper's mean over subspaces (dino_new_vq.py:447-450)."""
        jsd, ent = JSDLoss(), EntropyLoss()
        halves = [torch.chunk(p, chunks=2, dim=0) for p in self.probs]
        return (torch.stack([jsd(a, b) for a, b in halves]).mean(), torch.stack([ent(a, b) for a, b in halves]).mean())


def _dropout_quantize(z, cbn, gather_src, mode, norm_a, norm_b, M: int, q0):
    """The ``pq_dropout`` branch shared by Codebook and EMACodebook: draw + assign in host PyTorch, then the gather /
    loss kernel on the drawn indices (which address the FULL ``gather_src``, as in the reference).  Returns
    (idx int32 [M, n], z_q NCHW, mse_commit [M], mse_codebook [M], _Dropped)."""
    B, D, h, w = z.shape
    zr = core._normalize_rows(core._rows(z.float(), M), mode, norm_a, norm_b)         # (n, M, d), differentiable
    idx_d, probs, keeps = hp.dropout_assign(zr, cbn, q0.pq_dropout, q0.jsd_ts)
    idx, out, mse_commit, mse_cb, _ = core.pq_quantize(z, cbn, gather_src, mode, norm_a, norm_b, want_prob=False, idx=idx_d)
    if q0.use_weighted_sum:                                                          # soft sum over the kept codes only
        zq = torch.stack([p @ cbn[i][k] for i, (p, k) in enumerate(zip(probs, keeps))], dim=1)      # (n, M, d)
        mse_commit = ((zr - zq.detach()) ** 2).mean(dim=(0, 2))
        mse_cb = ((zq - zr.detach()) ** 2).mean(dim=(0, 2))
        out = zq.reshape(B, h, w, D).permute(0, 3, 1, 2).contiguous()
    return idx, out, mse_commit, mse_cb, _Dropped(probs, keeps)


def _soft_stats(z, cbn, mode, norm_a, norm_b, jsd_ts, prob, M, K):
    """jsd / entropy of the soft assignment (dino_new_vq.py:447-450): from the materialised differentiable tensor
    when there is one, else by the fused kernel that never writes the N x K*M probabilities."""
    if prob is not None:
        return core.soft_assignment_stats(prob, M, K)
    return ops.pq_soft_stats(z, cbn.detach(), None, mode, norm_a, norm_b, jsd_ts)


class EMACodebook(nn.Module):
    """EMA codebook, model/dino_new_vq.py:241-459.  ``forward(z, i, it)``; decay/eps are hard-coded to
    0.99 / 1e-5 in the reference (:267-268)."""

    def __init__(self, num_codebook_vectors: int, latent_dim: int, beta=0.25, normalize: str = "none",
                 use_restart: bool = False, use_weighted_sum: bool = False, need_initialized: str = "kmeans",
                 pq_dropout: float = 0.0, jsd_ts: float = 1.0, **_ignored):
        super().__init__()
        _unsupported(need_initialized=need_initialized not in ("none", "rand", "uni", "normal", "kmeans"))
        self.pq_dropout = pq_dropout
        self.use_weighted_sum = use_weighted_sum
        if use_weighted_sum:
            assert normalize == "none", "Weight_sum should be unnormalized"          # dino_new_vq.py:276-277
        self.latent_dim, self.beta = latent_dim, beta
        self.num_codebook_vectors = num_codebook_vectors
        self.codebook = EmbeddingEMA(num_codebook_vectors, latent_dim, decay=0.99, eps=1.0e-5)
        self.register_buffer("vq_count", torch.zeros(num_codebook_vectors), persistent=False)
        self.update_indices = self.update_candidates = None
        self.normalize = normalize
        if normalize == "z_trainable":
            self.z_mean = nn.Parameter(torch.zeros(self.latent_dim))
            self.z_log_var = nn.Parameter(torch.zeros(self.latent_dim))
        self.use_restart = use_restart
        self.need_initialized = need_initialized
        self.jsd_loss, self.entropy_loss = JSDLoss(), EntropyLoss()
        self.jsd_ts = jsd_ts

    @torch.no_grad()
    def prepare_restart(self, vq_current_count: torch.Tensor, z_flat: torch.Tensor) -> None:
        self.update_indices, self.update_candidates = draw_restart(vq_current_count, z_flat)

    @torch.no_grad()
    def restart(self) -> None:
        """dino_new_vq.py:317-325."""
        if self.update_indices is None or self.update_candidates is None:
            return
        self.codebook.weight.data[self.update_indices] = self.update_candidates
        self.codebook.reset()
        self.update_indices = self.update_candidates = None

    def forward(self, z: torch.Tensor, i: int = 0, it: int = 0):
        return _ema_codebook_group_forward([self], z)


def _ema_codebook_group_forward(mods: List[EMACodebook], z: torch.Tensor, want_prob: bool = True):
    q0 = mods[0]
    M, K, mode = len(mods), q0.num_codebook_vectors, q0.normalize
    B, D, h, w = z.shape
    d = D // M
    training = q0.training
    if q0.need_initialized != "none" and training:                                  # :336-364
        zf = z.detach().permute(0, 2, 3, 1).reshape(-1, D)
        for i, q in enumerate(mods):
            if q.need_initialized == "rand":
                q.prepare_restart(torch.zeros(K, dtype=torch.long, device=z.device), zf[:, i * d:(i + 1) * d])
                q.restart()
            elif q.need_initialized == "kmeans":                                    # :345-352
                centroids = hp.kmeans_centroids(zf[:, i * d:(i + 1) * d], K)
                q.codebook.weight.data.copy_(centroids); q.codebook.weight_avg.data.copy_(centroids)
            elif q.need_initialized in ("uni", "normal"):
                init = nn.init.xavier_uniform_ if q.need_initialized == "uni" else nn.init.xavier_normal_
                init(q.codebook.weight); init(q.codebook.weight_avg)
            q.need_initialized = "none"
    weight = torch.stack([q.codebook.weight for q in mods])
    norm_a = norm_b = None
    if mode == "z_trainable":
        norm_a = torch.cat([q.z_mean for q in mods])
        norm_b = torch.cat([q.z_log_var for q in mods]).exp().sqrt() + 1e-5
    with torch.no_grad():
        cbn = core.normalize_codebook(weight, mode, ema_style=True)
        src = weight.clone()                             # raw codebook gathered, as it is BEFORE this step's update (:403)
    grad_path = core._wants_grad(z, norm_a, norm_b)
    drop = None
    if q0.pq_dropout > 0.0:                                                          # dino_new_vq.py:388-391, every call
        idx, out, mse_commit, _, drop = _dropout_quantize(z, cbn, src, mode, norm_a, norm_b, M, q0)
        prob = torch.cat(drop.probs, dim=-1)                                         # (n, sum of kept codes)
    else:
        idx, out, mse_commit, _, prob = core.pq_quantize(z, cbn, src, mode, norm_a, norm_b,
                                                         want_prob=want_prob or grad_path or q0.use_weighted_sum,
                                                         temperature=q0.jsd_ts)
        if q0.use_weighted_sum:                                                      # dino_new_vq.py:400-401,438
            out, mse_commit, _ = _weighted_sum(z, prob, cbn, mode, norm_a, norm_b, M, K)
    output: Dict[str, torch.Tensor] = {}
    if training:
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
            if q0.use_restart:                                                      # :424-425 (drawn, applied by restart())
                rows = z.detach().float().permute(0, 2, 3, 1).reshape(B * h * w, M, d)
                for i, q in enumerate(mods):
                    q.prepare_restart(packed[i, :, d], rows[:, i])
            kept = drop.kept if drop is not None else K
            output["codebook-usage"] = ((kept - unused.float()) / kept).mean()      # :431-432
    output["vq-loss"] = q0.beta * mse_commit.mean()                                 # :435-436
    output["codebook-sum"] = torch.sum(torch.abs(torch.stack([q.codebook.weight for q in mods]))) / M
    output["jsd"], output["entropy"] = (drop.soft_stats() if drop is not None else
                                        _soft_stats(z, cbn, mode, norm_a, norm_b, q0.jsd_ts, prob, M, K))   # :447-450
    return out, output, (prob if want_prob else None)


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

    variant = "pqgo"

    def __init__(self, num_pq: int, num_codebook: int, embed_dim: int, beta: float = 0.25, book: float = 1.0,
                 normalize: Optional[str] = None, decay: float = 0.99, eps: float = 1e-5, use_restart: bool = False,
                 use_split: bool = False, use_gumbel: bool = False, use_weighted_sum: bool = False,
                 update_norm: bool = True, need_initialized: str = "none", pq_dropout: float = 0.0,
                 jsd_ts: float = 1.0, num_query: int = 3, num_pos: int = 10, quantizer_cls=Codebook) -> None:
        super().__init__(num_pq, embed_dim)
        extra = {"variant": self.variant} if quantizer_cls is Codebook else {}
        self.quantizers = nn.ModuleList([
            quantizer_cls(num_codebook, self.pq_dim, beta=beta, book=book, normalize=normalize, use_restart=use_restart,
                          use_split=use_split, use_weighted_sum=use_weighted_sum, need_initialized=need_initialized,
                          pq_dropout=pq_dropout, jsd_ts=jsd_ts, num_query=num_query, num_pos=num_pos, **extra)
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
                 jsd_ts: float = 1.0, quantizer_cls=Codebook) -> None:
        super().__init__(num_pq, embed_dim)
        extra = {"variant": "new_vq"} if quantizer_cls is Codebook else {}
        self.quantizers = nn.ModuleList([
            quantizer_cls(num_codebook, self.pq_dim, beta=beta, normalize=normalize, use_restart=use_restart,
                          use_weighted_sum=use_weighted_sum, need_initialized=need_initialized, pq_dropout=pq_dropout,
                          jsd_ts=jsd_ts, **extra)
            for _ in range(self.num_pq)
        ])

    def forward(self, z: torch.Tensor, it: int = 0):
        qs = list(self.quantizers)
        if all(isinstance(q, EMACodebook) for q in qs):
            return _ema_codebook_group_forward(qs, z, want_prob=self.materialize_prob)
        z_q, outputs, probs, _ = _codebook_group_forward(qs, z, want_prob=self.materialize_prob)
        if probs[0] is None:
            return z_q, outputs, None
        B, D, h, w = z.shape
        K = qs[0].num_codebook_vectors
        return z_q, outputs, torch.cat([p.reshape(B * h * w, -1) for p in probs], dim=-1)   # K columns each (fewer under pq_dropout)


class PQGOClsProductQuantizerWrapper(PQGOProductQuantizerWrapper):
    """model/dino_pqgo_cls.py:408-471: ``forward(z, it=-1)`` -> (z_q, outputs, distance_prob (B,h,w,K*M),
    pseudo_labels = [idx_i (n,)])."""

    variant = "pqgo_cls"

    def forward(self, z: torch.Tensor, it: int = -1):
        z_q, outputs, probs, idxs = _codebook_group_forward(list(self.quantizers), z, want_prob=self.materialize_prob)
        prob = torch.cat(probs, dim=-1) if probs[0] is not None else None
        return z_q, outputs, prob, [i.reshape(-1) for i in idxs]
