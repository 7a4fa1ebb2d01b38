"""Drop-in mirror of the reference's ``model/quantizer.py`` running on the sm_100a kernels.

Same class names, constructor arguments, forward signatures, output-dict keys and state-dict keys
(``quantizers.{i}.codebook.{weight,weight_avg,vq_count}``) as the reference, so wrappers, ``build.py``'s
isinstance checks and ``load_state_dict(strict=True)`` keep working.  Differences that are
visible to a caller are deliberate and small:

* all ``num_pq`` subspaces are quantised by one kernel per stage instead of a Python loop; the per-
  subspace modules still exist (``.quantizers[i]``) and can be called on their own;
* logging statistics (usage percentiles, ``codebook-usage``) come back as 0-dim tensors instead of
  Python floats, which removes ~6K host synchronisations per subspace per step (quantizer.py:23-29);
* the arithmetic always runs in fp32, also under autocast;
* ``use_gumbel`` / ``use_weighted_sum`` (stochastic / soft assignment research flags, quantizer.py:149-151,463-476)
  keep their sampling / soft sum in host PyTorch (``_host_paths.gumbel_indices``,
  ``_ema_group_forward_flags``) and use the kernels for everything else; ``F.gumbel_softmax`` is called once per
  subspace, in subspace order, on the reference's logits, so a seeded run draws the same noise.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa

from . import _pq_core as core
from . import ops
from ._host_paths import draw_restart, gumbel_indices, split_codes
from .dist_utils import all_reduce_packed_, all_reduce_tensor, packed_peer_exchange

__all__ = ["VectorQuantizer", "EMAVectorQuantizer", "EmbeddingEMA", "ProductQuantizerWrapper", "get_histogram_count"]


@torch.no_grad()
def get_histogram_count(count: torch.Tensor, prefix: str = "") -> Dict:
    """model/quantizer.py:15-30 for a single (K,) count vector (values are 0-dim tensors)."""
    return core.percentile_stats(count.reshape(1, -1), prefix)


class _RestartMixin:
    """Dead-code restart shared by both quantiser flavours (model/quantizer.py:73-103, 298-328); the draw itself
    lives in _host_paths.draw_restart (Python's `random`, so seeds behave like the reference's)."""

    update_indices = None
    update_candidates = None

    @torch.no_grad()
    def prepare_restart(self, vq_current_count: torch.Tensor, z_flat: torch.Tensor) -> None:
        self.update_indices, self.update_candidates = draw_restart(vq_current_count, z_flat)


class VectorQuantizer(nn.Module, _RestartMixin):
    """Learned-codebook quantiser, NCHW in / NCHW out (model/quantizer.py:33-200)."""

    def __init__(self, num_codebook: int, embed_dim: int, beta: float = 0.25, normalize: Optional[str] = None,
                 use_restart: bool = False, use_gumbel: bool = False, use_split: bool = False,
                 use_weighted_sum: bool = False, update_norm: bool = True, need_initialized: str = "none",
                 **_ignored) -> None:
        super().__init__()
        self.num_codebook = num_codebook
        self.embed_dim = embed_dim
        self.beta = beta
        self.normalize = normalize
        if normalize == "z_trainable":
            self.z_mean = nn.Parameter(torch.zeros(self.embed_dim))
            self.z_log_var = nn.Parameter(torch.zeros(self.embed_dim))
        self.codebook = nn.Embedding(self.num_codebook, self.embed_dim)
        nn.init.xavier_normal_(self.codebook.weight)
        self.register_buffer("vq_count", torch.zeros(self.num_codebook), persistent=False)
        self.use_restart = use_restart
        self.use_gumbel = use_gumbel
        self.use_split = use_split
        self.use_weighted_sum = use_weighted_sum
        self.update_norm = update_norm
        self.need_initialized = need_initialized
        if use_split:
            raise NotImplementedError("NOT YET implemented. Currently only for EMA.")

    @torch.no_grad()
    def restart(self) -> None:
        if (self.update_indices is not None) and (self.update_candidates is not None):
            # NB: the reference indexes a copy here (quantizer.py:99), so its restart never changes the
            # codebook; only the exact-count reset has an effect.  Kept as is.
            self.codebook.weight.data[self.update_indices].copy_(self.update_candidates)
            self.vq_count.fill_(0)
            self.update_indices = None
            self.update_candidates = None

    def forward(self, z: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, torch.Tensor], torch.Tensor]:
        return _param_group_forward([self], z)

    def extra_repr(self) -> str:
        return (f"embed_dim={self.embed_dim}, num_codebook={self.num_codebook}, normalize={self.normalize}, "
                f"use_restart={self.use_restart}, update_norm={self.update_norm}")


class EmbeddingEMA(nn.Module):
    """EMA codebook state (model/quantizer.py:203-254)."""

    def __init__(self, num_codebook: int, embed_dim: int, decay: float = 0.99, eps: float = 1e-5) -> None:
        super().__init__()
        self.decay = decay
        self.eps = eps
        self.num_codebook = num_codebook
        weight = torch.randn(num_codebook, embed_dim)
        nn.init.uniform_(weight, -1.0 / num_codebook, 1.0 / num_codebook)
        self.register_buffer("weight", weight)
        self.register_buffer("weight_avg", weight.clone())
        self.register_buffer("vq_count", torch.zeros(num_codebook))

    @torch.no_grad()
    def reset(self) -> None:
        self.weight_avg.data.copy_(self.weight.data)
        self.vq_count.fill_(0)

    def forward(self, indices: torch.Tensor) -> torch.Tensor:
        return F.embedding(indices, self.weight)

    @torch.no_grad()
    def update(self, vq_current_count, vq_current_sum) -> None:
        """EMA count / average update and renormalisation (quantizer.py:233-254) through K6."""
        K, d = self.weight.shape
        packed = torch.cat([vq_current_sum.float(), vq_current_count.float().unsqueeze(1)], dim=1).unsqueeze(0).contiguous()
        ops.ema_update(packed, self.decay, self.eps, self.vq_count.view(1, K), self.weight_avg.view(1, K, d),
                       self.weight.view(1, K, d))

    # The three pieces of update() as separate calls, for API parity (quantizer.py:241-254); forward() never uses them.
    def _decay_towards(self, state: torch.Tensor, observed: torch.Tensor) -> None:
        torch.add(state.data * self.decay, observed, alpha=1 - self.decay, out=state.data)

    def vq_count_ema_update(self, vq_current_count) -> None:
        self._decay_towards(self.vq_count, vq_current_count)

    def weight_avg_ema_update(self, vq_current_sum) -> None:
        self._decay_towards(self.weight_avg, vq_current_sum)

    def weight_update(self) -> None:
        total = self.vq_count.sum()
        laplace = (self.vq_count + self.eps) / (total + self.num_codebook * self.eps) * total
        self.weight.data.copy_(self.weight_avg / laplace[:, None])


class EMAVectorQuantizer(nn.Module, _RestartMixin):
    """EMA quantiser on flat (n, d) input (model/quantizer.py:257-551)."""

    def __init__(self, num_codebook: int, embed_dim: int, beta: float = 0.25, normalize: Optional[str] = None,
                 decay: float = 0.99, eps: float = 1e-5, use_restart: bool = False, use_gumbel: bool = False,
                 use_split: bool = False, use_weighted_sum: bool = False, update_norm: bool = True,
                 need_initialized: str = "none") -> None:
        super().__init__()
        self.num_codebook = num_codebook
        self.embed_dim = embed_dim
        self.beta = beta
        self.decay = decay
        self.normalize = normalize
        if normalize == "z_trainable":
            self.register_buffer("z_mean", torch.zeros(self.embed_dim))
            self.register_buffer("z_log_var", torch.zeros(self.embed_dim))
        self.codebook = EmbeddingEMA(self.num_codebook, self.embed_dim, decay=decay, eps=eps)
        self.register_buffer("vq_count", torch.zeros(self.num_codebook), persistent=False)
        self.use_restart = use_restart
        self.use_split = use_split
        self.use_gumbel = use_gumbel
        self.need_initialized = need_initialized
        self.use_weighted_sum = use_weighted_sum
        self.update_norm = update_norm

    @torch.no_grad()
    def restart(self) -> None:
        if (self.update_indices is not None) and (self.update_candidates is not None):
            self.codebook.weight.data[self.update_indices] = self.update_candidates
            self.codebook.reset()
            self.update_indices = None
            self.update_candidates = None

    @torch.no_grad()
    def split(self, vq_current_count: torch.Tensor) -> int:
        """Replace unused codes by perturbed copies of the most used ones (quantizer.py:330-381)."""
        n = split_codes(vq_current_count, self.codebook.vq_count, self.codebook.weight, self.codebook.weight_avg)
        if n:
            self.vq_count.fill_(0)
        return n

    @torch.no_grad()
    def _maybe_initialize(self, z_flat: torch.Tensor) -> None:
        """First-training-call codebook initialisation (quantizer.py:399-414); host-side, runs once."""
        if self.need_initialized != "none" and self.training:
            if self.need_initialized == "rand":
                self.prepare_restart(torch.zeros(self.num_codebook, dtype=torch.long, device=z_flat.device), z_flat)
                self.restart()
            elif self.need_initialized == "kmeans":
                from sklearn.cluster import KMeans
                clustering = KMeans(init="k-means++", n_clusters=self.num_codebook, random_state=0)
                clustering.fit(z_flat.detach().cpu().numpy())
                centroids = torch.from_numpy(np.array(clustering.cluster_centers_)).float().to(z_flat.device)
                self.codebook.weight.data.copy_(centroids)
                self.codebook.weight_avg.data.copy_(centroids)
            self.need_initialized = False

    def forward(self, z: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, torch.Tensor], torch.Tensor]:
        return _ema_group_forward([self], z)

    def extra_repr(self) -> str:
        return (f"embed_dim={self.embed_dim}, num_codebook={self.num_codebook}, normalize={self.normalize}, "
                f"use_split={self.use_split}, use_restart={self.use_restart}")


# ---------------------------------------------------------------------------------------------------
# fused group forwards (one call for all subspaces)
# ---------------------------------------------------------------------------------------------------
def _stack(ts: List[torch.Tensor]) -> torch.Tensor:
    """[M, ...] tensor of the per-subspace tensors WITHOUT a copy when they are consecutive slices of one
    storage (the wrapper arranges that for EMA buffers), else torch.stack."""
    t0 = ts[0]
    if len(ts) == 1:
        return t0.unsqueeze(0)
    step = t0.numel()
    base = t0.untyped_storage().data_ptr()
    if all(t.is_contiguous() and t.untyped_storage().data_ptr() == base and
           t.storage_offset() == t0.storage_offset() + i * step for i, t in enumerate(ts)):
        return torch.as_strided(t0, (len(ts),) + tuple(t0.shape), (step,) + tuple(t0.stride()), t0.storage_offset())
    return torch.stack(ts)


def _z_trainable_prelude(mods, z: torch.Tensor, d: int, training: bool):
    """Running statistics of z for normalize == "z_trainable" (quantizer.py:429-446): returns the per-channel
    (shift, scale) the kernels normalise with.  The std is taken BEFORE this step's running-statistics update, the
    mean is the buffer itself and already holds the updated value when z is normalised."""
    norm_b = torch.cat([q.z_log_var for q in mods]).exp().sqrt() + 1e-5
    if training:
        with torch.no_grad():
            # per-channel mean and mean of squares of z in ONE pass (K13) and ONE all-reduce of the [2, D] pair
            # (the reference: two reductions + two all_reduce_tensor("mean") per subspace, :433-438).  It cannot
            # ride in the packed EMA buffer: the assignment below depends on the updated mean.
            mom = all_reduce_tensor(ops.channel_moments(z), op="mean")
            logvar = (mom[1] - mom[0] * mom[0]).log()
            for i, q in enumerate(mods):
                q.z_mean.data.mul_(q.decay).add_(mom[0, i * d:(i + 1) * d], alpha=1 - q.decay)
                q.z_log_var.data.mul_(q.decay).add_(logvar[i * d:(i + 1) * d], alpha=1 - q.decay)
    return torch.cat([q.z_mean for q in mods]), norm_b


def _ema_group_forward_flags(mods: List["EMAVectorQuantizer"], z: torch.Tensor, want_prob: bool = True):
    """EMAVectorQuantizer.forward with ``use_gumbel`` (training) and / or ``use_weighted_sum`` (quantizer.py:463-476,
    483-486,534-536).  The stochastic index draw and the soft sum are host PyTorch on device tensors;
    normalisation, distances / soft assignment, gather + loss, scatter-add, EMA update and the usage statistics are
    the same kernels as the top-1 path (one launch each for all M subspaces, no fused tail)."""
    q0 = mods[0]
    M, K, mode, beta = len(mods), q0.num_codebook, q0.normalize, q0.beta
    d = z.shape[1] // M
    n = z.shape[0]
    training = q0.training
    weight = _stack([q.codebook.weight for q in mods])
    norm_a, norm_b = _z_trainable_prelude(mods, z, d, training) if mode == "z_trainable" else (None, None)
    with torch.no_grad():
        if mode in ("l2", "z_norm", "none"):
            cbn, cn2 = ops.pq_prepare_codebook(weight, mode)
        else:
            cbn = core.normalize_codebook(weight, mode, ema_style=True)
            cn2 = ops.pq_cnorm2(cbn)
        src = cbn if q0.update_norm else weight.clone()
    z32 = z if z.dtype == torch.float32 else z.float()
    grad = core._wants_grad(z)
    wsum = q0.use_weighted_sum
    prob = None
    if want_prob or wsum:                                                               # :468
        prob = (core.DistanceProb.apply(z32, cbn, cn2, mode, norm_a, norm_b, 1.0) if grad
                else ops.pq_distance_prob(z32, cbn, cn2, mode, norm_a, norm_b))
    with torch.no_grad():
        if training and q0.use_gumbel:                                                  # :463-465
            z_norm = core._normalize_rows(core._rows(z32.detach(), M), mode, norm_a, norm_b)
            idx = gumbel_indices(z_norm, cbn, 0.01)
        else:
            idx = ops.pq_assign(z32.detach(), cbn, cn2, mode, norm_a, norm_b)            # :467
    if wsum:                                                                            # :470-471
        zr = core._normalize_rows(core._rows(z32, M), mode, norm_a, norm_b)             # (n, M, d), differentiable
        zq = torch.einsum("nmk,mkd->nmd", prob.view(n, M, K), cbn)
        mse_commit = ((zr - zq.detach()) ** 2).mean(dim=(0, 2))                         # :514, per subspace
        out = zq.reshape(n, M * d)                                                      # no straight-through (:534)
    elif grad:
        out, mse_commit, _ = core.PQGatherLoss.apply(z32, src, idx, mode, norm_a, norm_b)
    else:
        out, sqerr, _ = ops.pq_gather_loss(z32, src, idx, mode, norm_a, norm_b)
        mse_commit = (sqerr / max(n * d, 1)).to(torch.float32)
    output: Dict[str, torch.Tensor] = {}
    if training:
        with torch.no_grad():
            if wsum:                                                                    # soft counts / sums (:483-488)
                p = prob.detach().view(n, M, K)
                packed = torch.cat([torch.einsum("nmk,nmd->mkd", p, core._rows(z32.detach(), M)),
                                    p.sum(dim=0).unsqueeze(-1)], dim=-1).contiguous()
                packed = all_reduce_packed_(packed)
            else:
                packed = core.ema_statistics(z, idx, K)
            count = packed[:, :, d]
            state = [[q.vq_count for q in mods], [q.codebook.vq_count for q in mods],
                     [q.codebook.weight_avg for q in mods], [q.codebook.weight for q in mods]]
            exact_c, vqc_c, wavg_c, w_c = (_stack(ts).contiguous() for ts in state)
            if q0.use_restart:                                                          # :500-501, before the update
                for i, q in enumerate(mods):
                    q.prepare_restart(count[i], z[:, i * d:(i + 1) * d])
            unused = ops.ema_update(packed, q0.codebook.decay, q0.codebook.eps, vqc_c, wavg_c, w_c, exact_c)  # :493,504
            output.update(core.percentile_stats(exact_c, "total"))
            output.update(core.percentile_stats(count, "current"))
            for ts, st in zip(state, (exact_c, vqc_c, wavg_c, w_c)):
                if st.data_ptr() != ts[0].data_ptr():
                    for i, t in enumerate(ts):
                        t.copy_(st[i])
            if q0.use_split:
                n_split = torch.tensor([float(q.split(count[i])) for i, q in enumerate(mods)], device=z.device)
            else:
                n_split = unused.float()
            output["codebook-usage"] = ((K - n_split) / K).mean()                       # :510
    commitment = mse_commit.mean()
    output["loss"] = beta * commitment                                                  # :526
    output["commitment-loss"] = commitment
    output["codebook-sum"] = torch.sum(torch.abs(_stack([q.codebook.weight for q in mods]))) / M   # :532
    return out, output, (prob if want_prob else None)


def _ema_group_forward(mods: List[EMAVectorQuantizer], z: torch.Tensor, want_prob: bool = True, stacked=None):
    """EMAVectorQuantizer.forward (quantizer.py:383-542) for M = len(mods) subspaces at once; z is flat
    (n, M*d).  ``stacked``: the wrapper's cached [M, ...] views of (exact count, EMA count, weight_avg, weight) --
    verifying on every call that the per-subspace buffers are slices of one storage costs more host time than the
    whole step's launches at M = 64."""
    q0 = mods[0]
    M, K, mode, beta = len(mods), q0.num_codebook, q0.normalize, q0.beta
    if z.dim() != 2:
        raise ValueError("EMAVectorQuantizer expects a flat (n, d) input (model/quantizer.py:396)")
    d = z.shape[1] // M
    training = q0.training
    if q0.need_initialized != "none" and training:
        for i, q in enumerate(mods):
            q._maybe_initialize(z[:, i * d:(i + 1) * d])
    if q0.use_weighted_sum or (training and q0.use_gumbel):
        return _ema_group_forward_flags(mods, z, want_prob)
    weight = stacked[3] if stacked is not None else _stack([q.codebook.weight for q in mods])
    norm_a, norm_b = _z_trainable_prelude(mods, z, d, training) if mode == "z_trainable" else (None, None)
    per_code = mode in ("l2", "z_norm", "none")
    with torch.no_grad():
        if per_code:                    # normalised codebook + |c|^2 in one launch; a fresh tensor, never an alias
            cbn, cn2 = ops.pq_prepare_codebook(weight, mode)
        else:
            cbn, cn2 = core.normalize_codebook(weight, mode, ema_style=True), None
        # the gather source must survive the in-place EMA update below until backward() has run
        src = cbn if q0.update_norm else weight.clone()
    n = z.shape[0]
    sqerr = mse_commit = None
    if core._wants_grad(z):
        idx, out, mse_commit, _, prob = core.pq_quantize(z, cbn, src, mode, norm_a, norm_b, want_prob=want_prob, cnorm2=cn2)
    else:                               # no autograd bookkeeping: K1+K3 (one kernel where the shape allows), K2 on demand
        z32 = z if z.dtype == torch.float32 else z.float()
        idx, out, sqerr = ops.pq_assign_gather(z32, cbn, src, cn2, mode, norm_a, norm_b)
        prob = ops.pq_distance_prob(z32, cbn, cn2, mode, norm_a, norm_b) if want_prob else None
    output: Dict[str, torch.Tensor] = {}
    stats = None
    if training:
        with torch.no_grad():
            # K4 + K5 (:485-491).  With more than one NCCL rank and the fused tail kernel available the statistics are
            # accumulated into symmetric memory and summed over the ranks INSIDE the tail kernel (no all-reduce call).
            peer = None
            if not q0.use_split and K <= 1024:
                peer = packed_peer_exchange((M, K, d + 1), z.device)
            if peer is not None:
                local, handle, peer_ptrs, zero_next, packed = peer.acquire()   # `packed` receives the reduced statistics
                ops.pq_accumulate(z.detach().float(), idx, K, out=local)
                handle.barrier(channel=0)          # every rank's scatter-add is complete and visible to its peers
            else:
                packed = core.ema_statistics(z, idx, K)
            count = packed[:, :, d]
            # in-place EMA update on the stacked state; zero-copy when the buffers are slices of one
            # storage (ProductQuantizerWrapper._restack), otherwise stack -> update -> write back
            state = None
            if stacked is None:
                state = [[q.vq_count for q in mods], [q.codebook.vq_count for q in mods],
                         [q.codebook.weight_avg for q in mods], [q.codebook.weight for q in mods]]
                stacked = [_stack(ts) for ts in state]
            exact_c, vqc_c, wavg_c, w_c = (t.contiguous() for t in stacked)
            if not q0.use_split:
                # EMA update + percentiles x2 + usage + codebook-sum + loss scalars: ONE launch (:493-532)
                stats = ops.pq_train_tail(packed, q0.codebook.decay, q0.codebook.eps, vqc_c, wavg_c, w_c, exact_c,
                                          sqerr, n, beta,
                                          peers=None if peer is None else (peer_ptrs.data_ptr(), peer.world),
                                          zero_next=None if peer is None else zero_next)
            if stats is None:
                unused = ops.ema_update(packed, q0.codebook.decay, q0.codebook.eps, vqc_c, wavg_c, w_c, exact_c)  # :493,504
                output.update(core.percentile_stats(exact_c, "total"))                      # :496
                output.update(core.percentile_stats(count, "current"))                      # :495
            for ts, st in zip(state or (), (exact_c, vqc_c, wavg_c, w_c)):
                if st.data_ptr() != ts[0].data_ptr():
                    for i, t in enumerate(ts):
                        t.copy_(st[i])
            if q0.use_restart:
                for i, q in enumerate(mods):
                    q.prepare_restart(count[i], z[:, i * d:(i + 1) * d])                # :500-501
            if stats is not None:
                output.update(zip(ops.TAIL_KEYS[:7], stats.unbind(0)[:7]))
            else:
                if q0.use_split:
                    n_split = torch.tensor([float(q.split(count[i])) for i, q in enumerate(mods)], device=z.device)
                else:
                    n_split = unused.float()
                output["codebook-usage"] = ((K - n_split) / K).mean()                       # :510
    if stats is not None and sqerr is not None:
        output["loss"], output["commitment-loss"], output["codebook-sum"] = stats[9], stats[8], stats[7]
        return out, output, prob
    if mse_commit is None:
        mse_commit = (sqerr / max(n * d, 1)).to(torch.float32)
    commitment = mse_commit.mean()
    output["loss"] = beta * commitment                                                  # :526
    output["commitment-loss"] = commitment
    output["codebook-sum"] = stats[7] if stats is not None else torch.sum(torch.abs(weight)) / M   # :532 (buffers updated in place)
    return out, output, prob


def _param_group_forward(mods: List[VectorQuantizer], z: torch.Tensor, want_prob: bool = True):
    """VectorQuantizer.forward (quantizer.py:105-189) for M subspaces at once; z is NCHW (B, M*d, h, w)."""
    q0 = mods[0]
    M, K, mode, beta = len(mods), q0.num_codebook, q0.normalize, q0.beta
    if z.dim() != 4:
        raise ValueError("VectorQuantizer expects a (batch, embed_dim, h, w) input (model/quantizer.py:107)")
    d = z.shape[1] // M
    codebook = torch.stack([q.codebook.weight for q in mods])                 # autograd splits the grads back
    norm_a = norm_b = z_std = None
    if mode == "z_trainable":
        norm_a = torch.cat([q.z_mean for q in mods])
        z_std = torch.stack([q.z_log_var for q in mods]).exp().sqrt()
        norm_b = z_std.reshape(-1) + 1e-5
    cbn = core.normalize_codebook(codebook, mode, ema_style=False,
                                  z_mean=torch.stack([q.z_mean for q in mods]) if mode == "z_trainable" else None,
                                  z_std=z_std)
    idx_drawn = None
    if q0.training and q0.use_gumbel:                                         # :145-147 (plain -distance, no 0.01)
        with torch.no_grad():
            z_norm = core._normalize_rows(core._rows(z.detach().float(), M), mode, norm_a, norm_b)
            idx_drawn = gumbel_indices(z_norm, cbn.detach().float(), None)
    idx, out, mse_commit, mse_cb, prob = core.pq_quantize(z, cbn, cbn, mode, norm_a, norm_b, want_prob=want_prob,
                                                          idx=idx_drawn)
    output: Dict[str, torch.Tensor] = {}
    with torch.no_grad():                                                     # counts are updated in eval too (:158)
        packed = ops.pq_accumulate(z.detach().float(), idx, K)
        count = all_reduce_tensor(packed[:, :, d].contiguous(), op="sum")
        for i, q in enumerate(mods):
            q.vq_count += count[i]
        output.update(core.percentile_stats(torch.stack([q.vq_count for q in mods]), "total"))
        output.update(core.percentile_stats(count, "current"))
        if q0.use_restart:
            zf = z.detach().permute(0, 2, 3, 1).reshape(-1, z.shape[1])
            for i, q in enumerate(mods):
                q.prepare_restart(count[i], zf[:, i * d:(i + 1) * d])
    codebook_loss, commitment_loss = mse_cb.mean(), mse_commit.mean()
    output["loss"] = codebook_loss + beta * commitment_loss                   # :177
    output["codebook_loss"] = codebook_loss
    output["commitment_loss"] = commitment_loss
    return out, output, prob


class ProductQuantizerWrapper(nn.Module):
    """model/quantizer.py:554-614.  ``materialize_prob`` (default True, as the reference always returns
    it) controls whether the N x (K*num_pq) soft-assignment tensor is produced; set it to False when the
    caller ignores the third return value (e.g. DIONPQGO, model/dino_pqgo.py:140-154)."""

    def __init__(self, num_pq: int, num_codebook: int, embed_dim: int, beta: float = 0.25,
                 normalize: Optional[str] = None, decay: float = 0.99, eps: float = 1e-5, use_restart: bool = False,
                 use_gumbel: bool = False, use_split: bool = False, use_weighted_sum: bool = False,
                 update_norm: bool = True, need_initialized: str = "none", quantizer_cls=EMAVectorQuantizer) -> None:
        super().__init__()
        if embed_dim % num_pq != 0:
            raise ValueError(f"Embed dim {embed_dim} should be divisible by #PQ {num_pq}.")
        self.num_pq = num_pq
        self.pq_dim = embed_dim // num_pq
        self.materialize_prob = True
        self.quantizers = nn.ModuleList([
            quantizer_cls(num_codebook, self.pq_dim, beta=beta, normalize=normalize, decay=decay, eps=eps,
                          use_restart=use_restart, use_gumbel=use_gumbel, use_split=use_split,
                          use_weighted_sum=use_weighted_sum, update_norm=update_norm, need_initialized=need_initialized)
            for _ in range(self.num_pq)
        ])
        self._restack()

    # EMA buffers of all subspaces live in one [M, ...] storage each; the per-subspace buffers the state
    # dict exposes are views into it, so the kernels update every subspace in place with one launch.
    def _restack(self) -> None:
        qs = list(self.quantizers)
        self._stacked = None
        if not qs or not all(isinstance(q, EMAVectorQuantizer) for q in qs):
            return
        cache = {}
        with torch.no_grad():
            for key, owner, name in (("weight", lambda q: q.codebook, "weight"), ("weight_avg", lambda q: q.codebook, "weight_avg"),
                                     ("ema_count", lambda q: q.codebook, "vq_count"), ("exact", lambda q: q, "vq_count")):
                stacked = torch.stack([owner(q)._buffers[name] for q in qs]).contiguous()
                for i, q in enumerate(qs):
                    owner(q)._buffers[name] = stacked[i]
                cache[key] = stacked
        # order expected by _ema_group_forward; dropped (-> re-derived per call) if a buffer is ever re-assigned
        self._stacked = (cache["exact"], cache["ema_count"], cache["weight_avg"], cache["weight"])

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._restack()
        return out

    def forward(self, z: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, torch.Tensor], torch.Tensor]:
        qs = list(self.quantizers)
        st = self._stacked
        if st is not None:
            # the cached views stay valid as long as nobody re-assigned a buffer (load_state_dict / .to() keep or rebuild them)
            q0, qm = qs[0], qs[-1]
            if q0.codebook.weight.data_ptr() == st[3].data_ptr() and qm.codebook.weight.data_ptr() == st[3][-1].data_ptr() \
                    and q0.vq_count.data_ptr() == st[0].data_ptr():
                return _ema_group_forward(qs, z, want_prob=self.materialize_prob, stacked=st)
            self._restack()
            return _ema_group_forward(qs, z, want_prob=self.materialize_prob, stacked=self._stacked)
        if all(isinstance(q, EMAVectorQuantizer) for q in qs):
            return _ema_group_forward(qs, z, want_prob=self.materialize_prob)
        if all(isinstance(q, VectorQuantizer) for q in qs):
            return _param_group_forward(qs, z, want_prob=self.materialize_prob)
        # heterogeneous / user-supplied quantiser classes: the reference's loop (quantizer.py:589-611)
        z_split = torch.chunk(z, chunks=self.num_pq, dim=1)
        z_quantized, distance_prob, outputs = [], [], {}
        for i in range(self.num_pq):
            q_i, output_i, prob_i = self.quantizers[i](z_split[i])
            z_quantized.append(q_i)
            for k, v in output_i.items():
                outputs[k] = v if i == 0 else outputs[k] + v
            distance_prob.append(prob_i)
        for k in outputs:
            outputs[k] = outputs[k] / self.num_pq
        return torch.cat(z_quantized, dim=1), outputs, torch.cat(distance_prob, dim=-1)

    def extra_repr(self) -> str:
        return f"num_pq={self.num_pq}"
