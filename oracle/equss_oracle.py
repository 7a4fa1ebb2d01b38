"""CPU oracle for the EQUSS product-quantization hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain PyTorch-on-CPU restatement of the reference's algorithm for the path in SURVEY.md section 8
(the reference is pure Python/PyTorch, so the restatement uses the same ATen CPU primitives the
reference calls: matmul, argmin, softmax, embedding, one_hot, interpolate, einsum, bincount, topk --
torch 2.11.0 in this image; the reference pins no versions).  Every function cites the reference lines
it follows (paths relative to the reference root).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg
may import this module, and only as the checker or the timed CPU baseline -- never as a product path.

Parity pinning: the reference ships no tests, golden vectors or fixtures for this path, so
this oracle is pinned against outputs of the reference ITSELF: ``oracle/make_golden.py`` imports the
reference modules from /root/reference (with stubs for the off-path torchmetrics / pydensecrf imports),
runs them on seeded inputs, asserts this oracle reproduces them bit for bit, and writes the vectors to
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` re-checks the oracle against those files.
``data/precompute_knns.py`` cannot be imported (hydra / pytorch_lightning are absent): ``oracle/make_golden_knn.py``
executes its kNN statements (:307-317) straight out of the reference file's AST and pins ``knn`` below to their result.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------
# normalisation  (model/quantizer.py:419-455)
# --------------------------------------------------------------------------------------------------


def normalize_pair(z_flat: torch.Tensor, codebook: torch.Tensor, mode: Optional[str],
                   z_mean: Optional[torch.Tensor] = None, z_log_var: Optional[torch.Tensor] = None,
                   ema_style: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (z_norm, codebook_norm).  ``ema_style`` selects the z_trainable flavour of
    EMAVectorQuantizer (:428-450, codebook standardised over dim 0) versus VectorQuantizer (:129-133)."""
    if mode == "l2":
        return F.normalize(z_flat, dim=1), F.normalize(codebook, dim=1)          # :420-421
    if mode == "z_norm":
        zs, zm = torch.std_mean(z_flat, dim=1, keepdim=True)                      # :423
        cs, cm = torch.std_mean(codebook, dim=1, keepdim=True)                    # :426
        return (z_flat - zm) / (zs + 1e-5), (codebook - cm) / (cs + 1e-5)
    if mode == "z_trainable":
        std = z_log_var.exp().sqrt()                                              # :430
        zn = (z_flat - z_mean) / (std + 1e-5)                                     # :446
        if ema_style:
            cs, cm = torch.std_mean(codebook, dim=0)                              # :449
            return zn, (codebook - cm) / (cs + 1e-5)
        return zn, (codebook - z_mean) / (std + 1e-5)                             # :133
    if mode == "none":
        return z_flat, codebook
    raise ValueError(f"Unsupported normalize type {mode}")                        # :455


def sq_distance(z_norm: torch.Tensor, codebook_norm: torch.Tensor) -> torch.Tensor:
    """(n, K) squared L2 distance in the reference's association order (model/quantizer.py:457-461)."""
    return (torch.sum(z_norm ** 2, dim=1, keepdim=True)
            + torch.sum(codebook_norm ** 2, dim=1)
            - 2 * torch.matmul(z_norm, codebook_norm.t()))


def histogram_percentiles(count: torch.Tensor, prefix: str) -> Dict[str, Optional[float]]:
    """Usage percentiles p10/p50/p90 (model/quantizer.py:15-30): first rank whose cumulative sorted
    probability reaches the level, divided by K; None if never reached."""
    prob = count.float() / (count.sum() + 1)
    K = prob.numel()
    csum = torch.cumsum(torch.sort(prob, dim=0, descending=True)[0], dim=0)
    out: Dict[str, Optional[float]] = {}
    for tag, level in (("p10", 0.1), ("p50", 0.5), ("p90", 0.9)):
        hit = torch.nonzero(csum >= level)
        out[f"{prefix}-{tag}"] = (int(hit[0, 0]) / K) if hit.numel() else None
    return out


# --------------------------------------------------------------------------------------------------
# EMA state  (model/quantizer.py:203-254)
# --------------------------------------------------------------------------------------------------


class EmaState:
    """weight / weight_avg / vq_count of one EmbeddingEMA (model/quantizer.py:215-221)."""

    def __init__(self, weight: torch.Tensor, weight_avg: Optional[torch.Tensor] = None,
                 vq_count: Optional[torch.Tensor] = None, decay: float = 0.99, eps: float = 1e-5):
        self.weight = weight.clone()
        self.weight_avg = weight.clone() if weight_avg is None else weight_avg.clone()
        self.vq_count = torch.zeros(weight.shape[0]) if vq_count is None else vq_count.clone()
        self.decay, self.eps = decay, eps

    def update(self, count: torch.Tensor, total: torch.Tensor) -> None:
        K = self.weight.shape[0]
        self.vq_count.mul_(self.decay).add_(count, alpha=1 - self.decay)          # :242
        self.weight_avg.mul_(self.decay).add_(total, alpha=1 - self.decay)        # :245
        n = self.vq_count.sum()                                                   # :248
        smoothed = (self.vq_count + self.eps) / (n + K * self.eps) * n            # :249-251
        self.weight.copy_(self.weight_avg / smoothed.unsqueeze(1))                # :253-254


def ema_vq_forward(z_flat: torch.Tensor, state: EmaState, exact_count: torch.Tensor, *, normalize: Optional[str],
                   beta: float = 0.25, training: bool = True, update_norm: bool = True,
                   allreduce=None, allreduce_mean=None, z_mean: Optional[torch.Tensor] = None,
                   z_log_var: Optional[torch.Tensor] = None, temperature: float = 1.0, ema_decay: float = 0.99,
                   use_weighted_sum: bool = False, gumbel_noise: Optional[torch.Tensor] = None
                   ) -> Tuple[torch.Tensor, Dict, torch.Tensor, torch.Tensor]:
    """One subspace of EMAVectorQuantizer.forward (model/quantizer.py:383-542; all four normalisation modes, no
    restart/split).  ``use_weighted_sum`` (:470-471,483-484,534): soft-assignment-weighted sum of the codes, soft
    statistics, no straight-through.  ``gumbel_noise`` (n, K), training only: the use_gumbel draw (:463-465) with the
    noise given -- F.gumbel_softmax(l, tau=1, hard=True) is one_hot(argmax(l + g)) with g = -log(Exp(1)).  ``allreduce`` stands for all_reduce_tensor(:490-491),
    ``allreduce_mean`` for the op="mean" calls of the z_trainable statistics (:437-438); ``z_mean`` /
    ``z_log_var`` are that mode's buffers, updated in place in training.
    Returns (z_q_ste, outputs, distance_prob, indices)."""
    if normalize == "z_trainable":
        # NB (:429-446): the std is evaluated BEFORE the running statistics are updated, the mean is the
        # buffer itself and therefore already holds the updated value when z is normalised.
        std_before = z_log_var.exp().sqrt()                                       # :430
        if training:                                                              # :432-444 running stats of z
            m0 = torch.mean(z_flat, dim=0)
            sq0 = torch.mean(z_flat * z_flat, dim=0)
            if allreduce_mean is not None:
                m0, sq0 = allreduce_mean(m0), allreduce_mean(sq0)
            z_mean.mul_(ema_decay).add_(m0, alpha=1 - ema_decay)
            z_log_var.mul_(ema_decay).add_((sq0 - m0 * m0).log(), alpha=1 - ema_decay)
        z_norm = (z_flat - z_mean) / (std_before + 1e-5)                          # :446
        cs, cm = torch.std_mean(state.weight, dim=0)                              # :449
        cb_norm = (state.weight - cm) / (cs + 1e-5)
    else:
        z_norm, cb_norm = normalize_pair(z_flat, state.weight, normalize)
    dist = sq_distance(z_norm, cb_norm)
    if training and gumbel_noise is not None:
        idx = torch.argmax(-dist / 0.01 + gumbel_noise, dim=1)                    # :463-465
    else:
        idx = torch.argmin(dist, dim=1)                                           # :467
    prob = F.softmax(-dist / temperature, dim=1)                                  # :468 / dino_new_vq.py:398
    src = cb_norm if update_norm else state.weight                                # :473-476
    q = torch.matmul(prob, cb_norm) if use_weighted_sum else F.embedding(idx, src)   # :470-476
    out: Dict = {}
    if training:
        K = state.weight.shape[0]
        onehot = prob if use_weighted_sum else F.one_hot(idx, K).to(z_flat.dtype)   # :483-485
        count = onehot.sum(dim=0)                                                 # :487
        total = torch.matmul(onehot.t(), z_flat)                                  # :488 (raw z, not z_norm)
        if allreduce is not None:
            count, total = allreduce(count), allreduce(total)
        exact_count += count                                                      # :493
        out.update(histogram_percentiles(exact_count, "total"))                   # :496
        out.update(histogram_percentiles(count, "current"))                       # :495
        state.update(count, total)                                                # :504
        unused = int((count == 0).sum())                                          # :509
        out["codebook-usage"] = (K - unused) / K
    commitment = F.mse_loss(z_norm, q)                                            # :514
    out["loss"] = beta * commitment                                               # :526
    out["commitment-loss"] = commitment
    out["codebook-sum"] = torch.sum(torch.abs(state.weight))                      # :532
    q_ste = q if use_weighted_sum else z_norm + (q - z_norm)                      # :534-536 (value of the STE)
    return q_ste, out, prob, idx


def pq_forward_ema(z: torch.Tensor, states: Sequence[EmaState], exact_counts: Sequence[torch.Tensor], **kw
                   ) -> Tuple[torch.Tensor, Dict, torch.Tensor, torch.Tensor]:
    """ProductQuantizerWrapper.forward over flat (n, D) input (model/quantizer.py:587-611).
    Returns (z_q, outputs averaged over subspaces, distance_prob (n, K*M), indices (M, n))."""
    M = len(states)
    chunks = torch.chunk(z, chunks=M, dim=1)                                      # :589
    qs, probs, idxs = [], [], []
    outputs: Dict = {}
    for i in range(M):
        q, o, p, ix = ema_vq_forward(chunks[i], states[i], exact_counts[i], **kw)
        qs.append(q); probs.append(p); idxs.append(ix)
        for k, v in o.items():
            outputs[k] = v if i == 0 else outputs[k] + v                          # :598-603
    for k in outputs:
        outputs[k] = outputs[k] / M                                               # :607-608
    return torch.cat(qs, dim=1), outputs, torch.cat(probs, dim=-1), torch.stack(idxs)


def param_vq_forward(z_nchw: torch.Tensor, codebook: torch.Tensor, *, normalize: Optional[str], beta: float = 0.25,
                     book: float = 1.0, gather_raw: bool = False, temperature: float = 1.0,
                     gumbel_noise: Optional[torch.Tensor] = None
                     ) -> Tuple[torch.Tensor, Dict, torch.Tensor, torch.Tensor]:
    """One subspace of the learned-codebook quantiser on NCHW input: VectorQuantizer.forward
    (model/quantizer.py:105-189) when ``gather_raw`` is False, dino_pqgo.Codebook.forward
    (model/dino_pqgo.py:579-705) when True (gathers the raw embedding, `book` weight, softmax / jsd_ts)."""
    b, d, h, w = z_nchw.shape
    z_flat = z_nchw.permute(0, 2, 3, 1).contiguous().view(-1, d)                  # :112-115
    z_norm, cb_norm = normalize_pair(z_flat, codebook, normalize, ema_style=False)
    dist = sq_distance(z_norm, cb_norm)
    if gumbel_noise is not None:
        idx = torch.argmax(-dist + gumbel_noise, dim=1)                           # :145-147 (training, use_gumbel)
    else:
        idx = torch.argmin(dist, dim=1)                                           # :150
    prob = F.softmax(-dist / temperature, dim=1)                                  # :151 / dino_pqgo.py:655
    q = F.embedding(idx, codebook if gather_raw else cb_norm)                     # :153 / dino_pqgo.py:665
    codebook_loss = F.mse_loss(q, z_norm)                                         # :175
    commitment = F.mse_loss(z_norm, q)                                            # :176
    out = {"loss": book * codebook_loss + beta * commitment, "codebook_loss": codebook_loss,
           "commitment_loss": commitment}
    q_ste = z_norm + (q - z_norm)                                                 # :184
    return q_ste.view(b, h, w, d).permute(0, 3, 1, 2).contiguous(), out, prob, idx


def v2_ema_vq_forward(z_nchw: torch.Tensor, embeddings: torch.Tensor, beta: float = 0.25
                      ) -> Tuple[torch.Tensor, Dict, torch.Tensor, torch.Tensor]:
    """Eval-mode quantizer_v2.EMAVectorQuantizer.forward (model/quantizer_v2.py:253-308): always l2, and
    -- a quirk kept for parity -- the output rows are gathered from z_norm, not the codebook (:274)."""
    b, d, h, w = z_nchw.shape
    flat = z_nchw.permute(0, 2, 3, 1).contiguous().view(-1, d)
    z_norm, cb_norm = F.normalize(flat, dim=1), F.normalize(embeddings, dim=1)
    dist = sq_distance(z_norm, cb_norm)
    prob = F.softmax(-dist * 1.0, dim=1)
    idx = torch.argmin(dist, dim=1)
    emb = F.embedding(idx, z_norm)                                                # :274
    commitment = F.mse_loss(z_norm, emb)
    out = {"commitment-loss": commitment, "loss": beta * commitment, "codebook-sum": torch.sum(torch.abs(embeddings))}
    return emb.view(b, h, w, -1).permute(0, 3, 1, 2).contiguous(), out, prob, idx


def v2_ema_vq_train_step(z_nchw: torch.Tensor, embeddings: torch.Tensor, N: torch.Tensor, z_avg: torch.Tensor,
                         beta: float = 0.25, decay: float = 0.99, eps: float = 1e-5
                         ) -> Tuple[torch.Tensor, Dict, torch.Tensor, torch.Tensor]:
    """Training-mode quantizer_v2.EMAVectorQuantizer.forward (model/quantizer_v2.py:253-308): as the eval
    function above, then the EMA update from sums of **z_norm** (:279-281); the two all_reduce_tensor calls
    discard their result (:283-284), so the statistics are per rank.  Buffers are updated in place; the
    returned ``codebook-sum`` is evaluated after the update (:304)."""
    b, d, h, w = z_nchw.shape
    K = embeddings.shape[0]
    flat = z_nchw.permute(0, 2, 3, 1).contiguous().view(-1, d)
    z_norm, cb_norm = F.normalize(flat, dim=1), F.normalize(embeddings, dim=1)
    dist = sq_distance(z_norm, cb_norm)
    prob = F.softmax(-dist * 1.0, dim=1)
    idx = torch.argmin(dist, dim=1)
    emb = F.embedding(idx, z_norm)                                                # :274
    onehot = F.one_hot(idx, K).to(z_nchw.dtype)
    N.mul_(decay).add_(onehot.sum(dim=0), alpha=1 - decay)                        # :286
    z_avg.mul_(decay).add_(torch.matmul(onehot.t(), z_norm), alpha=1 - decay)     # :287
    n = N.sum()
    weights = (N + eps) / (n + K * eps) * n                                       # :290
    embeddings.copy_(z_avg / weights.unsqueeze(1))                                # :291-292
    commitment = F.mse_loss(z_norm, emb)
    out = {"commitment-loss": commitment, "loss": beta * commitment, "codebook-sum": torch.sum(torch.abs(embeddings))}
    return emb.view(b, h, w, -1).permute(0, 3, 1, 2).contiguous(), out, prob, idx


# --------------------------------------------------------------------------------------------------
# inline variants V4 / V6  (model/dino_new_vq.py:241-671, model/dino_pqgo_cls.py:191-405) and the two
# distance_prob consumers they call (model/loss.py:490-525)
# --------------------------------------------------------------------------------------------------


def entropy_loss(p: torch.Tensor) -> torch.Tensor:
    """EntropyLoss.forward (model/loss.py:490-505): MINUS the entropy of the batch-averaged assignment."""
    avg_p = p.mean(0)
    return -torch.sum(-avg_p * torch.log(avg_p + 1e-8), dim=-1)


def jsd_loss(p: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """JSDLoss.forward (model/loss.py:508-525): KLDivLoss(batchmean, log_target) of both inputs against the
    smoothed mixture."""
    m = (0.5 * (p + q).add(1e-6)).log()
    lp, lq = p.add(1e-6).log(), q.add(1e-6).log()
    n = p.shape[0]
    kl = lambda a, t: (t.exp() * (t - a)).sum() / n     # noqa: E731  F.kl_div(a, t, "batchmean", log_target=True)
    return 0.5 * (kl(m, lp) + kl(m, lq))


def new_vq_ema_forward(z_nchw: torch.Tensor, state: EmaState, exact_count: torch.Tensor, *, normalize: str,
                       beta: float = 0.25, jsd_ts: float = 1.0, training: bool = True,
                       z_mean: Optional[torch.Tensor] = None, z_log_var: Optional[torch.Tensor] = None,
                       allreduce=None, use_weighted_sum: bool = False, dropout_keep: Optional[torch.Tensor] = None
                       ) -> Tuple[torch.Tensor, Dict, torch.Tensor, torch.Tensor]:
    """dino_new_vq.EMACodebook.forward (model/dino_new_vq.py:327-459), top-1 path without init:
    NCHW in, the quantised rows come from the RAW codebook as it was before this step's update (:403),
    EMA sums of raw z (:411), softmax(-d / jsd_ts), JSD / entropy between the two batch halves (:447-450).
    ``dropout_keep``: the boolean keep mask of pq_dropout (:388-391, drawn in training AND evaluation): distances,
    argmin and soft assignment run over the kept codes only, while the indices -- positions in the KEPT list --
    still address the full codebook in the gather (:403), the one-hot (:408) and the EMA update; the usage ratio
    is taken over the kept codes (:431-432)."""
    b, d, h, w = z_nchw.shape
    K = state.weight.shape[0]
    z_flat = z_nchw.permute(0, 2, 3, 1).contiguous().view(-1, d)                  # :333-334
    z_norm, cb_norm = normalize_pair(z_flat, state.weight, normalize, z_mean, z_log_var)   # :366-387
    if dropout_keep is not None:
        cb_norm = cb_norm[dropout_keep]                                           # :390-391
    kept = cb_norm.shape[0]
    dist = sq_distance(z_norm, cb_norm)
    idx = torch.argmin(dist, dim=1)
    prob = F.softmax(-dist / jsd_ts, dim=1)                                       # :398
    if use_weighted_sum:
        q = torch.matmul(prob, cb_norm)                                           # :400-401 (hard statistics below)
    else:
        q = F.embedding(idx, state.weight)                                        # :403 (a copy: pre-update rows)
    out: Dict = {}
    if training:
        onehot = F.one_hot(idx, K).to(z_flat.dtype)
        count, total = onehot.sum(dim=0), torch.matmul(onehot.t(), z_flat)        # :410-411
        if allreduce is not None:
            count, total = allreduce(count), allreduce(total)
        exact_count += count                                                      # :415
        state.update(count, total)                                                # :422
        out["codebook-usage"] = (kept - int((count == 0).sum())) / kept           # :431-432
    commitment = F.mse_loss(z_norm, q)
    out["vq-loss"] = beta * commitment                                            # :435-436
    out["codebook-sum"] = torch.sum(torch.abs(state.weight))                      # :445 (after the update)
    p1, p2 = torch.chunk(prob, chunks=2, dim=0)
    out["jsd"] = jsd_loss(p1, p2)                                                 # :449
    out["entropy"] = entropy_loss(p1)                                             # :450
    q_ste = q if use_weighted_sum else z_norm + (q - z_norm)                      # :438-439
    return q_ste.view(b, h, w, d).permute(0, 3, 1, 2).contiguous(), out, prob, idx


def inline_codebook_forward(z_nchw: torch.Tensor, codebook: torch.Tensor, exact_count: torch.Tensor, *,
                            variant: str, normalize: str, beta: float = 0.25, book: float = 1.0,
                            jsd_ts: float = 1.0, training: bool = True, z_mean: Optional[torch.Tensor] = None,
                            z_log_var: Optional[torch.Tensor] = None, use_weighted_sum: bool = False,
                            dropout_keep: Optional[torch.Tensor] = None
                            ) -> Tuple[torch.Tensor, Dict, torch.Tensor, torch.Tensor]:
    """The learned-codebook ``Codebook.forward`` of dino_new_vq.py:537-671 (variant "new_vq"), dino_pqgo.py:579-705
    ("pqgo") and dino_pqgo_cls.py:303-405 ("pqgo_cls"): raw embedding gathered, counts in training only,
    ``vq-loss`` = [book *] codebook + beta * commitment; new_vq adds jsd / entropy.  prob is returned flat (n, K);
    the callers reshape it (pqgo / pqgo_cls: (b, h, w, K)).  ``dropout_keep``: pq_dropout's keep mask
    (dino_new_vq.py:600-603, dino_pqgo.py:641-644), semantics as in new_vq_ema_forward."""
    b, d, h, w = z_nchw.shape
    K = codebook.shape[0]
    z_flat = z_nchw.permute(0, 2, 3, 1).contiguous().view(-1, d)
    z_norm, cb_norm = normalize_pair(z_flat, codebook, normalize, z_mean, z_log_var)
    if dropout_keep is not None:
        cb_norm = cb_norm[dropout_keep]
    kept = cb_norm.shape[0]
    dist = sq_distance(z_norm, cb_norm)
    idx = torch.argmin(dist, dim=1)
    prob = F.softmax(-dist / jsd_ts, dim=1)
    q = torch.matmul(prob, cb_norm) if use_weighted_sum else F.embedding(idx, codebook)   # dino_pqgo.py:658-665
    out: Dict = {}
    if training:
        count = F.one_hot(idx, K).to(z_flat.dtype).sum(dim=0)
        exact_count += count
        out["codebook-usage"] = (kept - int((count == 0).sum())) / kept
    cb_loss, commit = F.mse_loss(q, z_norm), F.mse_loss(z_norm, q)
    out["vq-loss"] = (book if variant == "pqgo" else 1.0) * cb_loss + beta * commit
    if variant == "new_vq":
        p1, p2 = torch.chunk(prob, chunks=2, dim=0)
        out["jsd"] = jsd_loss(p1, p2)
        out["entropy"] = entropy_loss(p1)
    q_ste = q if use_weighted_sum else z_norm + (q - z_norm)                      # dino_pqgo.py:690-691
    return q_ste.view(b, h, w, d).permute(0, 3, 1, 2).contiguous(), out, prob, idx


def restart_candidates(count: torch.Tensor, z_rows: torch.Tensor, rng) -> Tuple[torch.Tensor, torch.Tensor]:
    """prepare_restart (model/quantizer.py:298-319 and its inline twins): dead codes and the rows of z that
    replace them, drawn with Python's ``random`` (``rng`` is a random.Random or the module)."""
    n_data = z_rows.shape[0]
    dead = torch.nonzero(count == 0, as_tuple=True)[0]
    z_indices = list(range(n_data))
    rng.shuffle(z_indices)
    if len(dead) <= n_data:
        z_indices = z_indices[:len(dead)]
    else:
        dead = dead.tolist()
        rng.shuffle(dead)
        dead = dead[:n_data]
    return dead, z_rows[z_indices]


# --------------------------------------------------------------------------------------------------
# evaluation  (model/evaluator.py:46-111, model/metric.py:44-125)
# --------------------------------------------------------------------------------------------------


def evaluator_forward(out: torch.Tensor, label: torch.Tensor, clusters: torch.Tensor, lin_w: torch.Tensor,
                      lin_b: torch.Tensor, num_classes: int
                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """UnSegEvaluator.forward, non-CRF branch (model/evaluator.py:46-82) with ClusterLookup alpha=None
    (:93-111).  Returns (linear_loss, linear_preds, cluster_loss, cluster_preds)."""
    if out.shape[-2:] != label.shape[-2:]:
        out = F.interpolate(out, label.shape[-2:], mode="bilinear", align_corners=False)   # :53-54
    lin_logits = F.conv2d(out, lin_w.view(lin_w.shape[0], -1, 1, 1), lin_b)                # :67
    lin_preds = lin_logits.argmax(1)                                                       # :68
    nc = F.normalize(clusters, dim=1)                                                      # :95
    nf = F.normalize(out, dim=1)                                                           # :96
    inner = torch.einsum("bchw,nc->bnhw", nf, nc)                                          # :98
    probs = F.one_hot(torch.argmax(inner, dim=1), clusters.shape[0]).permute(0, 3, 1, 2).to(torch.float32)
    cluster_loss = -(probs * inner).sum(1).mean()                                          # :106
    cluster_preds = probs.argmax(1)                                                        # :70
    lab = label.reshape(-1)
    mask = (lab >= 0) & (lab < num_classes)                                                # :73
    flat = lin_logits.permute(0, 2, 3, 1).reshape(-1, num_classes)                         # :76
    lin_loss = F.cross_entropy(flat[mask], lab[mask])                                      # :80
    return lin_loss, lin_preds, cluster_loss, cluster_preds


def confusion_update(confusion: torch.Tensor, preds: torch.Tensor, label: torch.Tensor, num_classes: int,
                     extra_classes: int = 0) -> torch.Tensor:
    """UnSegMetrics.update (model/metric.py:44-58): rows = prediction, cols = label; predictions >=
    num_classes are dropped even when extra classes exist (:49)."""
    p, l = preds.reshape(-1), label.reshape(-1)
    m = (l >= 0) & (l < num_classes) & (p >= 0) & (p < num_classes)
    rows = num_classes + extra_classes
    binc = torch.bincount(l[m] * rows + p[m], minlength=num_classes * rows)
    return confusion + binc.reshape(num_classes, rows).t()


def metrics_compute(confusion: torch.Tensor, hungarian: bool) -> Dict[str, torch.Tensor]:
    """UnSegMetrics.compute for extra_classes == 0 (model/metric.py:60-98) without the CSV side effect."""
    from scipy.optimize import linear_sum_assignment
    if hungarian:
        assign = linear_sum_assignment(confusion.cpu(), maximize=True)                     # :66
        hist = confusion[np.argsort(assign[1]), :]                                         # :72
    else:
        hist = confusion
    tp = torch.diag(hist)
    fp = hist.sum(0) - tp
    fn = hist.sum(1) - tp
    iou = tp / (tp + fp + fn)
    iou = iou[~torch.isnan(iou)].mean()
    accuracy = tp.sum() / hist.sum()                                                      # :94
    return {"iou": 100 * iou, "accuracy": 100 * accuracy}


# --------------------------------------------------------------------------------------------------
# kNN  (data/precompute_knns.py:165-171, 305-319)
# --------------------------------------------------------------------------------------------------


def knn(normed_feats: torch.Tensor, k: int = 30, queries: Optional[torch.Tensor] = None
        ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Indices (and values) of the k largest cosine similarities per query (precompute_knns.py:313-315)."""
    q = normed_feats if queries is None else queries
    sims = torch.einsum("nf,mf->nm", q, normed_feats)
    vals, idx = torch.topk(sims, k)
    return idx, vals


# --------------------------------------------------------------------------------------------------
# Expansion head  (model/dino_pqgo.py:104-112,127-128; model/blocks/module.py:20-44)
# --------------------------------------------------------------------------------------------------


def expansion_head(x: torch.Tensor, w1: torch.Tensor, b1: Optional[torch.Tensor], w2: torch.Tensor,
                   b2: Optional[torch.Tensor], w3: torch.Tensor, b3: Optional[torch.Tensor]) -> torch.Tensor:
    """code = cluster1(x) + cluster2(x): Conv1x1(C->D) plus Conv1x1(C->C), ReLU, Conv1x1(C->D), in the dtype of x
    (module.py:39-43: ``out = self.cluster1(x); out += self.cluster2(x)``).  Weights are Conv2d-shaped (o, c, 1, 1)."""
    out = F.conv2d(x, w1, b1)
    out += F.conv2d(F.relu(F.conv2d(x, w2, b2)), w3, b3)
    return out


# --------------------------------------------------------------------------------------------------
# STEGO correspondence loss  (model/loss.py:647-739)
# --------------------------------------------------------------------------------------------------


def stego_helper(f1: torch.Tensor, f2: torch.Tensor, c1: torch.Tensor, c2: torch.Tensor, shift: float, cfg: Dict
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """STEGOLoss.helper (model/loss.py:677-700): feature correlation (no grad, optionally row-centred and re-centred on
    its original mean), code correlation, ``-clamp(cd) * (fd - shift)``."""
    with torch.no_grad():
        fd = torch.einsum("nchw,ncij->nhwij", F.normalize(f1, dim=1, eps=1e-10), F.normalize(f2, dim=1, eps=1e-10))
        if cfg["pointwise"]:
            old_mean = fd.mean()
            fd -= fd.mean([3, 4], keepdim=True)
            fd = fd - fd.mean() + old_mean
    cd = torch.einsum("nchw,ncij->nhwij", F.normalize(c1, dim=1, eps=1e-10), F.normalize(c2, dim=1, eps=1e-10))
    min_val = 0.0 if cfg["zero_clamp"] else -9999.0
    if cfg["stabilize"]:
        loss = -cd.clamp(min_val, .8) * (fd - shift)
    else:
        loss = -cd.clamp(min_val) * (fd - shift)
    return loss, cd


def stego_forward(cfg: Dict, feats: torch.Tensor, feats_pos: torch.Tensor, code: torch.Tensor, code_pos: torch.Tensor
                  ) -> torch.Tensor:
    """STEGOLoss.forward (model/loss.py:702-739), drawing its random numbers in the reference's order on the tensors'
    device (two coordinate sets, then one fixed-point-free permutation per negative sample)."""
    def samp(t, coords):
        return F.grid_sample(t, coords.permute(0, 2, 1, 3), padding_mode="border", align_corners=True)
    n, S, dev = feats.shape[0], cfg["feature_samples"], feats.device
    coords1 = torch.rand([n, S, S, 2], device=dev) * 2 - 1
    coords2 = torch.rand([n, S, S, 2], device=dev) * 2 - 1
    f, c = samp(feats, coords1), samp(code, coords1)
    fp, cp = samp(feats_pos, coords2), samp(code_pos, coords2)
    intra, _ = stego_helper(f, f, c, c, cfg["pos_intra_shift"], cfg)
    inter, _ = stego_helper(f, fp, c, cp, cfg["pos_inter_shift"], cfg)
    negs = []
    for _ in range(cfg["neg_samples"]):
        perm = torch.randperm(n, device=dev, dtype=torch.long)
        perm[perm == torch.arange(n, device=dev)] += 1
        perm = perm % n
        neg, _ = stego_helper(f, samp(feats[perm], coords2), c, samp(code[perm], coords2), cfg["neg_inter_shift"], cfg)
        negs.append(neg)
    return (cfg["pos_intra_weight"] * intra.mean() + cfg["pos_inter_weight"] * inter.mean() +
            cfg["neg_inter_weight"] * torch.cat(negs, dim=0).mean())


# --------------------------------------------------------------------------------------------------
# near-tie audit helpers used by the parity tests
# --------------------------------------------------------------------------------------------------


def top2_margin_fp64(z_norm: torch.Tensor, cb_norm: torch.Tensor, idx_a: torch.Tensor, idx_b: torch.Tensor
                     ) -> torch.Tensor:
    """For rows where two implementations picked different codes, the fp64 relative distance gap
    |d(a) - d(b)| / max(d(a), d(b), tiny) between the two picks."""
    zd, cd = z_norm.double(), cb_norm.double()
    da = ((zd - cd[idx_a]) ** 2).sum(1)
    db = ((zd - cd[idx_b]) ** 2).sum(1)
    return (da - db).abs() / torch.clamp(torch.maximum(da, db), min=1e-300)
