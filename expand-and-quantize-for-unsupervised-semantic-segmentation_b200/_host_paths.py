"""Rare host-side paths of the PQ head that depend on random number generators and therefore stay in Python
(SURVEY.md 7.5 / 8a row a6): dead-code restart, most-used-code splitting, the Gumbel index draw and ``pq_dropout``.

They are written once here and shared by every quantiser flavour (the reference repeats them per class:
model/quantizer.py:73-103,298-381; dino_pqgo.py:546-577; dino_new_vq.py:293-325,516-535).  The random draws are made
with the same generators, in the same order and with the same arguments as the reference, so a seeded run
replaces the same codes by the same rows (tests/test_gpu_variants.py::test_restart_paths).
"""
from __future__ import annotations

import random
from typing import List, Optional, Tuple, Union

import torch
import torch.nn.functional as F

__all__ = ["draw_restart", "split_codes", "gumbel_indices", "dropout_keep_mask", "dropout_assign", "kmeans_centroids"]


@torch.no_grad()
def draw_restart(count: torch.Tensor, rows: torch.Tensor) -> Tuple[Union[torch.Tensor, List[int]], torch.Tensor]:
    """Pick replacement rows for the codes nobody selected.

    count: (K,) selections of each code in this step;  rows: (n, d) candidate rows (z or z_norm, per variant).
    Returns (dead code ids, replacement rows).  One ``random.shuffle`` of range(n) decides which rows are used; when
    there are more dead codes than rows, a second shuffle decides which dead codes are served
    (model/quantizer.py:298-319)."""
    n_rows = rows.shape[0]
    dead = torch.nonzero(count == 0, as_tuple=True)[0]
    perm = list(range(n_rows))
    random.shuffle(perm)
    n_dead = int(dead.numel())
    if n_dead > n_rows:
        served = dead.tolist()
        random.shuffle(served)
        return served[:n_rows], rows[perm]
    return dead, rows[perm[:n_dead]]


@torch.no_grad()
def split_codes(count: torch.Tensor, ema_count: torch.Tensor, weight: torch.Tensor, weight_avg: torch.Tensor,
                sigma: float = 0.02) -> int:
    """Give every dead code half of a busy code (model/quantizer.py:330-381), in place on the EMA state.

    The j-th dead code (in a random order drawn with ``torch.randperm``) is paired with the j-th most used code
    (by EMA count).  The pair shares the busy code's EMA count and running sum in halves, and their vectors
    become ``w + e`` (dead) and ``w - e`` (busy) with ``e ~ N(0, sigma^2)``.  Returns the number of codes replaced."""
    dead = torch.nonzero(count == 0, as_tuple=True)[0]
    n = int(dead.numel())
    if n == 0:
        return 0
    dead = dead[torch.randperm(n)]             # CPU generator, like the reference (:340)
    busy = torch.argsort(ema_count, dim=0, descending=True)[:n]
    jitter = sigma * torch.randn(n, weight.shape[1], dtype=weight.dtype, device=weight.device)
    w_busy, half_cnt, half_avg = weight[busy], ema_count[busy] / 2.0, weight_avg[busy] / 2.0
    # dead codes first, busy codes second: a code that is both (all codes dead) ends up with the busy value
    for ids, w_new in ((dead, w_busy + jitter), (busy, w_busy - jitter)):
        weight[ids] = w_new
        ema_count[ids] = half_cnt
        weight_avg[ids] = half_avg
    return n


@torch.no_grad()
def gumbel_indices(z_norm: torch.Tensor, codebook_norm: torch.Tensor, divisor: Optional[float]) -> torch.Tensor:
    """Stochastic assignment of the ``use_gumbel`` research flag (training only).

    z_norm: (n, M, d) normalised rows;  codebook_norm: (M, K, d).  Returns int32 (M, n) indices.
    Per subspace, in subspace order like the reference's loop (model/quantizer.py:595-604), the reference's distance
    (:457-461), then ``argmax(F.gumbel_softmax(-distance / divisor, tau=1.0, hard=True, dim=1))`` (:463-465;
    ``divisor`` = 0.01 for EMAVectorQuantizer, None = plain ``-distance`` for VectorQuantizer :145-147).  The noise
    comes from torch's generator of the device inside ``F.gumbel_softmax`` -- one call per subspace with an (n, K)
    argument, so a seeded run consumes the generator exactly like the reference does on the same device."""
    n, M, _ = z_norm.shape
    idx = torch.empty((M, n), dtype=torch.int32, device=z_norm.device)
    for i in range(M):
        zi, ci = z_norm[:, i, :], codebook_norm[i]
        distance = (torch.sum(zi ** 2, dim=1, keepdim=True) + torch.sum(ci ** 2, dim=1)
                    - 2 * torch.matmul(zi, ci.t()))
        logits = -distance if divisor is None else -distance / divisor
        hard = F.gumbel_softmax(logits, tau=1.0, hard=True, dim=1)
        idx[i] = torch.argmax(hard, dim=1).to(torch.int32)
    return idx


def dropout_keep_mask(num_codes: int, p: float, device) -> torch.Tensor:
    """The keep mask of ``pq_dropout``: ``uniform(0, 1) > p`` per code (model/dino_new_vq.py:388-389).  The reference
    draws with ``torch.cuda.FloatTensor(K).uniform_()``, i.e. K floats from the default generator of the current CUDA
    device; an fp32 ``uniform_`` on a fresh K-element tensor of the activations' device consumes that generator in the
    same way.  Tests replace this function to inject the reference's draws."""
    return torch.empty(num_codes, dtype=torch.float32, device=device).uniform_() > p


def dropout_assign(z_norm: torch.Tensor, codebook_norm: torch.Tensor, p: float, temperature: float
                   ) -> Tuple[torch.Tensor, List[torch.Tensor], List[torch.Tensor]]:
    """Assignment under the ``pq_dropout`` research flag (model/dino_new_vq.py:387-399,599-609; dino_pqgo.py:640-657).

    z_norm: (n, M, d) normalised rows (may carry a graph);  codebook_norm: (M, K, d) (may carry a graph).
    Per subspace, in subspace order like the reference's loop, one keep mask is drawn, the reference's distance is
    taken to the kept codes only, and ``argmin`` / ``softmax(-distance / temperature)`` follow.  Returns
    (int32 (M, n) indices -- positions in the KEPT list, which is how the reference then addresses the full
    codebook --, [prob_i (n, kept_i)], [keep_i (K,) bool]).  The width of the soft assignment differs per subspace,
    so this stays a host loop over library matmuls."""
    n, M, _ = z_norm.shape
    K = codebook_norm.shape[1]
    idx = torch.empty((M, n), dtype=torch.int32, device=z_norm.device)
    probs, keeps = [], []
    for i in range(M):
        keep = dropout_keep_mask(K, p, z_norm.device)
        zi, ci = z_norm[:, i, :], codebook_norm[i][keep]
        distance = (torch.sum(zi ** 2, dim=1, keepdim=True) + torch.sum(ci ** 2, dim=1)
                    - 2 * torch.matmul(zi, ci.t()))
        idx[i] = torch.argmin(distance.detach(), dim=1).to(torch.int32)
        probs.append(F.softmax(-distance / temperature, dim=1))
        keeps.append(keep)
    return idx, probs, keeps


@torch.no_grad()
def kmeans_centroids(rows: torch.Tensor, num_codes: int) -> torch.Tensor:
    """``need_initialized == "kmeans"`` (model/dino_pqgo.py:595-601, dino_new_vq.py:345-352, dino_pqgo_cls.py:317-323,
    quantizer.py:405-412): scikit-learn k-means++ / Lloyd with ``random_state=0`` on the first training batch's rows,
    on the host, once.  Returns the (num_codes, d) centroids as fp32 on the rows' device."""
    from sklearn.cluster import KMeans
    clustering = KMeans(init="k-means++", n_clusters=num_codes, random_state=0)
    clustering.fit(rows.detach().cpu().numpy())
    return torch.from_numpy(clustering.cluster_centers_).float().to(rows.device)
