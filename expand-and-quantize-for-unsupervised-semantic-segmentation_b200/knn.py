"""Global-feature kNN precompute (K11), the arithmetic of ``data/precompute_knns.py:305-319``.

``precompute_knns`` takes the (N_img, F) mean-pooled, L2-normalised DINO features the reference's
``get_feats`` produces (precompute_knns.py:165-171) and returns the int64 (N_img, k) neighbour table the
reference stores under key ``nns``; ``save_nns`` / ``load_nns`` keep the ``.npz`` contract that
``UnSegDataset`` reads (data/dataset_aug.py:488-497,520).  With a process group, queries are sharded
across ranks (independent rows, no reduction) and gathered on every rank.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .dist_utils import get_rank, get_world_size, is_distributed_set, shard_range

__all__ = ["precompute_knns", "save_nns", "load_nns"]


@torch.no_grad()
def precompute_knns(normed_feats: torch.Tensor, k: int = 30, queries: Optional[torch.Tensor] = None,
                    sharded: bool = True) -> torch.Tensor:
    """Top-k most similar rows (column 0 is the query itself when queries is None)."""
    db = normed_feats
    q = db if queries is None else queries
    if sharded and is_distributed_set() and get_world_size() > 1:
        lo, hi = shard_range(q.shape[0])
        local = ops.knn_topk(q[lo:hi], db, k)
        per = (q.shape[0] + get_world_size() - 1) // get_world_size()
        pad = torch.full((per, k), -1, dtype=torch.int64, device=local.device)
        pad[: hi - lo] = local
        gathered = [torch.empty_like(pad) for _ in range(get_world_size())]
        dist.all_gather(gathered, pad)
        return torch.cat(gathered, dim=0)[: q.shape[0]]
    return ops.knn_topk(q, db, k)


def save_nns(path: str, nearest_neighbors: torch.Tensor) -> None:
    np.savez_compressed(path, nns=nearest_neighbors.cpu().numpy())        # precompute_knns.py:319


def load_nns(path: str) -> np.ndarray:
    return np.load(path)["nns"]                                           # dataset_aug.py:494-495
