"""Host mirror of the reference's ``model/metric.py`` (``UnSegMetrics``): same constructor, buffer, methods and
results; the accumulation runs on the K9 histogram kernel, the 27x27 matching stays on scipy."""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .dist_utils import all_reduce_tensor

__all__ = ["UnSegMetrics"]


def _match(conf: torch.Tensor) -> Tuple[np.ndarray, np.ndarray]:
    """Maximum-weight assignment of the rows of ``conf`` to its columns (scipy's Hungarian solver)."""
    from scipy.optimize import linear_sum_assignment
    return linear_sum_assignment(conf.detach().cpu(), maximize=True)


def _scores(hist: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(mean IoU over the classes that occur, pixel accuracy, per-row hit rate) of a matched histogram whose rows are
    predictions and columns labels (metric.py:85-94)."""
    hit = hist.diagonal()
    per_label, per_pred = hist.sum(dim=0), hist.sum(dim=1)
    iou = hit / (per_label + per_pred - hit)
    miou = iou[~iou.isnan()].mean()
    return miou, hit.sum() / hist.sum(), hit / per_pred


class UnSegMetrics(nn.Module):
    """model/metric.py:13-125.  ``confusion_matrix`` is int64 [num_classes + extra_classes, num_classes]: rows are
    predictions (cluster ids), columns labels."""

    def __init__(self, num_classes: int, extra_classes: int, compute_hungarian: bool, device: torch.device) -> None:
        super().__init__()
        if extra_classes != 0 and not compute_hungarian:
            raise ValueError("No hungarian means that all classes are in order, so extra classes should be 0.")
        self.num_classes, self.extra_classes = num_classes, extra_classes
        self.compute_hungarian = compute_hungarian
        self.device = device
        rows = num_classes + extra_classes
        self.register_buffer("confusion_matrix", torch.zeros(rows, num_classes, dtype=torch.long, device=device))
        self.assignments: Optional[tuple] = None
        self.histogram: Optional[torch.Tensor] = None
        self.write_csv = True       # the reference writes a CSV on every compute() (metric.py:100-108)

    def reset(self):
        self.confusion_matrix.zero_()
        self.assignments = self.histogram = None

    @torch.no_grad()
    def update(self, preds: torch.Tensor, label: torch.Tensor):
        """metric.py:44-58 in one pass over (preds, label): warp-privatised shared-memory bins, pairs outside
        [0, num_classes) on either side are dropped."""
        ops.confusion_update(preds, label, self.num_classes, self.confusion_matrix)

    def _unassigned_rows(self):
        """Cluster rows the matching left without a class (only possible with extra classes)."""
        taken = set(int(r) for r in self.assignments[0])
        return [r for r in range(self.num_classes + self.extra_classes) if r not in taken]

    def _matched_histogram(self) -> torch.Tensor:
        conf, C = self.confusion_matrix, self.num_classes
        if not self.compute_hungarian:                       # linear probe: prediction i already means class i
            ident = torch.arange(C).unsqueeze(1)
            self.assignments = (ident, ident.clone())
            return conf
        self.assignments = _match(conf)
        if self.extra_classes == 0:
            # row k of the histogram = the cluster matched to class k (argsort inverts the class <- cluster map)
            return conf[np.argsort(self.assignments[1]), :]
        # more clusters than classes (metric.py:73-81): match classes to clusters on the transpose, pool the
        # unmatched clusters into one extra row and pad a zero column so the histogram stays square
        cluster_of_class = _match(conf.t())[1]
        pooled = conf[self._unassigned_rows(), :].sum(dim=0, keepdim=True)
        square = torch.cat([conf[cluster_of_class, :], pooled], dim=0)
        return torch.cat([square, torch.zeros(C + 1, 1, device=square.device)], dim=1)

    @torch.no_grad()
    def compute(self, prefix: str = None) -> Dict[str, torch.Tensor]:
        """mIoU and pixel accuracy in percent (metric.py:60-110); the confusion matrix is summed over ranks first."""
        self.confusion_matrix = all_reduce_tensor(self.confusion_matrix, op="sum")          # K10, :63
        self.histogram = self._matched_histogram()
        miou, accuracy, hit_rate = _scores(self.histogram)
        if self.write_csv:
            self._dump_csv(prefix, hit_rate)
        return {"iou": 100 * miou, "accuracy": 100 * accuracy}

    def _dump_csv(self, prefix, hit_rate: torch.Tensor) -> None:
        """The reference's side effect: histogram plus a percentage column, as a pandas CSV under ./class_matrix."""
        import pandas as pd
        folder = f"./class_matrix/Cityscapes/STEGO/{prefix}/"
        os.makedirs(folder, exist_ok=True)
        table = torch.cat([self.histogram, (100 * hit_rate).unsqueeze(-1)], dim=1)
        pd.DataFrame(table.cpu().numpy()).to_csv(f"{folder}{prefix}_7.csv")

    @torch.no_grad()
    def map_clusters(self, clusters):
        """Class id of every cluster id in ``clusters`` under the last matching; unmatched clusters map to -1
        (metric.py:112-125, including its insertion rule for the -1 entries)."""
        lookup = self.assignments[1]
        if self.extra_classes != 0:
            for row in self._unassigned_rows():                       # ascending
                at = len(lookup) if row == lookup.shape[0] else row + 1
                lookup = np.insert(lookup, at, -1)
        return torch.tensor(lookup)[clusters]
