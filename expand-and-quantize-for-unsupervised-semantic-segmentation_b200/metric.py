"""Drop-in mirror of the reference's ``model/metric.py`` (UnSegMetrics) on the K9 histogram kernel."""
from __future__ import annotations

import os
from typing import Dict

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .dist_utils import all_reduce_tensor

__all__ = ["UnSegMetrics"]


class UnSegMetrics(nn.Module):
    """model/metric.py:13-125.  ``confusion_matrix`` is int64 [num_classes+extra, num_classes], rows =
    prediction, cols = label."""

    def __init__(self, num_classes: int, extra_classes: int, compute_hungarian: bool, device: torch.device) -> None:
        super().__init__()
        self.num_classes = num_classes
        if (not compute_hungarian) and (extra_classes != 0):
            raise ValueError("No hungarian means that all classes are in order, so extra classes should be 0.")
        self.compute_hungarian = compute_hungarian
        self.extra_classes = extra_classes
        self.device = device
        self.register_buffer("confusion_matrix",
                             torch.zeros(num_classes + extra_classes, num_classes, dtype=torch.long, device=device))
        self.assignments = None
        self.histogram = None
        self.write_csv = True    # the reference dumps a CSV on every compute() (metric.py:100-108)

    def reset(self):
        self.confusion_matrix.fill_(0)
        self.assignments = None
        self.histogram = None

    @torch.no_grad()
    def update(self, preds: torch.Tensor, label: torch.Tensor):
        """Accumulate the confusion matrix (metric.py:44-58): one pass, warp-privatised shared-memory bins."""
        ops.confusion_update(preds, label, self.num_classes, self.confusion_matrix)

    @torch.no_grad()
    def compute(self, prefix: str = None) -> Dict[str, torch.Tensor]:
        """mIoU and accuracy (metric.py:60-110).  The 27x27 Hungarian stays on scipy, as in the reference."""
        from scipy.optimize import linear_sum_assignment
        self.confusion_matrix = all_reduce_tensor(self.confusion_matrix, op="sum")          # K10, :63
        if self.compute_hungarian:
            self.assignments = linear_sum_assignment(self.confusion_matrix.detach().cpu(), maximize=True)
            if self.extra_classes == 0:
                self.histogram = self.confusion_matrix[np.argsort(self.assignments[1]), :]
            else:
                assignments_t = linear_sum_assignment(self.confusion_matrix.detach().cpu().t(), maximize=True)
                histogram = self.confusion_matrix[assignments_t[1], :]
                missing = list(set(range(self.num_classes + self.extra_classes)) - set(self.assignments[0]))
                new_row = self.confusion_matrix[missing, :].sum(0, keepdim=True)
                histogram = torch.cat([histogram, new_row], dim=0)
                new_col = torch.zeros(self.num_classes + 1, 1, device=histogram.device)
                self.histogram = torch.cat([histogram, new_col], dim=1)
        else:
            self.assignments = (torch.arange(self.num_classes).unsqueeze(1), torch.arange(self.num_classes).unsqueeze(1))
            self.histogram = self.confusion_matrix
        tp = torch.diag(self.histogram)
        fp = torch.sum(self.histogram, dim=0) - tp
        fn = torch.sum(self.histogram, dim=1) - tp
        iou = tp / (tp + fp + fn)
        iou = iou[~torch.isnan(iou)].mean()
        precision = tp / (tp + fn)
        accuracy = torch.sum(tp) / torch.sum(self.histogram)
        output = dict(iou=100 * iou, accuracy=100 * accuracy)
        if self.write_csv:
            import pandas as pd
            os.makedirs(f'./class_matrix/Cityscapes/STEGO/{prefix}/', exist_ok=True)
            tmp = torch.cat([self.histogram, (precision * 100).unsqueeze(-1)], dim=1)
            pd.DataFrame(tmp.cpu().numpy()).to_csv(f'./class_matrix/Cityscapes/STEGO/{prefix}/{prefix}_7.csv')
        return output

    @torch.no_grad()
    def map_clusters(self, clusters):
        if self.extra_classes == 0:
            return torch.tensor(self.assignments[1])[clusters]
        missing = sorted(list(set(range(self.num_classes + self.extra_classes)) - set(self.assignments[0])))
        cluster_to_class = self.assignments[1]
        for missing_entry in missing:
            if missing_entry == cluster_to_class.shape[0]:
                cluster_to_class = np.append(cluster_to_class, -1)
            else:
                cluster_to_class = np.insert(cluster_to_class, missing_entry + 1, -1)
        return torch.tensor(cluster_to_class)[clusters]
