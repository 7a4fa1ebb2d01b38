"""kNN timing at the BASELINE config-5 shape: one 6250-query shard and the full 50000 x 50000 problem (CUDA events)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import equss_b200  # noqa: F401
from equss_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
db = F.normalize(torch.randn(50000, 768, device=dev), dim=1)
for nq in (6250, 50000):
    for k in (8, 30):
        ops.knn_topk(db[:nq], db, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            idx = ops.knn_topk(db[:nq], db, k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        ok = bool((idx[:, 0] == torch.arange(nq, device=dev)).all())
        print(f"knn nq={nq} k={k}: {ms:.3f} ms  ({2 * nq * 50000 * 768 / ms / 1e9:.0f} TFLOP/s useful)  self-first={ok}", flush=True)
