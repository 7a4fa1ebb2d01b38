"""kNN timing: fused top-k epilogue vs the unfused GEMM + select path (EQUSS_KNN_UNFUSED=1), shard and full shapes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from equss_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
db = F.normalize(torch.randn(50000, 768, device=dev), dim=1)
for nq in (6250, 50000):
    for k in (8, 30):
        for mode in ("fused", "unfused"):
            if mode == "unfused":
                os.environ["EQUSS_KNN_UNFUSED"] = "1"
            else:
                os.environ.pop("EQUSS_KNN_UNFUSED", None)
            q = db[:nq]
            for _ in range(2):
                ops.knn_topk(q, db, k)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.knn_topk(q, db, k)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"nq={nq} k={k} {mode}: {ms:.3f} ms  ({2.0 * nq * 50000 * 768 / ms / 1e9:.0f} TFLOP/s useful)")
