"""One eval step (fused assign+gather, probe logits, probe argmax+confusion) at the C2 shape, for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from equss_b200 import ops
dev = torch.device("cuda:0")
B, D, h, w, H, W, C, M, K = 32, 1024, 40, 40, 320, 320, 27, 64, 256
z = torch.randn(B, D, h, w, device=dev)
cbn = F.normalize(torch.randn(M, K, D // M, device=dev), dim=2).contiguous()
cn2 = ops.pq_cnorm2(cbn)
wmat = torch.randn(28 + C, D, device=dev); bias = torch.zeros(28 + C, device=dev); wpack = ops.probe_pack(wmat)
label = torch.randint(-1, C, (B, H, W), device=dev)
cc = torch.zeros(C, C, dtype=torch.long, device=dev); lc = torch.zeros(C, C, dtype=torch.long, device=dev)
for _ in range(3):
    idx, zq, sq = ops.pq_assign_gather(z, cbn, None, cn2, "l2")
    logits = ops.probe_logits(zq, wpack, bias)
    ops.probe_argmax_confusion(logits, B, h, w, 28 + C, label, C, [(0, C), (28, C)], want_preds=False, confusions=[cc, lc])
torch.cuda.synchronize(); print("ok")
