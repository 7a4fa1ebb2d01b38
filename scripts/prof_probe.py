import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from equss_b200 import ops
dev = torch.device("cuda:0")
B, D, h, w, H, W, C = 32, 1024, 40, 40, 320, 320, 27
feat = torch.randn(B, D, h, w, device=dev)
wmat = torch.randn(28 + C, D, device=dev); bias = torch.zeros(28 + C, device=dev); wpack = ops.probe_pack(wmat)
label = torch.randint(-1, C, (B, H, W), device=dev)
cc = torch.zeros(C, C, dtype=torch.long, device=dev); lc = torch.zeros(C, C, dtype=torch.long, device=dev)
for _ in range(2):
    logits = ops.probe_logits(feat, wpack, bias)
    ops.probe_argmax_confusion(logits, B, h, w, 28 + C, label, C, [(0, C), (28, C)], want_preds=False, confusions=[cc, lc])
torch.cuda.synchronize(); print("ok")
