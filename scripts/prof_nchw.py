import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from equss_b200 import ops
dev = torch.device("cuda:0")
M, K, d = 64, 256, 16
z = torch.randn(32, M * d, 40, 40, device=dev)
cbn = F.normalize(torch.randn(M, K, d, device=dev), dim=2).contiguous()
cn2 = ops.pq_cnorm2(cbn)
for _ in range(3):
    idx = ops.pq_assign(z, cbn, cn2, "l2", algo=2)
torch.cuda.synchronize()
print("ok", int(idx.sum()))
