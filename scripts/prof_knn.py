import sys, os
sys.path.insert(0, "/root/repo")
import torch, torch.nn.functional as F
import equss_b200
from equss_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
db = F.normalize(torch.randn(50000, 768, device=dev), dim=1)
for k in (8, 30):
    ops.knn_topk(db[:6250], db, k)
torch.cuda.synchronize()
