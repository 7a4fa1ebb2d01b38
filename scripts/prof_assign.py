"""One assign launch per shape for ncu (C2 flat by default)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from equss_b200 import ops
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
algo = int(sys.argv[2]) if len(sys.argv) > 2 else 2
shape, M, K, d = {"c2": ((51200, 1024), 64, 256, 16), "c2nchw": ((32, 1024, 40, 40), 64, 256, 16),
                  "c4": ((50176, 1024), 16, 512, 64), "d32": ((51200, 1024), 32, 256, 32)}[which]
z = torch.randn(*shape, device=dev)
cbn = F.normalize(torch.randn(M, K, d, device=dev), dim=2).contiguous()
cn2 = ops.pq_cnorm2(cbn)
for _ in range(3):
    idx = ops.pq_assign(z, cbn, cn2, "l2", algo=algo)
torch.cuda.synchronize(); print("ok")
