import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from equss_b200 import ops
dev = torch.device("cuda:0")
nchw = len(sys.argv) > 1 and sys.argv[1] == "nchw"
z = torch.randn(*((32, 1024, 40, 40) if nchw else (51200, 1024)), device=dev)
cbn = F.normalize(torch.randn(64, 256, 16, device=dev), dim=2).contiguous()
cn2 = ops.pq_cnorm2(cbn)
for _ in range(3):
    ops.pq_assign_gather(z, cbn, None, cn2, "l2")
torch.cuda.synchronize(); print("ok")
