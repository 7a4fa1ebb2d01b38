import sys, os
sys.path.insert(0, "/root/repo")
import torch, torch.nn.functional as F
from equss_b200 import ops
dev = torch.device("cuda:0")
def timeit(fn, iters=10, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for shape, M, K, d in [((51200, 1024), 64, 256, 16), ((32, 1024, 40, 40), 64, 256, 16), ((3136, 512), 8, 256, 64)]:
    z = torch.randn(*shape, device=dev); cbn = F.normalize(torch.randn(M, K, d, device=dev), dim=2).contiguous(); cn2 = ops.pq_cnorm2(cbn)
    ms = timeit(lambda: ops.pq_distance_prob(z, cbn, cn2, "l2"))
    N = z.numel() // (M * d)
    print(f"distance_prob {shape} M{M} K{K} d{d}: {ms*1e3:.0f} us, write {4*N*M*K/ms/1e6:.0f} GB/s")
