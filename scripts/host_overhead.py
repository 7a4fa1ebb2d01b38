"""CPU-side cost of the eager module calls: wall time of a loop WITHOUT synchronisation inside (the GPU queue absorbs the
launches) vs the GPU time of the same loop.  If host time per step approaches GPU time per step the step is launch-bound."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from equss_b200 import ops
from equss_b200.quantizer import ProductQuantizerWrapper
dev = torch.device("cuda:0")
M, K, D, N = 64, 256, 1024, 51200
pq = ProductQuantizerWrapper(M, K, D, normalize="l2").to(dev).train()
pq.materialize_prob = False
with torch.no_grad():
    for q in pq.quantizers:
        q.codebook.weight.copy_(torch.randn(K, D // M, device=dev)); q.codebook.weight_avg.copy_(q.codebook.weight)
zs = [torch.randn(N, D, device=dev) for _ in range(3)]
def run(n):
    with torch.no_grad():
        for i in range(n):
            pq(zs[i % 3])
run(10); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record(); run(50); e1.record(); t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"train step: host {1e6 * (t1 - t0) / 50:.0f} us/step issued, GPU {1e3 * e0.elapsed_time(e1) / 50:.0f} us/step")
# small problem: pure host cost visible
pq1 = ProductQuantizerWrapper(8, 256, 512, normalize="l2").to(dev).train(); pq1.materialize_prob = False
z1 = torch.randn(3136, 512, device=dev)
with torch.no_grad():
    for _ in range(10): pq1(z1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200): pq1(z1)
    torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"C1 train step: {1e6 * (t1 - t0) / 200:.0f} us/step wall (launch-bound)")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
with torch.no_grad():
    for _ in range(200): pq1(z1)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
