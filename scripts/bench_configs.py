#!/usr/bin/env python
"""Secondary measurements: the BASELINE.json configs that are not bench.py's default workload, one JSON line each.
(bench.py itself measures configs[1] (default) and configs[2] (--workload pq_train).)  CUDA events, inputs resident in HBM,
rotating buffers larger than L2 where the shape allows it; `cpu` = the oracle port on all host threads, bounded sample."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, torch.nn.functional as F
import equss_b200
from equss_b200 import ops
from equss_b200.quantizer import ProductQuantizerWrapper
import equss_oracle as O

dev = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def gpu_time(fn, bufs, iters=30, warm=5):
    for i in range(warm):
        fn(bufs[i % len(bufs)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(bufs[i % len(bufs)])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def cpu_time(fn, n=3):
    fn(); t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n


def emit(**kw):
    print(json.dumps(kw), flush=True)


torch.set_num_threads(os.cpu_count() or 1)
# ---- C1: pq_baseline ProductQuantizer forward, batch 4 x 28x28 tokens, 512-d, 8 subspaces x 256 codewords ----------------------
N1, D1, M1, K1 = 4 * 28 * 28, 512, 8, 256
pq = ProductQuantizerWrapper(M1, K1, D1, normalize="l2").to(dev).eval(); pq.materialize_prob = False
with torch.no_grad():
    for q in pq.quantizers: q.codebook.weight.copy_(torch.randn(K1, D1 // M1, device=dev))
z1 = [torch.randn(N1, D1, device=dev) for _ in range(4)]
with torch.no_grad():
    t_eager = gpu_time(lambda z: pq(z), z1)
    g = torch.cuda.CUDAGraph(); zs = z1[0]
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): pq(zs)
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    with torch.cuda.graph(g): pq(zs)
    t_graph = gpu_time(lambda z: g.replay(), z1)
cb1 = torch.randn(M1, K1, D1 // M1); zc = torch.randn(N1, D1)
t_cpu = cpu_time(lambda: [O.param_vq_forward(zc[:, m * 64:(m + 1) * 64].reshape(4, 28, 28, 64).permute(0, 3, 1, 2), cb1[m], normalize="l2") for m in range(M1)])
emit(config="C1 pq_baseline ProductQuantizerWrapper.forward (eval), 3136 x 512, M=8 K=256 d=64", pixels=N1,
     gpu_us_eager=round(t_eager * 1e6, 1), gpu_us_cuda_graph=round(t_graph * 1e6, 1), gpu_Mpx_s=round(N1 / t_graph / 1e6, 2),
     cpu_us=round(t_cpu * 1e6, 1), cpu_Mpx_s=round(N1 / t_cpu / 1e6, 4), cpu_threads=torch.get_num_threads())

# ---- C4: cityscapes hi-res shape, batch 16 x 56x56 tokens, 16 subspaces x 512 codewords (d = 64) ---------------------------------
for tag, Bc in (("full batch (1 GPU)", 16), ("one of 8 pixel shards", 2)):
    M4, K4, d4 = 16, 512, 64
    z4 = [torch.randn(Bc, 1024, 56, 56, device=dev) for _ in range(3)]
    cbn = F.normalize(torch.randn(M4, K4, d4, device=dev), dim=2).contiguous(); cn2 = ops.pq_cnorm2(cbn)
    n4 = Bc * 56 * 56
    t_a = gpu_time(lambda z: ops.pq_assign(z, cbn, cn2, "l2"), z4)
    t_ag = gpu_time(lambda z: ops.pq_assign_gather(z, cbn, None, cn2, "l2"), z4)
    emit(config=f"C4 cityscapes hi-res, {Bc} x 1024 x 56 x 56, M=16 K=512 d=64, {tag}", pixels=n4,
         assign_us=round(t_a * 1e6, 1), assign_Mpx_s=round(n4 / t_a / 1e6, 1), assign_useful_TFLOPs=round(2 * n4 * K4 * 1024 / t_a / 1e12, 1),
         assign_gather_us=round(t_ag * 1e6, 1), assign_gather_GBps=round((8 * n4 * 1024 + 4 * n4 * M4) / t_ag / 1e9, 1),
         frac_of_hbm_peak=round((8 * n4 * 1024 + 4 * n4 * M4) / t_ag / 1e9 / PEAK, 3))

# ---- C5: precompute_knns global-feature kNN, 50k x 768, one of 8 query shards ---------------------------------------------------
db = F.normalize(torch.randn(50000, 768, device=dev), dim=1)
for k in (8, 30):
    t = gpu_time(lambda _: ops.knn_topk(db[:6250], db, k), [0], iters=5, warm=2)
    emit(config=f"C5 kNN, 6250 queries (1/8 shard) x 50000 x 768, k={k}", queries=6250, gpu_ms=round(t * 1e3, 2),
         queries_per_s=round(6250 / t), useful_TFLOPs=round(2 * 6250 * 50000 * 768 / t / 1e12, 1))
dbc = db[:5000].cpu()
t_cpu = cpu_time(lambda: O.knn(dbc, 30, queries=dbc[:625]), n=2)
emit(config="C5 kNN CPU oracle (einsum + topk), 625 queries x 5000 x 768, k=30 (bounded sample)", cpu_ms=round(t_cpu * 1e3, 1),
     cpu_useful_TFLOPs=round(2 * 625 * 5000 * 768 / t_cpu / 1e12, 3), cpu_threads=torch.get_num_threads())
