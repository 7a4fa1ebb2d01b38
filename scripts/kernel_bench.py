"""Per-kernel timings at the BASELINE shapes (CUDA events, inputs > L2 via rotation). Not a test."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import equss_b200
from equss_b200 import ops

dev = torch.device("cuda:0")
PEAK = 6453.1


def timeit(fn, bufs, iters=20, warm=3):
    for i in range(warm):
        fn(bufs[i % len(bufs)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(bufs[i % len(bufs)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(tag, shape, M, K, d, nbuf=3, only=None):
    D = M * d
    zs = [torch.randn(*shape, device=dev) for _ in range(nbuf)]
    N = zs[0].numel() // D
    cbn = F.normalize(torch.randn(M, K, d, device=dev), dim=2).contiguous()
    cn2 = ops.pq_cnorm2(cbn)
    idx = ops.pq_assign(zs[0], cbn, cn2, "l2", algo=1)
    res = {}
    if not only or "tc" in only:
        res["assign_tc"] = (timeit(lambda z: ops.pq_assign(z, cbn, cn2, "l2", algo=2), zs), 4 * N * D + 4 * N * M)
    if not only or "fused" in only:
        res["assign_gather"] = (timeit(lambda z: ops.pq_assign_gather(z, cbn, None, cn2, "l2"), zs), 8 * N * D + 4 * N * M)
    if only and "tf32" in only:
        res["assign_tf32"] = (timeit(lambda z: ops.pq_assign(z, cbn, cn2, "l2", algo=3), zs), 4 * N * D + 4 * N * M)
    if not only or "simt" in only:
        res["assign_simt"] = (timeit(lambda z: ops.pq_assign(z, cbn, cn2, "l2", algo=1), zs, iters=5, warm=1), 4 * N * D + 4 * N * M)
    if not only or "rows" in only:
        res["gather_loss"] = (timeit(lambda z: ops.pq_gather_loss(z, cbn, idx, "l2"), zs), 8 * N * D + 4 * N * M)
        res["accumulate"] = (timeit(lambda z: ops.pq_accumulate(z, idx, K), zs), 4 * N * D + 4 * N * M)
    for k, (ms, by) in res.items():
        print(f"{tag:28s} {k:14s} {ms*1e3:9.1f} us  {by/ms/1e6:8.1f} GB/s  {by/ms/1e6/PEAK*100:5.1f}% of HBM  {N/ms/1e3:8.1f} Mpx/s", flush=True)


if __name__ == "__main__":
    only = sys.argv[1:] or None
    run("C2 flat  N=51200 M64 K256 d16", (51200, 1024), 64, 256, 16, only=only)
    run("C2 nchw  32x1024x40x40", (32, 1024, 40, 40), 64, 256, 16, only=only)
    run("C4 flat  N=50176 M16 K512 d64", (50176, 1024), 16, 512, 64, only=only)
    run("C1 flat  N=3136 M8 K256 d64", (3136, 512), 8, 256, 64, nbuf=1, only=only)
    run("d32 flat N=51200 M32 K256", (51200, 1024), 32, 256, 32, only=only)
    if not only or "probe" in only:
        B, D, h, w, H, W, C = 32, 1024, 40, 40, 320, 320, 27
        feats = [torch.randn(B, D, h, w, device=dev) for _ in range(3)]
        Cp = 28; wmat = torch.randn(Cp + C, D, device=dev); bias = torch.zeros(Cp + C, device=dev); wpack = ops.probe_pack(wmat)
        label = torch.randint(-1, C, (B, H, W), device=dev)
        ms = timeit(lambda f: ops.probe_logits(f, wpack, bias), feats)
        print(f"probe_logits(tc) {ms*1e3:.1f} us  {4*B*D*h*w/ms/1e6:.1f} GB/s")
        ms = timeit(lambda f: ops.probe_logits(f, wpack, bias, algo=1), feats)
        print(f"probe_logits(simt) {ms*1e3:.1f} us  {4*B*D*h*w/ms/1e6:.1f} GB/s")
        logits = ops.probe_logits(feats[0], wpack, bias)
        cc = torch.zeros(C, C, dtype=torch.long, device=dev); lc = torch.zeros(C, C, dtype=torch.long, device=dev)
        ms = timeit(lambda f: ops.probe_argmax_confusion(logits, B, h, w, Cp + C, label, C, [(0, C), (Cp, C)], want_preds=False, confusions=[cc, lc]), feats)
        print(f"probe_argmax_confusion(no preds) {ms*1e3:.1f} us  {8*B*H*W/ms/1e6:.1f} GB/s")
        ms = timeit(lambda f: ops.probe_argmax_confusion(logits, B, h, w, Cp + C, label, C, [(0, C), (Cp, C)], want_preds=True), feats)
        print(f"probe_argmax(preds) {ms*1e3:.1f} us  {24*B*H*W/ms/1e6:.1f} GB/s")
        preds = torch.randint(0, C, (B, H, W), device=dev)
        ms = timeit(lambda f: ops.confusion_update(preds, label, C, cc), feats)
        print(f"confusion_update {ms*1e3:.1f} us  {16*B*H*W/ms/1e6:.1f} GB/s")
    if not only or "head" in only:
        from equss_b200.head import SegmentationHead
        for (B, C, h, w, D) in [(32, 384, 40, 40, 1024), (32, 768, 40, 40, 1024), (16, 768, 56, 56, 1024)]:
            head = SegmentationHead(C, D).to(dev).eval()
            xs = [torch.randn(B, C, h, w, device=dev) for _ in range(3)]
            n = B * h * w
            fl = 2.0 * n * (C * D * 2 + C * C)
            with torch.no_grad():
                ms = timeit(lambda x: head(x), xs)
                ms_h = timeit(lambda x: ops.head_gemm(x, head.cluster2[0].weight, head.cluster2[0].bias, relu=True), xs)
                out = head(xs[0])
                ref = (head.cluster1(xs[0].double()) if False else None)
                h64 = head.double()
                ref = h64.cluster1(xs[0].double()) + h64.cluster2(xs[0].double())
                err = float((out.double() - ref).abs().max() / ref.abs().max())
                head.float()
                torch.backends.cudnn.allow_tf32 = True
                ms_t = timeit(lambda x: head.cluster1(x) + head.cluster2(x), xs)
                torch.backends.cudnn.allow_tf32 = False
                ms_f = timeit(lambda x: head.cluster1(x) + head.cluster2(x), xs, iters=5, warm=2)
            print(f"head B={B} C={C} {h}x{w} D={D}: {ms*1e3:8.1f} us (hidden {ms_h*1e3:.1f}) {fl/ms/1e9:7.1f} TFLOP/s useful, "
                  f"x3 issued = {3*fl/ms/1e9/1405.3*100:.1f}% of bf16-equivalent peak... rel err {err:.2e} | "
                  f"torch cudnn tf32 {ms_t*1e3:.1f} us, torch fp32 {ms_f*1e3:.1f} us", flush=True)
    if not only or "knn" in only:
        db = F.normalize(torch.randn(50000, 768, device=dev), dim=1)
        ms = timeit(lambda f: ops.knn_topk(db[:6250], db, 30), [0], iters=2, warm=1)
        print(f"knn 6250x50000x768 k=30: {ms:.1f} ms  {2*6250*50000*768/ms/1e9:.1f} TFLOP/s")
