"""Every kernel of libequss_b200.so once, at the BASELINE shapes, for one `ncu --set full` capture:

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_all \
        python scripts/prof_all_kernels.py

Each op runs once untimed (warm-up, outside the profiler range) and once between cudaProfilerStart/Stop."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import equss_b200  # noqa: F401
from equss_b200 import ops
from equss_b200.codebooks import EMACodebook, NewVQProductQuantizerWrapper
from equss_b200.evaluator import UnSegEvaluator
from equss_b200.head import SegmentationHead
from equss_b200.quantizer import ProductQuantizerWrapper

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, D, h, w, H, W, C, M, K = 32, 1024, 40, 40, 320, 320, 27, 64, 256
d = D // M
z_nchw = torch.randn(B, D, h, w, device=dev)
z_flat = torch.randn(B * h * w, D, device=dev)
cb = torch.randn(M, K, d, device=dev)
label = torch.randint(-1, C, (B, H, W), device=dev)
wmat = torch.randn(28 + C, D, device=dev)
bias = torch.zeros(28 + C, device=dev)
cc = torch.zeros(C, C, dtype=torch.long, device=dev)
lc = torch.zeros(C, C, dtype=torch.long, device=dev)
z4 = torch.randn(16 * 56 * 56, 1024, device=dev)
cb4 = F.normalize(torch.randn(16, 512, 64, device=dev), dim=2).contiguous()
db = F.normalize(torch.randn(50000, 768, device=dev), dim=1)
pq = ProductQuantizerWrapper(M, K, D, normalize="l2").to(dev).train()
pq.materialize_prob = False
pq_zn = ProductQuantizerWrapper(8, K, 512, normalize="z_norm").to(dev).train()
pq_zn.materialize_prob = False
nv = NewVQProductQuantizerWrapper(M, K, D, normalize="l2", jsd_ts=0.5, quantizer_cls=EMACodebook).to(dev).eval()
nv.materialize_prob = False
ev = UnSegEvaluator(D, C).to(dev).train()
head = SegmentationHead(768, D).to(dev).eval()
x_head = torch.randn(B, 768, h, w, device=dev)
with torch.no_grad():
    for q in list(pq.quantizers) + list(pq_zn.quantizers):
        q.codebook.weight.copy_(torch.randn_like(q.codebook.weight)); q.codebook.weight_avg.copy_(q.codebook.weight)
    for q in nv.quantizers:
        q.codebook.weight.copy_(torch.randn_like(q.codebook.weight))


def one_pass():
    with torch.no_grad():
        cbn, cn2 = ops.pq_prepare_codebook(cb, "l2")                                   # codebook_prepare_kernel
        idx, zq, sq = ops.pq_assign_gather(z_nchw, cbn, None, cn2, "l2")               # build_image + assign_f16x2<FUSE> (NCHW)
        ops.pq_assign_gather(z_flat, cbn, None, cn2, "l2")                             # flat, G = 2
        idx_u = ops.pq_assign(z_flat, cbn, cn2, "l2")                                  # unfused assign
        ops.pq_gather_loss(z_flat, cbn, idx_u, "l2")                                   # gather_loss_flat_l2
        ops.pq_gather_loss(z_nchw, cbn, idx, "l2")                                     # gather_loss_nchw_l2
        ops.pq_accumulate(z_flat, idx_u, K)                                            # accumulate_flat
        ops.pq_accumulate(z_nchw, idx, K)                                              # accumulate_rows
        ops.pq_distance_prob(z_flat[:6400], cbn, cn2, "l2")                            # distance_prob_tiled (1/8 of C2: 420 MB out)
        ops.pq_soft_stats(z_flat, cbn, cn2, "l2", None, None, 0.5)                     # soft_stats_kernel
        ops.channel_moments(z_flat); ops.channel_moments(z_nchw)                       # channel_moments_{flat,nchw}
        pq(z_flat)                                                                     # train step: + ema_train_tail_kernel
        pq_zn(z_flat[:3136, :512].contiguous())                                        # split-tf32 assign_tc (z_norm rows) at C1
        i4, q4, s4 = ops.pq_assign_gather(z4, cb4, None, None, "l2")                   # C4: two chunks, merge, rescan, gather
        ops.pq_accumulate(z4, i4, 512)
        ops.pq_assign(z_flat[:4096], cbn, cn2, "l2", algo=1)                           # assign_simt (exact validator)
        logits = ops.probe_logits(zq, ops.probe_pack(wmat), bias)                      # probe image + probe_logits_tc
        ops.probe_argmax_confusion(logits, B, h, w, 28 + C, label, C, [(0, C), (28, C)], want_preds=False, confusions=[cc, lc])
        preds = ops.probe_argmax_confusion(logits, B, h, w, 28 + C, label, C, [(0, C), (28, C)])
        ops.confusion_update(preds[0], label, C, cc)                                   # confusion_kernel
        head(x_head)                                                                   # head_gemm_tc x2
        ops.knn_topk(db[:6250], db, 8)                                                 # knn_gemm_tc<topk> + knn_merge
        nv(z_nchw, 0)                                                                  # V4 eval: fused jsd / entropy
    ll, _, cl, _ = ev(zq, None, label)                                                 # token_gram + probe_losses (fwd + grad)
    (ll + cl).backward()
    zg = z_flat[:12800].clone().requires_grad_(True)
    o, out, _ = pq(zg)                                                                 # gather_loss_bwd
    (o.sum() + out["loss"]).backward()
    torch.cuda.synchronize()


one_pass()
torch.cuda.profiler.start()
one_pass()
torch.cuda.profiler.stop()
print("ok")
