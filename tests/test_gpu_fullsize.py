"""BASELINE.json's full shapes against the CPU ORACLE (not against another kernel of this repo).

C2 (51 200 px, M=64, K=256, d=16; flat and NCHW) and C4 (50 176 px, M=16, K=512, d=64) assignments against the
oracle's fp32 distance + argmin per subspace, every disagreement audited in fp64; the evaluator at B=2, D=1024,
40x40 -> 320x320; one query shard of the 50k x 768 kNN against einsum + topk on the host with an fp64 tie audit.
The oracle needs a few seconds per case on the GPU box's host cores."""
import pytest
import torch
import torch.nn.functional as F

import equss_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NEAR_TIE = 1e-6


def _oracle_assign_audit(z_cpu, cbn_cpu, idx_gpu, M, d):
    """z_cpu: flat (n, D) or NCHW; idx_gpu int32 [M, n] (CPU).  Returns the number of audited near-ties."""
    zf = z_cpu if z_cpu.dim() == 2 else z_cpu.permute(0, 2, 3, 1).reshape(-1, z_cpu.shape[1])
    ties = 0
    for m in range(M):
        zn = F.normalize(zf[:, m * d:(m + 1) * d], dim=1)
        ref = torch.argmin(O.sq_distance(zn, cbn_cpu[m]), dim=1)
        got = idx_gpu[m].long()
        bad = (ref != got).nonzero().flatten()
        if bad.numel():
            marg = O.top2_margin_fp64(zn[bad], cbn_cpu[m], ref[bad], got[bad])
            assert float(marg.max()) < NEAR_TIE, f"subspace {m}: {bad.numel()} mismatches, worst fp64 margin {float(marg.max()):.3e}"
            ties += int(bad.numel())
    return ties


@pytest.mark.parametrize("name,shape,M,K,d", [
    ("C2 flat", (51200, 1024), 64, 256, 16),
    ("C2 NCHW", (32, 1024, 40, 40), 64, 256, 16),
    ("C4 NCHW", (16, 1024, 56, 56), 16, 512, 64),
    ("C4 shard flat", (6272, 1024), 16, 512, 64),
    ("C1", (3136, 512), 8, 256, 64),
])
def test_assign_gather_full_size_vs_oracle(name, shape, M, K, d):
    from equss_b200 import ops
    torch.set_num_threads(max(1, torch.get_num_threads()))
    g = torch.Generator().manual_seed(2024)
    z = torch.randn(*shape, generator=g)
    cbn = F.normalize(torch.randn(M, K, d, generator=g), dim=2)
    zd, cd = z.to(DEV), cbn.to(DEV)
    idx, out, sqerr = ops.pq_assign_gather(zd, cd, None, None, "l2")
    torch.cuda.synchronize()
    idx_c = idx.cpu()
    n = idx_c.shape[1]
    ties = _oracle_assign_audit(z, cbn, idx_c, M, d)
    assert ties <= max(4, int(2e-5 * M * n)), f"{name}: {ties} near-ties of {M * n} rows is implausibly many"
    # gathered / straight-through rows and the commitment loss against the oracle's arithmetic on the GPU's indices
    zf = z if z.dim() == 2 else z.permute(0, 2, 3, 1).reshape(-1, z.shape[1])
    of = out.cpu() if z.dim() == 2 else out.cpu().permute(0, 2, 3, 1).reshape(-1, z.shape[1])
    for m in (0, M // 2, M - 1):
        zn = F.normalize(zf[:, m * d:(m + 1) * d], dim=1)
        q = cbn[m][idx_c[m].long()]
        torch.testing.assert_close(of[:, m * d:(m + 1) * d], zn + (q - zn), rtol=1e-5, atol=1e-6)
        mse = float(sqerr[m]) / (n * d)
        assert mse == pytest.approx(float(F.mse_loss(zn, q)), rel=1e-5)
    # counts / sums of the scatter-add against one_hot arithmetic (exact counts)
    packed = ops.pq_accumulate(zd, idx, K).cpu()
    for m in (0, M - 1):
        cnt = torch.bincount(idx_c[m].long(), minlength=K).float()
        assert torch.equal(packed[m, :, d], cnt)
        ref_sum = torch.zeros(K, d).index_add_(0, idx_c[m].long(), zf[:, m * d:(m + 1) * d])
        torch.testing.assert_close(packed[m, :, :d], ref_sum, rtol=1e-4, atol=1e-4)


def test_evaluator_full_width_vs_oracle():
    """UnSegEvaluator predictions at D = 1024, 40x40 tokens -> 320x320 labels (the cocostuff27 eval geometry, 2 images)."""
    from equss_b200 import ops
    torch.manual_seed(5)
    B, D, h, w, H, W, C = 2, 1024, 40, 40, 320, 320, 27
    feat = torch.randn(B, D, h, w)
    clusters, lin_w, lin_b = torch.randn(C, D), torch.randn(C, D) * 0.03, torch.randn(C) * 0.1
    label = torch.randint(-1, C, (B, H, W))
    _, lp_ref, _, cp_ref = O.evaluator_forward(feat, label, clusters, lin_w, lin_b, C)
    Cp = 28
    wmat = torch.zeros(Cp + C, D); wmat[:C] = F.normalize(clusters, dim=1); wmat[Cp:] = lin_w
    bias = torch.zeros(Cp + C); bias[Cp:] = lin_b
    logits = ops.probe_logits(feat.to(DEV), wmat.to(DEV), bias.to(DEV))
    conf_c = torch.zeros(C, C, dtype=torch.long, device=DEV)
    conf_l = torch.zeros(C, C, dtype=torch.long, device=DEV)
    cp, lp = ops.probe_argmax_confusion(logits, B, h, w, Cp + C, label.to(DEV), C, [(0, C), (Cp, C)], confusions=[conf_c, conf_l])
    cp, lp = cp.cpu(), lp.cpu()
    up = F.interpolate(feat, (H, W), mode="bilinear", align_corners=False)
    inner = torch.einsum("bchw,nc->bnhw", F.normalize(up, dim=1), F.normalize(clusters, dim=1))
    lin = F.conv2d(up, lin_w.view(C, D, 1, 1), lin_b)
    nbad = 0
    for got, ref, lg in ((cp, cp_ref, inner), (lp, lp_ref, lin)):
        bad = got != ref
        nbad += int(bad.sum())
        if bad.any():
            lr = lg.permute(0, 2, 3, 1)[bad]
            gap = (lr.gather(1, got[bad][:, None]) - lr.gather(1, ref[bad][:, None])).abs().squeeze(1) / lr.abs().max(dim=1)[0]
            assert float(gap.max()) < 1e-5, f"worst relative logit gap {float(gap.max()):.3e}"
    assert nbad <= max(2, int(2e-5 * B * H * W))
    assert torch.equal(conf_c.cpu(), O.confusion_update(torch.zeros(C, C, dtype=torch.long), cp, label, C))
    assert torch.equal(conf_l.cpu(), O.confusion_update(torch.zeros(C, C, dtype=torch.long), lp, label, C))


@pytest.mark.parametrize("k", [8, 30])
def test_knn_query_shard_of_50k_vs_oracle(k):
    """One of the eight query shards of BASELINE configs[4] (6 250 queries x 50 000 x 768) against einsum + topk on
    the host; membership differences must be fp64 ties at the k-th similarity, column 0 must be the query itself."""
    from equss_b200 import ops
    torch.manual_seed(50)
    n, Fd, nq, lo = 50000, 768, 6250, 12500            # shard 2 of 8
    db = F.normalize(torch.randn(n, Fd), dim=1)
    q = db[lo:lo + nq]
    idx, sims = ops.knn_topk(q.to(DEV), db.to(DEV), k, return_sims=True)
    idx, sims = idx.cpu(), sims.cpu()
    ridx, rvals = O.knn(db, k, queries=q)
    torch.testing.assert_close(sims, rvals, rtol=1e-5, atol=2e-6)
    assert torch.equal(idx[:, 0], torch.arange(lo, lo + nq))
    differ = (idx.sort(dim=1)[0] != ridx.sort(dim=1)[0]).any(dim=1).nonzero().flatten()
    assert differ.numel() <= 10, f"{differ.numel()} rows with a different neighbour set"
    for r in differ.tolist():
        exact = (q[r].double() @ db.double().t())
        kth = torch.topk(exact, k)[0][-1]
        for j in set(idx[r].tolist()) ^ set(ridx[r].tolist()):
            assert abs(float(exact[j] - kth)) < 5e-6, (r, j)
    # ordering inside a row: non-increasing similarity
    assert bool((sims[:, :-1] >= sims[:, 1:]).all())
