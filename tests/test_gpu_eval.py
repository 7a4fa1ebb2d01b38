"""GPU parity of the cluster/linear probe + confusion histogram and the kNN against fixtures and oracle."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import equss_oracle as O

pytestmark = pytest.mark.gpu


def _ops():
    from equss_b200 import ops
    return ops


def _probe(feat, clusters, lin_w, lin_b, label, C, rows=None):
    ops = _ops()
    dev = torch.device("cuda:0")
    B, D, h, w = feat.shape
    Cc = clusters.shape[0]
    wmat = torch.cat([F.normalize(clusters, dim=1), lin_w], dim=0).to(dev)
    bias = torch.cat([torch.zeros(Cc), lin_b]).to(dev)
    logits = ops.probe_logits(feat.to(dev), wmat, bias)
    conf_c = torch.zeros(rows or C, C, dtype=torch.long, device=dev)
    conf_l = torch.zeros(C, C, dtype=torch.long, device=dev)
    preds = ops.probe_argmax_confusion(logits, B, h, w, Cc + C, label.to(dev), C, [(0, Cc), (Cc, C)],
                                       confusions=[conf_c, conf_l])
    torch.cuda.synchronize()
    return preds[0].cpu(), preds[1].cpu(), conf_c.cpu(), conf_l.cpu()


def _audit_preds(pg, pr, logits_ref):
    """Disagreements must be fp32 near-ties of the reference's own logits (relative gap < 1e-5 of the
    logit scale: interpolate-then-dot vs dot-then-interpolate differ by fp32 reassociation only)."""
    bad = (pg != pr)
    nb = int(bad.sum())
    if nb == 0:
        return 0
    lr = logits_ref.permute(0, 2, 3, 1)[bad]              # (nb, C)
    a = lr.gather(1, pg[bad][:, None]).squeeze(1)
    b = lr.gather(1, pr[bad][:, None]).squeeze(1)
    gap = (a - b).abs() / lr.abs().max(dim=1)[0].clamp_min(1e-30)
    assert float(gap.max()) < 1e-5, f"{nb} pred mismatches, worst relative logit gap {float(gap.max()):.3e}"
    return nb


def test_golden_probe_and_confusion(golden_dir):
    g = np.load(os.path.join(golden_dir, "eval_probe.npz"))
    C = 27
    feat, label = torch.from_numpy(g["feat"]), torch.from_numpy(g["label"])
    cp, lp, conf_c, conf_l = _probe(feat, torch.from_numpy(g["clusters"]), torch.from_numpy(g["lin_w"]),
                                    torch.from_numpy(g["lin_b"]), label, C)
    assert np.array_equal(cp.numpy(), g["cluster_preds"])
    assert np.array_equal(lp.numpy(), g["linear_preds"])
    assert np.array_equal(conf_c.numpy(), g["cluster_confusion"])
    assert np.array_equal(conf_l.numpy(), g["linear_confusion"])
    res = O.metrics_compute(conf_c, True)
    assert float(res["iou"]) == pytest.approx(float(g["cluster_iou"]), rel=1e-6)


@pytest.mark.parametrize("B,D,h,w,H,W", [(2, 64, 28, 28, 224, 224), (3, 96, 7, 5, 23, 31), (1, 32, 10, 10, 10, 10)])
def test_probe_vs_oracle(B, D, h, w, H, W):
    torch.manual_seed(3)
    C = 27
    feat = torch.randn(B, D, h, w)
    clusters, lin_w, lin_b = torch.randn(C, D), torch.randn(C, D) * 0.1, torch.randn(C) * 0.1
    label = torch.randint(-1, C, (B, H, W))
    ll, lp_ref, cl, cp_ref = O.evaluator_forward(feat, label, clusters, lin_w, lin_b, C)
    cp, lp, conf_c, conf_l = _probe(feat, clusters, lin_w, lin_b, label, C)
    up = F.interpolate(feat, (H, W), mode="bilinear", align_corners=False) if (h, w) != (H, W) else feat
    inner = torch.einsum("bchw,nc->bnhw", F.normalize(up, dim=1), F.normalize(clusters, dim=1))
    lin = F.conv2d(up, lin_w.view(C, D, 1, 1), lin_b)
    n1 = _audit_preds(cp, cp_ref, inner)
    n2 = _audit_preds(lp, lp_ref, lin)
    assert n1 + n2 <= max(2, int(2e-5 * B * H * W))
    # the fused histogram must equal the reference bincount evaluated on the kernel's own predictions
    assert torch.equal(conf_c, O.confusion_update(torch.zeros(C, C, dtype=torch.long), cp, label, C))
    assert torch.equal(conf_l, O.confusion_update(torch.zeros(C, C, dtype=torch.long), lp, label, C))


@pytest.mark.parametrize("B,D,h,w,Ct", [(32, 1024, 40, 40, 54), (2, 64, 28, 28, 54), (3, 96, 6, 6, 27), (1, 512, 56, 56, 19)])
def test_probe_logits_tensor_core_accuracy(B, D, h, w, Ct):
    """tcgen05 split-tf32 logits: fp32-level accuracy against an fp64 contraction, and agreement with the
    CUDA-core kernel to fp32 summation-order noise."""
    ops = _ops()
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    feat = torch.randn(B, D, h, w, device=dev)
    wmat = torch.randn(Ct, D, device=dev)
    bias = torch.randn(Ct, device=dev)
    pack = ops.probe_pack(wmat)
    assert pack[2] is not None, "shape should be supported by the tensor-core kernel"
    lt = ops.probe_logits(feat, pack, bias, algo=0)
    ls = ops.probe_logits(feat, pack, bias, algo=1)
    ref = torch.einsum("bchw,jc->bhwj", feat.double(), wmat.double()).reshape(-1, Ct) + bias.double()
    scale = float(ref.abs().max())
    err_t = float((lt[:, :Ct].double() - ref).abs().max()) / scale
    err_s = float((ls[:, :Ct].double() - ref).abs().max()) / scale
    assert err_t < 2e-6, f"tensor-core logits off by {err_t:.2e} of the logit scale (CUDA-core kernel: {err_s:.2e})"
    assert torch.equal(lt[:, Ct:], torch.zeros_like(lt[:, Ct:]))


@pytest.mark.parametrize("n,C,extra", [(0, 27, 0), (1, 27, 0), (100003, 27, 0), (50000, 19, 3), (4096, 300, 0)])
def test_confusion_update_vs_bincount(n, C, extra):
    ops = _ops()
    dev = torch.device("cuda:0")
    torch.manual_seed(n + C)
    preds = torch.randint(-2, C + extra + 2, (n,))
    label = torch.randint(-1, C + 1, (n,))
    label[::7] = 255 if n else 0
    conf = torch.zeros(C + extra, C, dtype=torch.long, device=dev)
    ops.confusion_update(preds.to(dev), label.to(dev), C, conf)
    ops.confusion_update(preds.to(dev), label.to(dev), C, conf)     # accumulates in place
    ref = O.confusion_update(torch.zeros(C + extra, C, dtype=torch.long), preds, label, C, extra)
    assert torch.equal(conf.cpu(), 2 * ref)


def test_confusion_coherent_labels():
    """Long runs of identical (pred, label) pairs exercise the warp run-length merge."""
    ops = _ops()
    dev = torch.device("cuda:0")
    C = 27
    label = torch.arange(200000) // 997 % C
    preds = torch.arange(200000) // 1500 % C
    label[1000:1100] = -1
    conf = torch.zeros(C, C, dtype=torch.long, device=dev)
    ops.confusion_update(preds.to(dev), label.to(dev), C, conf)
    assert torch.equal(conf.cpu(), O.confusion_update(torch.zeros(C, C, dtype=torch.long), preds, label, C))


def test_knn_golden(golden_dir):
    ops = _ops()
    g = np.load(os.path.join(golden_dir, "knn.npz"))
    feats = torch.from_numpy(g["feats"]).cuda()
    idx, sims = ops.knn_topk(feats, feats, 8, return_sims=True)
    assert np.array_equal(idx.cpu().numpy(), g["idx"])
    np.testing.assert_allclose(sims.cpu().numpy(), g["vals"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("nq,n,Fd,k", [(257, 1000, 768, 30), (64, 5000, 100, 8), (5, 33, 7, 1), (130, 130, 64, 32)])
def test_knn_vs_topk(nq, n, Fd, k):
    ops = _ops()
    torch.manual_seed(nq)
    db = F.normalize(torch.randn(n, Fd), dim=1)
    q = db[:nq] if nq <= n else F.normalize(torch.randn(nq, Fd), dim=1)
    idx, sims = ops.knn_topk(q.cuda(), db.cuda(), k, return_sims=True)
    ridx, rvals = O.knn(db, k, queries=q)
    idx, sims = idx.cpu(), sims.cpu()
    torch.testing.assert_close(sims, rvals, rtol=1e-5, atol=2e-6)
    # identical neighbour sets, except members whose similarity ties the k-th value within fp32 noise
    exact = torch.einsum("nf,mf->nm", q.double(), db.double())
    kth = rvals[:, -1:].double()
    for r in range(nq):
        diff = set(idx[r].tolist()) ^ set(ridx[r].tolist())
        for j in diff:
            assert abs(float(exact[r, j] - kth[r])) < 5e-6, (r, j)
    if nq <= n:
        assert torch.equal(idx[:, 0], torch.arange(nq))     # column 0 is the query itself (dataset_aug.py:520)


@pytest.mark.parametrize("scale_q,scale_db", [(1.0, 1.0), (1000.0, 3e-4), (1e-3, 250.0)])
def test_knn_screen_path_is_exact_for_any_scale(scale_q, scale_db):
    """F % 64 == 0 takes the fp16 screen + exact fp32 decision (csrc/knn_h.cu): un-normalised rows of very different
    magnitudes (the fp16 copies are rescaled by a power of two, the margin follows the measured norms) must give the
    same neighbours as an exact top-k."""
    ops = _ops()
    torch.manual_seed(11)
    n, nq, Fd, k = 4000, 300, 128, 12
    db = torch.randn(n, Fd) * scale_db * (0.2 + torch.rand(n, 1))          # row norms spread over 5x
    q = torch.randn(nq, Fd) * scale_q
    idx, sims = ops.knn_topk(q.cuda(), db.cuda(), k, return_sims=True)
    exact = torch.einsum("nf,mf->nm", q.double(), db.double())
    rvals, ridx = exact.topk(k, dim=1)
    idx, sims = idx.cpu(), sims.cpu()
    torch.testing.assert_close(sims.double(), rvals, rtol=2e-5, atol=1e-6 * scale_q * scale_db)
    for r in range(nq):
        for j in set(idx[r].tolist()) ^ set(ridx[r].tolist()):          # only fp32-level ties of the k-th value may differ
            assert abs(float(exact[r, j] - rvals[r, -1])) <= 2e-5 * abs(float(rvals[r, -1])), (r, j)


def test_knn_screen_path_with_duplicate_rows():
    """Hundreds of identical database rows tie within the screen's margin: the survivor lists overflow and those queries
    are scanned exhaustively -- the result is still the exact top-k with the lowest indices among equal similarities."""
    ops = _ops()
    torch.manual_seed(12)
    n, Fd, k = 3000, 64, 8
    db = F.normalize(torch.randn(n, Fd), dim=1)
    db[500:900] = db[100]                                   # 401 copies of row 100
    q = torch.cat([db[100:101], db[:40]])
    idx, sims = ops.knn_topk(q.cuda(), db.cuda(), k, return_sims=True)
    idx = idx.cpu()
    want0 = [100] + list(range(500, 500 + k - 1))             # the duplicated row itself: ties broken by the lower index
    assert idx[0].tolist() == want0, idx[0].tolist()
    exact = torch.einsum("nf,mf->nm", q.double(), db.double())
    rvals, ridx = exact.topk(k, dim=1)
    for r in range(1, q.shape[0]):
        for j in set(idx[r].tolist()) ^ set(ridx[r].tolist()):
            assert abs(float(exact[r, j] - rvals[r, -1])) < 5e-6, (r, j)


@pytest.mark.parametrize("B,D,h,w,H,W,extra", [(2, 32, 10, 10, 40, 40, 0), (3, 48, 7, 5, 23, 31, 0), (1, 64, 28, 28, 224, 224, 3),
                                                (2, 16, 6, 6, 6, 6, 0)])
def test_probe_losses_and_gradients_vs_reference_formulation(B, D, h, w, H, W, extra):
    """K8b: linear (masked CE) and cluster (mean cosine of the winning centre) losses of UnSegEvaluator.forward and their
    gradients w.r.t. the probe parameters, against the oracle's formulation on upsampled features with CPU autograd
    (model/evaluator.py:46-82,95-106).  Integer and non-integer scale factors, identity size, extra clusters."""
    from equss_b200.evaluator import UnSegEvaluator
    torch.manual_seed(B * 100 + D)
    C = 27
    ev = UnSegEvaluator(D, C, extra_classes=extra)
    with torch.no_grad():
        ev.linear_probe.weight.mul_(3.0)
    feat = torch.randn(B, D, h, w)
    label = torch.randint(-1, C, (B, H, W))
    # oracle on CPU with autograd through its own parameters
    cl = ev.cluster_probe.clusters.detach().clone().requires_grad_(True)
    lw = ev.linear_probe.weight.detach().view(C, D).clone().requires_grad_(True)
    lb = ev.linear_probe.bias.detach().clone().requires_grad_(True)
    ll_ref, lp_ref, cl_ref, cp_ref = O.evaluator_forward(feat, label, cl, lw, lb, C)
    (ll_ref + 0.7 * cl_ref).backward()
    ev = ev.to("cuda:0").train()
    ll, lp, closs, cp = ev(feat.cuda(), None, label.cuda())
    assert float(ll) == pytest.approx(float(ll_ref), rel=1e-5)
    assert float(closs) == pytest.approx(float(cl_ref), rel=1e-5)
    (ll + 0.7 * closs).backward()
    for got, ref, name in ((ev.cluster_probe.clusters.grad, cl.grad, "clusters"),
                           (ev.linear_probe.weight.grad.view(C, D), lw.grad, "linear weight"),
                           (ev.linear_probe.bias.grad, lb.grad, "linear bias")):
        scale = float(ref.abs().max())
        err = float((got.cpu() - ref).abs().max())
        assert err <= 3e-5 * scale + 1e-9, f"{name}: max err {err:.3e} vs scale {scale:.3e}"
    # no-grad call: forward-only kernel, same values
    ev.eval()
    with torch.no_grad():
        ll2, _, cl2, _ = ev(feat.cuda(), None, label.cuda())
    assert float(ll2) == pytest.approx(float(ll_ref), rel=1e-5) and float(cl2) == pytest.approx(float(cl_ref), rel=1e-5)
    # all-ignored labels: the reference's cross-entropy over an empty selection is NaN
    with torch.no_grad():
        ll3, _, _, _ = ev(feat.cuda(), None, torch.full((B, H, W), -1, device="cuda:0"))
    assert torch.isnan(ll3)


@pytest.mark.parametrize("cnt", [3, 4, 7, 10, 16, 19, 21, 27, 30, 32])
def test_probe_argmax_tournament_equals_sequential_kernel(cnt, monkeypatch):
    """The default probe kernel finds the argmax with a max3 tournament (csrc/probe_argmax_core.cuh) and walks a
    persistent schedule whose last round is split into row ranges; EQUSS_PROBE_ARGMAX_T=0 selects the sequential
    kernel.  Predictions and confusion matrices must be bit-identical on random, heavily tied and NaN / inf logits,
    for every channel count the kernel is instantiated for, ragged shapes and shapes with a split remainder."""
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(100 + cnt)
    cm = (cnt + 3) & ~3
    ct = 2 * cm
    heads = [(0, cnt), (cm, cnt)]
    for (B, h, w, H, W) in [(3, 9, 11, 70, 83), (40, 6, 5, 48, 40), (2, 4, 80, 9, 640)]:
        lab = torch.randint(-1, cnt, (B, H, W), generator=g).to(dev)
        rnd = torch.randn(B * h * w, ct, generator=g)
        tied = torch.randint(-2, 3, (B * h * w, ct), generator=g).float()
        bad = rnd.clone()
        bad.view(-1)[torch.randint(0, bad.numel(), (bad.numel() // 20,), generator=g)] = float("nan")
        bad[::7] = float("nan")
        bad[3::11] = float("-inf")
        bad[5::13, : ct // 2] = float("inf")
        for logits in (rnd, tied, bad):
            out = {}
            for mode in ("0", "1"):
                monkeypatch.setenv("EQUSS_PROBE_ARGMAX_T", mode)
                confs = [torch.zeros(cnt, cnt, dtype=torch.long, device=dev) for _ in heads]
                preds = ops.probe_argmax_confusion(logits.to(dev), B, h, w, ct, lab, cnt, heads, confusions=confs)
                torch.cuda.synchronize()
                out[mode] = (preds, confs)
            for a, b in zip(out["0"][0] + out["0"][1], out["1"][0] + out["1"][1]):
                assert torch.equal(a, b)
