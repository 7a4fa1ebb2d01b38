"""The nn.Module mirrors (same names / signatures / state-dict keys as the reference) on the GPU, against
the golden fixtures produced by the reference modules themselves."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import equss_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _val(v):
    return float(v) if v is not None else float("nan")


@pytest.mark.parametrize("mode", ["l2", "z_norm", "none"])
def test_product_quantizer_wrapper_ema_matches_reference(golden_dir, mode):
    from equss_b200.quantizer import EMAVectorQuantizer, ProductQuantizerWrapper
    g = np.load(os.path.join(golden_dir, f"pq_ema_{mode}.npz"))
    M, K = int(g["M"]), int(g["K"])
    D = g["z0"].shape[1]
    pq = ProductQuantizerWrapper(M, K, D, beta=0.25, normalize=mode, decay=0.99, eps=1e-5,
                                 quantizer_cls=EMAVectorQuantizer)
    sd = pq.state_dict()
    assert sorted(sd.keys()) == sorted(f"quantizers.{i}.codebook.{n}" for i in range(M)
                                       for n in ("weight", "weight_avg", "vq_count"))
    w0 = torch.from_numpy(g["weight0"])
    for i in range(M):
        sd[f"quantizers.{i}.codebook.weight"] = w0[i].clone()
        sd[f"quantizers.{i}.codebook.weight_avg"] = w0[i].clone()
    pq.load_state_dict(sd, strict=True)
    pq = pq.to(DEV)
    pq.train()
    for s in range(4):
        if s == 3:
            pq.eval()
        z = torch.from_numpy(g[f"z{s}"]).to(DEV)
        zq, out, prob = pq(z)
        np.testing.assert_allclose(zq.cpu().numpy(), g[f"zq{s}"], rtol=1e-5, atol=1e-6)
        keys = {k[len(f"out{s}/"):] for k in g.files if k.startswith(f"out{s}/")}
        assert set(out.keys()) == keys
        for k in keys:
            assert _val(out[k]) == pytest.approx(float(g[f"out{s}/{k}"]), rel=1e-5, abs=1e-7, nan_ok=True), (s, k)
        w = torch.stack([q.codebook.weight for q in pq.quantizers]).cpu().numpy()
        np.testing.assert_allclose(w, g[f"weight_after{s}"], rtol=1e-5, atol=1e-7)
        wa = torch.stack([q.codebook.weight_avg for q in pq.quantizers]).cpu().numpy()
        np.testing.assert_allclose(wa, g[f"weight_avg_after{s}"], rtol=1e-5, atol=1e-7)
        c = torch.stack([q.codebook.vq_count for q in pq.quantizers]).cpu().numpy()
        np.testing.assert_allclose(c, g[f"vq_count_after{s}"], rtol=1e-6, atol=1e-7)
        assert np.array_equal(torch.stack([q.vq_count for q in pq.quantizers]).cpu().numpy(), g[f"exact_after{s}"])
    np.testing.assert_allclose(prob.cpu().numpy(), g["prob3"], rtol=2e-5, atol=1e-7)
    # the per-subspace module is callable on its own, like quantizers[i](z_i) in the reference loop
    q1, o1, p1 = pq.quantizers[1](z[:, D // M:2 * D // M].contiguous())
    np.testing.assert_allclose(q1.cpu().numpy(), g["zq3"][:, D // M:2 * D // M], rtol=1e-5, atol=1e-6)
    # state dict round trip after training keeps the stacked storage consistent
    sd2 = {k: v.cpu() for k, v in pq.state_dict().items()}
    np.testing.assert_allclose(sd2["quantizers.2.codebook.weight"].numpy(), g["weight_after3"][2], rtol=1e-5, atol=1e-7)


def test_learned_codebook_variants_match_reference(golden_dir):
    from equss_b200.codebooks import Codebook, PQGOProductQuantizerWrapper
    from equss_b200.quantizer import VectorQuantizer
    from equss_b200.quantizer_v2 import EMAVectorQuantizer as V2EMA
    g = np.load(os.path.join(golden_dir, "pq_param_nchw.npz"))
    z = torch.from_numpy(g["z"]).to(DEV)
    K, d = int(g["K"]), z.shape[1]
    vq = VectorQuantizer(K, d, beta=0.25, normalize="l2").to(DEV).eval()
    with torch.no_grad():
        vq.codebook.weight.copy_(torch.from_numpy(g["v1_codebook"]))
    q, out, prob = vq(z)
    np.testing.assert_allclose(q.detach().cpu().numpy(), g["v1_q"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(prob.detach().cpu().numpy(), g["v1_prob"], rtol=2e-5, atol=1e-7)
    assert float(out["loss"]) == pytest.approx(float(g["v1_loss"]), rel=1e-5)
    assert float(out["codebook_loss"]) == pytest.approx(float(g["v1_codebook_loss"]), rel=1e-5)
    cb = Codebook(K, d, beta=0.25, book=1.0, normalize="none", need_initialized="none").to(DEV).eval()
    with torch.no_grad():
        cb.embedding.weight.copy_(torch.from_numpy(g["v5_codebook"]))
    q5, out5, prob5, idx5 = cb(z, torch.zeros_like(z))
    assert np.array_equal(idx5.cpu().numpy(), g["v5_idx"])
    np.testing.assert_allclose(q5.detach().cpu().numpy(), g["v5_q"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(prob5.detach().cpu().numpy(), g["v5_prob"], rtol=2e-5, atol=1e-7)
    assert float(out5["vq-loss"]) == pytest.approx(float(g["v5_vq_loss"]), rel=1e-5)
    w = PQGOProductQuantizerWrapper(1, K, d, normalize="none").to(DEV).eval()
    with torch.no_grad():
        w.quantizers[0].embedding.weight.copy_(torch.from_numpy(g["v5_codebook"]))
    zq, (zs, zqs, idxs), outs, probs = w(z)
    assert np.array_equal(idxs[0].cpu().numpy(), g["v5_idx"]) and tuple(probs.shape) == tuple(g["v5_prob"].shape)
    g2 = np.load(os.path.join(golden_dir, "pq_v2_nchw.npz"))
    v2 = V2EMA(K, d, beta=0.25).to(DEV).eval()
    with torch.no_grad():
        v2.embeddings.copy_(torch.from_numpy(g2["embeddings"]))
    q2, out2, _ = v2(z)
    np.testing.assert_allclose(q2.cpu().numpy(), g2["q"], rtol=1e-5, atol=1e-6)
    assert float(out2["loss"]) == pytest.approx(float(g2["loss"]), rel=1e-5)


def test_learned_codebook_gradients_match_autograd():
    """Gradients w.r.t. the activations and the codebook parameters of the fused module equal torch
    autograd on the oracle's formulation (STE + codebook loss + beta * commitment loss)."""
    from equss_b200.quantizer import ProductQuantizerWrapper, VectorQuantizer
    torch.manual_seed(3)
    M, K, D = 2, 16, 32
    pq = ProductQuantizerWrapper(M, K, D, beta=0.25, normalize="l2", quantizer_cls=VectorQuantizer).to(DEV).train()
    z = torch.randn(2, D, 5, 4, device=DEV, requires_grad=True)
    zq, out, _ = pq(z)
    go = torch.randn_like(zq)
    ((zq * go).sum() + out["loss"]).backward()
    zc = z.detach().cpu().double().requires_grad_(True)
    d = D // M
    total = 0.0
    cbs = [q.codebook.weight.detach().cpu().double().requires_grad_(True) for q in pq.quantizers]
    for m in range(M):
        q_ste, o, _, _ = O.param_vq_forward_grad(zc[:, m * d:(m + 1) * d], cbs[m], "l2", 0.25) \
            if hasattr(O, "param_vq_forward_grad") else _oracle_param_grad(zc[:, m * d:(m + 1) * d], cbs[m])
        total = total + (q_ste * go.cpu().double()[:, m * d:(m + 1) * d]).sum() + o / M
    total.backward()
    torch.testing.assert_close(z.grad.cpu().double(), zc.grad, rtol=2e-4, atol=2e-6)
    for m in range(M):
        torch.testing.assert_close(pq.quantizers[m].codebook.weight.grad.cpu().double(), cbs[m].grad, rtol=2e-4, atol=2e-6)


def _oracle_param_grad(z_nchw, codebook, beta=0.25):
    """Differentiable restatement of model/quantizer.py:112-187 (fp64)."""
    b, d, h, w = z_nchw.shape
    zf = z_nchw.permute(0, 2, 3, 1).reshape(-1, d)
    zn, cn = F.normalize(zf, dim=1), F.normalize(codebook, dim=1)
    idx = torch.argmin(O.sq_distance(zn, cn), dim=1)
    q = F.embedding(idx, cn)
    loss = F.mse_loss(q, zn.detach()) + beta * F.mse_loss(zn, q.detach())
    ste = zn + (q - zn).detach()
    return ste.view(b, h, w, d).permute(0, 3, 1, 2), loss, None, None


def test_evaluator_and_metrics_modules(golden_dir, tmp_path, monkeypatch):
    from equss_b200.evaluator import UnSegEvaluator
    from equss_b200.metric import UnSegMetrics
    g = np.load(os.path.join(golden_dir, "eval_probe.npz"))
    C, D = 27, g["feat"].shape[1]
    ev = UnSegEvaluator(D, C).to(DEV).eval()
    with torch.no_grad():
        ev.cluster_probe.clusters.copy_(torch.from_numpy(g["clusters"]))
        ev.linear_probe.weight.copy_(torch.from_numpy(g["lin_w"]).view(C, D, 1, 1))
        ev.linear_probe.bias.copy_(torch.from_numpy(g["lin_b"]))
    feat, label = torch.from_numpy(g["feat"]).to(DEV), torch.from_numpy(g["label"]).to(DEV)
    ll, lp, cl, cp = ev(feat, None, label, is_crf=False)
    assert np.array_equal(lp.cpu().numpy(), g["linear_preds"]) and np.array_equal(cp.cpu().numpy(), g["cluster_preds"])
    assert float(ll) == pytest.approx(float(g["linear_loss"]), rel=1e-5)
    assert float(cl) == pytest.approx(float(g["cluster_loss"]), rel=1e-5)
    monkeypatch.chdir(tmp_path)                      # compute() writes ./class_matrix/... like the reference
    for name, preds, hung in (("cluster", cp, True), ("linear", lp, False)):
        mt = UnSegMetrics(C, 0, hung, torch.device(DEV))
        mt.update(preds, label)
        assert np.array_equal(mt.confusion_matrix.cpu().numpy(), g[f"{name}_confusion"])
        res = mt.compute(prefix="t")
        assert float(res["iou"]) == pytest.approx(float(g[f"{name}_iou"]), rel=1e-6)
        assert float(res["accuracy"]) == pytest.approx(float(g[f"{name}_accuracy"]), rel=1e-6)
        mt.reset()
        assert int(mt.confusion_matrix.sum()) == 0
    # fused path: confusion matrices accumulated inside the probe kernel, no prediction tensors
    cc = torch.zeros(C, C, dtype=torch.long, device=DEV)
    lc = torch.zeros(C, C, dtype=torch.long, device=DEV)
    ev.predict(feat, label, cc, lc, want_preds=False)
    assert np.array_equal(cc.cpu().numpy(), g["cluster_confusion"]) and np.array_equal(lc.cpu().numpy(), g["linear_confusion"])
    # probe losses are differentiable w.r.t. the probe parameters (they train the probes)
    ev.train()
    ll, _, cl, _ = ev(feat, None, label)
    (ll + cl).backward()
    assert ev.linear_probe.weight.grad is not None and ev.cluster_probe.clusters.grad is not None


def test_knn_module_and_npz_contract(golden_dir, tmp_path):
    from equss_b200.knn import load_nns, precompute_knns, save_nns
    g = np.load(os.path.join(golden_dir, "knn.npz"))
    nns = precompute_knns(torch.from_numpy(g["feats"]).to(DEV), k=8)
    assert np.array_equal(nns.cpu().numpy(), g["idx"])
    save_nns(str(tmp_path / "nns_test.npz"), nns)
    back = load_nns(str(tmp_path / "nns_test.npz"))
    assert back.dtype == np.int64 and np.array_equal(back, g["idx"])
