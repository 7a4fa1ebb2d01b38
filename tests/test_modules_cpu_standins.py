"""Host logic of the nn.Module mirrors on CPU: the reference's fixtures replayed through the modules with every kernel
entry point replaced by its torch definition (tests/kernel_standins.py).  What is under test here is everything AROUND
the kernels -- stacked state and its views, which statistics feed which update, output keys and return tuples, the
write-back into the state-dict buffers, evaluation vs training branches.  The kernels themselves are held to the same
fixtures on the device (tests/test_gpu_modules.py, test_gpu_variants.py); without the stand-ins a CPU tensor raises."""
import os

import numpy as np
import pytest
import torch

import kernel_standins


def _val(v):
    return float(v) if v is not None else float("nan")


@pytest.mark.parametrize("mode", ["l2", "z_norm", "none"])
def test_ema_wrapper_trajectory_on_cpu(golden_dir, monkeypatch, mode):
    """ProductQuantizerWrapper(EMAVectorQuantizer), 3 training steps + 1 evaluation step (model/quantizer.py:383-611)."""
    import equss_b200  # noqa: F401
    from equss_b200.quantizer import EMAVectorQuantizer, ProductQuantizerWrapper
    kernel_standins.install(monkeypatch)
    g = np.load(os.path.join(golden_dir, f"pq_ema_{mode}.npz"))
    M, K = int(g["M"]), int(g["K"])
    D = g["z0"].shape[1]
    pq = ProductQuantizerWrapper(M, K, D, beta=0.25, normalize=mode, decay=0.99, eps=1e-5, quantizer_cls=EMAVectorQuantizer)
    sd = pq.state_dict()
    w0 = torch.from_numpy(g["weight0"])
    for i in range(M):
        sd[f"quantizers.{i}.codebook.weight"] = w0[i].clone()
        sd[f"quantizers.{i}.codebook.weight_avg"] = w0[i].clone()
    pq.load_state_dict(sd, strict=True)
    pq.train()
    for s in range(4):
        if s == 3:
            pq.eval()
        with torch.no_grad():
            zq, out, prob = pq(torch.from_numpy(g[f"z{s}"]))
        np.testing.assert_allclose(zq.numpy(), g[f"zq{s}"], rtol=1e-5, atol=1e-6)
        keys = {k[len(f"out{s}/"):] for k in g.files if k.startswith(f"out{s}/")}
        assert set(out.keys()) == keys
        for k in keys:
            assert _val(out[k]) == pytest.approx(float(g[f"out{s}/{k}"]), rel=1e-5, abs=1e-7, nan_ok=True), (s, k)
        for name, get in (("weight", lambda q: q.codebook.weight), ("weight_avg", lambda q: q.codebook.weight_avg),
                          ("vq_count", lambda q: q.codebook.vq_count), ("exact", lambda q: q.vq_count)):
            got = torch.stack([get(q) for q in pq.quantizers]).numpy()
            np.testing.assert_allclose(got, g[f"{name}_after{s}"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(prob.numpy(), g["prob3"], rtol=1e-5, atol=1e-7)
    # the state dict exposes per-subspace views of the stacked storage: a round trip keeps both consistent
    sd2 = pq.state_dict()
    np.testing.assert_allclose(sd2["quantizers.2.codebook.weight"].numpy(), g["weight_after3"][2], rtol=1e-5, atol=1e-6)
    pq2 = ProductQuantizerWrapper(M, K, D, beta=0.25, normalize=mode, quantizer_cls=EMAVectorQuantizer)
    pq2.load_state_dict(sd2, strict=True)
    with torch.no_grad():
        zq2, _, _ = pq2.eval()(torch.from_numpy(g["z3"]))
    assert torch.equal(zq2, zq)
    # materialize_prob = False: same outputs, no soft assignment
    pq2.materialize_prob = False
    with torch.no_grad():
        zq3, _, none = pq2(torch.from_numpy(g["z3"]))
    assert none is None and torch.equal(zq3, zq)


def test_learned_and_v2_variants_on_cpu(golden_dir, monkeypatch):
    """VectorQuantizer (V1), dino_pqgo.Codebook (V5) alone and inside its wrapper, quantizer_v2 (V3), evaluation mode."""
    import equss_b200  # noqa: F401
    from equss_b200.codebooks import Codebook, PQGOProductQuantizerWrapper
    from equss_b200.quantizer import VectorQuantizer
    from equss_b200.quantizer_v2 import EMAVectorQuantizer as V2EMA
    kernel_standins.install(monkeypatch)
    g = np.load(os.path.join(golden_dir, "pq_param_nchw.npz"))
    z = torch.from_numpy(g["z"])
    K, d = int(g["K"]), z.shape[1]
    vq = VectorQuantizer(K, d, beta=0.25, normalize="l2").eval()
    with torch.no_grad():
        vq.codebook.weight.copy_(torch.from_numpy(g["v1_codebook"]))
        q, out, prob = vq(z)
    np.testing.assert_allclose(q.numpy(), g["v1_q"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(prob.numpy(), g["v1_prob"], rtol=1e-5, atol=1e-7)
    assert float(out["loss"]) == pytest.approx(float(g["v1_loss"]), rel=1e-5)
    assert float(out["codebook_loss"]) == pytest.approx(float(g["v1_codebook_loss"]), rel=1e-5)
    cb = Codebook(K, d, beta=0.25, book=1.0, normalize="none", need_initialized="none").eval()
    with torch.no_grad():
        cb.embedding.weight.copy_(torch.from_numpy(g["v5_codebook"]))
        q5, out5, prob5, idx5 = cb(z, torch.zeros_like(z))
    assert np.array_equal(idx5.numpy(), g["v5_idx"])
    np.testing.assert_allclose(q5.numpy(), g["v5_q"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(prob5.numpy(), g["v5_prob"], rtol=1e-5, atol=1e-7)
    assert float(out5["vq-loss"]) == pytest.approx(float(g["v5_vq_loss"]), rel=1e-5)
    w = PQGOProductQuantizerWrapper(1, K, d, normalize="none").eval()
    with torch.no_grad():
        w.quantizers[0].embedding.weight.copy_(torch.from_numpy(g["v5_codebook"]))
        zq, (zs, zqs, idxs), outs, probs = w(z)
    assert np.array_equal(idxs[0].numpy(), g["v5_idx"]) and tuple(probs.shape) == tuple(g["v5_prob"].shape)
    assert len(zs) == 1 and len(zqs) == 1 and torch.equal(zqs[0], zq)
    g2 = np.load(os.path.join(golden_dir, "pq_v2_nchw.npz"))
    v2 = V2EMA(K, d, beta=0.25).eval()
    with torch.no_grad():
        v2.embeddings.copy_(torch.from_numpy(g2["embeddings"]))
        q2, out2, _ = v2(z)
    np.testing.assert_allclose(q2.numpy(), g2["q"], rtol=1e-5, atol=1e-6)
    assert float(out2["loss"]) == pytest.approx(float(g2["loss"]), rel=1e-5)


@pytest.mark.parametrize("mode", ["l2", "none"])
def test_new_vq_ema_wrapper_trajectory_on_cpu(golden_dir, monkeypatch, mode):
    """dino_new_vq.ProductQuantizerWrapper(EMACodebook), V4: raw codebook gathered, EMA sums of raw z, jsd / entropy."""
    import equss_b200  # noqa: F401
    from equss_b200.codebooks import EMACodebook, NewVQProductQuantizerWrapper
    kernel_standins.install(monkeypatch)
    g = np.load(os.path.join(golden_dir, f"pq_newvq_ema_{mode}.npz"))
    M, K, ts = int(g["M"]), int(g["K"]), float(g["jsd_ts"])
    steps = sorted(int(k[1:]) for k in g.files if k[0] == "z" and k[1:].isdigit())
    D = g["z0"].shape[1]
    pq = NewVQProductQuantizerWrapper(M, K, D, beta=0.25, normalize=mode, need_initialized="none", jsd_ts=ts, quantizer_cls=EMACodebook)
    with torch.no_grad():
        for i, q in enumerate(pq.quantizers):
            q.codebook.weight.copy_(torch.from_numpy(g["weight0"][i])); q.codebook.weight_avg.copy_(q.codebook.weight)
    pq.train()
    for s in steps:
        if f"out{s}/codebook-usage" not in g.files:
            pq.eval()
        with torch.no_grad():
            zq, out, prob = pq(torch.from_numpy(g[f"z{s}"]), s)
        np.testing.assert_allclose(zq.numpy(), g[f"zq{s}"], rtol=1e-5, atol=1e-6)
        keys = {k[len(f"out{s}/"):] for k in g.files if k.startswith(f"out{s}/")}
        assert set(out.keys()) == keys
        for k in keys:
            assert _val(out[k]) == pytest.approx(float(g[f"out{s}/{k}"]), rel=1e-5, abs=1e-7), (s, k)
        w = torch.stack([q.codebook.weight for q in pq.quantizers]).numpy()
        np.testing.assert_allclose(w, g[f"weight_after{s}"], rtol=1e-5, atol=1e-6)


def test_without_the_standins_a_cpu_tensor_is_refused():
    import equss_b200
    from equss_b200.quantizer import ProductQuantizerWrapper
    pq = ProductQuantizerWrapper(2, 8, 8, normalize="l2").eval()
    with pytest.raises(equss_b200._native.EqussNativeError):
        pq(torch.randn(4, 8))


def _run_against_oracle(pq, O, M, K, d, mode, zs, *, update_norm=True, z_stats=None):
    """Replays `zs` (training steps, the last one in evaluation mode) through the wrapper and through the oracle's
    per-subspace restatement of EMAVectorQuantizer.forward; every output and every buffer must agree."""
    w0 = torch.stack([q.codebook.weight.clone() for q in pq.quantizers])
    states = [O.EmaState(w0[i]) for i in range(M)]
    exact = [torch.zeros(K) for _ in range(M)]
    pq.train()
    for s, z in enumerate(zs):
        training = s < len(zs) - 1
        if not training:
            pq.eval()
        with torch.no_grad():
            zq, out, prob = pq(z)
        qs, oout = [], {}
        for i in range(M):
            kw = {}
            if z_stats is not None:
                kw = {"z_mean": z_stats[0][i], "z_log_var": z_stats[1][i]}
            q, o, p, ix = O.ema_vq_forward(z[:, i * d:(i + 1) * d], states[i], exact[i], normalize=mode, beta=0.25,
                                           training=training, update_norm=update_norm, **kw)
            qs.append(q)
            for k, v in o.items():
                oout[k] = v if i == 0 else oout[k] + v
        oout = {k: v / M for k, v in oout.items()}
        np.testing.assert_allclose(zq.numpy(), torch.cat(qs, dim=1).numpy(), rtol=1e-5, atol=1e-6)
        assert set(out.keys()) == set(oout.keys()), (s, sorted(out), sorted(oout))
        for k, v in oout.items():
            assert float(out[k]) == pytest.approx(float(v), rel=1e-5, abs=1e-7), (s, k)
        for i, q in enumerate(pq.quantizers):
            np.testing.assert_allclose(q.codebook.weight.numpy(), states[i].weight.numpy(), rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(q.codebook.weight_avg.numpy(), states[i].weight_avg.numpy(), rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(q.codebook.vq_count.numpy(), states[i].vq_count.numpy(), rtol=1e-5, atol=1e-7)
            assert torch.equal(q.vq_count, exact[i])


def test_rare_branches_of_the_ema_wrapper_on_cpu(monkeypatch):
    """Branches the fixtures do not reach, against the oracle (itself pinned to the reference bit for bit): more than
    1024 codes (no fused tail: EMA update + percentile calls), ``update_norm=False`` (raw codebook gathered from a
    snapshot taken before the in-place update), and the z_trainable running statistics (model/quantizer.py:428-450)."""
    import equss_oracle as O
    import equss_b200  # noqa: F401
    from equss_b200.quantizer import EMAVectorQuantizer, ProductQuantizerWrapper
    kernel_standins.install(monkeypatch)
    torch.manual_seed(12)

    def make(M, K, D, mode, **kw):
        pq = ProductQuantizerWrapper(M, K, D, beta=0.25, normalize=mode, quantizer_cls=EMAVectorQuantizer, **kw)
        with torch.no_grad():
            for q in pq.quantizers:
                q.codebook.weight.copy_(torch.randn(K, D // M)); q.codebook.weight_avg.copy_(q.codebook.weight)
        return pq

    M, K, d = 2, 1030, 4
    _run_against_oracle(make(M, K, M * d, "l2"), O, M, K, d, "l2", [torch.randn(96, M * d) for _ in range(3)])
    M, K, d = 3, 16, 8
    _run_against_oracle(make(M, K, M * d, "none", update_norm=False), O, M, K, d, "none",
                        [torch.randn(80, M * d) for _ in range(3)], update_norm=False)
    pq = make(M, K, M * d, "z_trainable")
    stats = ([q.z_mean.clone() for q in pq.quantizers], [q.z_log_var.clone() for q in pq.quantizers])
    _run_against_oracle(pq, O, M, K, d, "z_trainable", [torch.randn(80, M * d) * 1.5 + 0.3 for _ in range(3)], z_stats=stats)
    for i, q in enumerate(pq.quantizers):
        np.testing.assert_allclose(q.z_mean.numpy(), stats[0][i].numpy(), rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(q.z_log_var.numpy(), stats[1][i].numpy(), rtol=1e-5, atol=1e-7)


def _close_grad(got, want, what):
    want = torch.from_numpy(want)
    assert got is not None and got.shape == want.shape, what
    assert torch.allclose(got, want, rtol=1e-3, atol=2e-5 * float(want.abs().max()) + 1e-9), \
        (what, float((got - want).abs().max()), float(want.abs().max()))


@pytest.mark.parametrize("variant,mode", [("new_vq", "l2"), ("new_vq", "z_norm"), ("pqgo_cls", "l2"),
                                          ("pqgo_cls", "z_trainable"), ("pqgo", "z_norm")])
def test_inline_learned_codebooks_with_gradients_on_cpu(golden_dir, monkeypatch, variant, mode):
    """V4 / V5 / V6 learned codebooks through their wrappers (training + evaluation call): return tuples, counts, losses,
    and the gradients of the reference's own graph w.r.t. z, the embedding and the z_trainable parameters -- i.e. the
    autograd plumbing of PQGatherLoss / DistanceProb around the (here: stand-in) forward and backward kernels."""
    import equss_b200  # noqa: F401
    from equss_b200 import codebooks as CB
    kernel_standins.install(monkeypatch)
    g = np.load(os.path.join(golden_dir, f"pq_inline_{variant}_{mode}.npz"))
    M, K, ts = int(g["M"]), int(g["K"]), float(g["jsd_ts"])
    D = g["z0"].shape[1]
    if variant == "new_vq":
        pq = CB.NewVQProductQuantizerWrapper(M, K, D, beta=0.25, normalize=mode, need_initialized="none", jsd_ts=ts)
    elif variant == "pqgo_cls":
        pq = CB.PQGOClsProductQuantizerWrapper(M, K, D, beta=0.25, normalize=mode, need_initialized="none", jsd_ts=ts)
    else:
        pq = CB.PQGOProductQuantizerWrapper(M, K, D, beta=0.25, book=0.6, normalize=mode, need_initialized="none", jsd_ts=ts)
    with torch.no_grad():
        for i, q in enumerate(pq.quantizers):
            q.embedding.weight.copy_(torch.from_numpy(g["codebook"][i]))
            if mode == "z_trainable":
                q.z_mean.copy_(torch.from_numpy(g["z_mean"][i])); q.z_log_var.copy_(torch.from_numpy(g["z_log_var"][i]))
    for s, training in ((0, True), (1, False)):
        pq.train(training)
        pq.zero_grad()
        z = torch.from_numpy(g[f"z{s}"]).requires_grad_(True)
        B, _, h, w = z.shape
        if variant == "new_vq":
            zq, out, prob = pq(z, s)
            idxs = None
        elif variant == "pqgo_cls":
            zq, out, prob, idxs = pq(z)
            assert all(i.shape == (B * h * w,) for i in idxs)
        else:
            zq, (zs, zqs, idxs), out, prob = pq(z, torch.zeros_like(z))
            assert all(i.shape == (B, h, w) for i in idxs) and len(zs) == M and len(zqs) == M
        assert tuple(prob.shape) == tuple(g[f"prob{s}"].shape)
        if idxs is not None:
            assert np.array_equal(torch.stack([i.reshape(-1) for i in idxs]).numpy().astype(np.int32), g[f"idx{s}"])
        np.testing.assert_allclose(zq.detach().numpy(), g[f"zq{s}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(prob.detach().numpy(), g[f"prob{s}"], rtol=1e-5, atol=1e-7)
        keys = {k[len(f"out{s}/"):] for k in g.files if k.startswith(f"out{s}/")}
        assert set(out.keys()) == keys
        for k in keys:
            assert float(out[k].detach() if torch.is_tensor(out[k]) else out[k]) == pytest.approx(float(g[f"out{s}/{k}"]), rel=1e-5, abs=1e-7), (s, k)
        assert np.array_equal(torch.stack([q.vq_count for q in pq.quantizers]).numpy(), g[f"exact_after{s}"])
        total = (zq * torch.from_numpy(g[f"go{s}"])).sum() + out["vq-loss"] + (prob * torch.from_numpy(g[f"gp{s}"])).sum()
        if variant == "new_vq":
            total = total + 0.3 * out["jsd"] + 0.2 * out["entropy"]
        total.backward()
        _close_grad(z.grad, g[f"grad_z{s}"], f"{variant}/{mode} dz step {s}")
        _close_grad(torch.stack([q.embedding.weight.grad for q in pq.quantizers]), g[f"grad_cb{s}"], f"{variant}/{mode} dcodebook")
        if mode == "z_trainable":
            _close_grad(torch.stack([q.z_mean.grad for q in pq.quantizers]), g[f"grad_zmean{s}"], "dz_mean")
            _close_grad(torch.stack([q.z_log_var.grad for q in pq.quantizers]), g[f"grad_zlogvar{s}"], "dz_log_var")
