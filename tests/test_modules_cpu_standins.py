"""Host logic of the nn.Module mirrors on CPU: the reference's fixtures replayed through the modules with every kernel
entry point replaced by its torch definition (tests/kernel_standins.py).  What is under test here is everything AROUND
the kernels -- stacked state and its views, which statistics feed which update, output keys and return tuples, the
write-back into the state-dict buffers, evaluation vs training branches.  The kernels themselves are held to the same
fixtures on the device (tests/test_gpu_modules.py, test_gpu_variants.py); without the stand-ins a CPU tensor raises.
tests/test_gpu_tests_replayed_on_cpu.py re-runs the module-level GPU tests themselves the same way; this file holds what
those do not reach (state-dict round trip, materialize_prob, rare branches against the oracle)."""
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
    M, K, d = 2, 1024, 256                     # config/pq_baseline.yaml:32-33,41 (embed_dims 512, num_pq 2, 1024 codes)
    _run_against_oracle(make(M, K, M * d, "l2"), O, M, K, d, "l2", [torch.randn(120, M * d) for _ in range(3)])
    M, K, d = 3, 16, 8
    _run_against_oracle(make(M, K, M * d, "none", update_norm=False), O, M, K, d, "none",
                        [torch.randn(80, M * d) for _ in range(3)], update_norm=False)
    pq = make(M, K, M * d, "z_trainable")
    stats = ([q.z_mean.clone() for q in pq.quantizers], [q.z_log_var.clone() for q in pq.quantizers])
    _run_against_oracle(pq, O, M, K, d, "z_trainable", [torch.randn(80, M * d) * 1.5 + 0.3 for _ in range(3)], z_stats=stats)
    for i, q in enumerate(pq.quantizers):
        np.testing.assert_allclose(q.z_mean.numpy(), stats[0][i].numpy(), rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(q.z_log_var.numpy(), stats[1][i].numpy(), rtol=1e-5, atol=1e-7)


def test_evaluator_probe_weight_cache_on_cpu(monkeypatch):
    """UnSegEvaluator packs both probes into one operand.  In eval() mode the packed operand is cached on the
    parameters' storage + version counters: an in-place update (optimizer step, ``copy_`` under no_grad) is seen, a write
    through ``.data`` is not until ``invalidate_cache()``; in train() mode the operand is rebuilt on every call."""
    import equss_b200  # noqa: F401
    from equss_b200 import ops
    from equss_b200.evaluator import UnSegEvaluator
    kernel_standins.install(monkeypatch)
    packs = []
    real_pack = ops.probe_pack
    monkeypatch.setattr(ops, "probe_pack", lambda w: (packs.append(1), real_pack(w))[1])
    torch.manual_seed(13)
    D, C = 16, 5
    ev = UnSegEvaluator(D, C).eval()
    feat, label = torch.randn(2, D, 6, 6), torch.randint(-1, C, (2, 24, 24))

    def fresh_preds():
        other = UnSegEvaluator(D, C).eval()
        other.load_state_dict(ev.state_dict())
        return other.predict(feat, label)

    lp0, cp0 = ev.predict(feat, label)
    ev.predict(feat, label)
    assert len(packs) == 1                                            # second call served from the cache
    with torch.no_grad():
        ev.cluster_probe.clusters.copy_(torch.randn(C, D))            # in-place: version counter moves
    lp1, cp1 = ev.predict(feat, label)
    assert len(packs) == 2 and not torch.equal(cp1, cp0) and torch.equal(lp1, lp0)
    want = fresh_preds()
    assert torch.equal(cp1, want[1]) and torch.equal(lp1, want[0])
    ev.linear_probe.weight.data = torch.randn(C, D, 1, 1)             # new storage: seen through data_ptr
    lp2, _ = ev.predict(feat, label)
    assert not torch.equal(lp2, lp1) and torch.equal(lp2, fresh_preds()[0])
    n = len(packs)
    ev.linear_probe.bias.data.add_(torch.tensor([50.0, 0, 0, 0, 0]))   # .data write: no version bump, same storage
    stale, _ = ev.predict(feat, label)
    assert len(packs) == n and torch.equal(stale, lp2)                # documented: stale until invalidated
    ev.invalidate_cache()
    lp3, _ = ev.predict(feat, label)
    assert len(packs) == n + 1 and (lp3 == 0).float().mean() > (lp2 == 0).float().mean() and torch.equal(lp3, fresh_preds()[0])
    ev.train()
    n = len(packs)
    ev.predict(feat, label); ev.predict(feat, label)
    assert len(packs) == n + 2                                        # training: rebuilt every call


def test_expansion_head_kernel_path_on_cpu(golden_dir, monkeypatch):
    """SegmentationHead without gradients: hidden = relu(W2 x + b2), code = [W1 | W3] [x ; hidden] + (b1 + b3) -- two GEMM
    calls, result a (B, D, h, w) view of NHWC memory -- against the reference module's output (fixture of
    oracle/make_golden_head.py); the packed operand follows parameter updates in eval() mode like the probe cache."""
    import equss_b200  # noqa: F401
    from equss_b200 import ops
    from equss_b200.head import SegmentationHead
    kernel_standins.install(monkeypatch)
    calls = []
    real = ops.head_gemm
    monkeypatch.setattr(ops, "head_gemm", lambda *a, **k: (calls.append((tuple(a[1].shape), k.get("relu", False))), real(*a, **k))[1])
    g = np.load(os.path.join(golden_dir, "expansion_head.npz"))
    head = SegmentationHead(64, 96).eval()
    head.load_state_dict({k.replace("_", ".", 2) if k.startswith("cluster") else k: torch.from_numpy(v)
                          for k, v in g.items() if k.startswith("cluster")}, strict=True)
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        code = head(x)
    assert calls == [((64, 64, 1, 1), True), ((96, 128), False)]                 # hidden layer, then both branches at once
    assert code.shape == (2, 96, 8, 8) and code.permute(0, 2, 3, 1).is_contiguous()      # NHWC memory: flat rows for the PQ ops
    np.testing.assert_allclose(code.numpy(), g["out"], rtol=1e-5, atol=1e-6)
    with torch.no_grad():
        head.cluster1[0].bias.add_(1.0)                                           # in-place update: seen by the cache
        np.testing.assert_allclose(head(x).numpy(), g["out"] + 1.0, rtol=1e-5, atol=1e-5)
        head.cluster2[2].bias.data.sub_(1.0)                                      # .data write: needs invalidate_cache()
        np.testing.assert_allclose(head(x).numpy(), g["out"] + 1.0, rtol=1e-5, atol=1e-5)
        head.invalidate_cache()
        np.testing.assert_allclose(head(x).numpy(), g["out"], rtol=1e-5, atol=1e-5)
    # with gradients the reference's convolutions run (autograd), same values
    xg = x.clone().requires_grad_(True)
    np.testing.assert_allclose(head(xg).detach().numpy(), g["out"], rtol=1e-5, atol=1e-5)


def test_new_vq_soft_statistics_paths_agree_on_cpu(golden_dir, monkeypatch):
    """dino_new_vq learned codebooks report jsd / entropy of the soft assignment.  Three routes must give the same
    numbers: the materialised differentiable tensor (materialize_prob=True), the differentiable tensor built only for
    the statistics (materialize_prob=False with gradients: nothing returned, but jsd / entropy carry a graph), and the
    fused statistics entry point (materialize_prob=False without gradients: no N x K*M tensor at all)."""
    import equss_b200  # noqa: F401
    from equss_b200 import codebooks as CB
    from equss_b200 import ops
    kernel_standins.install(monkeypatch)
    fused_calls = []
    real = ops.pq_soft_stats
    monkeypatch.setattr(ops, "pq_soft_stats", lambda *a, **k: (fused_calls.append(1), real(*a, **k))[1])
    g = np.load(os.path.join(golden_dir, "pq_inline_new_vq_l2.npz"))
    M, K, ts = int(g["M"]), int(g["K"]), float(g["jsd_ts"])
    D = g["z0"].shape[1]
    pq = CB.NewVQProductQuantizerWrapper(M, K, D, beta=0.25, normalize="l2", need_initialized="none", jsd_ts=ts).eval()
    with torch.no_grad():
        for i, q in enumerate(pq.quantizers):
            q.embedding.weight.copy_(torch.from_numpy(g["codebook"][i]))
    z = torch.from_numpy(g["z1"])
    zq_a, out_a, prob_a = pq(z.clone().requires_grad_(True), 1)
    assert prob_a is not None and prob_a.requires_grad and not fused_calls
    pq.materialize_prob = False
    zq_b, out_b, prob_b = pq(z.clone().requires_grad_(True), 1)
    assert prob_b is None and out_b["jsd"].requires_grad and out_b["entropy"].requires_grad and not fused_calls
    with torch.no_grad():
        zq_c, out_c, prob_c = pq(z, 1)
    assert prob_c is None and len(fused_calls) == 1
    for k in ("jsd", "entropy", "vq-loss"):
        ref = float(g[f"out1/{k}"])
        for out in (out_a, out_b, out_c):
            assert float(out[k].detach()) == pytest.approx(ref, rel=1e-5, abs=1e-8), k
    assert torch.equal(zq_a.detach(), zq_b.detach()) and torch.equal(zq_a.detach(), zq_c)


@pytest.mark.parametrize("b1,b2,b3", [(True, True, True), (False, True, False), (True, False, False), (False, False, True)])
def test_expansion_head_bias_combinations_on_cpu(monkeypatch, b1, b2, b3):
    """expansion_head on user-built 1x1-convolution stacks with any subset of biases: the packed [W1 | W3] operand and
    b1 + b3 equal the convolutions' sum (model/dino_pqgo.py:127-128); other stacks are refused with the reference layout."""
    import torch.nn as nn
    import equss_b200  # noqa: F401
    from equss_b200.head import expansion_head
    kernel_standins.install(monkeypatch)
    torch.manual_seed(53)
    C, D = 12, 20
    c1 = nn.Sequential(nn.Conv2d(C, D, (1, 1), bias=b1)).eval()
    c2 = nn.Sequential(nn.Conv2d(C, C, (1, 1), bias=b2), nn.ReLU(), nn.Conv2d(C, D, (1, 1), bias=b3)).eval()
    x = torch.randn(2, C, 5, 6)
    with torch.no_grad():
        want = c1(x) + c2(x)
        got = expansion_head(x, c1, c2)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        expansion_head(x, c2, c1)
    with pytest.raises(ValueError):
        expansion_head(x, nn.Sequential(nn.Conv2d(C, D, (3, 3), padding=1)), c2)
    with pytest.raises(ValueError):
        expansion_head(x[0], c1, c2)
