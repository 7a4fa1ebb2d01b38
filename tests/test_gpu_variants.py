"""Every PQ variant of SURVEY.md 8a (V3 train, V4 EMA + learned, V5, V6, z_trainable for V1/V2, restart / split) on the
GPU against fixtures produced by the UNMODIFIED reference modules (oracle/make_golden_variants.py), including the
gradients the reference's autograd graph yields for the soft-assignment losses.

Tolerances (BASELINE north_star): indices / counts bit-exact (fixtures contain no fp32 near-ties), floats 1e-5
relative; gradients 2e-5 of the gradient's scale (fp32 contractions in a different association order)."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _val(v):
    return float(v) if v is not None else float("nan")


def _outs(out, g, s):
    keys = {k[len(f"out{s}/"):] for k in g.files if k.startswith(f"out{s}/")}
    assert set(out.keys()) == keys, (sorted(out.keys()), sorted(keys))
    for k in keys:
        assert _val(out[k]) == pytest.approx(float(g[f"out{s}/{k}"]), rel=1e-5, abs=1e-7, nan_ok=True), (s, k)


def _close_grad(got, ref, what):
    ref = torch.from_numpy(ref)
    scale = float(ref.abs().max())
    err = float((got.detach().cpu() - ref).abs().max())
    assert err <= 2e-5 * scale + 1e-9, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("mode", ["l2", "none"])
def test_new_vq_ema_codebook_matches_reference(golden_dir, mode):
    """V4 EMA (model/dino_new_vq.py:241-459) in its wrapper: 3 train steps + eval, vq-loss / codebook-usage /
    codebook-sum / jsd / entropy, the EMA buffers, and d(total)/dz through jsd + entropy + prob + STE."""
    from equss_b200.codebooks import EMACodebook, NewVQProductQuantizerWrapper
    g = np.load(os.path.join(golden_dir, f"pq_newvq_ema_{mode}.npz"))
    M, K, ts = int(g["M"]), int(g["K"]), float(g["jsd_ts"])
    D = g["z0"].shape[1]
    pq = NewVQProductQuantizerWrapper(M, K, D, beta=0.25, normalize=mode, need_initialized="none", jsd_ts=ts,
                                      quantizer_cls=EMACodebook)
    assert sorted(pq.state_dict().keys()) == sorted(f"quantizers.{i}.codebook.{n}" for i in range(M)
                                                    for n in ("weight", "weight_avg", "vq_count"))
    with torch.no_grad():
        for i, q in enumerate(pq.quantizers):
            q.codebook.weight.copy_(torch.from_numpy(g["weight0"][i])); q.codebook.weight_avg.copy_(q.codebook.weight)
    pq = pq.to(DEV).train()
    for s in range(4):
        if s == 3:
            pq.eval()
        z = torch.from_numpy(g[f"z{s}"]).to(DEV)
        with torch.no_grad():
            zq, out, prob = pq(z, s)
        np.testing.assert_allclose(zq.cpu().numpy(), g[f"zq{s}"], rtol=1e-5, atol=1e-6)
        _outs(out, g, s)
        for name, attr in (("weight", "weight"), ("weight_avg", "weight_avg"), ("vq_count", "vq_count")):
            t = torch.stack([getattr(q.codebook, attr) for q in pq.quantizers]).cpu().numpy()
            np.testing.assert_allclose(t, g[f"{name}_after{s}"], rtol=1e-5, atol=1e-7)
        assert np.array_equal(torch.stack([q.vq_count for q in pq.quantizers]).cpu().numpy(), g[f"exact_after{s}"])
    np.testing.assert_allclose(prob.cpu().numpy(), g["prob3"], rtol=2e-5, atol=1e-7)
    # fused jsd / entropy (no N x K*M tensor) agrees with the materialised path
    pq.materialize_prob = False
    with torch.no_grad():
        _, out_f, prob_f = pq(z, 3)
    assert prob_f is None
    _outs(out_f, g, 3)
    pq.materialize_prob = True
    # gradients w.r.t. the activations (distance_prob must be differentiable)
    zg = torch.from_numpy(g["zg"]).to(DEV).requires_grad_(True)
    zq, out, prob = pq(zg, 0)
    total = ((zq * torch.from_numpy(g["go"]).to(DEV)).sum() + out["vq-loss"] + 0.3 * out["jsd"] + 0.2 * out["entropy"]
             + (prob * torch.from_numpy(g["gp"]).to(DEV)).sum())
    assert float(total) == pytest.approx(float(g["grad_total"]), rel=1e-5)
    total.backward()
    _close_grad(zg.grad, g["grad_z"], "V4 EMA dz")


@pytest.mark.parametrize("variant,mode", [("new_vq", "l2"), ("new_vq", "z_norm"), ("pqgo_cls", "l2"),
                                          ("pqgo_cls", "z_trainable"), ("pqgo", "z_norm")])
def test_inline_learned_codebooks_match_reference(golden_dir, variant, mode):
    """V4 / V5 / V6 learned codebooks through their wrappers: train + eval forward, return tuples, counts, and the
    reference's gradients w.r.t. z, the embedding and (z_trainable) the normalisation parameters."""
    from equss_b200 import codebooks as CB
    g = np.load(os.path.join(golden_dir, f"pq_inline_{variant}_{mode}.npz"))
    M, K, ts = int(g["M"]), int(g["K"]), float(g["jsd_ts"])
    D = g["z0"].shape[1]
    d = D // M
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
    pq = pq.to(DEV)
    for s, training in ((0, True), (1, False)):
        pq.train(training)
        pq.zero_grad()
        z = torch.from_numpy(g[f"z{s}"]).to(DEV).requires_grad_(True)
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
            got = torch.stack([i.reshape(-1) for i in idxs]).cpu().numpy().astype(np.int32)
            assert np.array_equal(got, g[f"idx{s}"])
        np.testing.assert_allclose(zq.detach().cpu().numpy(), g[f"zq{s}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(prob.detach().cpu().numpy(), g[f"prob{s}"], rtol=2e-5, atol=1e-7)
        _outs(out, g, s)
        exact = torch.stack([q.vq_count.to(DEV) for q in pq.quantizers]).cpu().numpy()
        assert np.array_equal(exact, g[f"exact_after{s}"])
        total = (zq * torch.from_numpy(g[f"go{s}"]).to(DEV)).sum() + out["vq-loss"] + (prob * torch.from_numpy(g[f"gp{s}"]).to(DEV)).sum()
        if variant == "new_vq":
            total = total + 0.3 * out["jsd"] + 0.2 * out["entropy"]
        total.backward()
        _close_grad(z.grad, g[f"grad_z{s}"], f"{variant}/{mode} dz step {s}")
        _close_grad(torch.stack([q.embedding.weight.grad for q in pq.quantizers]), g[f"grad_cb{s}"], f"{variant}/{mode} dcodebook")
        if mode == "z_trainable":
            _close_grad(torch.stack([q.z_mean.grad for q in pq.quantizers]), g[f"grad_zmean{s}"], "dz_mean")
            _close_grad(torch.stack([q.z_log_var.grad for q in pq.quantizers]), g[f"grad_zlogvar{s}"], "dz_log_var")
    # single module call keeps the per-variant arity (forward(z, i, it) / forward(z, z_pos) / forward(z))
    q0 = pq.quantizers[0].eval()
    z0 = torch.from_numpy(g["z1"]).to(DEV)[:, :d].contiguous()
    with torch.no_grad():
        r = q0(z0, 0, 0) if variant == "new_vq" else (q0(z0) if variant == "pqgo_cls" else q0(z0, torch.zeros_like(z0)))
    assert len(r) == (3 if variant == "new_vq" else 4)
    np.testing.assert_allclose(r[0].cpu().numpy(), g["zq1"][:, :d], rtol=1e-5, atol=1e-6)


def test_quantizer_v2_ema_training_trajectory(golden_dir):
    """V3 (model/quantizer_v2.py:253-308) in training: output gathered from z_norm rows, EMA of z_norm sums with
    per-rank statistics, embeddings / N / z_avg after each of 3 steps."""
    from equss_b200.quantizer_v2 import ProductQuantizerWrapper
    g = np.load(os.path.join(golden_dir, "pq_v2_train.npz"))
    M, K = int(g["M"]), int(g["K"])
    D = g["z0"].shape[1]
    pq = ProductQuantizerWrapper(M, K, D, beta=0.25, normalize="l2", decay=float(g["decay"]), eps=1e-5)
    with torch.no_grad():
        for i, q in enumerate(pq.quantizers):
            q.embeddings.copy_(torch.from_numpy(g["embeddings0"][i])); q.z_avg.copy_(q.embeddings)
    pq = pq.to(DEV).train()
    for s in range(3):
        with torch.no_grad():
            q, out, prob = pq(torch.from_numpy(g[f"z{s}"]).to(DEV))
        np.testing.assert_allclose(q.cpu().numpy(), g[f"q{s}"], rtol=1e-5, atol=1e-6)
        _outs(out, g, s)
        for name, attr in (("embeddings", "embeddings"), ("N", "N"), ("z_avg", "z_avg")):
            t = torch.stack([getattr(m, attr) for m in pq.quantizers]).cpu().numpy()
            np.testing.assert_allclose(t, g[f"{name}_after{s}"], rtol=1e-5, atol=1e-7)


def test_z_trainable_modes_match_reference(golden_dir):
    """normalize="z_trainable": V2 EMA with running statistics of z (model/quantizer.py:428-450; the std is taken
    before the update, the mean after) over 3 train steps + eval, and the V1 learned codebook (:129-133) with the
    reference's gradients for z, codebook, z_mean and z_log_var."""
    from equss_b200.quantizer import EMAVectorQuantizer, ProductQuantizerWrapper, VectorQuantizer
    g = np.load(os.path.join(golden_dir, "pq_ztrainable.npz"))
    M, K = int(g["M"]), int(g["K"])
    D = g["z0"].shape[1]
    pq = ProductQuantizerWrapper(M, K, D, beta=0.25, normalize="z_trainable", decay=float(g["decay"]), eps=1e-5,
                                 quantizer_cls=EMAVectorQuantizer)
    with torch.no_grad():
        for i, q in enumerate(pq.quantizers):
            q.codebook.weight.copy_(torch.from_numpy(g["weight0"][i])); q.codebook.weight_avg.copy_(q.codebook.weight)
            q.z_mean.copy_(torch.from_numpy(g["z_mean0"][i])); q.z_log_var.copy_(torch.from_numpy(g["z_log_var0"][i]))
    pq = pq.to(DEV).train()
    for s in range(4):
        pq.train(s < 3)
        z = torch.from_numpy(g[f"z{s}"]).to(DEV)
        with torch.no_grad():
            zq, out, prob = pq(z)
        np.testing.assert_allclose(zq.cpu().numpy(), g[f"zq{s}"], rtol=1e-5, atol=2e-6)
        _outs(out, g, s)
        np.testing.assert_allclose(torch.stack([q.z_mean for q in pq.quantizers]).cpu().numpy(), g[f"z_mean_after{s}"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(torch.stack([q.z_log_var for q in pq.quantizers]).cpu().numpy(), g[f"z_log_var_after{s}"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(torch.stack([q.codebook.weight for q in pq.quantizers]).cpu().numpy(), g[f"weight_after{s}"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(prob.cpu().numpy(), g["prob3"], rtol=5e-5, atol=1e-7)
    # V1
    d = g["v1_z"].shape[1]
    vq = VectorQuantizer(K, d, beta=0.25, normalize="z_trainable")
    with torch.no_grad():
        vq.codebook.weight.copy_(torch.from_numpy(g["v1_codebook"])); vq.z_mean.copy_(torch.from_numpy(g["v1_z_mean"]))
        vq.z_log_var.copy_(torch.from_numpy(g["v1_z_log_var"]))
    vq = vq.to(DEV).eval()
    zv = torch.from_numpy(g["v1_z"]).to(DEV).requires_grad_(True)
    q, out, prob = vq(zv)
    np.testing.assert_allclose(q.detach().cpu().numpy(), g["v1_q"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(prob.detach().cpu().numpy(), g["v1_prob"], rtol=5e-5, atol=1e-7)
    assert float(out["loss"]) == pytest.approx(float(g["v1_loss"]), rel=1e-5)
    (out["loss"] + (q * torch.from_numpy(g["v1_go"]).to(DEV)).sum()).backward()
    _close_grad(zv.grad, g["v1_grad_z"], "V1 z_trainable dz")
    _close_grad(vq.codebook.weight.grad, g["v1_grad_cb"], "V1 z_trainable dcodebook")
    _close_grad(vq.z_mean.grad, g["v1_grad_zmean"], "V1 dz_mean")
    _close_grad(vq.z_log_var.grad, g["v1_grad_zlogvar"], "V1 dz_log_var")


def test_restart_paths(golden_dir):
    """Host-RNG maintenance paths with Python's `random` seeded like the fixture run: rand init (V2), in-forward
    restart of dead codes (V5 from z_norm rows, V4 from raw z rows), prepare_restart + restart() (V2)."""
    from equss_b200.codebooks import Codebook
    from equss_b200.quantizer import EMAVectorQuantizer
    g = np.load(os.path.join(golden_dir, "pq_restart.npz"))
    K = int(g["K"])
    d = g["a_z"].shape[1]
    ema = EMAVectorQuantizer(K, d, beta=0.25, normalize="l2", need_initialized="rand").to(DEV).train()
    random.seed(5)
    with torch.no_grad():
        q, out, _ = ema(torch.from_numpy(g["a_z"]).to(DEV))
    np.testing.assert_allclose(q.cpu().numpy(), g["a_q"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ema.codebook.weight.cpu().numpy(), g["a_weight_after"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(ema.codebook.weight_avg.cpu().numpy(), g["a_weight_avg_after"], rtol=1e-5, atol=1e-7)
    assert float(out["loss"]) == pytest.approx(float(g["a_loss"]), rel=1e-5)
    zb = torch.from_numpy(g["b_z"]).to(DEV)
    for variant, key in (("pqgo", "b"), ("new_vq", "c")):
        cb = Codebook(K, d, beta=0.25, book=1.0, normalize="l2", use_restart=True, need_initialized="none", variant=variant)
        with torch.no_grad():
            cb.embedding.weight.copy_(torch.from_numpy(g["b_weight0"]))
        cb = cb.to(DEV).train()
        random.seed(9)
        with torch.no_grad():
            r = cb(zb, torch.zeros_like(zb)) if variant == "pqgo" else cb(zb, 0, 0)
        np.testing.assert_allclose(r[0].cpu().numpy(), g[f"{key}_q"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(cb.embedding.weight.detach().cpu().numpy(), g[f"{key}_weight_after"], rtol=1e-6, atol=1e-7)
        assert float(r[1]["vq-loss"]) == pytest.approx(float(g[f"{key}_vq_loss"]), rel=1e-5)
        if variant == "pqgo":
            assert np.array_equal(r[3].cpu().numpy(), g["b_idx"])
            assert float(r[1]["codebook-usage"]) == pytest.approx(float(g["b_usage"]), rel=1e-6)
            assert np.array_equal(cb.vq_count.cpu().numpy(), g["b_count_after"])
    ema2 = EMAVectorQuantizer(K, d, beta=0.25, normalize="l2", use_restart=True)
    with torch.no_grad():
        ema2.codebook.weight.copy_(torch.from_numpy(g["b_weight0"])); ema2.codebook.weight_avg.copy_(ema2.codebook.weight)
    ema2 = ema2.to(DEV).train()
    random.seed(11)
    with torch.no_grad():
        ema2(torch.from_numpy(g["d_z"]).to(DEV))
    ema2.restart()
    np.testing.assert_allclose(ema2.codebook.weight.cpu().numpy(), g["d_weight_after"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(ema2.codebook.weight_avg.cpu().numpy(), g["d_weight_avg_after"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(ema2.codebook.vq_count.cpu().numpy(), g["d_vq_count_after"], rtol=1e-6, atol=1e-7)


def test_split_dead_codes_semantics():
    """use_split (model/quantizer.py:330-381): the device RNG differs from the fixture host's, so the check is on
    the invariants -- which codes move, halved counts / sums, +/- the same small jitter, exact counter cleared."""
    from equss_b200.quantizer import EMAVectorQuantizer
    torch.manual_seed(4)
    K, d = 16, 8
    q = EMAVectorQuantizer(K, d, normalize="l2", use_split=True).to(DEV).train()
    with torch.no_grad():
        q.codebook.weight.copy_(torch.randn(K, d)); q.codebook.weight_avg.copy_(torch.randn(K, d))
        q.codebook.vq_count.copy_(torch.arange(K, 0, -1).float())         # code 0 busiest
    w0, a0, c0 = q.codebook.weight.clone(), q.codebook.weight_avg.clone(), q.codebook.vq_count.clone()
    count = torch.ones(K, device=DEV); count[[5, 9, 12]] = 0
    q.vq_count.fill_(7)
    assert q.split(count) == 3
    w, a, c = q.codebook.weight, q.codebook.weight_avg, q.codebook.vq_count
    untouched = [k for k in range(K) if k not in (0, 1, 2, 5, 9, 12)]
    assert torch.equal(w[untouched], w0[untouched]) and torch.equal(c[untouched], c0[untouched])
    assert torch.equal(c[[0, 1, 2]], c0[[0, 1, 2]] / 2) and torch.equal(a[[0, 1, 2]], a0[[0, 1, 2]] / 2)
    dead_sorted = sorted([5, 9, 12], key=lambda k: float(c[k]), reverse=True)      # partner j has count c0[j]/2
    for j, k in enumerate(dead_sorted):
        assert float(c[k]) == float(c0[j]) / 2 and torch.equal(a[k], a0[j] / 2)
        jitter = w[k] - w0[j]
        torch.testing.assert_close(w[j], w0[j] - jitter, rtol=0, atol=1e-6)
        assert 0 < float(jitter.abs().max()) < 0.2
    assert float(q.vq_count.sum()) == 0.0
    assert q.split(torch.ones(K, device=DEV)) == 0


def test_ema_backward_uses_pre_update_codebook():
    """with normalize="none" the gather source aliased the live codebook, which the EMA update
    overwrites before backward().  The commitment gradient must use the codebook of the forward pass."""
    from equss_b200.quantizer import EMAVectorQuantizer, ProductQuantizerWrapper
    torch.manual_seed(8)
    M, K, D, n = 2, 16, 32, 300
    pq = ProductQuantizerWrapper(M, K, D, beta=0.25, normalize="none", quantizer_cls=EMAVectorQuantizer).to(DEV).train()
    with torch.no_grad():
        for q in pq.quantizers:
            q.codebook.weight.copy_(torch.randn(K, D // M) * 0.5); q.codebook.weight_avg.copy_(q.codebook.weight)
    w_before = torch.stack([q.codebook.weight.clone() for q in pq.quantizers])
    z = (torch.randn(n, D, device=DEV) * 0.5).requires_grad_(True)
    zq, out, _ = pq(z)
    assert not torch.equal(torch.stack([q.codebook.weight for q in pq.quantizers]), w_before)   # the update did run
    go = torch.randn_like(zq)
    ((zq * go).sum() + out["loss"]).backward()
    d = D // M
    zr = z.detach().view(n, M, d)
    dist = ((zr.unsqueeze(2) - w_before.unsqueeze(0)) ** 2).sum(-1)
    qsel = torch.gather(w_before.unsqueeze(0).expand(n, M, K, d), 2, dist.argmin(-1)[..., None, None].expand(n, M, 1, d)).squeeze(2)
    expect = go.view(n, M, d) + 0.25 * 2.0 * (zr - qsel) / (n * d) / M
    torch.testing.assert_close(z.grad.view(n, M, d), expect, rtol=1e-5, atol=1e-7)


def test_channel_moments_and_soft_stats_kernels():
    """K13 against torch reductions (flat and NCHW); fused jsd / entropy against the materialised formulation at
    d in {16, 64}, ragged tile counts, temperature != 1."""
    from equss_b200 import _pq_core as core
    from equss_b200 import ops
    torch.manual_seed(12)
    for shape in ((777, 96), (3, 96, 13, 11)):
        z = torch.randn(*shape, device=DEV) * 1.7 + 0.3
        mom = ops.channel_moments(z)
        zf = z if z.dim() == 2 else z.permute(0, 2, 3, 1).reshape(-1, shape[1])
        torch.testing.assert_close(mom[0], zf.double().mean(0).float(), rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(mom[1], (zf.double() ** 2).mean(0).float(), rtol=1e-6, atol=1e-6)
    for M, K, d, n, mode, ts in ((3, 200, 16, 1234, "l2", 0.5), (2, 256, 64, 70, "z_norm", 1.0), (4, 32, 8, 4096, "none", 2.0)):
        z = torch.randn(n, M * d, device=DEV) * (0.3 if mode == "none" else 1.0)
        cb = torch.randn(M, K, d, device=DEV) * (0.3 if mode == "none" else 1.0)
        cbn = core.normalize_codebook(cb, mode)
        jsd, ent = ops.pq_soft_stats(z, cbn, None, mode, None, None, ts)
        prob = ops.pq_distance_prob(z, cbn, None, mode, None, None, ts)
        jr, er = core.soft_assignment_stats(prob.double(), M, K)
        assert float(jsd) == pytest.approx(float(jr), rel=1e-5), (M, K, d)
        assert float(ent) == pytest.approx(float(er), rel=1e-5), (M, K, d)


def test_stego_loss_matches_reference(golden_dir):
    """STEGOLoss.helper on the reference's fixture (loss tensor, code correlation, gradients w.r.t. both
    code maps; pointwise / zero_clamp / stabilize variants), and the full forward against the oracle's formulation run
    on the same device with the same generator state (coordinate draws + fixed-point-free permutations)."""
    import equss_oracle as O
    from equss_b200.losses import STEGOLoss
    g = np.load(os.path.join(golden_dir, "stego.npz"))
    cfg = {"pointwise": True, "zero_clamp": True, "stabilize": False, "feature_samples": 11, "neg_samples": 2,
           "pos_intra_shift": float(g["pos_intra_shift"]), "pos_inter_shift": float(g["pos_inter_shift"]),
           "neg_inter_shift": float(g["neg_inter_shift"]), "pos_intra_weight": float(g["pos_intra_weight"]),
           "pos_inter_weight": float(g["pos_inter_weight"]), "neg_inter_weight": float(g["neg_inter_weight"])}
    f1, f2 = torch.from_numpy(g["f1"]).to(DEV), torch.from_numpy(g["f2"]).to(DEV)
    for tag, cfgv in (("a", cfg), ("b", dict(cfg, pointwise=False, zero_clamp=False, stabilize=True))):
        mod = STEGOLoss(cfgv)
        c1 = torch.from_numpy(g["c1"]).to(DEV).requires_grad_(True)
        c2 = torch.from_numpy(g["c2"]).to(DEV).requires_grad_(True)
        loss, cd = mod.helper(f1, f2, c1, c2, 0.2)
        np.testing.assert_allclose(cd.detach().cpu().numpy(), g[f"{tag}_cd"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(loss.detach().cpu().numpy(), g[f"{tag}_loss"], rtol=1e-5, atol=2e-6)
        loss.mean().backward()
        _close_grad(c1.grad, g[f"{tag}_g1"], f"stego {tag} d code")
        _close_grad(c2.grad, g[f"{tag}_g2"], f"stego {tag} d code_pos")
    torch.manual_seed(5)
    n, Cf, Cc = 4, 384, 70
    feats, feats_pos = torch.randn(n, Cf, 28, 28, device=DEV), torch.randn(n, Cf, 28, 28, device=DEV)
    code = torch.randn(n, Cc, 28, 28, device=DEV, requires_grad=True)
    code_pos = torch.randn(n, Cc, 28, 28, device=DEV)
    torch.manual_seed(6)
    got = STEGOLoss(cfg)(feats, feats_pos, code, code_pos)
    torch.manual_seed(6)
    ref = O.stego_forward(cfg, feats, feats_pos, code.detach(), code_pos)
    assert float(got) == pytest.approx(float(ref), rel=1e-5, abs=1e-7)
    got.backward()
    assert code.grad is not None and float(code.grad.abs().sum()) > 0
