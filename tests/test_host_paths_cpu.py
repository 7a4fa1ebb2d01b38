"""Host-side PyTorch pieces of the research flags on CPU tensors, against the fixtures produced by the
unmodified reference (oracle/make_golden_flags.py): the Gumbel index draw and the weighted sum are plain torch code and
need no device; the kernels around them are covered by tests/test_gpu_flags.py."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F


def _fixed_gumbel(noises):
    calls = [0]

    def fake(logits, tau=1.0, hard=True, dim=1):
        y = logits + noises[calls[0]]
        calls[0] += 1
        return F.one_hot(torch.argmax(y, dim=1), logits.shape[1]).to(logits.dtype)
    return fake, calls


def test_gumbel_indices_follow_the_reference_draw(golden_dir, monkeypatch):
    import equss_b200  # noqa: F401
    from equss_b200 import _pq_core as core
    from equss_b200._host_paths import gumbel_indices
    g = np.load(os.path.join(golden_dir, "pq_flag_ema_gumbel_l2.npz"))
    M = int(g["M"])
    z, w0 = torch.from_numpy(g["z0"]), torch.from_numpy(g["weight0"])
    fake, calls = _fixed_gumbel(list(torch.from_numpy(g["noise0"])))
    monkeypatch.setattr(torch.nn.functional, "gumbel_softmax", fake)
    z_norm = core._normalize_rows(core._rows(z, M), "l2", None, None)
    idx = gumbel_indices(z_norm, F.normalize(w0, dim=2), 0.01)                  # EMAVectorQuantizer: -distance / 0.01
    assert calls[0] == M and idx.dtype == torch.int32 and tuple(idx.shape) == (M, z.shape[0])
    assert np.array_equal(idx.numpy(), g["idx0"])
    # VectorQuantizer: plain -distance, one subspace, NCHW rows
    g = np.load(os.path.join(golden_dir, "pq_flag_param_gumbel.npz"))
    z = torch.from_numpy(g["z"])
    fake, calls = _fixed_gumbel([torch.from_numpy(g["noise"])])
    monkeypatch.setattr(torch.nn.functional, "gumbel_softmax", fake)
    z_norm = core._normalize_rows(core._rows(z, 1), "l2", None, None)
    idx = gumbel_indices(z_norm, F.normalize(torch.from_numpy(g["weight"]), dim=1).unsqueeze(0), None)
    assert calls[0] == 1 and np.array_equal(idx[0].numpy(), g["idx"])


def test_weighted_sum_matches_reference_values_and_gradients(golden_dir):
    import equss_b200  # noqa: F401
    from equss_b200.codebooks import _weighted_sum
    for variant, book in (("new_vq", 1.0), ("pqgo", 1.0)):
        g = np.load(os.path.join(golden_dir, f"pq_flag_inline_{variant}_weighted.npz"))
        K, ts = int(g["K"]), float(g["jsd_ts"])
        z = torch.from_numpy(g["z"]).requires_grad_(True)
        w = torch.from_numpy(g["weight"]).clone().requires_grad_(True)
        zf = z.permute(0, 2, 3, 1).reshape(-1, z.shape[1])
        dist = (zf ** 2).sum(1, keepdim=True) + (w ** 2).sum(1) - 2 * zf @ w.t()
        prob = F.softmax(-dist / ts, dim=1)                    # what the kernel's DistanceProb returns on the device
        out, commit, cb_loss = _weighted_sum(z, prob, w.unsqueeze(0), "none", None, None, 1, K)
        np.testing.assert_allclose(out.detach().numpy(), g["zq"], rtol=1e-5, atol=1e-7)
        loss = (book * cb_loss + 0.25 * commit).mean()
        assert abs(float(loss.detach()) - float(g["out/vq-loss"])) <= 1e-6 * abs(float(g["out/vq-loss"])) + 1e-9
        ((out * torch.from_numpy(g["go"])).sum() + loss).backward()
        np.testing.assert_allclose(z.grad.numpy(), g["grad_z"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(w.grad.numpy(), g["grad_w"], rtol=1e-4, atol=1e-6)


def _fixed_keep(draws, p):
    calls = [0]

    def fake(num_codes, prob, device):
        u = draws[calls[0]]
        calls[0] += 1
        assert u.shape == (num_codes,) and prob == p
        return u.to(device) > p
    return fake, calls


def test_pq_dropout_assignment_follows_the_reference(golden_dir, monkeypatch):
    """``pq_dropout`` (dino_new_vq.EMACodebook, fixture of oracle/make_golden_dropout.py): with the reference's uniform
    draws injected, the host assignment yields the reference's indices (positions in the kept list), its soft
    assignment (one column per kept code), its jsd / entropy and its usage ratio; gathering the FULL raw codebook at
    those indices gives the reference's quantised output."""
    import equss_b200  # noqa: F401
    from equss_b200 import _host_paths as hp
    from equss_b200 import _pq_core as core
    from equss_b200.codebooks import _Dropped
    g = np.load(os.path.join(golden_dir, "pq_flag_newvq_ema_dropout.npz"))
    M, K, p, ts = int(g["M"]), int(g["K"]), float(g["pq_dropout"]), float(g["jsd_ts"])
    weight = torch.from_numpy(g["weight0"])
    for s in range(3):
        z = torch.from_numpy(g[f"z{s}"])
        B, D, h, w = z.shape
        d = D // M
        fake, calls = _fixed_keep(list(torch.from_numpy(g[f"u{s}"])), p)
        monkeypatch.setattr(hp, "dropout_keep_mask", fake)
        zr = core._normalize_rows(core._rows(z, M), "l2", None, None)
        idx, probs, keeps = hp.dropout_assign(zr, F.normalize(weight, dim=2), p, ts)
        assert calls[0] == M and idx.dtype == torch.int32
        assert np.array_equal(idx.numpy(), g[f"idx{s}"])
        assert [int(k.sum()) for k in keeps] == [int(c) for c in (g[f"u{s}"] > p).sum(axis=1)]
        np.testing.assert_allclose(torch.cat(probs, dim=-1).numpy(), g[f"prob{s}"], rtol=1e-5, atol=1e-7)
        drop = _Dropped(probs, keeps)
        jsd, ent = drop.soft_stats()
        assert abs(float(jsd) - float(g[f"out{s}/jsd"])) <= 1e-5 * abs(float(g[f"out{s}/jsd"])) + 1e-8
        assert abs(float(ent) - float(g[f"out{s}/entropy"])) <= 1e-5 * abs(float(g[f"out{s}/entropy"])) + 1e-8
        q = torch.stack([weight[i][idx[i].long()] for i in range(M)], dim=1)          # full raw codebook, kept-list indices
        out = (zr + (q - zr)).reshape(B, h, w, D).permute(0, 3, 1, 2)
        np.testing.assert_allclose(out.numpy(), g[f"zq{s}"], rtol=1e-5, atol=1e-6)
        commit = ((zr - q) ** 2).mean(dim=(0, 2)).mean()
        assert abs(0.25 * float(commit) - float(g[f"out{s}/vq-loss"])) <= 1e-5 * float(g[f"out{s}/vq-loss"])
        if s < 2:
            count = torch.stack([torch.bincount(idx[i].long(), minlength=K) for i in range(M)])
            usage = ((drop.kept - (count == 0).sum(dim=1).float()) / drop.kept).mean()
            assert abs(float(usage) - float(g[f"out{s}/codebook-usage"])) <= 1e-6
            weight = torch.from_numpy(g[f"weight_after{s}"])                          # the EMA update itself is the kernel's


def test_pq_dropout_gradients_through_the_host_assignment(golden_dir, monkeypatch):
    """Learned Codebook of dino_new_vq / dino_pqgo with pq_dropout (and with the weighted sum): the host assignment's
    soft assignment carries the reference's gradients to z and to the codebook; the straight-through gather and the two
    losses (the kernels' part on the device) are written out in torch here."""
    import equss_b200  # noqa: F401
    from equss_b200 import _host_paths as hp
    from equss_b200 import _pq_core as core
    from equss_b200.codebooks import _Dropped
    for name in ("new_vq", "pqgo", "new_vq_weighted"):
        g = np.load(os.path.join(golden_dir, f"pq_flag_inline_{name}_dropout.npz"))
        K, p, ts, mode, variant = int(g["K"]), float(g["pq_dropout"]), float(g["jsd_ts"]), str(g["mode"]), str(g["variant"])
        weighted = bool(g["weighted"])
        z = torch.from_numpy(g["z"]).requires_grad_(True)
        wt = torch.from_numpy(g["weight"]).clone().requires_grad_(True)
        B, d, h, w = z.shape
        fake, calls = _fixed_keep([torch.from_numpy(g["u"])], p)
        monkeypatch.setattr(hp, "dropout_keep_mask", fake)
        zr = core._normalize_rows(core._rows(z, 1), mode, None, None)
        cbn = core.normalize_codebook(wt.unsqueeze(0), mode, ema_style=True)
        idx, probs, keeps = hp.dropout_assign(zr, cbn, p, ts)
        assert calls[0] == 1 and np.array_equal(idx[0].numpy(), g["idx"])
        np.testing.assert_allclose(probs[0].detach().numpy(), g["prob"].reshape(B * h * w, -1), rtol=1e-5, atol=1e-7)
        zn = zr[:, 0]
        q = probs[0] @ cbn[0][keeps[0]] if weighted else wt[idx[0].long()]
        cb_loss, commit = ((q - zn.detach()) ** 2).mean(), ((zn - q.detach()) ** 2).mean()
        out = q if weighted else zn + (q - zn).detach()
        out = out.reshape(B, h, w, d).permute(0, 3, 1, 2)
        np.testing.assert_allclose(out.detach().numpy(), g["zq"], rtol=1e-5, atol=1e-6)
        vq_loss = cb_loss + 0.25 * commit
        assert abs(float(vq_loss.detach()) - float(g["out/vq-loss"])) <= 1e-5 * abs(float(g["out/vq-loss"]))
        total = (out * torch.from_numpy(g["go"])).sum() + vq_loss + (probs[0] * torch.from_numpy(g["gp"])).sum()
        if variant == "new_vq":
            jsd, ent = _Dropped(probs, keeps).soft_stats()
            assert abs(float(jsd) - float(g["out/jsd"])) <= 1e-5 * abs(float(g["out/jsd"])) + 1e-8
            total = total + 0.3 * jsd + 0.2 * ent
        total.backward()
        np.testing.assert_allclose(z.grad.numpy(), g["grad_z"], rtol=2e-4, atol=2e-6)
        np.testing.assert_allclose(wt.grad.numpy(), g["grad_w"], rtol=2e-4, atol=2e-6)


def _emulate_kernels(monkeypatch):
    """Torch stand-ins for the four kernel entry points the pq_dropout branch reaches (gather + losses with the
    straight-through estimator, scatter-add, EMA update), so that the MODULE plumbing around the host assignment can
    run without a device.  The kernels themselves are pinned by the GPU tests."""
    from equss_b200 import _pq_core as core
    from equss_b200 import ops

    def pq_quantize(z, codebook_norm, gather_src, normalize, norm_a=None, norm_b=None, *, want_prob=True,
                    temperature=1.0, algo=0, cnorm2=None, idx=None):
        assert idx is not None and not want_prob          # the dropout branch always brings its own indices
        M = gather_src.shape[0]
        zr = core._normalize_rows(core._rows(z.float(), M), normalize, norm_a, norm_b)
        q = torch.stack([gather_src[i][idx[i].long()] for i in range(M)], dim=1)
        commit = ((zr - q.detach()) ** 2).mean(dim=(0, 2))
        cb = ((q - zr.detach()) ** 2).mean(dim=(0, 2))
        out = zr + (q - zr).detach()
        B, D, h, w = z.shape
        return idx, out.reshape(B, h, w, D).permute(0, 3, 1, 2).contiguous(), commit, cb, None

    def pq_accumulate(z, idx, K, **kw):
        M = idx.shape[0]
        zr = core._rows(z, M)
        packed = torch.zeros(M, K, zr.shape[2] + 1)
        for i in range(M):
            packed[i, :, :-1].index_add_(0, idx[i].long(), zr[:, i])
            packed[i, :, -1] = torch.bincount(idx[i].long(), minlength=K).float()
        return packed

    def ema_update(packed, decay, eps, vqc, wavg, wnew, exact):
        count, total = packed[:, :, -1], packed[:, :, :-1]
        K = count.shape[1]
        exact += count
        vqc.mul_(decay).add_(count, alpha=1 - decay)
        wavg.mul_(decay).add_(total, alpha=1 - decay)
        n = vqc.sum(dim=1, keepdim=True)
        wnew.copy_(wavg / ((vqc + eps) / (n + K * eps) * n).unsqueeze(-1))
        return (count == 0).sum(dim=1)

    monkeypatch.setattr(core, "pq_quantize", pq_quantize)
    monkeypatch.setattr(ops, "pq_accumulate", pq_accumulate)
    monkeypatch.setattr(ops, "ema_update", ema_update)


def test_pq_dropout_module_plumbing_with_emulated_kernels(golden_dir, monkeypatch):
    """NewVQProductQuantizerWrapper(EMACodebook, pq_dropout) and the learned Codebook variants end to end on CPU with
    the kernels replaced by torch stand-ins: return tuples, ragged soft-assignment widths, usage over the kept codes,
    jsd / entropy, the EMA trajectory and the gradients equal the unmodified reference's."""
    import equss_b200  # noqa: F401
    from equss_b200 import _host_paths as hp
    from equss_b200.codebooks import (Codebook, EMACodebook, NewVQProductQuantizerWrapper, PQGOProductQuantizerWrapper)
    _emulate_kernels(monkeypatch)
    g = np.load(os.path.join(golden_dir, "pq_flag_newvq_ema_dropout.npz"))
    M, K, p, ts = int(g["M"]), int(g["K"]), float(g["pq_dropout"]), float(g["jsd_ts"])
    D = g["z0"].shape[1]
    pq = NewVQProductQuantizerWrapper(M, K, D, beta=0.25, normalize="l2", jsd_ts=ts, pq_dropout=p, quantizer_cls=EMACodebook)
    with torch.no_grad():
        for i, q in enumerate(pq.quantizers):
            q.codebook.weight.copy_(torch.from_numpy(g["weight0"][i])); q.codebook.weight_avg.copy_(q.codebook.weight)
    pq.train()
    for s in range(3):
        if s == 2:
            pq.eval()
        fake, calls = _fixed_keep(list(torch.from_numpy(g[f"u{s}"])), p)
        monkeypatch.setattr(hp, "dropout_keep_mask", fake)
        with torch.no_grad():
            zq, out, prob = pq(torch.from_numpy(g[f"z{s}"]), s)
        assert calls[0] == M                                   # drawn in evaluation too, like the reference
        np.testing.assert_allclose(zq.numpy(), g[f"zq{s}"], rtol=1e-5, atol=1e-6)
        assert prob.shape == g[f"prob{s}"].shape
        np.testing.assert_allclose(prob.numpy(), g[f"prob{s}"], rtol=1e-5, atol=1e-7)
        keys = {k.split("/", 1)[1] for k in g.files if k.startswith(f"out{s}/")}
        assert set(out.keys()) == keys
        for k in keys:
            assert abs(float(out[k]) - float(g[f"out{s}/{k}"])) <= 1e-5 * abs(float(g[f"out{s}/{k}"])) + 1e-7, (s, k)
        for name, get in (("weight", lambda q: q.codebook.weight), ("weight_avg", lambda q: q.codebook.weight_avg),
                          ("vq_count", lambda q: q.codebook.vq_count), ("exact", lambda q: q.vq_count)):
            got = torch.stack([get(q) for q in pq.quantizers]).numpy()
            np.testing.assert_allclose(got, g[f"{name}_after{s}"], rtol=1e-5, atol=1e-6)

    for name in ("new_vq", "pqgo", "new_vq_weighted"):
        g = np.load(os.path.join(golden_dir, f"pq_flag_inline_{name}_dropout.npz"))
        K, p, ts, mode, variant = int(g["K"]), float(g["pq_dropout"]), float(g["jsd_ts"]), str(g["mode"]), str(g["variant"])
        z = torch.from_numpy(g["z"]).requires_grad_(True)
        B, d, h, w = z.shape
        cb = Codebook(K, d, beta=0.25, normalize=mode, need_initialized="none", jsd_ts=ts, pq_dropout=p,
                      use_weighted_sum=bool(g["weighted"]), variant=variant)
        with torch.no_grad():
            cb.embedding.weight.copy_(torch.from_numpy(g["weight"]))
        cb.train()
        fake, calls = _fixed_keep([torch.from_numpy(g["u"])], p)
        monkeypatch.setattr(hp, "dropout_keep_mask", fake)
        res = cb(z, 0, 0) if variant == "new_vq" else cb(z, torch.zeros_like(z))
        zq, out, prob = res[0], res[1], res[2]
        assert tuple(prob.shape) == tuple(g["prob"].shape)     # (n, kept) for new_vq, (b, h, w, kept) for pqgo
        np.testing.assert_allclose(zq.detach().numpy(), g["zq"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(prob.detach().numpy(), g["prob"], rtol=1e-5, atol=1e-7)
        if variant == "pqgo":
            assert np.array_equal(res[3].numpy(), g["ridx"])
        for k in {k.split("/", 1)[1] for k in g.files if k.startswith("out/")}:
            assert abs(float(out[k]) - float(g[f"out/{k}"])) <= 1e-5 * abs(float(g[f"out/{k}"])) + 1e-7, (name, k)
        np.testing.assert_allclose(cb.vq_count.numpy(), g["exact_after"])
        total = (zq * torch.from_numpy(g["go"])).sum() + out["vq-loss"] + (prob.reshape(B * h * w, -1) * torch.from_numpy(g["gp"])).sum()
        if variant == "new_vq":
            total = total + 0.3 * out["jsd"] + 0.2 * out["entropy"]
        total.backward()
        np.testing.assert_allclose(z.grad.numpy(), g["grad_z"], rtol=2e-4, atol=2e-6)
        np.testing.assert_allclose(cb.embedding.weight.grad.numpy(), g["grad_w"], rtol=2e-4, atol=2e-6)

    # the wrapper of dino_pqgo concatenates the ragged per-subspace soft assignments on the last axis
    fake, calls = _fixed_keep([torch.rand(8) for _ in range(2)], 0.25)
    monkeypatch.setattr(hp, "dropout_keep_mask", fake)
    wrap = PQGOProductQuantizerWrapper(2, 8, 8, normalize="l2", pq_dropout=0.25)
    zq, (z_split, zqs, idxs), out, prob = wrap(torch.randn(2, 8, 3, 3), torch.zeros(2, 8, 3, 3))
    assert calls[0] == 2 and zq.shape == (2, 8, 3, 3) and prob.shape[:3] == (2, 3, 3) and prob.shape[3] <= 16
    assert idxs[0].shape == (2, 3, 3) and "codebook-usage" in out


def test_kmeans_initialisation_matches_the_reference(golden_dir, monkeypatch):
    """need_initialized="kmeans" of the inline classes (fixture: oracle/make_golden_dropout.py::kmeans_init_case): the
    first training call replaces the codebook by scikit-learn's centroids of the batch rows (random_state=0), then the
    step runs on them; the flag is cleared."""
    import equss_b200  # noqa: F401
    from equss_b200 import _host_paths as hp
    from equss_b200.codebooks import Codebook, EMACodebook, _init_codebooks
    g = np.load(os.path.join(golden_dir, "pq_init_kmeans.npz"))
    K, z = int(g["K"]), torch.from_numpy(g["z"])
    d = z.shape[1]
    rows = z.permute(0, 2, 3, 1).reshape(-1, d)
    np.testing.assert_allclose(hp.kmeans_centroids(rows, K).numpy(), g["centroids"], rtol=1e-5, atol=1e-6)
    cb = Codebook(K, d, beta=0.25, normalize="l2", need_initialized="kmeans", variant="pqgo").train()
    _init_codebooks([cb], z, d)
    assert cb.need_initialized == "none"
    np.testing.assert_allclose(cb.embedding.weight.detach().numpy(), g["centroids"], rtol=1e-5, atol=1e-6)
    # in evaluation mode nothing is initialised (dino_pqgo.py:589)
    cb2 = Codebook(K, d, normalize="l2", need_initialized="kmeans", variant="pqgo").eval()
    assert cb2.need_initialized == "kmeans"
    # EMACodebook: centroids into weight and weight_avg, then the step's EMA update (kernels emulated in torch)
    _emulate_kernels(monkeypatch)
    from equss_b200 import _pq_core as core
    real_q = core.pq_quantize                                 # the emulation insists on external indices; give it the argmin

    def quantize_top1(zz, cbn, src, mode, na=None, nb=None, *, want_prob=True, temperature=1.0, **kw):
        zr = core._normalize_rows(core._rows(zz.float(), cbn.shape[0]), mode, na, nb)
        dist = ((zr.unsqueeze(2) - cbn.unsqueeze(0)) ** 2).sum(-1)                    # (n, M, K)
        idx = dist.argmin(dim=2).t().to(torch.int32).contiguous()
        res = real_q(zz, cbn, src, mode, na, nb, want_prob=False, idx=idx)
        prob = F.softmax(-dist / temperature, dim=2).reshape(zr.shape[0], -1) if want_prob else None
        return res[0], res[1], res[2], res[3], prob
    monkeypatch.setattr(core, "pq_quantize", quantize_top1)
    ema = EMACodebook(K, d, beta=0.25, normalize="l2", need_initialized="kmeans").train()
    with torch.no_grad():
        zq, out, prob = ema(z, 0, 0)
    assert ema.need_initialized == "none"
    np.testing.assert_allclose(zq.numpy(), g["ema_zq"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ema.codebook.weight.numpy(), g["ema_weight_after"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ema.codebook.weight_avg.numpy(), g["ema_weight_avg_after"], rtol=1e-5, atol=1e-6)
    for k in ("vq-loss", "codebook-usage", "jsd", "entropy", "codebook-sum"):
        assert abs(float(out[k]) - float(g[f"ema_out/{k}"])) <= 1e-5 * abs(float(g[f"ema_out/{k}"])) + 1e-7, k


@pytest.mark.parametrize("mode", ["l2", "z_norm", "none", "z_trainable"])
@pytest.mark.parametrize("nchw", [False, True])
def test_distance_prob_backward_matches_autograd(monkeypatch, mode, nchw):
    """_pq_core.DistanceProb: forward is the kernel, backward is hand-derived (softmax and distance Jacobians as batched
    contractions, routed through the row normalisation).  With the kernel's forward replaced by its torch definition,
    the gradients w.r.t. the activations, the (normalised) codebook and the z_trainable vectors must equal autograd's on
    the plain formula  softmax(-(|z_n|^2 + |c|^2 - 2 z_n.c) / T)  (model/quantizer.py:457-468), for flat and NCHW input."""
    import equss_b200  # noqa: F401
    from equss_b200 import _pq_core as core
    from equss_b200 import ops
    torch.manual_seed(8)
    M, K, d, T = 3, 7, 4, 0.7
    D = M * d

    def formula(z, cbn, a, b):
        zr = core._normalize_rows(core._rows(z.float(), M), mode, a, b)                 # (n, M, d)
        dist = (zr ** 2).sum(-1, keepdim=True) + (cbn ** 2).sum(-1).unsqueeze(0) - 2 * torch.einsum("nmd,mkd->nmk", zr, cbn)
        return torch.softmax(-dist / T, dim=2).reshape(zr.shape[0], M * K)

    def kernel_stub(z, cbn, cn2, normalize, norm_a=None, norm_b=None, temperature=1.0):
        with torch.no_grad():
            return formula(z, cbn, norm_a, norm_b)
    monkeypatch.setattr(ops, "pq_distance_prob", kernel_stub)

    z0 = torch.randn(2, D, 3, 5) if nchw else torch.randn(30, D)
    c0 = torch.randn(M, K, d) * 0.7
    a0 = torch.randn(D) * 0.1 if mode == "z_trainable" else None
    b0 = torch.rand(D) + 0.5 if mode == "z_trainable" else None
    g = torch.randn(z0.numel() // D, M * K)
    grads = []
    for fn in ("custom", "autograd"):
        z, c = z0.clone().requires_grad_(True), c0.clone().requires_grad_(True)
        a = a0.clone().requires_grad_(True) if a0 is not None else None
        b = b0.clone().requires_grad_(True) if b0 is not None else None
        if fn == "custom":
            prob = core.DistanceProb.apply(z, c, (c.detach() ** 2).sum(-1), mode, a, b, T)
        else:
            prob = formula(z, c, a, b)
        (prob * g).sum().backward()
        grads.append([t.grad for t in (z, c, a, b) if t is not None])
    for name, got, want in zip(("z", "codebook", "norm_a", "norm_b"), *grads):
        assert got is not None and got.shape == want.shape, name
        assert torch.allclose(got, want, rtol=2e-4, atol=1e-6 * float(want.abs().max()) + 1e-9), (mode, nchw, name)


def test_affine_parameter_gradients_from_the_activation_gradient():
    """z_trainable: z_norm = (z - a) / b per channel.  The backward kernel returns grad_z = g_znorm / b; the gradients
    of the two vectors follow from it on the host (_pq_core._affine_param_grads) -- checked against autograd."""
    import equss_b200  # noqa: F401
    from equss_b200 import _pq_core as core
    torch.manual_seed(9)
    M, d = 3, 4
    for z in (torch.randn(20, M * d), torch.randn(2, M * d, 3, 2)):
        z = z.clone().requires_grad_(True)
        a = (torch.randn(M * d) * 0.2).requires_grad_(True)
        b = (torch.rand(M * d) + 0.5).requires_grad_(True)
        zn = core._normalize_rows(core._rows(z, M), "z_trainable", a, b)
        (zn * torch.randn_like(zn)).sum().backward()
        ga, gb = core._affine_param_grads(z.grad, z.detach(), M, a.detach(), b.detach(), True, True)
        assert torch.allclose(ga, a.grad, rtol=1e-5, atol=1e-6) and torch.allclose(gb, b.grad, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag", ["f48", "f768"])
def test_knn_reference_table_test_body_on_cpu(golden_dir, tmp_path, monkeypatch, tag):
    """tests/test_gpu_knn_ref.py re-run on CPU tensors with ``ops.knn_topk`` replaced by its torch definition: the test
    body, the fixture handling and the ``precompute_knns`` / ``save_nns`` / ``load_nns`` plumbing (not the kernels)."""
    import test_gpu_knn_ref as T
    from equss_b200 import ops

    def knn_topk(queries, db, k, return_sims=False):
        sims, idx = torch.topk(torch.einsum("nf,mf->nm", queries, db), k)
        return (idx, sims) if return_sims else idx
    monkeypatch.setattr(ops, "knn_topk", knn_topk)
    monkeypatch.setattr(T, "DEV", "cpu")
    T.test_knn_equals_reference_statements(golden_dir, tmp_path, tag)
