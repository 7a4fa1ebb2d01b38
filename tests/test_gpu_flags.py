"""The research flags use_weighted_sum / use_gumbel of the nn.Module mirrors against fixtures produced
by the unmodified reference (oracle/make_golden_flags.py).  The Gumbel noise is pinned by replacing
torch.nn.functional.gumbel_softmax -- the function the mirror calls exactly like the reference does, once per subspace
on (n, K) logits -- with the fixture's pre-drawn noise."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _val(v):
    return float(v) if v is not None else float("nan")


class _FixedGumbel:
    def __init__(self, noises):
        self.noises, self.calls, self.shapes = noises, 0, []

    def __call__(self, logits, tau=1.0, hard=True, dim=1):
        assert tau == 1.0 and hard and dim == 1
        self.shapes.append(tuple(logits.shape))
        y = logits + self.noises[self.calls].to(logits.device)
        self.calls += 1
        return F.one_hot(torch.argmax(y, dim=1), logits.shape[1]).to(logits.dtype)


def _load_ema(pq, w0, M):
    sd = pq.state_dict()
    for i in range(M):
        sd[f"quantizers.{i}.codebook.weight"] = w0[i].clone()
        sd[f"quantizers.{i}.codebook.weight_avg"] = w0[i].clone()
    pq.load_state_dict(sd, strict=True)


def _check_ema_state(pq, g, s, count_tol):
    w = torch.stack([q.codebook.weight for q in pq.quantizers]).cpu().numpy()
    np.testing.assert_allclose(w, g[f"weight_after{s}"], rtol=2e-5, atol=1e-7)
    wa = torch.stack([q.codebook.weight_avg for q in pq.quantizers]).cpu().numpy()
    np.testing.assert_allclose(wa, g[f"weight_avg_after{s}"], rtol=2e-5, atol=1e-7)
    c = torch.stack([q.codebook.vq_count for q in pq.quantizers]).cpu().numpy()
    np.testing.assert_allclose(c, g[f"vq_count_after{s}"], rtol=count_tol, atol=1e-7)
    e = torch.stack([q.vq_count for q in pq.quantizers]).cpu().numpy()
    np.testing.assert_allclose(e, g[f"exact_after{s}"], rtol=count_tol, atol=1e-6)


@pytest.mark.parametrize("mode", ["none", "l2"])
def test_ema_weighted_sum_matches_reference(golden_dir, mode):
    """model/quantizer.py:470-471,483-484,534: soft sum of the codes, soft EMA statistics, no straight-through."""
    from equss_b200.quantizer import EMAVectorQuantizer, ProductQuantizerWrapper
    g = np.load(os.path.join(golden_dir, f"pq_flag_ema_weighted_{mode}.npz"))
    M, K = int(g["M"]), int(g["K"])
    D = g["z0"].shape[1]
    pq = ProductQuantizerWrapper(M, K, D, beta=0.25, normalize=mode, decay=0.99, eps=1e-5, use_weighted_sum=True,
                                 quantizer_cls=EMAVectorQuantizer)
    _load_ema(pq, torch.from_numpy(g["weight0"]), M)
    pq = pq.to(DEV).train()
    for s in range(4):
        if s == 3:
            pq.eval()
        zq, out, prob = pq(torch.from_numpy(g[f"z{s}"]).to(DEV))
        np.testing.assert_allclose(zq.cpu().numpy(), g[f"zq{s}"], rtol=2e-5, atol=2e-5 * float(np.abs(g[f"zq{s}"]).max()))
        keys = {k[len(f"out{s}/"):] for k in g.files if k.startswith(f"out{s}/")}
        assert set(out.keys()) == keys
        for k in keys:
            # the usage ranks of fractional counts may move by one position at a threshold crossing
            tol = dict(abs=1.0 / K + 1e-7) if k.startswith(("total-", "current-")) else dict(rel=2e-5, abs=1e-7)
            assert _val(out[k]) == pytest.approx(float(g[f"out{s}/{k}"]), nan_ok=True, **tol), (s, k)
        _check_ema_state(pq, g, s, 2e-5)
    # gradients reach z through the soft assignment (no straight-through on this path)
    zg = torch.from_numpy(g["zg"]).to(DEV).requires_grad_(True)
    zq, out, _ = pq(zg)
    np.testing.assert_allclose(zq.detach().cpu().numpy(), g["zq_g"], rtol=2e-5, atol=1e-6)
    ((zq * torch.from_numpy(g["go"]).to(DEV)).sum() + out["loss"]).backward()
    ref = g["grad_z"]
    np.testing.assert_allclose(zg.grad.cpu().numpy(), ref, rtol=1e-3, atol=2e-5 * float(np.abs(ref).max()))


def test_ema_gumbel_matches_reference(golden_dir, monkeypatch):
    """model/quantizer.py:463-465: training indices drawn through F.gumbel_softmax(-distance / 0.01), per subspace."""
    from equss_b200.quantizer import EMAVectorQuantizer, ProductQuantizerWrapper
    g = np.load(os.path.join(golden_dir, "pq_flag_ema_gumbel_l2.npz"))
    M, K, steps = int(g["M"]), int(g["K"]), int(g["steps"])
    n, D = g["z0"].shape
    pq = ProductQuantizerWrapper(M, K, D, beta=0.25, normalize="l2", decay=0.99, eps=1e-5, use_gumbel=True,
                                 quantizer_cls=EMAVectorQuantizer)
    _load_ema(pq, torch.from_numpy(g["weight0"]), M)
    pq = pq.to(DEV).train()
    for s in range(steps + 1):
        if s == steps:
            pq.eval()
        else:
            fake = _FixedGumbel(list(torch.from_numpy(g[f"noise{s}"])))
            monkeypatch.setattr(torch.nn.functional, "gumbel_softmax", fake)
        zq, out, prob = pq(torch.from_numpy(g[f"z{s}"]).to(DEV))
        monkeypatch.undo()
        if s < steps:
            assert fake.calls == M and fake.shapes == [(n, K)] * M          # one draw per subspace, like the reference loop
        np.testing.assert_allclose(zq.cpu().numpy(), g[f"zq{s}"], rtol=1e-5, atol=1e-6)
        for k in {k[len(f"out{s}/"):] for k in g.files if k.startswith(f"out{s}/")}:
            assert _val(out[k]) == pytest.approx(float(g[f"out{s}/{k}"]), rel=1e-5, abs=1e-7, nan_ok=True), (s, k)
        _check_ema_state(pq, g, s, 1e-6)
    assert int(g["flips"]) > 0
    # unseeded: the draw really is stochastic and stays close to the argmin (logits are distances / 0.01)
    pq.train()
    z = torch.from_numpy(g["z0"]).to(DEV)
    a = pq(z)[0]
    b = pq(z)[0]
    assert not torch.equal(a, b)


def test_param_gumbel_matches_reference(golden_dir, monkeypatch):
    """model/quantizer.py:145-147: VectorQuantizer draws from plain -distance in training."""
    from equss_b200.quantizer import VectorQuantizer
    g = np.load(os.path.join(golden_dir, "pq_flag_param_gumbel.npz"))
    z = torch.from_numpy(g["z"]).to(DEV)
    K, d = int(g["K"]), z.shape[1]
    vq = VectorQuantizer(K, d, beta=0.25, normalize="l2", use_gumbel=True).to(DEV).train()
    with torch.no_grad():
        vq.codebook.weight.copy_(torch.from_numpy(g["weight"]))
    fake = _FixedGumbel([torch.from_numpy(g["noise"])])
    monkeypatch.setattr(torch.nn.functional, "gumbel_softmax", fake)
    q, out, prob = vq(z)
    monkeypatch.undo()
    assert fake.calls == 1
    np.testing.assert_allclose(q.detach().cpu().numpy(), g["zq"], rtol=1e-5, atol=1e-6)
    for k in ("loss", "codebook_loss", "commitment_loss"):
        assert float(out[k]) == pytest.approx(float(g[f"out/{k}"]), rel=1e-5)
    assert np.array_equal(vq.vq_count.cpu().numpy(), g["exact_after"])
    vq.eval()                                           # eval ignores the flag
    q_eval = vq(z)[0]
    assert not torch.equal(q_eval, q.detach())


def test_new_vq_ema_weighted_sum_matches_reference(golden_dir):
    """dino_new_vq.EMACodebook with use_weighted_sum (:400-401,438): soft sum out, HARD statistics into the EMA."""
    from equss_b200.codebooks import EMACodebook, NewVQProductQuantizerWrapper
    g = np.load(os.path.join(golden_dir, "pq_flag_newvq_ema_weighted.npz"))
    M, K, ts = int(g["M"]), int(g["K"]), float(g["jsd_ts"])
    D = g["z0"].shape[1]
    pq = NewVQProductQuantizerWrapper(M, K, D, beta=0.25, normalize="none", need_initialized="none", jsd_ts=ts,
                                      use_weighted_sum=True, quantizer_cls=EMACodebook).to(DEV)
    w0 = torch.from_numpy(g["weight0"]).to(DEV)
    with torch.no_grad():
        for i, q in enumerate(pq.quantizers):
            q.codebook.weight.copy_(w0[i]); q.codebook.weight_avg.copy_(w0[i])
    pq.train()
    for s in range(3):
        if s == 2:
            pq.eval()
        with torch.no_grad():
            zq, out, prob = pq(torch.from_numpy(g[f"z{s}"]).to(DEV), s)
        # a soft sum of K codes: fp32 reassociation noise relative to the scale of the rows, not to each element
        np.testing.assert_allclose(zq.cpu().numpy(), g[f"zq{s}"], rtol=2e-5, atol=2e-5 * float(np.abs(g[f"zq{s}"]).max()))
        for k in {k[len(f"out{s}/"):] for k in g.files if k.startswith(f"out{s}/")}:
            assert _val(out[k]) == pytest.approx(float(g[f"out{s}/{k}"]), rel=2e-5, abs=1e-7), (s, k)
        w = torch.stack([q.codebook.weight for q in pq.quantizers]).cpu().numpy()
        np.testing.assert_allclose(w, g[f"weight_after{s}"], rtol=2e-5, atol=1e-8)


@pytest.mark.parametrize("variant", ["new_vq", "pqgo"])
def test_inline_codebook_weighted_sum_matches_reference(golden_dir, variant):
    """Learned inline Codebook with use_weighted_sum (dino_new_vq.py:616-617, dino_pqgo.py:661-662): output, loss and the
    gradients w.r.t. z and the embedding from the reference's own autograd graph."""
    from equss_b200.codebooks import Codebook
    g = np.load(os.path.join(golden_dir, f"pq_flag_inline_{variant}_weighted.npz"))
    K, ts = int(g["K"]), float(g["jsd_ts"])
    z = torch.from_numpy(g["z"]).to(DEV).requires_grad_(True)
    cb = Codebook(K, z.shape[1], beta=0.25, normalize="none", need_initialized="none", jsd_ts=ts, use_weighted_sum=True,
                  variant=variant).to(DEV).train()
    with torch.no_grad():
        cb.embedding.weight.copy_(torch.from_numpy(g["weight"]))
    res = cb(z, 0, 0) if variant == "new_vq" else cb(z, torch.zeros_like(z))
    q, out = res[0], res[1]
    np.testing.assert_allclose(q.detach().cpu().numpy(), g["zq"], rtol=2e-5, atol=2e-5 * float(np.abs(g["zq"]).max()))
    for k in {k[len("out/"):] for k in g.files if k.startswith("out/")}:
        assert _val(out[k]) == pytest.approx(float(g[f"out/{k}"]), rel=2e-5, abs=1e-7), k
    ((q * torch.from_numpy(g["go"]).to(DEV)).sum() + out["vq-loss"]).backward()
    for got, ref in ((z.grad, g["grad_z"]), (cb.embedding.weight.grad, g["grad_w"])):
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=1e-3, atol=2e-5 * float(np.abs(ref).max()))


def test_unreproducible_flags_still_fail_loudly():
    from equss_b200.codebooks import Codebook, EMACodebook
    with pytest.raises(NotImplementedError):
        Codebook(8, 4, pq_dropout=0.1, variant="pqgo_cls")           # dino_pqgo_cls.py has no pq_dropout
    with pytest.raises(NotImplementedError):
        Codebook(8, 4, need_initialized="faiss")
    assert Codebook(8, 4, use_split=True).use_split                 # accepted and unused, as in the reference
    with pytest.raises(AssertionError):
        EMACodebook(8, 4, normalize="l2", use_weighted_sum=True)     # dino_new_vq.py:276-277
    with pytest.raises(AssertionError):
        Codebook(8, 4, use_gumbel=True)                              # dino_pqgo.py:502-503
