"""Differential test of the module mirrors' host logic against the LIVE reference modules on CPU, over configurations the
fixtures do not enumerate (normalisation x update_norm x weighted sum x restart / split x layouts x sizes).

Both sides get the same state dict, the same inputs and the same seeds of ``random`` / torch; the mirror runs with the
kernel entry points replaced by their torch definitions (tests/kernel_standins.py), so every difference found here is a
difference in the Python around the kernels.  Needs /root/reference (build container only)."""
import itertools
import os
import random
import sys

import numpy as np
import pytest
import torch

import kernel_standins

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")


@pytest.fixture()
def ref(monkeypatch):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    from make_golden import import_reference
    import_reference()
    import model.dino_new_vq as nv
    import model.dino_pqgo as pqgo
    import model.dino_pqgo_cls as pcls
    import model.quantizer as q1
    import equss_b200  # noqa: F401
    kernel_standins.install(monkeypatch)
    return {"q1": q1, "nv": nv, "pqgo": pqgo, "pcls": pcls}


def _f(v):
    if v is None:
        return float("nan")
    return float(v.detach()) if torch.is_tensor(v) else float(v)


def _same_outputs(ro, mo, what):
    assert set(ro.keys()) == set(mo.keys()), (what, sorted(ro), sorted(mo))
    for k in ro:
        a, b = _f(ro[k]), _f(mo[k])
        assert (np.isnan(a) and np.isnan(b)) or b == pytest.approx(a, rel=2e-5, abs=1e-7), (what, k, a, b)


def _same_state(r, m, what):
    rs, ms = r.state_dict(), m.state_dict()
    assert sorted(rs) == sorted(ms), what
    for k in rs:
        assert torch.allclose(ms[k], rs[k], rtol=2e-5, atol=1e-6), (what, k, float((ms[k] - rs[k]).abs().max()))


@pytest.mark.parametrize("mode,update_norm,weighted,restart,split", [
    ("l2", True, False, False, False), ("l2", False, False, True, False), ("z_norm", True, False, False, True),
    ("none", False, True, False, False), ("none", True, True, False, False), ("z_trainable", True, False, True, True),
    ("z_norm", False, False, True, True), ("l2", True, True, False, False)])
def test_ema_wrapper_against_the_live_reference(ref, mode, update_norm, weighted, restart, split):
    """model/quantizer.py ProductQuantizerWrapper(EMAVectorQuantizer): 3 training steps + 1 evaluation step."""
    from equss_b200.quantizer import EMAVectorQuantizer, ProductQuantizerWrapper
    q1 = ref["q1"]
    torch.manual_seed(21)
    M, K, d, n = 3, 12, 6, 90
    kw = dict(beta=0.3, normalize=mode, decay=0.95, eps=1e-4, use_restart=restart, use_split=split,
              use_weighted_sum=weighted, update_norm=update_norm)
    r = q1.ProductQuantizerWrapper(M, K, M * d, quantizer_cls=q1.EMAVectorQuantizer, **kw)
    m = ProductQuantizerWrapper(M, K, M * d, quantizer_cls=EMAVectorQuantizer, **kw)
    with torch.no_grad():
        for q in r.quantizers:
            q.codebook.weight.copy_(torch.randn(K, d) * (0.05 if mode == "none" else 1.0)); q.codebook.weight_avg.copy_(q.codebook.weight)
    m.load_state_dict(r.state_dict(), strict=True)
    r.train(); m.train()
    for step in range(4):
        if step == 3:
            r.eval(); m.eval()
        z = torch.randn(n, M * d) * (0.05 if mode == "none" else 1.0) + 0.01 * step
        outs = []
        for mod in (r, m):
            random.seed(100 + step); torch.manual_seed(200 + step)
            with torch.no_grad():
                outs.append(mod(z))
        (rq, ro, rp), (mq, mo, mp) = outs
        assert torch.allclose(mq, rq, rtol=2e-5, atol=2e-6), (step, float((mq - rq).abs().max()))
        assert torch.allclose(mp, rp, rtol=2e-5, atol=1e-7), step
        _same_outputs(ro, mo, f"step {step}")
        _same_state(r, m, f"step {step}")
        for qr, qm in zip(r.quantizers, m.quantizers):
            if weighted:                                     # soft counts: fp32 sums in a different order
                assert torch.allclose(qm.vq_count, qr.vq_count, rtol=1e-5), step
            else:
                assert torch.equal(qm.vq_count, qr.vq_count), step
            if restart and step < 3:                         # drawn by the forward, applied by the trainer (restart())
                random.seed(300 + step)
                qr.restart(); qm.restart()
        _same_state(r, m, f"step {step} after restart")


@pytest.mark.parametrize("variant,cls,mode,restart,weighted", [
    ("nv", "EMACodebook", "l2", True, False), ("nv", "EMACodebook", "z_norm", False, False), ("nv", "EMACodebook", "none", False, True),
    ("nv", "Codebook", "z_norm", True, False), ("nv", "Codebook", "none", False, True),
    ("pqgo", "Codebook", "l2", True, False), ("pqgo", "Codebook", "none", False, True), ("pcls", "Codebook", "z_trainable", False, False),
    # (dino_pqgo.Codebook with z_trainable cannot run in the reference: z_pos_norm is never assigned, dino_pqgo.py:650)
    ("pcls", "Codebook", "z_norm", True, False), ("pcls", "Codebook", "l2", False, False)])
def test_inline_wrappers_against_the_live_reference(ref, variant, cls, mode, restart, weighted):
    """The inline ProductQuantizerWrapper copies of dino_new_vq / dino_pqgo / dino_pqgo_cls: 2 training calls + 1
    evaluation call, every element of the variant's return tuple."""
    from equss_b200 import codebooks as CB
    R = ref[variant]
    Mir = {"nv": CB.NewVQProductQuantizerWrapper, "pqgo": CB.PQGOProductQuantizerWrapper, "pcls": CB.PQGOClsProductQuantizerWrapper}[variant]
    torch.manual_seed(23)
    M, K, d, B, h, w = 2, 10, 6, 2, 5, 4
    kw = dict(beta=0.3, normalize=mode, use_restart=restart, use_weighted_sum=weighted, need_initialized="none", jsd_ts=0.6)
    if variant == "pqgo":
        kw["book"] = 0.7
    r = R.ProductQuantizerWrapper(M, K, M * d, quantizer_cls=getattr(R, cls), **kw)
    m = Mir(M, K, M * d, quantizer_cls=getattr(CB, cls), **kw)
    scale = 0.3 if mode == "none" else 1.0
    with torch.no_grad():
        for q in r.quantizers:
            if cls == "EMACodebook":
                q.codebook.weight.copy_(torch.randn(K, d) * scale); q.codebook.weight_avg.copy_(q.codebook.weight)
            else:
                q.embedding.weight.copy_(torch.randn(K, d) * scale)
    m.load_state_dict(r.state_dict(), strict=True)
    r.train(); m.train()
    for step in range(3):
        if step == 2:
            r.eval(); m.eval()
        z = torch.randn(B, M * d, h, w) * scale
        res = []
        for mod in (r, m):
            random.seed(400 + step); torch.manual_seed(500 + step)
            with torch.no_grad():
                res.append(mod(z, step) if variant != "pqgo" else mod(z, torch.zeros_like(z), step))
        rr, mr = res
        assert len(rr) == len(mr)
        assert torch.allclose(mr[0], rr[0], rtol=2e-5, atol=2e-6), step
        if variant == "nv":
            (_, ro, rp), (_, mo, mp) = rr, mr
        elif variant == "pcls":
            (_, ro, rp, ri), (_, mo, mp, mi) = rr, mr
            assert all(torch.equal(a, b) for a, b in zip(ri, mi))
        else:
            (_, (rs, rqs, ri), ro, rp), (_, (ms, mqs, mi), mo, mp) = rr, mr
            assert all(torch.equal(a, b) for a, b in zip(ri, mi)) and all(torch.equal(a, b) for a, b in zip(rs, ms))
            assert all(torch.allclose(a, b, rtol=2e-5, atol=2e-6) for a, b in zip(rqs, mqs))
        assert mp.shape == rp.shape and torch.allclose(mp, rp, rtol=2e-5, atol=1e-7), step
        _same_outputs(ro, mo, f"{variant}/{cls} step {step}")
        _same_state(r, m, f"{variant}/{cls} step {step}")
        for qr, qm in zip(r.quantizers, m.quantizers):
            assert torch.equal(qm.vq_count.cpu(), qr.vq_count.cpu()), step
            if restart and cls == "EMACodebook" and step < 2:
                qr.restart(); qm.restart()
        _same_state(r, m, f"{variant}/{cls} step {step} after restart")


@pytest.mark.parametrize("variant,cls,mode,weighted", [("nv", "EMACodebook", "z_norm", False), ("nv", "Codebook", "l2", False),
                                                       ("pqgo", "Codebook", "none", True), ("nv", "EMACodebook", "none", True)])
def test_pq_dropout_wrappers_against_the_live_reference(ref, monkeypatch, variant, cls, mode, weighted):
    """pq_dropout through the wrappers: the reference's ``torch.cuda.FloatTensor(K).uniform_()`` and the mirror's own draw
    are fed the same uniform numbers (one (K,) vector per subspace per call, evaluation calls included)."""
    from equss_b200 import _host_paths as hp
    from equss_b200 import codebooks as CB
    R = ref[variant]
    Mir = {"nv": CB.NewVQProductQuantizerWrapper, "pqgo": CB.PQGOProductQuantizerWrapper}[variant]
    torch.manual_seed(29)
    M, K, d, B, h, w, p = 3, 14, 6, 2, 5, 4, 0.3
    kw = dict(beta=0.3, normalize=mode, use_weighted_sum=weighted, need_initialized="none", jsd_ts=0.6, pq_dropout=p)
    r = R.ProductQuantizerWrapper(M, K, M * d, quantizer_cls=getattr(R, cls), **kw)
    m = Mir(M, K, M * d, quantizer_cls=getattr(CB, cls), **kw)
    scale = 0.3 if mode == "none" else 1.0
    with torch.no_grad():
        for q in r.quantizers:
            if cls == "EMACodebook":
                q.codebook.weight.copy_(torch.randn(K, d) * scale); q.codebook.weight_avg.copy_(q.codebook.weight)
            else:
                q.embedding.weight.copy_(torch.randn(K, d) * scale)
    m.load_state_dict(r.state_dict(), strict=True)
    r.train(); m.train()

    class _Pending:
        def __init__(self, u):
            self.u = u

        def uniform_(self):
            return self.u

    for step in range(3):
        if step == 2:
            r.eval(); m.eval()
        z = torch.randn(B, M * d, h, w) * scale
        draws = list(torch.rand(M, K))
        it_r, it_m = iter(draws), iter(draws)
        monkeypatch.setattr(torch.cuda, "FloatTensor", lambda n: _Pending(next(it_r)), raising=False)
        monkeypatch.setattr(hp, "dropout_keep_mask", lambda n, prob, device: next(it_m).to(device) > prob)
        with torch.no_grad():
            rr = r(z, step) if variant != "pqgo" else r(z, torch.zeros_like(z), step)
            mr = m(z, step) if variant != "pqgo" else m(z, torch.zeros_like(z), step)
        assert next(it_r, None) is None and next(it_m, None) is None          # M draws each, evaluation included
        assert torch.allclose(mr[0], rr[0], rtol=2e-5, atol=2e-6), step
        ro, mo = (rr[1], mr[1]) if variant == "nv" else (rr[2], mr[2])
        rp, mp = (rr[2], mr[2]) if variant == "nv" else (rr[3], mr[3])
        assert mp.shape == rp.shape and torch.allclose(mp, rp, rtol=2e-5, atol=1e-7), step      # ragged widths concatenated
        if variant == "pqgo":
            assert all(torch.equal(a, b) for a, b in zip(rr[1][2], mr[1][2]))
        _same_outputs(ro, mo, f"dropout {variant}/{cls} step {step}")
        _same_state(r, m, f"dropout {variant}/{cls} step {step}")
        for qr, qm in zip(r.quantizers, m.quantizers):
            assert torch.equal(qm.vq_count.cpu(), qr.vq_count.cpu()), step


@pytest.mark.parametrize("B,D,h,w,H,W,extra", [(2, 24, 7, 5, 23, 31, 0), (1, 16, 6, 6, 6, 6, 0), (2, 32, 5, 8, 40, 64, 3)])
def test_evaluator_and_metrics_against_the_live_reference(ref, monkeypatch, tmp_path, B, D, h, w, H, W, extra):
    """UnSegEvaluator.forward + UnSegMetrics.update / compute (Hungarian matching, extra cluster rows) next to the
    reference at label resolution: non-integer scale factors, equal sizes, extra classes."""
    import model.evaluator as ref_eval
    import model.metric as ref_metric
    from equss_b200.evaluator import UnSegEvaluator
    from equss_b200.metric import UnSegMetrics
    monkeypatch.chdir(tmp_path)                                # compute() writes ./class_matrix/...
    torch.manual_seed(31)
    C = 9
    r = ref_eval.UnSegEvaluator(D, C, extra).eval()
    m = UnSegEvaluator(D, C, extra).eval()
    m.load_state_dict(r.state_dict(), strict=True)
    out = torch.randn(B, D, h, w)
    label = torch.randint(-1, C, (B, H, W))
    with torch.no_grad():
        rl, rlp, rc, rcp = r(out, None, label)
        ml, mlp, mc, mcp = m(out, None, label)
    assert float((mlp != rlp).float().mean()) < 2e-3 and float((mcp != rcp).float().mean()) < 2e-3     # fp32 near-ties only
    assert float(ml) == pytest.approx(float(rl), rel=2e-5) and float(mc) == pytest.approx(float(rc), rel=2e-4)
    for hung, preds_r, preds_m, ex in ((True, rcp, rcp, extra), (False, rlp, rlp, 0)):
        rm = ref_metric.UnSegMetrics(C, ex, hung, torch.device("cpu"))
        mm = UnSegMetrics(C, ex, hung, torch.device("cpu"))
        rm.update(preds_r, label); mm.update(preds_m, label)
        assert torch.equal(mm.confusion_matrix, rm.confusion_matrix)
        rr, mr = rm.compute("t"), mm.compute("t")
        for k in rr:
            assert float(mr[k]) == pytest.approx(float(rr[k]), rel=1e-6), k
        if hung:
            assert torch.equal(mm.map_clusters(preds_m), rm.map_clusters(preds_r))


@pytest.mark.parametrize("mode,restart,gumbel", [("l2", False, False), ("z_norm", True, False), ("z_trainable", False, False),
                                                 ("none", False, False), ("l2", False, True)])
def test_learned_vector_quantizer_against_the_live_reference(ref, mode, restart, gumbel):
    """model/quantizer.py VectorQuantizer (NCHW, learned codebook): training call with the gradients of the reference's
    graph (straight-through output, both losses), exact counters, evaluation call."""
    from equss_b200.quantizer import VectorQuantizer
    q1 = ref["q1"]
    torch.manual_seed(37)
    K, d, B, h, w = 11, 8, 2, 5, 4
    r = q1.VectorQuantizer(K, d, beta=0.3, normalize=mode, use_restart=restart, use_gumbel=gumbel)
    m = VectorQuantizer(K, d, beta=0.3, normalize=mode, use_restart=restart, use_gumbel=gumbel)
    m.load_state_dict(r.state_dict(), strict=True)
    go = torch.randn(B, d, h, w)
    for step, training in enumerate((True, False)):
        r.train(training); m.train(training)
        z = torch.randn(B, d, h, w) * (0.3 if mode == "none" else 1.0)
        res = []
        for mod in (r, m):
            mod.zero_grad()
            zi = z.clone().requires_grad_(True)
            random.seed(600 + step); torch.manual_seed(700 + step)
            q, out, prob = mod(zi)
            (out["loss"] + (q * go).sum() + (prob * prob).sum()).backward()
            res.append((q, out, prob, zi.grad))
        (rq, ro, rp, rg), (mq, mo, mp, mg) = res
        assert torch.allclose(mq, rq, rtol=2e-5, atol=2e-6) and torch.allclose(mp, rp, rtol=2e-5, atol=1e-7), step
        _same_outputs(ro, mo, f"V1 {mode} step {step}")
        assert torch.allclose(mg, rg, rtol=1e-3, atol=1e-6 * float(rg.abs().max()) + 1e-9), (step, float((mg - rg).abs().max()))
        for (kn, pr), (_, pm) in zip(r.named_parameters(), m.named_parameters()):
            assert (pr.grad is None) == (pm.grad is None), kn
            if pr.grad is not None:
                assert torch.allclose(pm.grad, pr.grad, rtol=1e-3, atol=1e-6 * float(pr.grad.abs().max()) + 1e-9), (step, kn)
        assert torch.equal(m.vq_count, r.vq_count)
        if restart and training:
            r.restart(); m.restart()
        _same_state(r, m, f"V1 {mode} step {step}")


def test_quantizer_v2_against_the_live_reference(ref):
    """model/quantizer_v2.py (V3): EMA quantiser that gathers rows of z_norm (a quirk kept for parity), 3 training steps +
    evaluation through its wrapper."""
    import model.quantizer_v2 as q2
    from equss_b200 import quantizer_v2 as m2
    torch.manual_seed(41)
    M, K, d, B, h, w = 2, 9, 6, 2, 5, 4
    r = q2.ProductQuantizerWrapper(M, K, M * d, beta=0.3, normalize="l2", decay=0.9, eps=1e-4)
    m = m2.ProductQuantizerWrapper(M, K, M * d, beta=0.3, normalize="l2", decay=0.9, eps=1e-4)
    with torch.no_grad():
        for q in r.quantizers:
            q.embeddings.copy_(torch.randn(K, d)); q.z_avg.copy_(q.embeddings)
    m.load_state_dict(r.state_dict(), strict=True)
    r.train(); m.train()
    for step in range(4):
        if step == 3:
            r.eval(); m.eval()
        z = torch.randn(B, M * d, h, w)
        with torch.no_grad():
            (rq, ro, rp), (mq, mo, mp) = r(z), m(z)
        assert torch.allclose(mq, rq, rtol=2e-5, atol=2e-6) and torch.allclose(mp, rp, rtol=2e-5, atol=1e-7), step
        _same_outputs(ro, mo, f"V3 step {step}")
        _same_state(r, m, f"V3 step {step}")


@pytest.mark.parametrize("pointwise,zero_clamp,stabilize", [(True, True, False), (False, False, True)])
def test_stego_loss_against_the_live_reference(ref, pointwise, zero_clamp, stabilize):
    """model/loss.py STEGOLoss: same sampled coordinates and negative permutations (torch seeds), loss value and the
    gradient that reaches the code maps."""
    import model.loss as ref_loss
    from equss_b200.losses import STEGOLoss
    cfg = {"pointwise": pointwise, "zero_clamp": zero_clamp, "stabilize": stabilize, "feature_samples": 7, "neg_samples": 2,
           "pos_intra_shift": 0.18, "pos_inter_shift": 0.12, "neg_inter_shift": 0.46,
           "pos_intra_weight": 0.67, "pos_inter_weight": 0.25, "neg_inter_weight": 0.63}
    torch.manual_seed(43)
    n, C, Cc, h, w = 4, 12, 10, 9, 9
    feats, feats_pos = torch.randn(n, C, h, w), torch.randn(n, C, h, w)
    code, code_pos = torch.randn(n, Cc, h, w), torch.randn(n, Cc, h, w)
    res = []
    for L in (ref_loss.STEGOLoss(cfg), STEGOLoss(cfg)):
        c1, c2 = code.clone().requires_grad_(True), code_pos.clone().requires_grad_(True)
        torch.manual_seed(800)
        loss = L(feats, feats_pos, c1, c2)
        loss.backward()
        res.append((loss.detach(), c1.grad, c2.grad))
    (rl, rg1, rg2), (ml, mg1, mg2) = res
    assert float(ml) == pytest.approx(float(rl), rel=2e-5)
    assert torch.allclose(mg1, rg1, rtol=1e-3, atol=1e-8) and torch.allclose(mg2, rg2, rtol=1e-3, atol=1e-8)


def test_user_supplied_quantizer_class_takes_the_reference_loop(ref):
    """``quantizer_cls`` may be any module with the quantiser signature (model/quantizer.py:577-611): the mirror then runs
    the reference's per-subspace loop -- chunk, call, concatenate, mean of the output dictionaries."""
    from equss_b200.quantizer import ProductQuantizerWrapper
    q1 = ref["q1"]

    class Halver(torch.nn.Module):
        def __init__(self, num_codebook, embed_dim, **kw):
            super().__init__()
            self.scale = torch.nn.Parameter(torch.full((embed_dim,), 0.5))
            self.kw = kw

        def forward(self, z):
            return z * self.scale, {"loss": (z ** 2).mean(), "n": float(z.shape[1])}, torch.softmax(z[:, :3], dim=1)

    torch.manual_seed(47)
    r = q1.ProductQuantizerWrapper(3, 8, 12, normalize="l2", quantizer_cls=Halver)
    m = ProductQuantizerWrapper(3, 8, 12, normalize="l2", quantizer_cls=Halver)
    m.load_state_dict(r.state_dict(), strict=True)
    z = torch.randn(10, 12)
    (rq, ro, rp), (mq, mo, mp) = r(z), m(z)
    assert torch.equal(mq, rq) and torch.equal(mp, rp) and mp.shape == (10, 9)
    _same_outputs(ro, mo, "custom class")
    assert m.quantizers[0].kw.keys() == r.quantizers[0].kw.keys()            # same keyword arguments handed down


@pytest.mark.parametrize("yaml_file,variant,cls", [
    ("config/pqgo_baseline.yaml", "pqgo", "Codebook"), ("config/cityscapes/pqgo_baseline.yaml", "pqgo", "Codebook"),
    ("config/new_vq_baseline.yaml", "nv", "Codebook"), ("config/new_vq_baseline.yaml", "nv", "EMACodebook"),
    ("config/pqgo_cls.yaml", "pcls", "Codebook")])
def test_yaml_literal_inline_configuration_against_the_live_reference(ref, yaml_file, variant, cls):
    """The trainer paths exactly as the reference's YAML files configure them: the ``vq`` section is read with
    ``yaml.safe_load`` and turned into the wrapper's arguments the way the model constructors do it (``train.py``:
    model/dino_pqgo.py:38-76 -- cocostuff27 64 subspaces x 256 codes on 1024 channels, cityscapes 32 x 32; ``train_vq.py``:
    model/dino_new_vq.py:56-98 -- 32 x 512, with the learned codebook the file selects and the EMA one its comment offers;
    model/dino_pqgo_cls.py:41-68 -- 64 x 256), including ``need_initialized: "uni"`` (the first training call re-draws the
    codebook), a zero ``pq_dropout`` and the jsd temperature of the loss section.  2 training calls + 1 evaluation call on a
    small token grid; every element of the return tuple, the output dictionary and the state agree."""
    import yaml
    from equss_b200 import codebooks as CB
    cfg_all = yaml.safe_load(open(os.path.join(os.environ.get("EQUSS_REFERENCE", "/root/reference"), yaml_file)))
    vq, jsd = cfg_all["model"]["vq"], cfg_all["loss"]["jsd"]
    assert vq["vq_type"] == "param"
    kw = dict(beta=vq["beta"], normalize=vq.get("normalize", "none"), use_restart=vq.get("use_restart", False),
              use_weighted_sum=vq.get("use_weighted_sum", False), need_initialized=vq.get("need_initialized", False),
              pq_dropout=vq.get("pq_dropout", 0.0), jsd_ts=jsd.get("temperature", 1.0))
    if variant in ("pqgo", "pcls"):
        kw.update(use_split=vq.get("use_split", False), num_query=jsd.get("num_query", 3), num_pos=jsd.get("num_pos", 10))
    if variant == "pqgo":
        kw["book"] = vq.get("book", 1.0)
    M, K, D = vq["num_pq"][0], vq["num_codebooks"][0], vq["embed_dims"][0]
    R = ref[variant]
    Mir = {"nv": CB.NewVQProductQuantizerWrapper, "pqgo": CB.PQGOProductQuantizerWrapper, "pcls": CB.PQGOClsProductQuantizerWrapper}[variant]
    torch.manual_seed(31)
    r = R.ProductQuantizerWrapper(M, K, D, **kw, quantizer_cls=getattr(R, cls))
    m = Mir(M, K, D, **kw, quantizer_cls=getattr(CB, cls))
    m.load_state_dict(r.state_dict(), strict=True)
    r.train(); m.train()
    B, h, w = 2, 5, 4                                                     # even row count: jsd splits the batch in halves
    for step in range(3):
        if step == 2:
            r.eval(); m.eval()
        z = torch.randn(B, D, h, w)
        res = []
        for mod in (r, m):
            random.seed(600 + step); torch.manual_seed(700 + step)
            with torch.no_grad():
                res.append(mod(z, step) if variant != "pqgo" else mod(z, torch.zeros_like(z), step))
        rr, mr = res
        assert len(rr) == len(mr) and torch.allclose(mr[0], rr[0], rtol=2e-5, atol=2e-6), step
        if variant == "nv":
            (_, ro, rp), (_, mo, mp) = rr, mr
        elif variant == "pcls":
            (_, ro, rp, ri), (_, mo, mp, mi) = rr, mr
            assert len(ri) == len(mi) and all(torch.equal(a, b) for a, b in zip(ri, mi))
        else:
            (_, (rs, rqs, ri), ro, rp), (_, (ms, mqs, mi), mo, mp) = rr, mr
            assert len(ri) == len(mi) == M and all(torch.equal(a, b) for a, b in zip(ri, mi))
            assert all(torch.equal(a, b) for a, b in zip(rs, ms))
            assert all(torch.allclose(a, b, rtol=2e-5, atol=2e-6) for a, b in zip(rqs, mqs))
        assert mp.shape == rp.shape and torch.allclose(mp, rp, rtol=2e-5, atol=1e-7), step
        _same_outputs(ro, mo, f"{yaml_file} {cls} step {step}")
        _same_state(r, m, f"{yaml_file} {cls} step {step}")
        assert all(torch.equal(qm.vq_count.cpu(), qr.vq_count.cpu()) for qr, qm in zip(r.quantizers, m.quantizers)), step


@pytest.mark.parametrize("yaml_file,block", [("config/pq_baseline.yaml", 0), ("config/pq_baseline.yaml", 1),
                                             ("config/pq_vae.yaml", 0), ("config/pq_vae.yaml", 1)])
def test_yaml_literal_ema_configuration_against_the_live_reference(ref, yaml_file, block):
    """model/quantizer.py's EMA wrapper exactly as the reference's YAML files configure it (``vq`` section read with
    ``yaml.safe_load``, arguments assembled like ``DINOContra.__init__`` / ``DINOVae.__init__`` do, model/dino_contra.py:41-70):
    1024 codes on 512 channels, 2 / 16 subspaces (d = 256 / 32), ``use_split`` as the file says, ``need_initialized``
    left at the constructor's ``False`` when the file has no such key.  3 training steps + 1 evaluation step."""
    import yaml
    from equss_b200.quantizer import EMAVectorQuantizer, ProductQuantizerWrapper
    vq = yaml.safe_load(open(os.path.join(os.environ.get("EQUSS_REFERENCE", "/root/reference"), yaml_file)))["model"]["vq"]
    assert vq["vq_type"] == "ema"
    kw = dict(beta=vq["beta"], normalize=vq["normalize"], use_restart=vq.get("use_restart", False),
              use_gumbel=vq.get("use_gumbel", False), use_split=vq.get("use_split", False),
              use_weighted_sum=vq.get("use_weighted_sum", False), need_initialized=vq.get("need_initialized", False),
              decay=vq["decay"], eps=vq["eps"])
    if kw["need_initialized"] == "kmeans":
        pytest.skip("k-means initialisation of 1024 codes: pinned separately (tests/golden/pq_init_kmeans.npz)")
    M, K, D = vq["num_pq"][block], vq["num_codebooks"][block], vq["embed_dims"][block]
    q1 = ref["q1"]
    torch.manual_seed(41 + block)
    r = q1.ProductQuantizerWrapper(M, K, D, **kw, quantizer_cls=q1.EMAVectorQuantizer)
    m = ProductQuantizerWrapper(M, K, D, **kw, quantizer_cls=EMAVectorQuantizer)
    m.load_state_dict(r.state_dict(), strict=True)
    r.train(); m.train()
    for step in range(4):
        if step == 3:
            r.eval(); m.eval()
        z = torch.randn(150, D) + 0.01 * step
        outs = []
        for mod in (r, m):
            random.seed(800 + step); torch.manual_seed(900 + step)
            with torch.no_grad():
                outs.append(mod(z))
        (rq, ro, rp), (mq, mo, mp) = outs
        assert torch.allclose(mq, rq, rtol=2e-5, atol=2e-6), (step, float((mq - rq).abs().max()))
        assert torch.allclose(mp, rp, rtol=2e-5, atol=1e-7), step
        _same_outputs(ro, mo, f"{yaml_file}[{block}] step {step}")
        _same_state(r, m, f"{yaml_file}[{block}] step {step}")
        assert all(torch.equal(qm.vq_count, qr.vq_count) for qr, qm in zip(r.quantizers, m.quantizers)), step
