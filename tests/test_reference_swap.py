"""Import-swap on the build box: the reference's OWN wrapper code constructs and holds the mirror
modules when ``model.evaluator`` / ``model.quantizer`` / ``model.metric`` are replaced in ``sys.modules``.

Needs /root/reference (skipped on the GPU box, where the kernels run but the reference tree is absent; the forward
arithmetic of the swapped-in modules is what the ``-m gpu`` tests check).  No kernel runs here: a CPU call must fail
loudly instead of falling back."""
import importlib
import os
import sys

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")


@pytest.fixture()
def swapped():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    from make_golden import import_reference
    import_reference()                                   # stubs torchmetrics / pydensecrf, puts REF on sys.path
    import equss_b200
    saved = {k: sys.modules.get(k) for k in ("model.evaluator", "model.quantizer", "model.metric", "wrapper.PQGOWrapper")}
    ref_eval = importlib.import_module("model.evaluator")
    ref_quant = importlib.import_module("model.quantizer")
    sys.modules["model.evaluator"] = equss_b200.evaluator
    sys.modules["model.quantizer"] = equss_b200.quantizer
    sys.modules["model.metric"] = equss_b200.metric
    sys.modules.pop("wrapper.PQGOWrapper", None)
    try:
        yield ref_eval, ref_quant, equss_b200
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_reference_wrapper_builds_on_the_mirrors(swapped):
    ref_eval, ref_quant, eq = swapped
    wrapper = importlib.import_module("wrapper.PQGOWrapper")
    cfg = {"num_classes": 27, "eval": {"extra_classes": 0, "output_type": "vq0"}, "dataset_name": "cocostuff27",
           "model": {"vq": {"num_pq": [4], "embed_dims": [64]}}, "loss": {"stego_weight": 1.0, "vq_weight": 1.0}}
    w = wrapper.PQGOWrapper(cfg, torch.nn.Identity())
    assert type(w.evaluator) is eq.evaluator.UnSegEvaluator
    ref_keys = sorted(ref_eval.UnSegEvaluator(64, 27, 0).state_dict().keys())
    assert sorted(w.evaluator.state_dict().keys()) == ref_keys
    w.evaluator.load_state_dict(ref_eval.UnSegEvaluator(64, 27, 0).state_dict(), strict=True)
    # the mirror never computes on the CPU
    with pytest.raises(eq._native.EqussNativeError):
        w.evaluator(torch.randn(1, 64, 4, 4), None, torch.zeros(1, 8, 8, dtype=torch.long))


def test_quantizer_state_dicts_and_signatures_match(swapped):
    import inspect
    ref_eval, ref_quant, eq = swapped
    for name in ("VectorQuantizer", "EMAVectorQuantizer", "EmbeddingEMA", "ProductQuantizerWrapper"):
        r, m = getattr(ref_quant, name), getattr(eq.quantizer, name)
        rp = list(inspect.signature(r.__init__).parameters)
        mp = [p for p in inspect.signature(m.__init__).parameters if not p.startswith("_")]
        assert mp[:len(rp)] == rp, (name, rp, mp)
    ref = ref_quant.ProductQuantizerWrapper(4, 16, 32, normalize="l2")
    mir = eq.quantizer.ProductQuantizerWrapper(4, 16, 32, normalize="l2")
    assert sorted(ref.state_dict().keys()) == sorted(mir.state_dict().keys())
    mir.load_state_dict(ref.state_dict(), strict=True)
    for k, v in ref.state_dict().items():
        assert torch.equal(mir.state_dict()[k], v)
    # build.py:77 sorts parameters by isinstance on these three class names
    from equss_b200.quantizer import EMAVectorQuantizer, EmbeddingEMA, VectorQuantizer
    kinds = {type(mod) for mod in mir.modules()}
    assert EMAVectorQuantizer in kinds and EmbeddingEMA in kinds
    # (the reference's wrapper cannot build the learned-codebook flavour itself: it passes decay= to a ctor without it)
    assert sorted(VectorQuantizer(8, 16, normalize="l2").state_dict().keys()) == \
        sorted(ref_quant.VectorQuantizer(8, 16, normalize="l2").state_dict().keys())


def test_every_mirrored_class_keeps_the_reference_constructor_and_forward(swapped):
    """Drop-in surface, class by class: every constructor parameter of the reference exists in the mirror in
    the same order with the same default (``quantizer_cls`` defaults to the mirror's own class), ``forward`` takes the
    reference's arguments, and state dicts of the inline wrappers load with strict=True in both directions."""
    import inspect
    _, _, eq = swapped
    pairs = []

    def add(ref_mod, mir_mod, names):
        r, m = importlib.import_module(ref_mod), importlib.import_module(mir_mod)
        assert os.path.abspath(r.__file__).startswith(REF) and not os.path.abspath(m.__file__).startswith(REF)
        for n in names:
            rn, mn = n if isinstance(n, tuple) else (n, n)
            pairs.append((f"{ref_mod}.{rn}", getattr(r, rn), getattr(m, mn)))

    for k in ("model.quantizer", "model.evaluator", "model.metric"):       # compare against the REAL reference modules
        sys.modules.pop(k, None)
    add("model.quantizer", "equss_b200.quantizer", ["VectorQuantizer", "EMAVectorQuantizer", "EmbeddingEMA", "ProductQuantizerWrapper"])
    add("model.quantizer_v2", "equss_b200.quantizer_v2", ["EMAVectorQuantizer", "ProductQuantizerWrapper"])
    add("model.evaluator", "equss_b200.evaluator", ["UnSegEvaluator", "ClusterLookup"])
    add("model.metric", "equss_b200.metric", ["UnSegMetrics"])
    add("model.blocks.module", "equss_b200.head", ["SegmentationHead"])
    add("model.loss", "equss_b200.losses", ["STEGOLoss"])
    add("model.loss", "equss_b200.codebooks", ["JSDLoss", "EntropyLoss"])
    add("model.dino_pqgo", "equss_b200.codebooks", ["Codebook", ("ProductQuantizerWrapper", "PQGOProductQuantizerWrapper")])
    add("model.dino_new_vq", "equss_b200.codebooks", ["Codebook", "EMACodebook", ("ProductQuantizerWrapper", "NewVQProductQuantizerWrapper")])
    add("model.dino_pqgo_cls", "equss_b200.codebooks", ["Codebook", ("ProductQuantizerWrapper", "PQGOClsProductQuantizerWrapper")])
    assert len(pairs) == 20
    for name, r, m in pairs:
        rs, ms = inspect.signature(r.__init__).parameters, inspect.signature(m.__init__).parameters
        assert [p for p in rs if p not in ms] == [], name
        assert [p for p in ms if p in rs] == list(rs), name                               # same order
        for k in rs:
            if k != "quantizer_cls" and rs[k].default is not inspect.Parameter.empty:
                assert rs[k].default == ms[k].default, (name, k, rs[k].default, ms[k].default)
        if "Codebook" in name and "Wrapper" not in name:
            continue                          # one mirror class serves three files: forward(z, *args), arity per variant
        rf, mf = list(inspect.signature(r.forward).parameters), list(inspect.signature(m.forward).parameters)
        assert mf[:len(rf)] == rf, (name, rf, mf)
    import model.dino_new_vq as nv
    import model.dino_pqgo as pqgo
    import model.dino_pqgo_cls as pcls
    cbm = eq.codebooks
    for R, Mi, kw_r, kw_m in ((pqgo.ProductQuantizerWrapper, cbm.PQGOProductQuantizerWrapper, {}, {}),
                              (pcls.ProductQuantizerWrapper, cbm.PQGOClsProductQuantizerWrapper, {}, {}),
                              (nv.ProductQuantizerWrapper, cbm.NewVQProductQuantizerWrapper, {"quantizer_cls": nv.EMACodebook}, {"quantizer_cls": cbm.EMACodebook}),
                              (nv.ProductQuantizerWrapper, cbm.NewVQProductQuantizerWrapper, {"quantizer_cls": nv.Codebook}, {})):
        ref = R(4, 16, 32, normalize="z_trainable", need_initialized="none", **kw_r)
        mir = Mi(4, 16, 32, normalize="z_trainable", need_initialized="none", **kw_m)
        assert sorted(ref.state_dict().keys()) == sorted(mir.state_dict().keys())
        mir.load_state_dict(ref.state_dict(), strict=True)
        ref.load_state_dict(mir.state_dict(), strict=True)
        for k, v in ref.state_dict().items():
            assert torch.equal(mir.state_dict()[k], v)


def test_cluster_lookup_forward_equals_the_reference(swapped):
    """The stand-alone ClusterLookup.forward (hard assignment, temperature softmax, log-probabilities) against the
    unmodified reference class on the same parameter and input -- plain torch on both sides, bit for bit."""
    ref_eval, _, eq = swapped
    torch.manual_seed(0)
    ref = ref_eval.ClusterLookup(12, 7)
    mir = eq.evaluator.ClusterLookup(12, 7)
    mir.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(2, 12, 5, 4)
    x[0, :, 0, 0] = 0.0                                              # a zero feature vector: all cosines tie at 0
    for alpha, log_probs in ((None, False), (2.0, False), (0.5, False), (3.0, True)):
        (rl, rp), (ml, mp) = ref(x, alpha=alpha, log_probs=log_probs), mir(x, alpha=alpha, log_probs=log_probs)
        assert torch.equal(rl, ml) and torch.equal(rp, mp) and rp.dtype == mp.dtype, (alpha, log_probs)
    xg = x.clone().requires_grad_(True)
    mir(xg, alpha=2.0)[0].backward()
    xr = x.clone().requires_grad_(True)
    ref(xr, alpha=2.0)[0].backward()
    assert torch.allclose(xg.grad, xr.grad, rtol=1e-6, atol=1e-8)
    assert torch.allclose(mir.clusters.grad, ref.clusters.grad, rtol=1e-6, atol=1e-8)


def test_embedding_ema_helper_calls_equal_the_reference(swapped):
    """EmbeddingEMA's three public update pieces (vq_count_ema_update / weight_avg_ema_update / weight_update,
    model/quantizer.py:241-254) and reset() are plain torch in the mirror too (update() itself is the kernel): same
    buffers after the same calls, bit for bit."""
    _, ref_quant, eq = swapped
    torch.manual_seed(1)
    ref = ref_quant.EmbeddingEMA(9, 6, decay=0.9, eps=1e-4)
    mir = eq.quantizer.EmbeddingEMA(9, 6, decay=0.9, eps=1e-4)
    mir.load_state_dict(ref.state_dict(), strict=True)
    for step in range(3):
        count = torch.randint(0, 5, (9,)).float()
        total = torch.randn(9, 6)
        for e in (ref, mir):
            e.vq_count_ema_update(count)
            e.weight_avg_ema_update(total)
            e.weight_update()
        for k, v in ref.state_dict().items():
            assert torch.equal(mir.state_dict()[k], v), (step, k)
    idx = torch.tensor([[0, 3], [8, 8]])
    assert torch.equal(ref(idx), mir(idx))
    ref.reset(); mir.reset()
    for k, v in ref.state_dict().items():
        assert torch.equal(mir.state_dict()[k], v), k
    with pytest.raises(eq._native.EqussNativeError):
        mir.update(torch.ones(9), torch.randn(9, 6))               # the fused update is the kernel: no CPU fallback


def test_restart_and_split_draw_like_the_reference(swapped):
    """prepare_restart / restart / split (model/quantizer.py:298-381) are host code on both sides: with the same seeds
    of Python's ``random`` and of torch's CPU generator the mirror replaces the same codes by the same rows and splits
    the same codes with the same noise -- compared with the live reference object, buffer by buffer."""
    import random
    _, ref_quant, eq = swapped
    K, d = 12, 5
    torch.manual_seed(4)
    w0 = torch.randn(K, d)

    def pair():
        r = ref_quant.EMAVectorQuantizer(K, d, normalize="l2", use_restart=True, use_split=True)
        m = eq.quantizer.EMAVectorQuantizer(K, d, normalize="l2", use_restart=True, use_split=True)
        for q in (r, m):
            q.codebook.weight.copy_(w0); q.codebook.weight_avg.copy_(w0 * 0.5)
            q.codebook.vq_count.copy_(torch.arange(K, dtype=torch.float32).flip(0) + 1.0)
            q.vq_count.fill_(3.0)
        return r, m

    def same_state(r, m, what):
        for k, v in r.state_dict().items():
            assert torch.equal(m.state_dict()[k], v), (what, k)
        assert torch.equal(r.vq_count, m.vq_count), what

    # (a) fewer dead codes than rows, (b) more dead codes than rows (the served codes are drawn too), (c) none dead
    for name, count, n_rows in (("few_dead", torch.tensor([3, 0, 1, 0, 0, 2, 5, 0, 1, 1, 0, 4.0]), 40),
                                ("more_dead_than_rows", torch.tensor([1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.0]), 4),
                                ("none_dead", torch.ones(K), 10)):
        r, m = pair()
        rows = torch.randn(n_rows, d)
        for q in (r, m):
            random.seed(7)
            q.prepare_restart(count, rows)
        ri = r.update_indices.tolist() if torch.is_tensor(r.update_indices) else list(r.update_indices)
        mi = m.update_indices.tolist() if torch.is_tensor(m.update_indices) else list(m.update_indices)
        assert ri == mi and torch.equal(r.update_candidates, m.update_candidates), name
        r.restart(); m.restart()
        assert r.update_indices is None and m.update_indices is None
        same_state(r, m, name)
    # split: dead codes paired with the busiest codes, +- N(0, 0.02^2) noise, halved counts and running sums
    for name, count in (("some_dead", torch.tensor([3, 0, 1, 0, 0, 2, 5, 0, 1, 1, 0, 4.0])), ("none_dead", torch.ones(K)),
                        ("all_dead", torch.zeros(K))):
        r, m = pair()
        torch.manual_seed(11)
        nr = r.split(count)
        torch.manual_seed(11)
        nm = m.split(count)
        assert nr == nm == int((count == 0).sum()), name
        same_state(r, m, "split/" + name)


def test_evaluator_torch_losses_equal_the_reference(swapped):
    """The differentiable PyTorch formulation of the evaluator losses (``UnSegEvaluator._losses_torch``: probes at token
    resolution, interpolated logits, the norm of the upsampled feature from 2x2 Gram terms) against the unmodified
    reference's forward at label resolution (model/evaluator.py:46-82,95-106): losses to 1e-5, gradients of the probe
    parameters and of the features through the reference's own graph."""
    import torch.nn.functional as F
    ref_eval, _, eq = swapped
    from equss_b200.evaluator import _upsampled_feature_norm
    torch.manual_seed(2)
    for (h, w, H, W) in ((10, 10, 40, 40), (7, 5, 20, 18), (6, 6, 6, 6)):
        x = torch.randn(2, 16, h, w)
        direct = F.interpolate(x, (H, W), mode="bilinear", align_corners=False).norm(dim=1) if (h, w) != (H, W) else x.norm(dim=1)
        with torch.no_grad():
            assert torch.allclose(_upsampled_feature_norm(x, H, W), direct, rtol=1e-5, atol=1e-6), (h, w, H, W)
    D, C = 24, 27
    ref = ref_eval.UnSegEvaluator(D, C, 0)
    mir = eq.evaluator.UnSegEvaluator(D, C, 0)
    mir.load_state_dict(ref.state_dict(), strict=True)
    out = torch.randn(2, D, 10, 10)
    label = torch.randint(-1, C, (2, 40, 40))
    xr = out.clone().requires_grad_(True)
    rl, rlp, rc, rcp = ref(xr, None, label)
    xm = out.clone().requires_grad_(True)
    ml, mc = mir._losses_torch(xm, label, rcp)
    assert float(ml.detach()) == pytest.approx(float(rl.detach()), rel=1e-5) and float(mc.detach()) == pytest.approx(float(rc.detach()), rel=1e-5)
    (rl + rc).backward()
    (ml + mc).backward()
    assert torch.allclose(mir.linear_probe.weight.grad, ref.linear_probe.weight.grad, rtol=1e-4, atol=1e-7)
    assert torch.allclose(mir.linear_probe.bias.grad, ref.linear_probe.bias.grad, rtol=1e-4, atol=1e-7)
    assert torch.allclose(mir.cluster_probe.clusters.grad, ref.cluster_probe.clusters.grad, rtol=1e-4, atol=1e-7)
    # features that carry a gradient themselves: also through the norm of the upsampled feature vector
    assert torch.allclose(xm.grad, xr.grad, rtol=1e-3, atol=1e-6 * float(xr.grad.abs().max()) + 1e-9)
    xz = out.clone()
    xz[0, :, 0, 0] = 0.0                                       # zero feature vector: finite gradients
    xz.requires_grad_(True)
    sum(mir._losses_torch(xz, label, rcp)).backward()
    assert torch.isfinite(xz.grad).all()


def test_jsd_entropy_and_soft_assignment_stats_equal_the_reference(swapped):
    """codebooks.JSDLoss / EntropyLoss and _pq_core.soft_assignment_stats (the per-subspace mean of both, what the V4
    wrapper reports) against model/loss.py:490-525 on random soft assignments, values and gradients; out-of-range
    inputs raise like the reference."""
    _, _, eq = swapped
    import model.loss as ref_loss
    from equss_b200 import _pq_core as core
    torch.manual_seed(6)
    n, M, K = 64, 3, 11
    logits = torch.randn(n, M, K) * 3
    pr = torch.softmax(logits, dim=2).reshape(n, M * K).requires_grad_(True)
    pm = pr.detach().clone().requires_grad_(True)
    rj, re = ref_loss.JSDLoss(), ref_loss.EntropyLoss()
    want_j = torch.stack([rj(*torch.chunk(pr.view(n, M, K)[:, i], 2, dim=0)) for i in range(M)]).mean()
    want_e = torch.stack([re(*torch.chunk(pr.view(n, M, K)[:, i], 2, dim=0)) for i in range(M)]).mean()
    got_j, got_e = core.soft_assignment_stats(pm, M, K)
    assert float(got_j.detach()) == pytest.approx(float(want_j.detach()), rel=1e-5)
    assert float(got_e.detach()) == pytest.approx(float(want_e.detach()), rel=1e-5)
    (want_j + 0.5 * want_e).backward()
    (got_j + 0.5 * got_e).backward()
    assert torch.allclose(pm.grad, pr.grad, rtol=1e-4, atol=1e-8)
    a, b = torch.chunk(pr.detach().view(n, M, K)[:, 0], 2, dim=0)
    mj, me = eq.codebooks.JSDLoss(), eq.codebooks.EntropyLoss()
    assert float(mj(a, b)) == pytest.approx(float(rj(a, b)), rel=1e-5)
    assert float(me(a, b)) == pytest.approx(float(re(a, b)), rel=1e-6)
    assert float(mj(a, a)) == pytest.approx(float(rj(a, a)), rel=1e-4, abs=1e-9)     # not 0: the 1e-6 smoothing is asymmetric
    for bad in (a - 0.5, a + 0.9):
        with pytest.raises(ValueError):
            rj(bad, b)
        with pytest.raises(ValueError):
            mj(bad, b)
    with pytest.raises(ValueError):
        core.soft_assignment_stats(pm.detach()[:-1], M, K)          # odd number of rows: no two halves
