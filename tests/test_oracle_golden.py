"""The oracle restatement reproduces the reference outputs stored in tests/golden (CPU only).

The fixtures were produced by oracle/make_golden.py running the unmodified reference modules; these
tests replay the oracle on the stored inputs.  Indices must match exactly (fixtures contain no fp32
near-ties); float outputs are compared to 1e-6 because MKL's summation order may differ between hosts."""
import os

import numpy as np
import pytest
import torch

import equss_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


@pytest.mark.parametrize("mode", ["l2", "z_norm", "none"])
def test_ema_pq_multi_step(golden_dir, mode):
    g = _load(golden_dir, f"pq_ema_{mode}.npz")
    M, K = int(g["M"]), int(g["K"])
    w0 = torch.from_numpy(g["weight0"])
    states = [O.EmaState(w0[i]) for i in range(M)]
    exact = [torch.zeros(K) for _ in range(M)]
    for s in range(4):
        z = torch.from_numpy(g[f"z{s}"])
        q, out, prob, idx = O.pq_forward_ema(z, states, exact, normalize=mode, beta=0.25, training=s < 3)
        assert np.array_equal(idx.numpy().astype(np.int32), g[f"idx{s}"])
        np.testing.assert_allclose(q.numpy(), g[f"zq{s}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(torch.stack([st.weight for st in states]).numpy(), g[f"weight_after{s}"],
                                   rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(torch.stack([st.vq_count for st in states]).numpy(), g[f"vq_count_after{s}"],
                                   rtol=1e-6, atol=1e-7)
        assert np.array_equal(torch.stack(exact).numpy(), g[f"exact_after{s}"])
        for k, v in out.items():
            ref = float(g[f"out{s}/{k}"])
            if v is None:
                assert np.isnan(ref)
            else:
                assert float(v) == pytest.approx(ref, rel=1e-5, abs=1e-7), k
    np.testing.assert_allclose(prob.numpy(), g["prob3"], rtol=1e-5, atol=1e-8)


def test_param_and_v2_variants(golden_dir):
    g = _load(golden_dir, "pq_param_nchw.npz")
    z = torch.from_numpy(g["z"])
    q, out, prob, idx = O.param_vq_forward(z, torch.from_numpy(g["v1_codebook"]), normalize="l2", beta=0.25)
    assert np.array_equal(idx.numpy().astype(np.int32), g["v1_idx"])
    np.testing.assert_allclose(q.numpy(), g["v1_q"], rtol=1e-6, atol=1e-7)
    assert float(out["loss"]) == pytest.approx(float(g["v1_loss"]), rel=1e-5)
    q, out, prob, idx = O.param_vq_forward(z, torch.from_numpy(g["v5_codebook"]), normalize="none", beta=0.25,
                                           gather_raw=True)
    assert np.array_equal(idx.view(z.shape[0], z.shape[2], z.shape[3]).numpy().astype(np.int32), g["v5_idx"])
    np.testing.assert_allclose(q.numpy(), g["v5_q"], rtol=1e-6, atol=1e-7)
    assert float(out["loss"]) == pytest.approx(float(g["v5_vq_loss"]), rel=1e-5)
    g2 = _load(golden_dir, "pq_v2_nchw.npz")
    q, out, prob, idx = O.v2_ema_vq_forward(torch.from_numpy(g2["z"]), torch.from_numpy(g2["embeddings"]))
    assert np.array_equal(idx.numpy().astype(np.int32), g2["idx"])
    np.testing.assert_allclose(q.numpy(), g2["q"], rtol=1e-6, atol=1e-7)


def test_evaluator_and_metrics(golden_dir):
    g = _load(golden_dir, "eval_probe.npz")
    C = 27
    ll, lp, cl, cp = O.evaluator_forward(torch.from_numpy(g["feat"]), torch.from_numpy(g["label"]),
                                         torch.from_numpy(g["clusters"]), torch.from_numpy(g["lin_w"]),
                                         torch.from_numpy(g["lin_b"]), C)
    assert np.array_equal(lp.numpy(), g["linear_preds"])
    assert np.array_equal(cp.numpy(), g["cluster_preds"])
    assert float(ll) == pytest.approx(float(g["linear_loss"]), rel=1e-5)
    assert float(cl) == pytest.approx(float(g["cluster_loss"]), rel=1e-5)
    label = torch.from_numpy(g["label"])
    for name, preds, hung in (("cluster", cp, True), ("linear", lp, False)):
        conf = O.confusion_update(torch.zeros(C, C, dtype=torch.long), preds, label, C)
        assert np.array_equal(conf.numpy(), g[f"{name}_confusion"])
        res = O.metrics_compute(conf, hung)
        assert float(res["iou"]) == pytest.approx(float(g[f"{name}_iou"]), rel=1e-6)
        assert float(res["accuracy"]) == pytest.approx(float(g[f"{name}_accuracy"]), rel=1e-6)


def test_confusion_edge_cases():
    C = 5
    preds = torch.tensor([0, 4, 5, -1, 2, 6, 3])
    label = torch.tensor([0, -1, 2, 1, 255, 4, 3])
    conf = O.confusion_update(torch.zeros(C + 2, C, dtype=torch.long), preds, label, C, extra_classes=2)
    # only (0,0) and (3,3) survive the mask: preds >= C are dropped even with extra classes (metric.py:49)
    assert conf.sum() == 2 and conf[0, 0] == 1 and conf[3, 3] == 1
    empty = O.confusion_update(torch.zeros(C, C, dtype=torch.long), preds[:0], label[:0], C)
    assert empty.sum() == 0


def test_knn(golden_dir):
    g = _load(golden_dir, "knn.npz")
    idx, vals = O.knn(torch.from_numpy(g["feats"]), k=8)
    assert np.array_equal(idx.numpy(), g["idx"])
    assert np.array_equal(idx[:, 0].numpy(), np.arange(idx.shape[0]))   # column 0 is the query itself


def knn_ref_case(g, tag):
    """Features + the neighbour table the reference's own statements produced (oracle/make_golden_knn.py).  F = 48:
    features stored; F = 768: regenerated from the stored seed, checked against their stored float64 sum."""
    n, Fd = (int(v) for v in g[f"{tag}_shape"])
    if f"{tag}_feats" in g.files:
        feats = torch.from_numpy(g[f"{tag}_feats"])
    else:
        torch.manual_seed(int(g[f"{tag}_seed"]))
        feats = torch.nn.functional.normalize(torch.randn(n, Fd), dim=1)
        if float(feats.double().sum()) != float(g[f"{tag}_sum"]):
            pytest.skip("torch's CPU generator on this host does not reproduce the fixture's features")
    return feats, g[f"{tag}_nns"], g[f"{tag}_vals"]


@pytest.mark.parametrize("tag", ["f48", "f768"])
def test_knn_reference_statements(golden_dir, tag):
    """The table computed by the reference's OWN kNN statements (data/precompute_knns.py:307-317, executed from the
    reference file by oracle/make_golden_knn.py; top-30 as hard-coded there) == the oracle's, bit for bit; computing
    it in query chunks (the reference's n_batches loop) changes nothing."""
    g = _load(golden_dir, "knn_ref_loop.npz")
    feats, nns, vals = knn_ref_case(g, tag)
    idx, v = O.knn(feats, k=30)
    assert np.array_equal(idx.numpy(), nns) and idx.dtype == torch.int64
    np.testing.assert_allclose(v.numpy(), vals, rtol=0, atol=1e-6)
    parts = [O.knn(feats, k=30, queries=feats[a:a + 75])[0] for a in range(0, feats.shape[0], 75)]
    assert np.array_equal(torch.cat(parts).numpy(), nns)
    assert np.array_equal(nns[:, 0], np.arange(nns.shape[0]))            # column 0 is the query itself


def test_knn_reference_statements_live(golden_dir):
    """On the build box (reference present): re-extract the statements from the reference file and re-run them -- the
    stored table is what they produce today."""
    ref = os.environ.get("EQUSS_REFERENCE", "/root/reference")
    if not os.path.exists(os.path.join(ref, "data", "precompute_knns.py")):
        pytest.skip("reference tree not present (GPU box)")
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_make_golden_knn", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "make_golden_knn.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    picked, n_batches, path = mk.reference_knn_statements()
    g = _load(golden_dir, "knn_ref_loop.npz")
    assert n_batches == int(g["n_batches"])
    for tag in ("f48", "f768"):
        feats, nns, _ = knn_ref_case(g, tag)
        with torch.no_grad():
            for nb in (n_batches, 4):
                assert np.array_equal(mk.run_reference(picked, path, feats, nb).numpy(), nns)


def test_histogram_percentiles_none_when_unreached():
    out = O.histogram_percentiles(torch.zeros(8), "x")
    assert out == {"x-p10": None, "x-p50": None, "x-p90": None}
    out = O.histogram_percentiles(torch.tensor([100.0, 0, 0, 0]), "x")
    assert out["x-p10"] == 0.0 and out["x-p90"] == 0.0


def test_expansion_head(golden_dir):
    """oracle expansion_head == the reference SegmentationHead output stored by oracle/make_golden_head.py
    (model/blocks/module.py:20-44), and the fp64 yardstick of the fixture is what the GPU test measures against."""
    g = np.load(os.path.join(golden_dir, "expansion_head.npz"))
    t = {k: torch.from_numpy(g[k]) for k in g.files}
    args = (t["cluster1_0_weight"], t["cluster1_0_bias"], t["cluster2_0_weight"], t["cluster2_0_bias"],
            t["cluster2_2_weight"], t["cluster2_2_bias"])
    out = O.expansion_head(t["x"], *args)
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=1e-5, atol=1e-6)
    out64 = O.expansion_head(t["x"].double(), *[a.double() for a in args])
    np.testing.assert_allclose(out64.numpy(), g["out_fp64"], rtol=1e-12, atol=1e-12)
    assert float(np.abs(g["out"] - g["out_fp64"]).max() / np.abs(g["out_fp64"]).max()) < 1e-6


def test_head_mirror_host_logic():
    """Host side of the head mirror that needs no GPU: reference state_dict keys, the autograd path (PyTorch
    convolutions on the same parameters) equals the oracle, the packed [W1 | W3] operand follows in-place updates,
    channels-last activations are described as flat rows, and CUDA-only entry points refuse CPU tensors."""
    from equss_b200 import _native as N
    from equss_b200.head import SegmentationHead, _PackedHead
    torch.manual_seed(3)
    head = SegmentationHead(16, 24)
    assert list(head.state_dict().keys()) == ["cluster1.0.weight", "cluster1.0.bias", "cluster2.0.weight",
                                              "cluster2.0.bias", "cluster2.2.weight", "cluster2.2.bias"]
    x = torch.randn(2, 16, 3, 5)
    out = head(x)                                   # parameters require grad -> differentiable torch path, any device
    sd = head.state_dict()
    ref = O.expansion_head(x, *[sd[k] for k in sd])
    torch.testing.assert_close(out, ref, rtol=1e-6, atol=1e-6)
    out.sum().backward()
    assert head.cluster1[0].weight.grad is not None
    with torch.no_grad(), pytest.raises(N.EqussNativeError):
        head(x)                                     # kernel path: no CPU fallback
    pk = _PackedHead()
    w13, b13 = pk.get(head.cluster1[0], head.cluster2[2])
    assert w13.shape == (24, 32) and torch.equal(w13[:, :16], head.cluster1[0].weight.reshape(24, 16))
    torch.testing.assert_close(b13, head.cluster1[0].bias + head.cluster2[2].bias)
    assert pk.get(head.cluster1[0], head.cluster2[2])[0] is w13            # cached
    with torch.no_grad():
        head.cluster2[2].weight.add_(1.0)
    w13b, _ = pk.get(head.cluster1[0], head.cluster2[2])
    assert w13b is not w13 and torch.equal(w13b[:, 16:], head.cluster2[2].weight.reshape(24, 16))
    # (B, D, h, w) view of NHWC memory == flat rows for the PQ kernels; NCHW-dense stays NCHW
    z = torch.randn(2, 3, 5, 8).permute(0, 3, 1, 2)            # (2, 8, 3, 5), channels-last memory
    assert N.is_channels_last(z)
    zd, d, layout = N.zdesc_for(z, 4)
    assert layout == "flat" and d == 2 and zd.n_pixels == 30 and zd.stride_s == 8 and zd.stride_c == 1
    zd2, _, layout2 = N.zdesc_for(z.contiguous(), 4)
    assert layout2 == "nchw" and zd2.stride_c == 15 and zd2.stride_s == 1
    assert N.f32_dense(z) is z and N.f32_dense(z.contiguous(), like=z).stride() == z.stride()


def test_unseg_metrics_compute_host_side(tmp_path, monkeypatch):
    """UnSegMetrics.compute / map_clusters (host side, no GPU): equal to the oracle for extra_classes == 0, and -- where
    the reference checkout exists (the build container) -- equal to the unmodified model/metric.py on random confusion
    matrices, with and without extra classes."""
    import importlib.util
    from equss_b200.metric import UnSegMetrics
    monkeypatch.chdir(tmp_path)                       # compute() writes ./class_matrix/... like the reference
    torch.manual_seed(0)
    for hung in (True, False):
        conf = torch.randint(0, 40, (7, 7))
        m = UnSegMetrics(7, 0, hung, torch.device("cpu"))
        m.confusion_matrix.copy_(conf)
        got, ref = m.compute(prefix="t"), O.metrics_compute(conf, hung)
        assert torch.equal(got["iou"], ref["iou"]) and torch.equal(got["accuracy"], ref["accuracy"])
    with pytest.raises(ValueError):
        UnSegMetrics(7, 2, False, torch.device("cpu"))
    ref_file = os.path.join(os.environ.get("EQUSS_REFERENCE", "/root/reference"), "model", "metric.py")
    if not os.path.exists(ref_file):
        return
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(ref_file)))       # the reference imports utils.dist_utils
    try:
        spec = importlib.util.spec_from_file_location("ref_metric", ref_file)
        ref_mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_mod)
    finally:
        sys.path.pop(0)
    for trial in range(20):
        C = int(torch.randint(3, 10, (1,)))
        extra = int(torch.randint(1, 4, (1,))) if trial % 2 else 0
        conf = torch.randint(0, 50, (C + extra, C))
        a = UnSegMetrics(C, extra, True, torch.device("cpu"))
        b = ref_mod.UnSegMetrics(C, extra, True, torch.device("cpu"))
        a.confusion_matrix.copy_(conf); b.confusion_matrix.copy_(conf)
        ra, rb = a.compute(prefix="t"), b.compute(prefix="t")
        assert torch.equal(ra["iou"], rb["iou"]) and torch.equal(ra["accuracy"], rb["accuracy"])
        assert torch.equal(a.histogram, b.histogram)
        ids = torch.randint(0, C + extra, (3, 4))
        assert torch.equal(a.map_clusters(ids), b.map_clusters(ids))


# ----------------------------------------------------------------------------------------------------------------
# variants pinned in round 2 (oracle/make_golden_variants.py): V4 EMA, V4/V5/V6 learned, V3 train, z_trainable, restart
# ----------------------------------------------------------------------------------------------------------------
def _mean_outs(per_subspace):
    M = len(per_subspace)
    keys = per_subspace[0].keys()
    return {k: sum(o[k] for o in per_subspace) / M for k in keys}


def _check_outs(out, g, s):
    keys = {k[len(f"out{s}/"):] for k in g.files if k.startswith(f"out{s}/")}
    assert set(out.keys()) == keys
    for k in keys:
        assert float(out[k]) == pytest.approx(float(g[f"out{s}/{k}"]), rel=1e-5, abs=1e-7), (s, k)


@pytest.mark.parametrize("mode", ["l2", "none"])
def test_new_vq_ema_trajectory(golden_dir, mode):
    g = _load(golden_dir, f"pq_newvq_ema_{mode}.npz")
    M, K, ts = int(g["M"]), int(g["K"]), float(g["jsd_ts"])
    w0 = torch.from_numpy(g["weight0"])
    states = [O.EmaState(w0[i]) for i in range(M)]
    exact = [torch.zeros(K) for _ in range(M)]
    for s in range(4):
        z = torch.from_numpy(g[f"z{s}"])
        d = z.shape[1] // M
        res = [O.new_vq_ema_forward(z[:, i * d:(i + 1) * d], states[i], exact[i], normalize=mode, beta=0.25, jsd_ts=ts,
                                    training=s < 3) for i in range(M)]
        assert np.array_equal(torch.stack([r[3] for r in res]).numpy().astype(np.int32), g[f"idx{s}"])
        np.testing.assert_allclose(torch.cat([r[0] for r in res], dim=1).numpy(), g[f"zq{s}"], rtol=1e-6, atol=1e-7)
        _check_outs(_mean_outs([r[1] for r in res]), g, s)
        np.testing.assert_allclose(torch.stack([st.weight for st in states]).numpy(), g[f"weight_after{s}"], rtol=1e-5, atol=1e-7)
        assert np.array_equal(torch.stack(exact).numpy(), g[f"exact_after{s}"])


@pytest.mark.parametrize("variant,mode", [("new_vq", "l2"), ("new_vq", "z_norm"), ("pqgo_cls", "l2"),
                                          ("pqgo_cls", "z_trainable"), ("pqgo", "z_norm")])
def test_inline_codebooks(golden_dir, variant, mode):
    g = _load(golden_dir, f"pq_inline_{variant}_{mode}.npz")
    M, K, ts = int(g["M"]), int(g["K"]), float(g["jsd_ts"])
    cbs = torch.from_numpy(g["codebook"])
    zm = torch.from_numpy(g["z_mean"]) if mode == "z_trainable" else None
    zl = torch.from_numpy(g["z_log_var"]) if mode == "z_trainable" else None
    exact = [torch.zeros(K) for _ in range(M)]
    for s, training in ((0, True), (1, False)):
        z = torch.from_numpy(g[f"z{s}"])
        d = z.shape[1] // M
        res = [O.inline_codebook_forward(z[:, i * d:(i + 1) * d], cbs[i], exact[i], variant=variant, normalize=mode,
                                         beta=0.25, book=0.6, jsd_ts=ts, training=training,
                                         z_mean=None if zm is None else zm[i], z_log_var=None if zl is None else zl[i])
               for i in range(M)]
        assert np.array_equal(torch.stack([r[3] for r in res]).numpy().astype(np.int32), g[f"idx{s}"])
        np.testing.assert_allclose(torch.cat([r[0] for r in res], dim=1).numpy(), g[f"zq{s}"], rtol=1e-6, atol=1e-7)
        _check_outs(_mean_outs([r[1] for r in res]), g, s)
        assert np.array_equal(torch.stack(exact).numpy(), g[f"exact_after{s}"])


def test_v2_train_and_z_trainable_and_restart(golden_dir):
    g = _load(golden_dir, "pq_v2_train.npz")
    M, K, dec = int(g["M"]), int(g["K"]), float(g["decay"])
    emb = [torch.from_numpy(g["embeddings0"][i]).clone() for i in range(M)]
    Ns = [torch.zeros(K) for _ in range(M)]
    zav = [e.clone() for e in emb]
    for s in range(3):
        z = torch.from_numpy(g[f"z{s}"])
        d = z.shape[1] // M
        res = [O.v2_ema_vq_train_step(z[:, i * d:(i + 1) * d], emb[i], Ns[i], zav[i], beta=0.25, decay=dec) for i in range(M)]
        assert np.array_equal(torch.stack([r[3] for r in res]).numpy().astype(np.int32), g[f"idx{s}"])
        np.testing.assert_allclose(torch.cat([r[0] for r in res], dim=1).numpy(), g[f"q{s}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(torch.stack(emb).numpy(), g[f"embeddings_after{s}"], rtol=1e-5, atol=1e-7)
    g = _load(golden_dir, "pq_ztrainable.npz")
    M, K, dec = int(g["M"]), int(g["K"]), float(g["decay"])
    w0 = torch.from_numpy(g["weight0"])
    states = [O.EmaState(w0[i], decay=dec) for i in range(M)]
    exact = [torch.zeros(K) for _ in range(M)]
    zm = [torch.from_numpy(g["z_mean0"][i]).clone() for i in range(M)]
    zl = [torch.from_numpy(g["z_log_var0"][i]).clone() for i in range(M)]
    for s in range(4):
        z = torch.from_numpy(g[f"z{s}"])
        d = z.shape[1] // M
        res = [O.ema_vq_forward(z[:, i * d:(i + 1) * d], states[i], exact[i], normalize="z_trainable", beta=0.25,
                                training=s < 3, z_mean=zm[i], z_log_var=zl[i], ema_decay=dec) for i in range(M)]
        assert np.array_equal(torch.stack([r[3] for r in res]).numpy().astype(np.int32), g[f"idx{s}"])
        np.testing.assert_allclose(torch.stack(zm).numpy(), g[f"z_mean_after{s}"], rtol=1e-6, atol=1e-7)
    g = _load(golden_dir, "pq_restart.npz")
    import random
    dead, cand = O.restart_candidates(torch.zeros(int(g["K"]), dtype=torch.long), torch.from_numpy(g["a_z"]), random.Random(5))
    np.testing.assert_array_equal(cand.numpy(), g["a_init_rows"])
    assert len(dead) == int(g["K"])


def test_stego_helper_fixture(golden_dir):
    g = _load(golden_dir, "stego.npz")
    cfg = {"pointwise": True, "zero_clamp": True, "stabilize": False}
    f1, f2, c1, c2 = (torch.from_numpy(g[k]) for k in ("f1", "f2", "c1", "c2"))
    for tag, cfgv in (("a", cfg), ("b", {"pointwise": False, "zero_clamp": False, "stabilize": True})):
        loss, cd = O.stego_helper(f1, f2, c1, c2, 0.2, cfgv)
        np.testing.assert_allclose(loss.numpy(), g[f"{tag}_loss"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(cd.numpy(), g[f"{tag}_cd"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name,mode,weighted", [("pq_flag_ema_weighted_none", "none", True),
                                                ("pq_flag_ema_weighted_l2", "l2", True),
                                                ("pq_flag_ema_gumbel_l2", "l2", False)])
def test_research_flags_ema(golden_dir, name, mode, weighted):
    """use_weighted_sum / use_gumbel of model/quantizer.py (fixtures: oracle/make_golden_flags.py, the unmodified
    reference with the Gumbel noise pinned)."""
    g = _load(golden_dir, name + ".npz")
    M, K = int(g["M"]), int(g["K"])
    steps = int(g["steps"])
    w0 = torch.from_numpy(g["weight0"])
    states = [O.EmaState(w0[i]) for i in range(M)]
    exact = [torch.zeros(K) for _ in range(M)]
    for s in range(steps + 1):
        z = torch.from_numpy(g[f"z{s}"])
        d = z.shape[1] // M
        noise = torch.from_numpy(g[f"noise{s}"]) if f"noise{s}" in g.files else None
        res = [O.ema_vq_forward(z[:, i * d:(i + 1) * d], states[i], exact[i], normalize=mode, beta=0.25, training=s < steps,
                                use_weighted_sum=weighted, gumbel_noise=None if noise is None else noise[i]) for i in range(M)]
        assert np.array_equal(torch.stack([r[3] for r in res]).numpy().astype(np.int32), g[f"idx{s}"])
        np.testing.assert_allclose(torch.cat([r[0] for r in res], dim=1).numpy(), g[f"zq{s}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(torch.stack([st.weight for st in states]).numpy(), g[f"weight_after{s}"], rtol=1e-5, atol=1e-7)
        loss = sum(float(r[1]["loss"]) for r in res) / M
        assert loss == pytest.approx(float(g[f"out{s}/loss"]), rel=1e-5)
    if not weighted:
        assert int(g["flips"]) > 0 and float(g["min_gumbel_gap"]) > 2e-3


def test_research_flags_inline(golden_dir):
    g = _load(golden_dir, "pq_flag_param_gumbel.npz")
    q, out, prob, idx = O.param_vq_forward(torch.from_numpy(g["z"]), torch.from_numpy(g["weight"]), normalize="l2", beta=0.25,
                                           gumbel_noise=torch.from_numpy(g["noise"]))
    assert np.array_equal(idx.numpy().astype(np.int32), g["idx"]) and int(g["flips"]) > 0
    np.testing.assert_allclose(q.numpy(), g["zq"], rtol=1e-6, atol=1e-7)
    g = _load(golden_dir, "pq_flag_newvq_ema_weighted.npz")
    M, K, ts = int(g["M"]), int(g["K"]), float(g["jsd_ts"])
    w0 = torch.from_numpy(g["weight0"])
    states = [O.EmaState(w0[i]) for i in range(M)]
    exact = [torch.zeros(K) for _ in range(M)]
    for s in range(3):
        z = torch.from_numpy(g[f"z{s}"])
        d = z.shape[1] // M
        res = [O.new_vq_ema_forward(z[:, i * d:(i + 1) * d], states[i], exact[i], normalize="none", beta=0.25, jsd_ts=ts,
                                    training=s < 2, use_weighted_sum=True) for i in range(M)]
        np.testing.assert_allclose(torch.cat([r[0] for r in res], dim=1).numpy(), g[f"zq{s}"], rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(torch.stack([st.weight for st in states]).numpy(), g[f"weight_after{s}"], rtol=1e-5, atol=1e-8)
    for variant in ("new_vq", "pqgo"):
        g = _load(golden_dir, f"pq_flag_inline_{variant}_weighted.npz")
        q, out, _, _ = O.inline_codebook_forward(torch.from_numpy(g["z"]), torch.from_numpy(g["weight"]), torch.zeros(int(g["K"])),
                                                 variant=variant, normalize="none", beta=0.25, jsd_ts=float(g["jsd_ts"]),
                                                 training=True, use_weighted_sum=True)
        np.testing.assert_allclose(q.numpy(), g["zq"], rtol=1e-6, atol=1e-8)
        assert float(out["vq-loss"]) == pytest.approx(float(g["out/vq-loss"]), rel=1e-5)


def test_research_flag_pq_dropout(golden_dir):
    """pq_dropout of dino_new_vq / dino_pqgo (fixtures: oracle/make_golden_dropout.py, the unmodified reference with
    ``torch.cuda.FloatTensor`` replaced by the stored uniform draws)."""
    g = _load(golden_dir, "pq_flag_newvq_ema_dropout.npz")
    M, K, ts, p = int(g["M"]), int(g["K"]), float(g["jsd_ts"]), float(g["pq_dropout"])
    w0 = torch.from_numpy(g["weight0"])
    states = [O.EmaState(w0[i]) for i in range(M)]
    exact = [torch.zeros(K) for _ in range(M)]
    for s in range(3):
        z = torch.from_numpy(g[f"z{s}"])
        keep = torch.from_numpy(g[f"u{s}"]) > p
        d = z.shape[1] // M
        res = [O.new_vq_ema_forward(z[:, i * d:(i + 1) * d], states[i], exact[i], normalize="l2", beta=0.25, jsd_ts=ts,
                                    training=s < 2, dropout_keep=keep[i]) for i in range(M)]
        assert np.array_equal(torch.stack([r[3] for r in res]).numpy().astype(np.int32), g[f"idx{s}"])
        assert all(int(r[3].max()) < int(keep[i].sum()) for i, r in enumerate(res))       # positions in the kept list
        np.testing.assert_allclose(torch.cat([r[0] for r in res], dim=1).numpy(), g[f"zq{s}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(torch.cat([r[2] for r in res], dim=-1).numpy(), g[f"prob{s}"], rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(torch.stack([st.weight for st in states]).numpy(), g[f"weight_after{s}"], rtol=1e-5, atol=1e-7)
        for k in ("vq-loss", "jsd", "entropy") + (("codebook-usage",) if s < 2 else ()):
            got = sum(float(r[1][k]) for r in res) / M
            assert got == pytest.approx(float(g[f"out{s}/{k}"]), rel=1e-5, abs=1e-8), (s, k)
    for name in ("new_vq", "pqgo", "new_vq_weighted"):
        g = _load(golden_dir, f"pq_flag_inline_{name}_dropout.npz")
        K = int(g["K"])
        keep = torch.from_numpy(g["u"]) > float(g["pq_dropout"])
        q, out, prob, idx = O.inline_codebook_forward(torch.from_numpy(g["z"]), torch.from_numpy(g["weight"]), torch.zeros(K),
                                                      variant=str(g["variant"]), normalize=str(g["mode"]), beta=0.25,
                                                      jsd_ts=float(g["jsd_ts"]), training=True,
                                                      use_weighted_sum=bool(g["weighted"]), dropout_keep=keep)
        assert np.array_equal(idx.numpy().astype(np.int32), g["idx"]) and prob.shape[1] == int(keep.sum())
        np.testing.assert_allclose(q.numpy(), g["zq"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(prob.numpy(), g["prob"].reshape(prob.shape), rtol=1e-6, atol=1e-8)
        assert float(out["vq-loss"]) == pytest.approx(float(g["out/vq-loss"]), rel=1e-6)
        assert float(out["codebook-usage"]) == pytest.approx(float(g["out/codebook-usage"]), rel=1e-6)
