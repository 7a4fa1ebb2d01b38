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


def test_histogram_percentiles_none_when_unreached():
    out = O.histogram_percentiles(torch.zeros(8), "x")
    assert out == {"x-p10": None, "x-p50": None, "x-p90": None}
    out = O.histogram_percentiles(torch.tensor([100.0, 0, 0, 0]), "x")
    assert out["x-p10"] == 0.0 and out["x-p90"] == 0.0
