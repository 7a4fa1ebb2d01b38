"""The plain-C restatement of the path's index / integer side (oracle/equss_oracle_c.c) against the fixtures the
unmodified reference produced (tests/golden) and against the torch oracle on seeded inputs (CPU only).

Two independent restatements -- torch ops in oracle/equss_oracle.py, scalar C here -- have to agree with the reference's
stored outputs: indices, counts and confusion matrices exactly (the fixtures hold no fp32 near-ties), floats up to fp32
summation order.  Both are test infrastructure; the product path loads neither."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest
import torch

import equss_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODES = {"none": 0, "l2": 1, "z_norm": 2}


@pytest.fixture(scope="module")
def lib():
    spec = importlib.util.spec_from_file_location("_equss_oracle_build_c", os.path.join(ROOT, "oracle", "build_c.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    L = C.CDLL(mod.build())
    for name in ("eqo_pq_assign_gather", "eqo_counts_sums", "eqo_ema_update", "eqo_usage_percentiles",
                 "eqo_confusion_update", "eqo_knn_topk", "eqo_probe_argmax"):
        getattr(L, name).restype = C.c_int
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _assign(L, z, cb, mode, gather_raw=False, keep=None):
    z, cb = np.ascontiguousarray(z, np.float32), np.ascontiguousarray(cb, np.float32)
    n, (M, K, d) = z.shape[0], cb.shape
    idx = np.empty((M, n), np.int32)
    out = np.empty_like(z)
    sq = np.zeros(M, np.float64)
    keep8 = None if keep is None else np.ascontiguousarray(keep, np.uint8)
    rc = L.eqo_pq_assign_gather(_p(z), C.c_int64(n), M, K, d, _p(cb), MODES[mode], int(gather_raw), _p(keep8), _p(idx), _p(out), _p(sq))
    assert rc == 0
    return idx, out, sq


def _pct(L, count):
    p3 = np.empty(3, np.float32)
    assert L.eqo_usage_percentiles(_p(np.ascontiguousarray(count, np.float32)), int(count.shape[0]), _p(p3)) == 0
    return p3


@pytest.mark.parametrize("mode", ["l2", "z_norm", "none"])
def test_ema_trajectory_matches_the_reference(lib, golden_dir, mode):
    """ProductQuantizerWrapper(EMAVectorQuantizer), 3 training steps + 1 evaluation step (model/quantizer.py:383-542)."""
    g = np.load(os.path.join(golden_dir, f"pq_ema_{mode}.npz"))
    M, K = int(g["M"]), int(g["K"])
    weight = g["weight0"].copy()
    wavg, vqc, exact = weight.copy(), np.zeros((M, K), np.float32), np.zeros((M, K), np.float32)
    d = weight.shape[2]
    for s in range(4):
        z = g[f"z{s}"]
        n = z.shape[0]
        idx, out, sq = _assign(lib, z, weight, mode)
        assert np.array_equal(idx, g[f"idx{s}"])
        np.testing.assert_allclose(out, g[f"zq{s}"], rtol=1e-5, atol=1e-6)
        commit = float((sq / (n * d)).mean())
        assert commit == pytest.approx(float(g[f"out{s}/commitment-loss"]), rel=1e-5)
        assert 0.25 * commit == pytest.approx(float(g[f"out{s}/loss"]), rel=1e-5)
        if s < 3:
            count, total = np.empty((M, K), np.float32), np.empty((M, K, d), np.float32)
            assert lib.eqo_counts_sums(_p(np.ascontiguousarray(z)), C.c_int64(n), M, K, d, _p(idx), _p(count), _p(total)) == 0
            exact += count
            for m in range(M):
                assert lib.eqo_ema_update(_p(count[m]), _p(total[m]), K, d, C.c_float(0.99), C.c_float(1e-5),
                                          _p(vqc[m]), _p(wavg[m]), _p(weight[m])) == 0
            cur = np.stack([_pct(lib, count[m]) for m in range(M)]).mean(axis=0)
            tot = np.stack([_pct(lib, exact[m]) for m in range(M)]).mean(axis=0)
            for t, tag in enumerate(("p10", "p50", "p90")):
                for got, pre in ((cur, "current"), (tot, "total")):
                    ref = float(g[f"out{s}/{pre}-{tag}"])
                    assert (np.isnan(ref) and np.isnan(got[t])) or got[t] == pytest.approx(ref, abs=1e-6), (s, pre, tag)
            usage = np.mean([(K - int((count[m] == 0).sum())) / K for m in range(M)])
            assert usage == pytest.approx(float(g[f"out{s}/codebook-usage"]), abs=1e-6)
        assert np.array_equal(exact, g[f"exact_after{s}"])
        np.testing.assert_allclose(vqc, g[f"vq_count_after{s}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(wavg, g[f"weight_avg_after{s}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(weight, g[f"weight_after{s}"], rtol=1e-5, atol=1e-6)
        assert float(np.abs(weight).sum() / M) == pytest.approx(float(g[f"out{s}/codebook-sum"]), rel=1e-5)


def test_pq_dropout_indices_match_the_reference(lib, golden_dir):
    """dino_new_vq.EMACodebook with pq_dropout: indices are positions in the kept list, the gather reads the full RAW
    codebook at those positions (dino_new_vq.py:388-403)."""
    g = np.load(os.path.join(golden_dir, "pq_flag_newvq_ema_dropout.npz"))
    M, p = int(g["M"]), float(g["pq_dropout"])
    weight = g["weight0"]
    for s in range(3):
        z = g[f"z{s}"]
        B, D, h, w = z.shape
        rows = np.ascontiguousarray(z.transpose(0, 2, 3, 1).reshape(-1, D))
        idx, out, _ = _assign(lib, rows, weight, "l2", gather_raw=True, keep=g[f"u{s}"] > p)
        assert np.array_equal(idx, g[f"idx{s}"])
        np.testing.assert_allclose(out.reshape(B, h, w, D).transpose(0, 3, 1, 2), g[f"zq{s}"], rtol=1e-5, atol=1e-6)
        weight = g[f"weight_after{s}"]


def test_probe_argmax_and_confusion_match_the_reference(lib, golden_dir):
    """UnSegEvaluator.forward + UnSegMetrics.update (model/evaluator.py:46-82, model/metric.py:44-58): token-resolution
    logits, bilinear interpolation at label resolution, argmax, masked histogram."""
    g = np.load(os.path.join(golden_dir, "eval_probe.npz"))
    feat, label = g["feat"].astype(np.float64), np.ascontiguousarray(g["label"])
    B, D, h, w = feat.shape
    H, W = label.shape[1:]
    Cn = 27
    cl = g["clusters"].astype(np.float64)
    cl /= np.maximum(np.linalg.norm(cl, axis=1, keepdims=True), 1e-12)
    rows = feat.transpose(0, 2, 3, 1).reshape(-1, D)
    for name, wmat, bias in (("cluster", cl, np.zeros(Cn)), ("linear", g["lin_w"].astype(np.float64), g["lin_b"].astype(np.float64))):
        logits = np.ascontiguousarray((rows @ wmat.T + bias).astype(np.float32))
        pred = np.empty((B, H, W), np.int64)
        assert lib.eqo_probe_argmax(_p(logits), B, h, w, Cn, Cn, H, W, _p(pred)) == 0
        assert np.array_equal(pred, g[f"{name}_preds"])
        conf = np.zeros((Cn, Cn), np.int64)
        assert lib.eqo_confusion_update(_p(pred), _p(label), C.c_int64(label.size), Cn, 0, _p(conf)) == 0
        assert np.array_equal(conf, g[f"{name}_confusion"])
    # edge cases of the mask (metric.py:49): ignore labels, out-of-range labels, predictions in the extra rows
    preds = np.array([0, 4, 5, -1, 2, 6, 3], np.int64)
    lab = np.array([0, -1, 2, 1, 255, 4, 3], np.int64)
    conf = np.zeros((7, 5), np.int64)
    assert lib.eqo_confusion_update(_p(preds), _p(lab), C.c_int64(7), 5, 2, _p(conf)) == 0
    want = O.confusion_update(torch.zeros(7, 5, dtype=torch.long), torch.from_numpy(preds), torch.from_numpy(lab), 5, extra_classes=2)
    assert np.array_equal(conf, want.numpy()) and conf.sum() == 2
    assert lib.eqo_confusion_update(_p(preds), _p(lab), C.c_int64(0), 5, 0, _p(conf)) == 0 and conf.sum() == 2      # empty input


def test_knn_matches_the_fixture(lib, golden_dir):
    g = np.load(os.path.join(golden_dir, "knn.npz"))
    feats = np.ascontiguousarray(g["feats"])
    n, F = feats.shape
    idx = np.empty((n, 8), np.int64)
    assert lib.eqo_knn_topk(_p(feats), C.c_int64(n), _p(feats), C.c_int64(n), F, 8, _p(idx)) == 0
    assert np.array_equal(idx, g["idx"]) and np.array_equal(idx[:, 0], np.arange(n))


@pytest.mark.parametrize("tag", ["f48", "f768"])
def test_knn_matches_the_reference_statements(lib, golden_dir, tag):
    """k = 30 table produced by the reference's own statements (oracle/make_golden_knn.py), whole and in query chunks."""
    from test_oracle_golden import knn_ref_case
    g = np.load(os.path.join(golden_dir, "knn_ref_loop.npz"))
    feats, nns, _ = knn_ref_case(g, tag)
    feats = np.ascontiguousarray(feats.numpy())
    n, F = feats.shape
    idx = np.empty((n, 30), np.int64)
    assert lib.eqo_knn_topk(_p(feats), C.c_int64(n), _p(feats), C.c_int64(n), F, 30, _p(idx)) == 0
    assert np.array_equal(idx, nns)
    part = np.empty((75, 30), np.int64)
    q = np.ascontiguousarray(feats[150:225])
    assert lib.eqo_knn_topk(_p(q), C.c_int64(75), _p(feats), C.c_int64(n), F, 30, _p(part)) == 0
    assert np.array_equal(part, nns[150:225])


def test_c_and_torch_restatements_agree_on_random_inputs(lib):
    """Seeded random inputs beyond the fixtures: ragged sizes, d not a multiple of 4, K = 1, duplicate codes (the first
    of equal distances wins), percentiles that are never reached."""
    rng = np.random.default_rng(3)
    for n, M, K, d, mode in ((1, 1, 1, 3, "l2"), (77, 3, 19, 5, "z_norm"), (130, 2, 40, 16, "none"), (64, 4, 8, 8, "l2")):
        z = rng.standard_normal((n, M * d)).astype(np.float32)
        cb = rng.standard_normal((M, K, d)).astype(np.float32)
        if K > 2:
            cb[:, K - 1] = cb[:, 0]                               # exact duplicate: index 0 must win, never K - 1
        idx, out, sq = _assign(lib, z, cb, mode)
        for m in range(M):
            zn, cn = O.normalize_pair(torch.from_numpy(z[:, m * d:(m + 1) * d]), torch.from_numpy(cb[m]), mode)
            dist = O.sq_distance(zn, cn)
            ref = torch.argmin(dist, dim=1).numpy()
            bad = np.nonzero(ref != idx[m])[0]
            for r in bad:                                         # only genuine fp32 near-ties may differ
                assert abs(float(dist[r, ref[r]] - dist[r, idx[m][r]])) <= 1e-5 * float(dist[r].abs().max())
            assert (idx[m] != K - 1).all() or K <= 2
            ok = ref == idx[m]
            want = (zn + (cn[torch.from_numpy(ref)] - zn)).numpy()
            np.testing.assert_allclose(out[ok, m * d:(m + 1) * d], want[ok], rtol=1e-5, atol=1e-6)
    p3 = _pct(lib, np.zeros(16, np.float32))
    assert np.isnan(p3).all()                                     # nothing selected: no level is ever reached
    ref = O.histogram_percentiles(torch.tensor([5.0, 0.0, 1.0, 3.0]), "x")
    got = _pct(lib, np.array([5.0, 0.0, 1.0, 3.0], np.float32))
    for t, tag in enumerate(("p10", "p50", "p90")):
        r = ref[f"x-{tag}"]
        assert (r is None and np.isnan(got[t])) or got[t] == pytest.approx(r)
