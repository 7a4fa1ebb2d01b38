"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

For every case the script (1) runs the reference module on seeded CPU inputs, (2) asserts that the
oracle restatement (oracle/equss_oracle.py) reproduces the reference output BIT FOR BIT on the same
inputs and thread count, (3) checks that no row of the case is an fp32 near-tie (top-2 relative margin
> 2e-5), so the CUDA kernels must match the stored indices exactly, and (4) stores inputs + reference
outputs as a small fixture.  The reference imports torchmetrics and pydensecrf on paths that are never
executed here; both are absent from this image and are stubbed with empty modules.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("EQUSS_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.dont_write_bytecode = True


def import_reference():
    for name in ("torchmetrics", "torchmetrics.functional", "pydensecrf", "pydensecrf.densecrf", "pydensecrf.utils"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name == "torchmetrics.functional":
                m.pairwise_cosine_similarity = None
            if name == "pydensecrf.utils":
                m.unary_from_softmax = None
            sys.modules[name] = m
    sys.path.insert(0, REF)
    import model.quantizer as q1          # noqa
    import model.quantizer_v2 as q2       # noqa
    import model.dino_pqgo as pqgo        # noqa
    import model.evaluator as ev          # noqa
    import model.metric as metric         # noqa
    return q1, q2, pqgo, ev, metric


def np_(t):
    return t.detach().cpu().numpy()


def assert_same(a, b, what):
    if isinstance(a, torch.Tensor):
        assert torch.equal(a, b), f"oracle != reference for {what}: max abs diff {(a - b).abs().max()}"
    else:
        assert a == b or (a is None and b is None), f"oracle != reference for {what}: {a} vs {b}"


def min_margin(z_norm, cb_norm):
    d = ((z_norm.double()[:, None, :] - cb_norm.double()[None]) ** 2).sum(-1)
    top2 = torch.topk(d, 2, dim=1, largest=False)[0]
    return float(((top2[:, 1] - top2[:, 0]) / top2[:, 1].clamp_min(1e-300)).min())


def main():
    torch.set_num_threads(1)
    sys.path.insert(0, HERE)
    import equss_oracle as O
    q1, q2, pqgo, ev, metric = import_reference()
    os.makedirs(OUT, exist_ok=True)

    # ------------------------------------------------------------------ EMA PQ (model/quantizer.py), flat
    for mode in ("l2", "z_norm", "none"):
        torch.manual_seed(7)
        M, K, D, n, steps = 4, 32, 64, 256, 3
        ref = q1.ProductQuantizerWrapper(M, K, D, beta=0.25, normalize=mode, decay=0.99, eps=1e-5,
                                         quantizer_cls=q1.EMAVectorQuantizer)
        scale = 1.0 if mode != "none" else 0.05
        with torch.no_grad():
            for qz in ref.quantizers:        # spread the codes (default init is U(-1/K, 1/K))
                qz.codebook.weight.copy_(torch.randn(K, D // M) * scale)
                qz.codebook.weight_avg.copy_(qz.codebook.weight)
        w0 = torch.stack([qz.codebook.weight.clone() for qz in ref.quantizers])
        states = [O.EmaState(w0[i]) for i in range(M)]
        exact = [torch.zeros(K) for _ in range(M)]
        fix = {"weight0": np_(w0), "M": M, "K": K, "mode": mode}
        ref.train()
        for s in range(steps + 1):
            training = s < steps
            if not training:
                ref.eval()
            z = torch.randn(n, D) * scale * (1.0 + 0.1 * s)
            with torch.no_grad():
                rq, rout, rprob = ref(z)
            oq, oout, oprob, oidx = O.pq_forward_ema(z, states, exact, normalize=mode, beta=0.25, training=training)
            assert_same(oq, rq, f"ema/{mode}/step{s}/z_q")
            assert_same(oprob, rprob, f"ema/{mode}/step{s}/prob")
            for k, v in rout.items():
                assert_same(oout[k], v, f"ema/{mode}/step{s}/{k}")
            for i, qz in enumerate(ref.quantizers):
                assert_same(states[i].weight, qz.codebook.weight, f"ema/{mode}/step{s}/weight{i}")
                assert_same(states[i].weight_avg, qz.codebook.weight_avg, f"ema/{mode}/step{s}/weight_avg{i}")
                assert_same(states[i].vq_count, qz.codebook.vq_count, f"ema/{mode}/step{s}/vq_count{i}")
                assert_same(exact[i], qz.vq_count, f"ema/{mode}/step{s}/exact{i}")
            fix[f"z{s}"] = np_(z)
            fix[f"zq{s}"] = np_(rq)
            fix[f"idx{s}"] = np_(oidx).astype(np.int32)
            if not training:
                fix[f"prob{s}"] = np_(rprob)
            fix[f"weight_after{s}"] = np_(torch.stack([qz.codebook.weight for qz in ref.quantizers]))
            fix[f"weight_avg_after{s}"] = np_(torch.stack([qz.codebook.weight_avg for qz in ref.quantizers]))
            fix[f"vq_count_after{s}"] = np_(torch.stack([qz.codebook.vq_count for qz in ref.quantizers]))
            fix[f"exact_after{s}"] = np_(torch.stack([qz.vq_count for qz in ref.quantizers]))
            for k, v in rout.items():
                fix[f"out{s}/{k}"] = np.float64(float(v)) if v is not None else np.float64(np.nan)
        # margins are checked against the codebook each step used; recompute with a fresh replay
        states2 = [O.EmaState(w0[i]) for i in range(M)]
        exact2 = [torch.zeros(K) for _ in range(M)]
        worst = 1.0
        for s in range(steps + 1):
            z = torch.from_numpy(fix[f"z{s}"])
            for i, zc in enumerate(torch.chunk(z, M, dim=1)):
                zn, cn = O.normalize_pair(zc, states2[i].weight, mode)
                worst = min(worst, min_margin(zn, cn))
            O.pq_forward_ema(z, states2, exact2, normalize=mode, beta=0.25, training=s < steps)
        assert worst > 2e-5, f"golden case ema/{mode} contains a near-tie (margin {worst}); change the seed"
        fix["min_margin"] = np.float64(worst)
        np.savez_compressed(os.path.join(OUT, f"pq_ema_{mode}.npz"), **fix)
        print(f"pq_ema_{mode}: ok, min top-2 margin {worst:.3e}")

    # ------------------------------------------------------------------ learned codebook, NCHW (V1 and V5)
    torch.manual_seed(11)
    B, d, h, w, K = 2, 16, 9, 7, 24
    z = torch.randn(B, d, h, w)
    vq = q1.VectorQuantizer(K, d, beta=0.25, normalize="l2")
    vq.eval()
    with torch.no_grad():
        rq, rout, rprob = vq(z)
    oq, oout, oprob, oidx = O.param_vq_forward(z, vq.codebook.weight.detach(), normalize="l2", beta=0.25)
    assert_same(oq, rq, "param/v1/q"); assert_same(oprob, rprob, "param/v1/prob")
    for k in ("loss", "codebook_loss", "commitment_loss"):
        assert_same(oout[k], rout[k], f"param/v1/{k}")
    zf = z.permute(0, 2, 3, 1).reshape(-1, d)
    zn, cn = O.normalize_pair(zf, vq.codebook.weight.detach(), "l2")
    mm1 = min_margin(zn, cn)
    cb5 = pqgo.Codebook(K, d, beta=0.25, book=1.0, normalize="none", need_initialized="none")
    with torch.no_grad():
        cb5.embedding.weight.copy_(torch.randn(K, d) * 0.5)
    cb5.eval()
    with torch.no_grad():
        r5q, r5out, r5prob, r5idx = cb5(z, torch.zeros_like(z))
    o5q, o5out, o5prob, o5idx = O.param_vq_forward(z, cb5.embedding.weight.detach(), normalize="none", beta=0.25,
                                                   book=1.0, gather_raw=True, temperature=1.0)
    assert_same(o5q, r5q, "param/v5/q"); assert_same(o5idx.view(B, h, w), r5idx, "param/v5/idx")
    assert_same(o5prob.view(B, h, w, -1), r5prob, "param/v5/prob")
    assert_same(o5out["loss"], r5out["vq-loss"], "param/v5/vq-loss")
    mm5 = min_margin(zf, cb5.embedding.weight.detach())
    assert min(mm1, mm5) > 2e-5, (mm1, mm5)
    np.savez_compressed(os.path.join(OUT, "pq_param_nchw.npz"), z=np_(z), K=K,
                        v1_codebook=np_(vq.codebook.weight), v1_q=np_(rq), v1_idx=np_(oidx).astype(np.int32),
                        v1_prob=np_(rprob), v1_loss=np.float64(float(rout["loss"])),
                        v1_codebook_loss=np.float64(float(rout["codebook_loss"])),
                        v1_commitment_loss=np.float64(float(rout["commitment_loss"])),
                        v5_codebook=np_(cb5.embedding.weight), v5_q=np_(r5q), v5_idx=np_(r5idx).astype(np.int32),
                        v5_prob=np_(r5prob), v5_vq_loss=np.float64(float(r5out["vq-loss"])))
    print(f"pq_param_nchw: ok, margins {mm1:.3e} {mm5:.3e}")

    # ------------------------------------------------------------------ quantizer_v2 EMA (eval), NCHW
    torch.manual_seed(13)
    v2 = q2.EMAVectorQuantizer(K, d, beta=0.25)
    with torch.no_grad():
        v2.embeddings.copy_(torch.randn(K, d))
    v2.eval()
    with torch.no_grad():
        r2q, r2out, r2prob = v2(z)
    o2q, o2out, o2prob, o2idx = O.v2_ema_vq_forward(z, v2.embeddings, beta=0.25)
    assert_same(o2q, r2q, "v2/q"); assert_same(o2prob, r2prob, "v2/prob")
    assert_same(o2out["loss"], r2out["loss"], "v2/loss")
    zn, cn = O.normalize_pair(zf, v2.embeddings, "l2")
    mm2 = min_margin(zn, cn)
    assert mm2 > 2e-5, mm2
    np.savez_compressed(os.path.join(OUT, "pq_v2_nchw.npz"), z=np_(z), embeddings=np_(v2.embeddings), q=np_(r2q),
                        idx=np_(o2idx).astype(np.int32), loss=np.float64(float(r2out["loss"])))
    print(f"pq_v2_nchw: ok, margin {mm2:.3e}")

    # ------------------------------------------------------------------ evaluator + metrics
    torch.manual_seed(17)
    B, D, h, w, H, W, C = 2, 32, 10, 10, 40, 40, 27
    feat = torch.randn(B, D, h, w)
    label = torch.randint(-1, C, (B, H, W))
    evm = ev.UnSegEvaluator(D, C, extra_classes=0)
    evm.eval()
    with torch.no_grad():
        r_ll, r_lp, r_cl, r_cp = evm(feat, None, label, is_crf=False)
    o_ll, o_lp, o_cl, o_cp = O.evaluator_forward(feat, label, evm.cluster_probe.clusters.detach(),
                                                 evm.linear_probe.weight.detach().view(C, D),
                                                 evm.linear_probe.bias.detach(), C)
    assert_same(o_lp, r_lp, "eval/linear_preds"); assert_same(o_cp, r_cp, "eval/cluster_preds")
    assert_same(o_ll, r_ll, "eval/linear_loss"); assert_same(o_cl, r_cl, "eval/cluster_loss")
    mets = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:     # compute() writes a CSV into the CWD (metric.py:100-108)
        os.chdir(tmp)
        try:
            for name, preds, hung in (("cluster", r_cp, True), ("linear", r_lp, False)):
                mt = metric.UnSegMetrics(C, 0, hung, torch.device("cpu"))
                mt.update(preds, label)
                conf = mt.confusion_matrix.clone()
                oc = O.confusion_update(torch.zeros(C, C, dtype=torch.long), preds, label, C)
                assert_same(oc, conf, f"metric/{name}/confusion")
                res = mt.compute(prefix="golden")
                ores = O.metrics_compute(oc, hung)
                assert_same(ores["iou"], res["iou"], f"metric/{name}/iou")
                assert_same(ores["accuracy"], res["accuracy"], f"metric/{name}/accuracy")
                mets[f"{name}_confusion"] = np_(conf)
                mets[f"{name}_iou"] = np.float64(float(res["iou"]))
                mets[f"{name}_accuracy"] = np.float64(float(res["accuracy"]))
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "eval_probe.npz"), feat=np_(feat), label=np_(label),
                        clusters=np_(evm.cluster_probe.clusters), lin_w=np_(evm.linear_probe.weight.view(C, D)),
                        lin_b=np_(evm.linear_probe.bias), linear_preds=np_(r_lp), cluster_preds=np_(r_cp),
                        linear_loss=np.float64(float(r_ll)), cluster_loss=np.float64(float(r_cl)), **mets)
    print("eval_probe: ok")

    # ------------------------------------------------------------------ kNN
    # data/precompute_knns.py needs hydra / pytorch_lightning (absent) to import, so its three kNN lines
    # (:313-315) are the oracle's restatement here; the fixture pins the einsum+topk result on this torch build.
    # oracle/make_golden_knn.py executes the reference's own statements (taken from its AST) and pins the oracle to them.
    torch.manual_seed(26)   # first seed whose top-9 similarity gaps all exceed 1e-5 (no fp32 near-ties)
    feats = torch.nn.functional.normalize(torch.randn(300, 48), dim=1)
    idx, vals = O.knn(feats, k=8)
    gaps = (vals[:, :-1] - vals[:, 1:]).min()
    sims = feats @ feats.t()
    kth = torch.topk(sims, 9)[0]
    assert float(gaps) > 1e-5 and float((kth[:, 7] - kth[:, 8]).min()) > 1e-5
    np.savez_compressed(os.path.join(OUT, "knn.npz"), feats=np_(feats), idx=np_(idx), vals=np_(vals))
    print("knn: ok")


if __name__ == "__main__":
    main()
