"""GPU parity of the PQ kernels (through the C-ABI) against the golden fixtures and the CPU oracle.

Bar (BASELINE.json north_star): code indices and counts bit-exact, except fp32 near-ties whose fp64
top-2 relative margin is below 1e-6 (audited); losses / EMA codebooks within 1e-5 relative."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import equss_oracle as O

pytestmark = pytest.mark.gpu

NEAR_TIE = 1e-6


def _ops():
    from equss_b200 import ops
    return ops


def _audit_indices(idx_gpu, idx_ref, zn, cbn):
    """idx_*: (n,) long on CPU.  Every disagreement must be an fp32 near-tie.  Returns #near-ties."""
    bad = (idx_gpu != idx_ref).nonzero().flatten()
    if bad.numel() == 0:
        return 0
    marg = O.top2_margin_fp64(zn[bad], cbn, idx_ref[bad], idx_gpu[bad])
    assert float(marg.max()) < NEAR_TIE, f"{bad.numel()} index mismatches, worst fp64 margin {float(marg.max()):.3e}"
    return int(bad.numel())


@pytest.mark.parametrize("mode", ["l2", "z_norm", "none"])
@pytest.mark.parametrize("algo", [1, 0])   # exact SIMT kernel, then AUTO (tcgen05 when supported)
def test_golden_ema_multi_step(golden_dir, mode, algo):
    """Replays the reference's 3 training steps + 1 eval step (model/quantizer.py EMA PQ)."""
    ops = _ops()
    g = np.load(os.path.join(golden_dir, f"pq_ema_{mode}.npz"))
    M, K = int(g["M"]), int(g["K"])
    dev = torch.device("cuda:0")
    weight = torch.from_numpy(g["weight0"]).to(dev).contiguous()
    weight_avg = weight.clone()
    vq_count = torch.zeros(M, K, device=dev)
    exact = torch.zeros(M, K, device=dev)
    d = weight.shape[2]
    for s in range(4):
        z = torch.from_numpy(g[f"z{s}"]).to(dev)
        # codebook-side normalisation is host plumbing (M*K*d elements), same torch ops as the reference
        cbn = torch.stack([O.normalize_pair(z[:1, :d], weight[m], mode)[1] for m in range(M)]).contiguous()
        idx = ops.pq_assign(z, cbn, normalize=mode, algo=algo)
        assert np.array_equal(idx.cpu().numpy(), g[f"idx{s}"]), f"step {s}: indices differ from the reference"
        out, sqerr, _ = ops.pq_gather_loss(z, cbn, idx, normalize=mode)
        np.testing.assert_allclose(out.cpu().numpy(), g[f"zq{s}"], rtol=1e-5, atol=1e-6)
        n = z.shape[0]
        commit = (sqerr / (n * d)).float().mean().item()
        assert commit == pytest.approx(float(g[f"out{s}/commitment-loss"]), rel=1e-5)
        if s < 3:
            packed = ops.pq_accumulate(z, idx, K)
            unused = ops.ema_update(packed, 0.99, 1e-5, vq_count, weight_avg, weight, exact)
            np.testing.assert_allclose(weight.cpu().numpy(), g[f"weight_after{s}"], rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(weight_avg.cpu().numpy(), g[f"weight_avg_after{s}"], rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(vq_count.cpu().numpy(), g[f"vq_count_after{s}"], rtol=1e-6, atol=1e-7)
            assert np.array_equal(exact.cpu().numpy(), g[f"exact_after{s}"])
            usage = float(((K - unused.float()) / K).mean())
            assert usage == pytest.approx(float(g[f"out{s}/codebook-usage"]), rel=1e-6)
    prob = ops.pq_distance_prob(z, cbn, normalize=mode)
    np.testing.assert_allclose(prob.cpu().numpy(), g["prob3"], rtol=2e-5, atol=1e-7)


def test_golden_nchw_variants(golden_dir):
    """Learned-codebook variants on NCHW input: V1 (quantizer.VectorQuantizer), V5 (dino_pqgo.Codebook)
    and the quantizer_v2 quirk that gathers rows of z_norm."""
    ops = _ops()
    dev = torch.device("cuda:0")
    g = np.load(os.path.join(golden_dir, "pq_param_nchw.npz"))
    z = torch.from_numpy(g["z"]).to(dev)
    B, d, h, w = z.shape
    cb = torch.from_numpy(g["v1_codebook"]).to(dev)
    cbn = F.normalize(cb, dim=1)[None].contiguous()
    idx = ops.pq_assign(z, cbn, normalize="l2")
    assert np.array_equal(idx[0].cpu().numpy(), g["v1_idx"])
    out, sqerr, _ = ops.pq_gather_loss(z, cbn, idx, normalize="l2")
    np.testing.assert_allclose(out.cpu().numpy(), g["v1_q"], rtol=1e-5, atol=1e-6)
    mse = float(sqerr[0] / (B * h * w * d))
    assert mse == pytest.approx(float(g["v1_commitment_loss"]), rel=1e-5)
    assert mse * 1.25 == pytest.approx(float(g["v1_loss"]), rel=1e-5)
    cb5 = torch.from_numpy(g["v5_codebook"]).to(dev)[None].contiguous()
    idx5 = ops.pq_assign(z, cb5, normalize="none")
    assert np.array_equal(idx5[0].view(B, h, w).cpu().numpy(), g["v5_idx"])
    out5, sq5, _ = ops.pq_gather_loss(z, cb5, idx5, normalize="none")
    np.testing.assert_allclose(out5.cpu().numpy(), g["v5_q"], rtol=1e-5, atol=1e-6)
    assert float(sq5[0] / (B * h * w * d)) * 1.25 == pytest.approx(float(g["v5_vq_loss"]), rel=1e-5)
    g2 = np.load(os.path.join(golden_dir, "pq_v2_nchw.npz"))
    emb = torch.from_numpy(g2["embeddings"]).to(dev)
    cbn2 = F.normalize(emb, dim=1)[None].contiguous()
    idx2 = ops.pq_assign(z, cbn2, normalize="l2")
    assert np.array_equal(idx2[0].cpu().numpy(), g2["idx"])


SHAPES = [
    # (n_or_(B,h,w), M, K, d, mode)
    (3136, 8, 256, 64, "l2"),          # BASELINE config 1 (pq_baseline, flat)
    (1000, 64, 256, 16, "l2"),         # config-2 subspace shape, ragged N
    (777, 16, 512, 64, "l2"),          # config-4 subspace shape, ragged N
    (513, 4, 32, 32, "z_norm"),
    (300, 3, 40, 12, "none"),          # d not a power of two, K not a multiple of 16: generic kernels
    (1, 2, 8, 8, "l2"),                # single pixel
    ((2, 28, 28), 8, 256, 64, "l2"),   # NCHW
    ((3, 9, 7), 64, 256, 16, "l2"),    # NCHW, tiny odd grid
    ((2, 5, 5), 2, 16, 128, "z_norm"),
    # scatter-add variants: 128-bit shared CAS kernels (d in 4..64, flat and NCHW), one 1024-thread block per SM when
    # the accumulators of a subspace exceed 56 KB (K = 600, d = 64: 163 KB), scalar fallback otherwise
    (257, 4, 16, 4, "l2"),
    (640, 3, 24, 8, "z_norm"),
    ((2, 6, 5), 4, 32, 32, "z_norm"),
    (300, 2, 600, 64, "l2"),
    ((1, 20, 20), 2, 600, 64, "none"),
]


@pytest.mark.parametrize("shape,M,K,d,mode", SHAPES)
@pytest.mark.parametrize("algo", [1, 0])
def test_assign_gather_accumulate_vs_oracle(shape, M, K, d, mode, algo):
    ops = _ops()
    dev = torch.device("cuda:0")
    torch.manual_seed(1234)
    D = M * d
    if isinstance(shape, tuple):
        B, h, w = shape
        z = torch.randn(B, D, h, w)
        z_flat = z.permute(0, 2, 3, 1).reshape(-1, D)
    else:
        z = torch.randn(shape, D)
        z_flat = z
    n = z_flat.shape[0]
    cb = torch.randn(M, K, d) * (0.3 if mode == "none" else 1.0)
    if mode == "none":
        z, z_flat = z * 0.3, z_flat * 0.3
    zg = z.to(dev)
    cbn_list, zn_list = [], []
    for m in range(M):
        zn, cn = O.normalize_pair(z_flat[:, m * d:(m + 1) * d], cb[m], mode)
        cbn_list.append(cn); zn_list.append(zn)
    cbn = torch.stack(cbn_list).contiguous()
    idx = ops.pq_assign(zg, cbn.to(dev), normalize=mode, algo=algo)
    out, sqerr, znorm = ops.pq_gather_loss(zg, cbn.to(dev), idx, normalize=mode, want_znorm=True)
    packed = ops.pq_accumulate(zg, idx, K)
    packed_n = ops.pq_accumulate(zg, idx, K, use_norm=True, normalize=mode)
    torch.cuda.synchronize()
    idx_c = idx.cpu().long()
    out_flat = out.cpu() if not isinstance(shape, tuple) else out.cpu().permute(0, 2, 3, 1).reshape(-1, D)
    zn_flat = znorm.cpu() if not isinstance(shape, tuple) else znorm.cpu().permute(0, 2, 3, 1).reshape(-1, D)
    ties = 0
    for m in range(M):
        dist = O.sq_distance(zn_list[m], cbn[m])
        ref = torch.argmin(dist, dim=1)
        ties += _audit_indices(idx_c[m], ref, zn_list[m], cbn[m])
        # downstream quantities are checked against the oracle evaluated at the GPU's own indices
        q = cbn[m][idx_c[m]]
        ste = zn_list[m] + (q - zn_list[m])
        torch.testing.assert_close(zn_flat[:, m * d:(m + 1) * d], zn_list[m], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(out_flat[:, m * d:(m + 1) * d], ste, rtol=1e-5, atol=1e-6)
        mse_ref = F.mse_loss(zn_list[m], q).item()
        assert float(sqerr[m] / (n * d)) == pytest.approx(mse_ref, rel=1e-5, abs=1e-9)
        onehot = F.one_hot(idx_c[m], K).float()
        assert torch.equal(packed[m, :, d].cpu(), onehot.sum(0))                      # counts: bit-exact
        torch.testing.assert_close(packed[m, :, :d].cpu(), onehot.t() @ z_flat[:, m * d:(m + 1) * d],
                                   rtol=1e-5, atol=2e-5)
        torch.testing.assert_close(packed_n[m, :, :d].cpu(), onehot.t() @ zn_list[m], rtol=1e-5, atol=2e-5)
    assert ties <= max(2, int(1e-4 * n * M)), f"too many near-ties: {ties}"


def test_empty_input_is_a_noop():
    ops = _ops()
    dev = torch.device("cuda:0")
    cbn = F.normalize(torch.randn(4, 32, 16), dim=2).to(dev)
    z = torch.zeros(0, 64, device=dev)
    idx = ops.pq_assign(z, cbn, normalize="l2")
    assert tuple(idx.shape) == (4, 0)
    out, sqerr, _ = ops.pq_gather_loss(z, cbn, idx, normalize="l2")
    assert out.numel() == 0 and float(sqerr.sum()) == 0.0
    packed = ops.pq_accumulate(z, idx, 32)
    assert float(packed.abs().sum()) == 0.0


def test_collisions_all_pixels_one_code():
    """Every pixel maps to the same code: the scatter-add is fully contended, counts must stay exact."""
    ops = _ops()
    dev = torch.device("cuda:0")
    M, K, d, n = 2, 64, 16, 20000
    cb = F.normalize(torch.randn(M, K, d), dim=2)
    z = cb[:, 5, :].reshape(1, M * d).repeat(n, 1) * 3.0
    idx = ops.pq_assign(z.to(dev), cb.to(dev), normalize="l2")
    assert int((idx != 5).sum()) == 0
    packed = ops.pq_accumulate(z.to(dev), idx, K)
    assert float(packed[0, 5, d]) == n and float(packed[:, :, d].sum()) == n * M
    torch.testing.assert_close(packed[0, 5, :d].cpu(), z[:, :d].sum(0), rtol=1e-4, atol=1e-3)


def test_first_index_wins_on_exact_ties():
    """Duplicate codewords: torch.argmin returns the first minimal index."""
    ops = _ops()
    dev = torch.device("cuda:0")
    cb = F.normalize(torch.randn(1, 32, 16), dim=2)
    cb[0, 20] = cb[0, 3]
    cb[0, 31] = cb[0, 3]
    z = cb[0, 3].reshape(1, 16).repeat(300, 1) * 2.0
    for algo in (1, 0):
        idx = ops.pq_assign(z.to(dev), cb.to(dev), normalize="l2", algo=algo)
        assert int((idx != 3).sum()) == 0


@pytest.mark.parametrize("mode", ["l2", "none", "z_norm"])
@pytest.mark.parametrize("nchw", [False, True])
def test_backward_matches_autograd(mode, nchw):
    """K3' against torch autograd on the oracle's formulation of the STE output + losses."""
    ops = _ops()
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    M, K, d, beta = 4, 32, 16, 0.25
    D = M * d
    z = (torch.randn(2, D, 6, 5) if nchw else torch.randn(60, D)).double().requires_grad_(True)
    cb = torch.randn(M, K, d).double()
    z_flat = z.permute(0, 2, 3, 1).reshape(-1, D) if nchw else z
    n = z_flat.shape[0]
    idx = ops.pq_assign(z.detach().float().to(dev),
                        torch.stack([O.normalize_pair(z_flat[:1, :d].float(), cb[m].float(), mode)[1] for m in range(M)]).to(dev),
                        normalize=mode)
    idx_c = idx.cpu().long()
    go = torch.randn_like(z_flat)
    total = 0.0
    cbn_all = []
    for m in range(M):
        zn, cn = O.normalize_pair(z_flat[:, m * d:(m + 1) * d], cb[m], mode)
        cbn_all.append(cn)
        q = cn[idx_c[m]].detach()
        ste = zn + (q - zn).detach()
        total = total + (ste * go[:, m * d:(m + 1) * d]).sum() + beta * F.mse_loss(zn, q)
    total.backward()
    go_dev = go.float()
    if nchw:
        go_dev = go_dev.view(2, 6, 5, D).permute(0, 3, 1, 2).contiguous()
    coef = torch.full((M,), 2.0 * beta / (n * d))
    gz, _ = ops.pq_gather_loss_bwd(z.detach().float().to(dev), torch.stack(cbn_all).float().to(dev), idx, mode,
                                   go_dev.to(dev), coef.to(dev))
    torch.testing.assert_close(gz.cpu().double(), z.grad, rtol=2e-4, atol=2e-5)


TC_SHAPES = [
    # (shape, M, K, d, mode)  -- BASELINE.json shapes; the oracle is too slow here, the exact SIMT kernel
    # (itself oracle-checked above) is the comparator: the two kernels must agree on EVERY index.
    (3136, 8, 256, 64, "l2"),                 # config 1
    (51200, 64, 256, 16, "l2"),               # config 2 (flat view)
    ((32, 40, 40), 64, 256, 16, "l2"),        # config 2, NCHW as the evaluator receives it
    (6272, 16, 512, 64, "l2"),                # config 4, one rank's shard
    ((2, 56, 56), 16, 512, 64, "l2"),         # config 4 NCHW
    (4099, 16, 1024, 32, "none"),             # cityscapes-yaml-like K=1024, ragged N
    (2500, 32, 32, 32, "z_norm"),             # cityscapes yaml M=32,K=32
    (1000, 4, 40, 8, "l2"),                   # K not a multiple of 16 -> padded columns
]


@pytest.mark.parametrize("shape,M,K,d,mode", TC_SHAPES)
@pytest.mark.parametrize("algo", [2, 3])     # 2: fp16-split kernel where it applies (l2, d>=16), 3: split-tf32 kernel
def test_tcgen05_assign_equals_exact_kernel(shape, M, K, d, mode, algo):
    ops = _ops()
    dev = torch.device("cuda:0")
    torch.manual_seed(99)
    D = M * d
    z = torch.randn(*((shape[0], D, shape[1], shape[2]) if isinstance(shape, tuple) else (shape, D)), device=dev)
    cb = torch.randn(M, K, d, device=dev)
    if mode == "l2":
        cbn = F.normalize(cb, dim=2)
    elif mode == "z_norm":
        s, mu = torch.std_mean(cb, dim=2, keepdim=True)
        cbn = (cb - mu) / (s + 1e-5)
    else:
        cbn = cb * 0.5
    idx_simt = ops.pq_assign(z, cbn, normalize=mode, algo=1)
    idx_tc = ops.pq_assign(z, cbn, normalize=mode, algo=algo)
    torch.cuda.synchronize()
    nbad = int((idx_simt != idx_tc).sum())
    assert nbad == 0, f"{nbad} of {idx_tc.numel()} indices differ between the tcgen05 and the exact kernel"
    assert int(idx_tc.min()) >= 0 and int(idx_tc.max()) < K


H_SHAPES = [
    # (shape, M, K, d, codebook): cases aimed at the fp16-split kernel (l2 rows)
    (5000, 63, 256, 16, "unit"),              # odd M: no subspace pairing (G = 1)
    (5000, 64, 40, 16, "unit"),               # padded columns (K < NC)
    (4099, 16, 1024, 32, "unit"),             # four code chunks, ragged N
    ((3, 28, 28), 8, 512, 64, "unit"),        # NCHW, hw not a multiple of the 128-pixel tile, two chunks
    (6000, 8, 256, 64, "scaled"),             # un-normalised codebook (|c| ~ 7): power-of-two operand scaling
    (6000, 16, 256, 32, "tiny"),              # |c| ~ 1e-3
    (3000, 8, 512, 16, "duplicates"),         # exact ties inside and across chunks: first index must win
    (3000, 8, 256, 16, "zero_rows"),          # all-zero pixels (normalised row = 0)
]


@pytest.mark.parametrize("shape,M,K,d,kind", H_SHAPES)
def test_f16_split_assign_edge_cases(shape, M, K, d, kind):
    ops = _ops()
    dev = torch.device("cuda:0")
    torch.manual_seed(7)
    D = M * d
    z = torch.randn(*((shape[0], D, shape[1], shape[2]) if isinstance(shape, tuple) else (shape, D)), device=dev)
    cb = torch.randn(M, K, d, device=dev)
    cbn = F.normalize(cb, dim=2)
    if kind == "scaled":
        cbn = cb * 1.7
    elif kind == "tiny":
        cbn = cb * 1e-3
    elif kind == "duplicates":
        cbn[:, 5] = cbn[:, 3]
        cbn[:, K // 2:] = cbn[:, :K // 2]          # every code appears twice, NC apart when K = 2*NC
    elif kind == "zero_rows":
        z[::7] = 0
    cbn = cbn.contiguous()
    idx_simt = ops.pq_assign(z, cbn, normalize="l2", algo=1)
    idx_h = ops.pq_assign(z, cbn, normalize="l2", algo=2)
    torch.cuda.synchronize()
    nbad = int((idx_simt != idx_h).sum())
    assert nbad == 0, f"{nbad} of {idx_h.numel()} indices differ between the fp16-split and the exact kernel"
    if kind == "duplicates":
        assert int(idx_h.max()) < K // 2


FUSED_SHAPES = [
    (51200, 64, 256, 16),              # config 2, flat, subspace pairs
    ((32, 40, 40), 64, 256, 16),       # config 2, NCHW
    (5000, 63, 256, 16),               # odd M (no pairing), ragged N
    ((3, 28, 28), 16, 200, 32),        # d = 32, NCHW, hw not a multiple of the tile, padded columns
    (777, 32, 32, 32),                 # small codebook (NC = 32)
    (100, 64, 256, 16),                # fewer tiles than the gather lag
]


@pytest.mark.parametrize("shape,M,K,d", FUSED_SHAPES)
def test_fused_assign_gather_equals_two_kernel_path(shape, M, K, d):
    """equss_pq_assign_gather (K1 + K3 in one pass) must reproduce K1's indices and K3's output bit for bit;
    the squared error differs only by fp32 summation order."""
    ops = _ops()
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    D = M * d
    z = torch.randn(*((shape[0], D, shape[1], shape[2]) if isinstance(shape, tuple) else (shape, D)), device=dev)
    cbn = F.normalize(torch.randn(M, K, d, device=dev), dim=2).contiguous()
    src = torch.randn(M, K, d, device=dev)          # gather from a different table (dino_new_vq.py:403)
    idx_f, out_f, sq_f = ops.pq_assign_gather(z, cbn, src, normalize="l2", fused=True)
    idx_r = ops.pq_assign(z, cbn, normalize="l2", algo=1)
    out_r, sq_r, _ = ops.pq_gather_loss(z, src, idx_r, normalize="l2")
    torch.cuda.synchronize()
    assert torch.equal(idx_f, idx_r)
    assert torch.equal(out_f, out_r)
    torch.testing.assert_close(sq_f, sq_r, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("M,K", [(64, 256), (3, 40), (1, 1024), (16, 512), (2, 2048), (2, 1500)])
def test_usage_percentiles_vs_oracle(M, K):
    """equss_usage_percentiles == get_histogram_count (model/quantizer.py:15-30) per subspace, NaN for None."""
    ops = _ops()
    dev = torch.device("cuda:0")
    torch.manual_seed(M * K)
    count = torch.randint(0, 50, (M, K)).float() * (torch.rand(M, K) < 0.6).float()
    count[0] = 0                                   # never reaches any level -> None / NaN
    if M > 1:
        count[1] = 0; count[1, K // 3] = 7          # a single used code reaches every level at rank 0
    packed = torch.zeros(M, K, 5)
    packed[:, :, 4] = count                        # strided view, as the count column of the K4 buffer
    got = ops.usage_percentiles(packed.to(dev)[:, :, 4]).cpu()
    for m in range(M):
        ref = O.histogram_percentiles(count[m], "x")
        for j, tag in enumerate(("x-p10", "x-p50", "x-p90")):
            if ref[tag] is None:
                assert torch.isnan(got[m, j])
            else:
                assert float(got[m, j]) == pytest.approx(float(ref[tag]), abs=1e-7), (m, tag)
