"""GPU parity of the expansion head (cluster1(x) + cluster2(x), model/dino_pqgo.py:104-112,127-128,
model/blocks/module.py:20-44) through the C-ABI, against the golden fixture of the reference module and the oracle.

Tolerance: the kernel is a split-tf32 tensor-core contraction (hi.hi + lo.hi + hi.lo, fp32 accumulate); its error is
bounded by 1e-5 of the output scale (max |out|) against the fp64 evaluation -- the bar BASELINE.json sets for
floating-point results.  (The reference's own CPU fp32 result sits at ~3e-7, its default GPU path -- cuDNN TF32 -- at
~1e-3.)"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import equss_oracle as O

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5


def _rel(a: torch.Tensor, ref64: torch.Tensor) -> float:
    return float((a.double() - ref64).abs().max() / ref64.abs().max())


def _head_from_fixture(g, dev):
    from equss_b200.head import SegmentationHead
    C, D = g["cluster2_0_weight"].shape[0], g["cluster1_0_weight"].shape[0]
    head = SegmentationHead(C, D)
    sd = {k: torch.from_numpy(g[k.replace(".", "_")]) for k in head.state_dict().keys()}
    head.load_state_dict(sd)          # the reference's own state_dict keys
    return head.to(dev).eval()


def test_golden_segmentation_head(golden_dir):
    g = np.load(os.path.join(golden_dir, "expansion_head.npz"))
    dev = torch.device("cuda:0")
    head = _head_from_fixture(g, dev)
    x = torch.from_numpy(g["x"]).to(dev)
    with torch.no_grad():
        out = head(x)
    assert out.shape == g["out"].shape
    assert out.permute(0, 2, 3, 1).is_contiguous()            # flat (pixel, channel) rows for the PQ kernels
    ref64 = torch.from_numpy(g["out_fp64"])
    assert _rel(out.cpu(), ref64) < REL_TOL
    np.testing.assert_allclose(out.cpu().numpy(), g["out"], rtol=0, atol=REL_TOL * float(np.abs(g["out"]).max()))


SHAPES = [
    # (B, C, h, w, D)
    (2, 384, 40, 40, 1024),      # ViT-S features at the cocostuff27 eval grid, config-2 expanded dim
    (1, 768, 56, 56, 1024),      # ViT-B features at the cityscapes grid (config 4)
    (4, 384, 28, 28, 512),       # config 1 (28 x 28 = 784 tokens, not a multiple of 32: one 3-D TMA box per 32-pixel block)
    (2, 48, 6, 6, 64),           # 36 tokens: ragged block zero-filled by the TMA unit, C = 3 stages of 16
    (3, 64, 8, 4, 72),           # one 32-pixel block per image, D not a multiple of 4 x 32
    (1, 32, 5, 7, 33),           # 35 tokens (not a multiple of 4): NHWC copy + flat-row path, scalar stores
    (2, 96, 16, 18, 300),        # 288 tokens: partial 128-pixel tile inside every image, partial column tile
]


@pytest.mark.parametrize("B,C,h,w,D", SHAPES)
def test_expansion_head_vs_oracle(B, C, h, w, D):
    from equss_b200.head import SegmentationHead
    dev = torch.device("cuda:0")
    torch.manual_seed(B * 1000 + C + D)
    head = SegmentationHead(C, D).eval()
    with torch.no_grad():
        for p in head.parameters():
            p.mul_(3.0)                                   # default init is small; make the ReLU and biases matter
    x = torch.randn(B, C, h, w)
    p = {k: v.detach() for k, v in head.state_dict().items()}
    args = [p["cluster1.0.weight"], p["cluster1.0.bias"], p["cluster2.0.weight"], p["cluster2.0.bias"],
            p["cluster2.2.weight"], p["cluster2.2.bias"]]
    ref64 = O.expansion_head(x.double(), *[a.double() for a in args])
    head = head.to(dev)
    with torch.no_grad():
        out = head(x.to(dev))
    torch.cuda.synchronize()
    assert out.shape == (B, D, h, w)
    assert _rel(out.cpu(), ref64) < REL_TOL
    # the hidden layer alone (ReLU epilogue) and a flat-row input give the same numbers
    from equss_b200 import ops
    hid = ops.head_gemm(x.to(dev), head.cluster2[0].weight, head.cluster2[0].bias, relu=True)
    hid64 = F.relu(F.conv2d(x.double(), args[2].double(), args[3].double())).permute(0, 2, 3, 1).reshape(-1, C)
    assert _rel(hid.cpu(), hid64) < REL_TOL
    assert float(hid.min()) >= 0.0
    x_flat = x.permute(0, 2, 3, 1).reshape(-1, C).contiguous().to(dev)
    hid_flat = ops.head_gemm(x_flat, head.cluster2[0].weight, head.cluster2[0].bias, relu=True)
    assert _rel(hid_flat.cpu(), hid64) < REL_TOL


def test_head_output_feeds_pq_without_copy():
    """The (B, D, h, w) view of NHWC memory the head returns is consumed by the PQ ops as flat rows: same indices,
    same quantised values as the NCHW-contiguous copy of the same tensor."""
    from equss_b200 import ops
    from equss_b200.head import SegmentationHead
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    B, C, h, w, D, M, K = 2, 64, 8, 8, 128, 8, 32
    head = SegmentationHead(C, D).to(dev).eval()
    with torch.no_grad():
        code = head(torch.randn(B, C, h, w, device=dev))
    assert not code.is_contiguous()
    cbn = F.normalize(torch.randn(M, K, D // M, device=dev), dim=2).contiguous()
    idx_a, out_a, sq_a = ops.pq_assign_gather(code, cbn, None, None, "l2")
    idx_b, out_b, sq_b = ops.pq_assign_gather(code.contiguous(), cbn, None, None, "l2")
    assert torch.equal(idx_a, idx_b)
    assert out_a.shape == code.shape and out_a.stride() == code.stride()
    assert torch.equal(out_a.contiguous(), out_b)
    torch.testing.assert_close(sq_a, sq_b, rtol=1e-6, atol=0)
    packed_a = ops.pq_accumulate(code, idx_a, K)
    packed_b = ops.pq_accumulate(code.contiguous(), idx_b, K)
    torch.testing.assert_close(packed_a, packed_b, rtol=1e-5, atol=1e-6)


def test_head_training_path_uses_autograd():
    from equss_b200.head import SegmentationHead
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    head = SegmentationHead(32, 48).to(dev)
    x = torch.randn(2, 32, 4, 8, device=dev)
    out = head(x)                                  # parameters require grad: PyTorch convolutions, differentiable
    assert out.requires_grad
    out.sum().backward()
    assert head.cluster2[0].weight.grad is not None
    with torch.no_grad():
        fused = head(x)
    ref64 = O.expansion_head(x.double().cpu(), *[p.detach().double().cpu() for p in (
        head.cluster1[0].weight, head.cluster1[0].bias, head.cluster2[0].weight, head.cluster2[0].bias,
        head.cluster2[2].weight, head.cluster2[2].bias)])
    assert _rel(fused.cpu(), ref64) < REL_TOL


def test_head_rejects_bad_arguments():
    from equss_b200 import ops
    from equss_b200._native import EqussNativeError
    dev = torch.device("cuda:0")
    with pytest.raises(EqussNativeError):
        ops.head_gemm(torch.randn(4, 32), torch.randn(8, 32, device=dev))            # CPU tensor: no fallback
    with pytest.raises(ValueError):
        ops.head_gemm(torch.randn(4, 32, device=dev), torch.randn(8, 64, device=dev))  # channel mismatch
    with pytest.raises(EqussNativeError):
        ops.head_gemm(torch.randn(4, 40, device=dev), torch.randn(8, 40, device=dev))  # C % 32 != 0
