"""Functional host API: one thin function per C entry point (include/equss_b200.h).

Tensors in, tensors out; allocation and stream selection are PyTorch's, the arithmetic is the
sm_100a kernels'.  These functions do not record autograd history; the nn.Module mirrors in
``quantizer.py`` wrap them in ``torch.autograd.Function`` where the reference needs gradients.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _native as N

__all__ = [
    "pq_cnorm2", "pq_assign", "pq_assign_gather", "pq_gather_loss", "pq_gather_loss_bwd", "pq_accumulate", "ema_update",
    "pq_distance_prob", "pq_soft_stats", "channel_moments", "usage_percentiles", "pq_train_tail", "pq_prepare_codebook",
    "TAIL_KEYS", "probe_pack", "probe_logits", "probe_argmax_confusion", "probe_losses", "token_gram",
    "confusion_update", "knn_topk", "stego_feature_corr",
    "launch_count",
]


def launch_count() -> int:
    """Kernels launched by the library since load (bench.py's ``gpu_launches``)."""
    return int(N.lib().equss_launch_count())


def _norm_args(normalize, norm_a, norm_b, D, dev):
    if normalize not in N.NORM_MODES:
        raise ValueError(f"Unsupported normalize type {normalize}")
    mode = N.NORM_MODES[normalize]
    if mode == 3:
        if norm_a is None or norm_b is None:
            raise ValueError("z_trainable normalisation needs per-channel mean and denominator vectors")
        norm_a = N.f32c(norm_a.detach()).reshape(-1)
        norm_b = N.f32c(norm_b.detach()).reshape(-1)
        if norm_a.numel() != D or norm_b.numel() != D:
            raise ValueError(f"normalisation vectors must have {D} elements")
    else:
        norm_a = norm_b = None
    return mode, norm_a, norm_b


def _prep_z(z: torch.Tensor, M: int):
    dev = N.require_cuda(z)
    N.ensure_device(dev)
    z = N.f32_dense(z.detach())
    zd, d, layout = N.zdesc_for(z, M)
    return z, zd, d, dev


def pq_cnorm2(codebook_norm: torch.Tensor) -> torch.Tensor:
    """sum(c**2, dim=-1) for a stacked codebook [M, K, d]  (model/quantizer.py:459)."""
    dev = N.require_cuda(codebook_norm)
    N.ensure_device(dev)
    cb = N.f32c(codebook_norm.detach())
    M, K, d = cb.shape
    out = torch.empty((M, K), dtype=torch.float32, device=dev)
    N.check(N.lib().equss_pq_cnorm2(cb.data_ptr(), M, K, d, out.data_ptr(), N.stream_ptr(dev)), "equss_pq_cnorm2")
    return out


def pq_assign(z: torch.Tensor, codebook_norm: torch.Tensor, cnorm2: Optional[torch.Tensor] = None,
              normalize: Optional[str] = "l2", norm_a=None, norm_b=None, algo: int = N.ASSIGN_AUTO,
              return_margin: bool = False):
    """Nearest-codeword indices for every (pixel, subspace): int32 [M, N]  (model/quantizer.py:457-467)."""
    cb = N.f32c(codebook_norm.detach())
    M, K, d = cb.shape
    z, zd, dz, dev = _prep_z(z, M)
    N.require_cuda(z, cb)
    if dz != d:
        raise ValueError(f"codebook sub-dim {d} does not match activation sub-dim {dz}")
    if cnorm2 is None:
        cnorm2 = pq_cnorm2(cb)
    cnorm2 = N.f32c(cnorm2)
    mode, na, nb = _norm_args(normalize, norm_a, norm_b, M * d, dev)
    n = zd.n_pixels
    idx = torch.empty((M, n), dtype=torch.int32, device=dev)
    margin = torch.empty((M, n), dtype=torch.float32, device=dev) if return_margin else None
    L = N.lib()
    wsb = int(L.equss_pq_assign_workspace_bytes(n, M, K, d, algo))
    ws = torch.empty((max(wsb, 1),), dtype=torch.uint8, device=dev) if wsb > 0 else None
    rc = L.equss_pq_assign(z.data_ptr(), zd, cb.data_ptr(), cnorm2.data_ptr(), M, K, d, mode, N.ptr(na), N.ptr(nb),
                           idx.data_ptr(), N.ptr(margin), N.ptr(ws), wsb, algo, N.stream_ptr(dev))
    N.check(rc, "equss_pq_assign")
    return (idx, margin) if return_margin else idx


def pq_assign_gather(z: torch.Tensor, codebook_norm: torch.Tensor, gather_src: Optional[torch.Tensor] = None,
                     cnorm2: Optional[torch.Tensor] = None, normalize: Optional[str] = "l2", norm_a=None, norm_b=None,
                     fused: Optional[bool] = None):
    """Nearest-codeword indices AND the gathered / straight-through output in one call:
    (idx int32 [M, N], out like z, sqerr float64 [M]).  When the shape allows it (l2 rows, d in {16, 32}, K <= 256)
    one fused tcgen05 kernel reads the activations once; otherwise :func:`pq_assign` + :func:`pq_gather_loss` run
    back to back.  ``fused=False`` forces the two-kernel path, ``fused=True`` raises if fusion is unavailable."""
    cb = N.f32c(codebook_norm.detach())
    src = cb if gather_src is None else N.f32c(gather_src.detach())
    M, K, d = cb.shape
    z, zd, dz, dev = _prep_z(z, M)
    if dz != d or tuple(src.shape) != (M, K, d):
        raise ValueError("codebook / gather source shape does not match the activation sub-dim")
    mode, na, nb = _norm_args(normalize, norm_a, norm_b, M * d, dev)
    L = N.lib()
    can = bool(L.equss_pq_assign_gather_supported(zd, M, K, d, mode))
    if fused is True and not can:
        raise ValueError("pq_assign_gather: fused kernel not available for this shape")
    if not can or fused is False:
        idx = pq_assign(z, cb, cnorm2, normalize, norm_a, norm_b)
        out, sqerr, _ = pq_gather_loss(z, src, idx, normalize, norm_a, norm_b)
        return idx, out, sqerr
    if cnorm2 is None:
        cnorm2 = pq_cnorm2(cb)
    cnorm2 = N.f32c(cnorm2)
    n = zd.n_pixels
    idx = torch.empty((M, n), dtype=torch.int32, device=dev)
    out = torch.empty_like(z)
    sqerr = torch.empty((M,), dtype=torch.float64, device=dev)      # zeroed by the call (with the operand images)
    wsb = int(L.equss_pq_assign_workspace_bytes(n, M, K, d, N.ASSIGN_TCGEN05))
    ws = torch.empty((max(wsb, 1),), dtype=torch.uint8, device=dev)
    rc = L.equss_pq_assign_gather(z.data_ptr(), zd, cb.data_ptr(), cnorm2.data_ptr(), src.data_ptr(), M, K, d, mode,
                                  idx.data_ptr(), out.data_ptr(), sqerr.data_ptr(), ws.data_ptr(), wsb, N.stream_ptr(dev))
    N.check(rc, "equss_pq_assign_gather")
    return idx, out, sqerr


def pq_gather_loss(z: torch.Tensor, gather_src: torch.Tensor, idx: torch.Tensor, normalize: Optional[str] = "l2",
                   norm_a=None, norm_b=None, want_znorm: bool = False
                   ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """Gather + squared error + straight-through value (model/quantizer.py:474,514,534-536).

    Returns (out, sqerr, z_norm):  out has z's shape/layout, sqerr is float64 [M] with
    sum((z_norm - q)**2) per subspace, z_norm is returned only when ``want_znorm``."""
    src = N.f32c(gather_src.detach())
    M, K, d = src.shape
    z, zd, dz, dev = _prep_z(z, M)
    if dz != d:
        raise ValueError(f"gather source sub-dim {d} does not match activation sub-dim {dz}")
    mode, na, nb = _norm_args(normalize, norm_a, norm_b, M * d, dev)
    idx = idx.contiguous()
    assert idx.dtype == torch.int32 and tuple(idx.shape) == (M, zd.n_pixels)
    out = torch.empty_like(z)
    zn = torch.empty_like(z) if want_znorm else None
    sqerr = torch.zeros((M,), dtype=torch.float64, device=dev)
    rc = N.lib().equss_pq_gather_loss(z.data_ptr(), zd, src.data_ptr(), idx.data_ptr(), M, K, d, mode, N.ptr(na),
                                      N.ptr(nb), out.data_ptr(), N.ptr(zn), sqerr.data_ptr(), N.stream_ptr(dev))
    N.check(rc, "equss_pq_gather_loss")
    return out, sqerr, zn


def pq_gather_loss_bwd(z: torch.Tensor, gather_src: torch.Tensor, idx: torch.Tensor, normalize: Optional[str],
                       grad_out: Optional[torch.Tensor], coef: Optional[torch.Tensor],
                       norm_a=None, norm_b=None, want_grad_z: bool = True,
                       cb_coef: Optional[torch.Tensor] = None) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Backward of :func:`pq_gather_loss`: returns (grad_z, grad_gather_src)."""
    src = N.f32c(gather_src.detach())
    M, K, d = src.shape
    z, zd, dz, dev = _prep_z(z, M)
    mode, na, nb = _norm_args(normalize, norm_a, norm_b, M * d, dev)
    go = N.f32_dense(grad_out.detach(), like=z) if grad_out is not None else None
    if go is not None:
        assert go.shape == z.shape and go.stride() == z.stride()
    cf = N.f32c(coef.detach()).reshape(-1) if coef is not None else None
    cbf = N.f32c(cb_coef.detach()).reshape(-1) if cb_coef is not None else None
    gz = torch.empty_like(z) if want_grad_z else None
    gcb = torch.zeros_like(src) if cbf is not None else None
    rc = N.lib().equss_pq_gather_loss_bwd(z.data_ptr(), zd, src.data_ptr(), idx.contiguous().data_ptr(), M, K, d, mode,
                                          N.ptr(na), N.ptr(nb), N.ptr(go), N.ptr(cf), N.ptr(gz), N.ptr(cbf),
                                          N.ptr(gcb), N.stream_ptr(dev))
    N.check(rc, "equss_pq_gather_loss_bwd")
    return gz, gcb


def pq_accumulate(z: torch.Tensor, idx: torch.Tensor, num_codebook: int, use_norm: bool = False,
                  normalize: Optional[str] = "l2", norm_a=None, norm_b=None,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-code sums and counts packed as [M, K, d+1] (column d = count)  (model/quantizer.py:485-488)."""
    M = idx.shape[0]
    z, zd, d, dev = _prep_z(z, M)
    mode, na, nb = _norm_args(normalize, norm_a, norm_b, M * d, dev)
    K = int(num_codebook)
    if out is None:
        out = torch.zeros((M, K, d + 1), dtype=torch.float32, device=dev)
    else:
        assert out.is_contiguous() and tuple(out.shape) == (M, K, d + 1) and out.dtype == torch.float32
    rc = N.lib().equss_pq_accumulate(z.data_ptr(), zd, idx.contiguous().data_ptr(), M, K, d, int(bool(use_norm)), mode,
                                     N.ptr(na), N.ptr(nb), out.data_ptr(), N.stream_ptr(dev))
    N.check(rc, "equss_pq_accumulate")
    return out


def ema_update(packed: torch.Tensor, decay: float, eps: float, vq_count: torch.Tensor, weight_avg: torch.Tensor,
               weight: torch.Tensor, exact_count: Optional[torch.Tensor] = None) -> torch.Tensor:
    """In-place EMA codebook update on stacked state [M,K](,d)  (model/quantizer.py:233-254).
    Returns int32 [M]: number of codes that received no pixel this step (:509)."""
    dev = N.require_cuda(packed, vq_count, weight_avg, weight, exact_count)
    M, K, d1 = packed.shape
    d = d1 - 1
    for t, shp in ((vq_count, (M, K)), (weight_avg, (M, K, d)), (weight, (M, K, d))):
        assert t.is_contiguous() and t.dtype == torch.float32 and tuple(t.shape) == shp, (t.shape, shp)
    if exact_count is not None:
        assert exact_count.is_contiguous() and tuple(exact_count.shape) == (M, K)
    unused = torch.empty((M,), dtype=torch.int32, device=dev)
    rc = N.lib().equss_ema_update(packed.contiguous().data_ptr(), M, K, d, float(decay), float(eps), vq_count.data_ptr(),
                                  weight_avg.data_ptr(), weight.data_ptr(), N.ptr(exact_count), unused.data_ptr(),
                                  N.stream_ptr(dev))
    N.check(rc, "equss_ema_update")
    return unused


TAIL_KEYS = ("total-p10", "total-p50", "total-p90", "current-p10", "current-p50", "current-p90", "codebook-usage",
             "codebook-sum", "commitment-loss", "loss")
_tail_scratch = {}


def pq_train_tail(packed: torch.Tensor, decay: float, eps: float, vq_count: torch.Tensor, weight_avg: torch.Tensor,
                  weight: torch.Tensor, exact_count: torch.Tensor, sqerr: Optional[torch.Tensor], n_pixels: int,
                  beta: float, peers: Optional[Tuple[int, int]] = None,
                  zero_next: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """The tail of the EMA training step in one launch: in-place EMA update of the stacked state (as :func:`ema_update`)
    plus the ten scalar outputs of ``ProductQuantizerWrapper.forward`` (order: :data:`TAIL_KEYS`), averaged over the
    subspaces (model/quantizer.py:493-532,607-608).  Returns float32 [10], or None when K > 1024 (then call
    :func:`ema_update` / :func:`usage_percentiles`)."""
    dev = N.require_cuda(packed, vq_count, weight_avg, weight, exact_count, sqerr)
    M, K, d1 = packed.shape
    d = d1 - 1
    if K > 1024:
        return None
    for t, shp in ((vq_count, (M, K)), (weight_avg, (M, K, d)), (weight, (M, K, d)), (exact_count, (M, K))):
        assert t.is_contiguous() and t.dtype == torch.float32 and tuple(t.shape) == shp, (t.shape, shp)
    L = N.lib()
    key = (dev, M)
    scratch = _tail_scratch.get(key)
    if scratch is None:     # persistent per (device, M): the kernel's arrival counter lives in it and resets itself
        scratch = _tail_scratch[key] = torch.zeros((int(L.equss_pq_train_tail_scratch_floats(M)),), dtype=torch.float32, device=dev)
    stats = torch.empty((10,), dtype=torch.float32, device=dev)
    if sqerr is not None:
        assert sqerr.dtype == torch.float64 and sqerr.numel() == M
    if peers is not None:
        # `peers` = (device pointer to the array of every rank's packed-buffer pointer, world size): the cross-rank sum
        # happens inside the kernel over NVLink and the reduced statistics are written into `packed`
        rc = L.equss_pq_train_tail_peers(int(peers[0]), int(peers[1]), packed.data_ptr(), M, K, d, float(decay), float(eps),
                                         vq_count.data_ptr(), weight_avg.data_ptr(), weight.data_ptr(), exact_count.data_ptr(),
                                         N.ptr(sqerr), int(n_pixels), float(beta), scratch.data_ptr(), stats.data_ptr(),
                                         N.ptr(zero_next), N.stream_ptr(dev))
        N.check(rc, "equss_pq_train_tail_peers")
        return stats
    rc = L.equss_pq_train_tail(packed.contiguous().data_ptr(), M, K, d, float(decay), float(eps), vq_count.data_ptr(),
                               weight_avg.data_ptr(), weight.data_ptr(), exact_count.data_ptr(), N.ptr(sqerr), int(n_pixels),
                               float(beta), scratch.data_ptr(), stats.data_ptr(), N.stream_ptr(dev))
    N.check(rc, "equss_pq_train_tail")
    return stats


def pq_prepare_codebook(codebook: torch.Tensor, normalize: Optional[str]) -> Tuple[torch.Tensor, torch.Tensor]:
    """(codebook_norm [M, K, d], cnorm2 [M, K]) for the per-code normalisation modes "l2", "z_norm", "none"
    (model/quantizer.py:421,426,459) in one launch.  No autograd history (EMA codebooks are buffers)."""
    dev = N.require_cuda(codebook)
    N.ensure_device(dev)
    if normalize not in ("l2", "z_norm", "none"):
        raise ValueError(f"pq_prepare_codebook: per-code modes only, got {normalize}")
    cb = N.f32c(codebook.detach())
    M, K, d = cb.shape
    cbn = torch.empty_like(cb)
    cn2 = torch.empty((M, K), dtype=torch.float32, device=dev)
    rc = N.lib().equss_pq_prepare_codebook(cb.data_ptr(), M, K, d, N.NORM_MODES[normalize], cbn.data_ptr(), cn2.data_ptr(),
                                           N.stream_ptr(dev))
    N.check(rc, "equss_pq_prepare_codebook")
    return cbn, cn2


def usage_percentiles(count: torch.Tensor) -> torch.Tensor:
    """get_histogram_count (model/quantizer.py:15-30) for counts [M, K] (any strides): float32 [M, 3] with the
    p10 / p50 / p90 ranks divided by K, NaN where the reference returns None.  One launch, no host sync."""
    dev = N.require_cuda(count)
    N.ensure_device(dev)
    if count.dtype != torch.float32:
        count = count.float()
    M, K = count.shape
    out = torch.empty((M, 3), dtype=torch.float32, device=dev)
    rc = N.lib().equss_usage_percentiles(count.data_ptr(), count.stride(0), count.stride(1), M, K, out.data_ptr(),
                                         N.stream_ptr(dev))
    N.check(rc, "equss_usage_percentiles")
    return out


def pq_distance_prob(z: torch.Tensor, codebook_norm: torch.Tensor, cnorm2: Optional[torch.Tensor] = None,
                     normalize: Optional[str] = "l2", norm_a=None, norm_b=None,
                     temperature: float = 1.0) -> torch.Tensor:
    """softmax(-distance / temperature) for all subspaces, [N, M*K]  (model/quantizer.py:468,609)."""
    cb = N.f32c(codebook_norm.detach())
    M, K, d = cb.shape
    z, zd, dz, dev = _prep_z(z, M)
    if cnorm2 is None:
        cnorm2 = pq_cnorm2(cb)
    mode, na, nb = _norm_args(normalize, norm_a, norm_b, M * d, dev)
    prob = torch.empty((zd.n_pixels, M * K), dtype=torch.float32, device=dev)
    rc = N.lib().equss_pq_distance_prob(z.data_ptr(), zd, cb.data_ptr(), N.f32c(cnorm2).data_ptr(), M, K, d, mode,
                                        N.ptr(na), N.ptr(nb), float(temperature), prob.data_ptr(), N.stream_ptr(dev))
    N.check(rc, "equss_pq_distance_prob")
    return prob


def pq_soft_stats(z: torch.Tensor, codebook_norm: torch.Tensor, cnorm2: Optional[torch.Tensor] = None,
                  normalize: Optional[str] = "l2", norm_a=None, norm_b=None, temperature: float = 1.0
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(jsd, entropy) of model/dino_new_vq.py:447-450 -- JSDLoss between the soft assignments of the two batch halves
    and EntropyLoss of the first half's mean assignment (model/loss.py:490-525), each averaged over the subspaces --
    computed by one fused kernel that never writes the N x (K*M) probabilities.  No autograd history."""
    cb = N.f32c(codebook_norm.detach())
    M, K, d = cb.shape
    z, zd, dz, dev = _prep_z(z, M)
    if dz != d:
        raise ValueError(f"codebook sub-dim {d} does not match activation sub-dim {dz}")
    if cnorm2 is None:
        cnorm2 = pq_cnorm2(cb)
    mode, na, nb = _norm_args(normalize, norm_a, norm_b, M * d, dev)
    n = zd.n_pixels
    if n % 2 != 0:
        raise ValueError("JSD needs an even number of rows: the batch holds two views (model/dino_new_vq.py:447)")
    L = N.lib()
    if not L.equss_pq_soft_stats_supported(K, d):
        from ._pq_core import soft_assignment_stats
        return soft_assignment_stats(pq_distance_prob(z, cb, cnorm2, normalize, norm_a, norm_b, temperature), M, K)
    acc = torch.zeros((M * (K + 1),), dtype=torch.float64, device=dev)
    kl, ps = acc[:M], acc[M:].view(M, K)
    rc = L.equss_pq_soft_stats(z.data_ptr(), zd, cb.data_ptr(), N.f32c(cnorm2).data_ptr(), M, K, d, mode, N.ptr(na),
                               N.ptr(nb), float(temperature), kl.data_ptr(), ps.data_ptr(), N.stream_ptr(dev))
    N.check(rc, "equss_pq_soft_stats")
    half = n // 2
    jsd = (0.5 * kl / half).mean().float()
    avg = (ps / half).float()
    ent = (avg * torch.log(avg + 1e-8)).sum(dim=-1).mean()
    return jsd, ent


def channel_moments(z: torch.Tensor) -> torch.Tensor:
    """Per-channel mean and mean of squares of the activations, float32 [2, D], in one pass over z (flat (n, D) or
    NCHW (B, D, h, w))  (model/quantizer.py:433-434)."""
    dev = N.require_cuda(z)
    N.ensure_device(dev)
    z = N.f32_dense(z.detach())
    zd, _, _ = N.zdesc_for(z, 1)
    D = zd.dim
    sums = torch.zeros((2, D), dtype=torch.float64, device=dev)
    N.check(N.lib().equss_channel_moments(z.data_ptr(), zd, sums.data_ptr(), N.stream_ptr(dev)), "equss_channel_moments")
    return (sums / max(zd.n_pixels, 1)).float()


def probe_pack(wmat: torch.Tensor):
    """Pack probe weights [C_total, D] for :func:`probe_logits`: the K-major, zero-padded [D, C_pad] matrix and,
    when the shape allows it, the tensor-core operand image.  Do this once per weight update."""
    dev = N.require_cuda(wmat)
    N.ensure_device(dev)
    wmat = N.f32c(wmat.detach())
    Ct, D = wmat.shape
    L = N.lib()
    cpad = int(L.equss_probe_cpad(Ct))
    wmat_t = torch.zeros((D, cpad), dtype=torch.float32, device=dev)
    wmat_t[:, :Ct] = wmat.t()
    image = None
    nbytes = int(L.equss_probe_image_bytes(D, Ct))
    if nbytes > 0:
        image = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        N.check(L.equss_probe_build_image(wmat_t.data_ptr(), D, Ct, image.data_ptr(), N.stream_ptr(dev)),
                "equss_probe_build_image")
    return wmat_t, Ct, image


def probe_logits(feat: torch.Tensor, wmat, bias: Optional[torch.Tensor] = None, algo: int = 0) -> torch.Tensor:
    """Token-resolution probe logits [B*h*w, C_pad] from NCHW features (model/evaluator.py:67,98-100).
    ``wmat`` is either the [C_total, D] weight matrix or the result of :func:`probe_pack`.
    ``algo``: 0 = tcgen05 kernel when the shape is supported, 1 = CUDA-core kernel."""
    wmat_t, Ct, image = wmat if isinstance(wmat, tuple) else probe_pack(wmat)
    dev = N.require_cuda(feat, wmat_t, bias)
    N.ensure_device(dev)
    feat = N.f32c(feat.detach())
    B, D, h, w = feat.shape
    if wmat_t.shape[0] != D:
        raise ValueError(f"probe weight has {wmat_t.shape[0]} input channels, features have {D}")
    b = N.f32c(bias.detach()).reshape(-1) if bias is not None else None
    cpad = wmat_t.shape[1]
    logits = torch.empty((B * h * w, cpad), dtype=torch.float32, device=dev)
    L = N.lib()
    if algo == 0 and image is not None and L.equss_probe_logits_tc_supported(D, h, w, Ct):
        rc = L.equss_probe_logits_tc(feat.data_ptr(), B, D, h, w, image.data_ptr(), N.ptr(b), Ct, logits.data_ptr(),
                                     N.stream_ptr(dev))
        N.check(rc, "equss_probe_logits_tc")
        return logits
    rc = L.equss_probe_logits(feat.data_ptr(), B, D, h, w, wmat_t.data_ptr(), N.ptr(b), Ct, logits.data_ptr(),
                              N.stream_ptr(dev))
    N.check(rc, "equss_probe_logits")
    return logits


def probe_argmax_confusion(logits: torch.Tensor, B: int, h: int, w: int, c_total: int, label: torch.Tensor,
                           num_classes: int, heads: Sequence[Tuple[int, int]],
                           want_preds: bool = True,
                           confusions: Optional[Sequence[Optional[torch.Tensor]]] = None
                           ) -> List[Optional[torch.Tensor]]:
    """Bilinear interpolation of token logits + per-head argmax at label resolution, fused with the
    confusion histogram (model/evaluator.py:53-54,68-70 + model/metric.py:44-58).

    heads: [(first_channel, n_channels), ...];  confusions[i]: int64 [rows_i, num_classes] or None
    (accumulated in place).  Returns the per-head int64 (B, H, W) predictions (None if not wanted)."""
    dev = N.require_cuda(logits, label)
    label = label.contiguous()
    assert label.dtype == torch.int64 and label.dim() == 3 and label.shape[0] == B
    H, W = int(label.shape[1]), int(label.shape[2])
    nh = len(heads)
    preds = [torch.empty((B, H, W), dtype=torch.int64, device=dev) if want_preds else None for _ in range(nh)]
    confs = list(confusions) if confusions is not None else [None] * nh
    rows = []
    for c in confs:
        if c is not None:
            assert c.is_cuda and c.dtype == torch.int64 and c.is_contiguous() and c.shape[1] == num_classes
            rows.append(int(c.shape[0]))
        else:
            rows.append(0)
    rc = N.lib().equss_probe_argmax_confusion(
        logits.data_ptr(), B, h, w, c_total, label.data_ptr(), H, W, num_classes, nh,
        N.as_i32_array([o for o, _ in heads]), N.as_i32_array([c for _, c in heads]),
        N.as_voidp_array([N.ptr(p) for p in preds]), N.as_voidp_array([N.ptr(c) for c in confs]),
        N.as_i32_array(rows), N.stream_ptr(dev))
    N.check(rc, "equss_probe_argmax_confusion")
    return preds


def token_gram(feat: torch.Tensor) -> torch.Tensor:
    """2x2 Gram terms of the token grid, float32 [B*h*w, 5]: <x,x>, <x,right>, <x,down>, <x,down-right>, <x,down-left>
    over the channels of feat (B, D, h, w).  The norm of the bilinearly upsampled feature vector at any label pixel
    (model/evaluator.py:96 F.normalize of the upsampled map) follows from these without forming the map."""
    dev = N.require_cuda(feat)
    N.ensure_device(dev)
    feat = N.f32c(feat.detach())
    B, D, h, w = feat.shape
    gram = torch.empty((B * h * w, 5), dtype=torch.float32, device=dev)
    N.check(N.lib().equss_token_gram(feat.data_ptr(), B, D, h, w, gram.data_ptr(), N.stream_ptr(dev)), "equss_token_gram")
    return gram


def probe_losses(logits: torch.Tensor, gram: torch.Tensor, B: int, h: int, w: int, c_total: int, label: torch.Tensor,
                 num_classes: int, cluster_head: Tuple[int, int], linear_head: Tuple[int, int], want_grad: bool = False):
    """Sums behind the two evaluator losses (model/evaluator.py:65-80,95-106) from token-resolution logits, with the
    bilinear upsampling fused in.  Returns (sums float64 [2], n_valid int64 [1], grad_logits or None):
    linear_loss = sums[0] / n_valid, cluster_loss = -sums[1] / label.numel(); grad_logits [B*h*w, C_pad] holds the
    gradients of the two SUMS w.r.t. the token logits (see include/equss_b200.h)."""
    dev = N.require_cuda(logits, gram, label)
    label = label.contiguous()
    assert label.dtype == torch.int64 and label.dim() == 3 and label.shape[0] == B
    H, W = int(label.shape[1]), int(label.shape[2])
    L = N.lib()
    (oc, cc), (ol, cl) = cluster_head, linear_head
    if not L.equss_probe_losses_supported(h, w, H, W, c_total, cc, cl, oc, ol):
        raise ValueError("probe_losses: unsupported head layout / shape")
    acc = torch.zeros((3,), dtype=torch.float64, device=dev)
    sums, nv = acc[:2], acc[2:].view(torch.int64)
    g = torch.zeros_like(logits) if want_grad else None
    rc = L.equss_probe_losses(logits.data_ptr(), gram.data_ptr(), B, h, w, c_total, label.data_ptr(), H, W, num_classes,
                              oc, cc, ol, cl, sums.data_ptr(), nv.data_ptr(), N.ptr(g), N.stream_ptr(dev))
    N.check(rc, "equss_probe_losses")
    return sums, nv, g


def confusion_update(preds: torch.Tensor, label: torch.Tensor, num_classes: int, confusion: torch.Tensor) -> None:
    """confusion[pred, label] += 1 with the reference's mask (model/metric.py:44-58); in place."""
    dev = N.require_cuda(preds, label, confusion)
    N.ensure_device(dev)
    p = preds.reshape(-1).contiguous()
    l = label.reshape(-1).contiguous()
    if p.dtype != torch.int64:
        p = p.long()
    if l.dtype != torch.int64:
        l = l.long()
    assert p.numel() == l.numel()
    assert confusion.dtype == torch.int64 and confusion.is_contiguous() and confusion.shape[1] == num_classes
    rc = N.lib().equss_confusion_update(p.data_ptr(), l.data_ptr(), p.numel(), num_classes, int(confusion.shape[0]),
                                        confusion.data_ptr(), N.stream_ptr(dev))
    N.check(rc, "equss_confusion_update")


def head_gemm(a1: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, a2: Optional[torch.Tensor] = None,
              relu: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One layer of the expansion head (1x1 convolutions of model/dino_pqgo.py:104-112, model/blocks/module.py:27-36):
    ``out[r, o] = act(sum_k A[r, k] * w[o, k] + bias[o])`` with ``A = [a1 | a2]`` along k, on the tensor cores
    (split-tf32, fp32-level accuracy).

    a1: NCHW features (B, C1, h, w) -- read in place when h*w % 4 == 0 -- or flat rows (n, C1); a2: flat rows
    (n, C2) or None; w: (n_out, C1 + C2) (a Conv2d weight reshaped, branches concatenated along the input channels).
    Returns the flat (n, n_out) matrix, n = B*h*w: the (pixel, channel) layout the PQ ops take directly."""
    dev = N.require_cuda(a1, w, bias, a2)
    N.ensure_device(dev)
    L = N.lib()
    a1 = N.f32c(a1.detach())
    if a1.dim() == 4:
        B, C1, h, wd = a1.shape
        hw = h * wd
        nchw = 1
        if not L.equss_head_gemm_supported(C1, 0, hw, 1):     # token count not a multiple of 4: one NHWC copy
            a1 = a1.permute(0, 2, 3, 1).reshape(B * hw, C1).contiguous()
            B, hw, nchw = 1, B * hw, 0
    elif a1.dim() == 2:
        B, hw, C1, nchw = 1, a1.shape[0], a1.shape[1], 0
    else:
        raise ValueError(f"head_gemm: a1 must be (B, C, h, w) or (n, C), got {tuple(a1.shape)}")
    n = B * hw
    C2 = 0
    if a2 is not None:
        a2 = N.f32c(a2.detach())
        if a2.dim() != 2 or a2.shape[0] != n:
            raise ValueError(f"head_gemm: a2 must be ({n}, C2), got {tuple(a2.shape)}")
        C2 = a2.shape[1]
    w = N.f32c(w.detach()).reshape(w.shape[0], -1)
    n_out = w.shape[0]
    if w.shape[1] != C1 + C2:
        raise ValueError(f"head_gemm: weight has {w.shape[1]} input channels, activations have {C1 + C2}")
    b = N.f32c(bias.detach()).reshape(-1) if bias is not None else None
    if b is not None and b.numel() != n_out:
        raise ValueError(f"head_gemm: bias has {b.numel()} entries for {n_out} output channels")
    if out is None:
        out = torch.empty((n, n_out), dtype=torch.float32, device=dev)
    elif out.shape != (n, n_out) or out.dtype != torch.float32 or out.stride(1) != 1:
        raise ValueError("head_gemm: out must be a float32 (n, n_out) matrix with unit column stride")
    if n == 0:
        return out
    rc = L.equss_head_gemm(a1.data_ptr(), nchw, C1, N.ptr(a2), C2, B, hw, w.data_ptr(), N.ptr(b), n_out, int(relu),
                           out.data_ptr(), out.stride(0), N.stream_ptr(dev))
    N.check(rc, "equss_head_gemm")
    return out


def knn_topk(queries: torch.Tensor, db: torch.Tensor, k: int, return_sims: bool = False):
    """Indices of the k most similar database rows for each query, int64 [nq, k], best first
    (data/precompute_knns.py:313-315)."""
    dev = N.require_cuda(queries, db)
    N.ensure_device(dev)
    q = N.f32c(queries.detach())
    d = N.f32c(db.detach())
    nq, F = q.shape
    n, F2 = d.shape
    if F != F2:
        raise ValueError(f"feature dims differ: {F} vs {F2}")
    idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
    sims = torch.empty((nq, k), dtype=torch.float32, device=dev) if return_sims else None
    L = N.lib()
    wsb = int(L.equss_knn_workspace_bytes(nq, n, F, k))
    ws = torch.empty((max(wsb, 1),), dtype=torch.uint8, device=dev)
    rc = L.equss_knn_topk(q.data_ptr(), nq, d.data_ptr(), n, F, k, idx.data_ptr(), N.ptr(sims), ws.data_ptr(), wsb,
                          N.stream_ptr(dev))
    N.check(rc, "equss_knn_topk")
    return (idx, sims) if return_sims else idx


def stego_feature_corr(f1: torch.Tensor, f2: torch.Tensor, pointwise: bool = True) -> torch.Tensor:
    """Feature-correlation tensor of the STEGO loss (model/loss.py:679-687): cosine similarity of every sampled position
    of ``f1`` (n, C, S, S) with every sampled position of ``f2``, optionally row-centred and re-centred on the original
    global mean; returns (n, S, S, S, S), no autograd history (the backbone is frozen)."""
    dev = N.require_cuda(f1, f2)
    N.ensure_device(dev)
    a, b = N.f32c(f1.detach()), N.f32c(f2.detach())
    if a.shape != b.shape or a.dim() != 4:
        raise ValueError(f"stego_feature_corr: expected two (n, C, S, S) tensors, got {tuple(a.shape)} and {tuple(b.shape)}")
    n, C, S1, S2 = a.shape
    P = S1 * S2
    fd = torch.empty((n, P, P), dtype=torch.float32, device=dev)
    sums = torch.zeros((2,), dtype=torch.float64, device=dev)
    N.check(N.lib().equss_stego_feature_corr(a.data_ptr(), b.data_ptr(), n, C, P, int(bool(pointwise)), fd.data_ptr(),
                                             sums.data_ptr(), N.stream_ptr(dev)), "equss_stego_feature_corr")
    if pointwise:
        fd += ((sums[0] - sums[1]) / fd.numel()).float()       # fd - fd.mean() + old_mean  (:686)
    return fd.view(n, S1, S2, S1, S2)
