"""Torch definitions of the kernel entry points in ``equss_b200.ops`` -- TEST INFRASTRUCTURE for the CPU suite only.

The product path has no CPU fallback (a CPU tensor raises ``EqussNativeError``).  What the CPU suite can still exercise
is the HOST logic of the nn.Module mirrors around the kernels: stacking of the per-subspace state, which statistics go
where, output keys, return tuples, write-back into the state-dict buffers.  ``install(monkeypatch)`` replaces the
entry points the PQ modules call by the plain torch expression each kernel is specified to compute (the contract stated
in the docstrings of ``ops.py`` / ``include/equss_b200.h``), so that the reference's fixtures can be replayed through the
mirrors without a device.  The kernels themselves are held to the same fixtures by the ``-m gpu`` tests.
"""

import torch


def install(monkeypatch):
    from equss_b200 import _pq_core as core
    from equss_b200 import ops

    def rows(z, M, mode, a, b):
        return core._normalize_rows(core._rows(z.detach().float(), M), mode, a, b)          # (n, M, d)

    def like_z(t, z):                                                                        # (n, M, d) -> layout of z
        if z.dim() == 2:
            return t.reshape(z.shape[0], -1).contiguous()
        B, D, h, w = z.shape
        return t.reshape(B, h, w, D).permute(0, 3, 1, 2).contiguous()

    def distance(zr, cbn):
        return (zr ** 2).sum(-1, keepdim=True) + (cbn ** 2).sum(-1).unsqueeze(0) - 2 * torch.einsum("nmd,mkd->nmk", zr, cbn)

    def pq_cnorm2(codebook_norm):
        return (codebook_norm.detach().float() ** 2).sum(-1)

    def pq_prepare_codebook(codebook, normalize):
        cbn = core.normalize_codebook(codebook.detach().float(), normalize, ema_style=True).clone()
        return cbn, (cbn ** 2).sum(-1)

    def pq_assign(z, codebook_norm, cnorm2=None, normalize="l2", norm_a=None, norm_b=None, algo=0, return_margin=False):
        cbn = codebook_norm.detach().float()
        return distance(rows(z, cbn.shape[0], normalize, norm_a, norm_b), cbn).argmin(dim=2).t().to(torch.int32).contiguous()

    def pq_gather_loss(z, gather_src, idx, normalize="l2", norm_a=None, norm_b=None, want_znorm=False):
        src = gather_src.detach().float()
        M, K, d = src.shape
        zr = rows(z, M, normalize, norm_a, norm_b)
        q = torch.gather(src, 1, idx.long().unsqueeze(-1).expand(M, idx.shape[1], d)).permute(1, 0, 2)
        out = zr + (q - zr)                                                                  # model/quantizer.py:536
        sqerr = ((zr - q).double() ** 2).sum(dim=(0, 2))
        return like_z(out, z), sqerr, (like_z(zr, z) if want_znorm else None)

    def pq_gather_loss_bwd(z, gather_src, idx, normalize, grad_out, coef, norm_a=None, norm_b=None, want_grad_z=True,
                           cb_coef=None):
        """include/equss_b200.h: g_znorm = grad_out + coef[m] (z_norm - q), grad_z = J_norm(z)^T g_znorm;
        grad_gather_src += cb_coef[m] (q - z_norm) scattered by idx."""
        src = gather_src.detach().float()
        M, K, d = src.shape
        n = idx.shape[1]
        with torch.enable_grad():
            zl = z.detach().float().requires_grad_(True)
            zr = core._normalize_rows(core._rows(zl, M), normalize, norm_a, norm_b)
        q = torch.gather(src, 1, idx.long().unsqueeze(-1).expand(M, n, d)).permute(1, 0, 2)
        gz = gcb = None
        if want_grad_z:
            g_zn = torch.zeros_like(zr) if grad_out is None else core._rows(grad_out.detach().float(), M).clone()
            if coef is not None:
                g_zn = g_zn + coef.detach().float().reshape(1, M, 1) * (zr.detach() - q)
            (gz,) = torch.autograd.grad(zr, zl, g_zn)
        if cb_coef is not None:
            gcb = torch.zeros_like(src)
            contrib = cb_coef.detach().float().reshape(1, M, 1) * (q - zr.detach())
            for m in range(M):
                gcb[m].index_add_(0, idx[m].long(), contrib[:, m])
        return gz, gcb

    def pq_assign_gather(z, codebook_norm, gather_src=None, cnorm2=None, normalize="l2", norm_a=None, norm_b=None, fused=None):
        idx = pq_assign(z, codebook_norm, cnorm2, normalize, norm_a, norm_b)
        out, sqerr, _ = pq_gather_loss(z, codebook_norm if gather_src is None else gather_src, idx, normalize, norm_a, norm_b)
        return idx, out, sqerr

    def pq_distance_prob(z, codebook_norm, cnorm2=None, normalize="l2", norm_a=None, norm_b=None, temperature=1.0):
        cbn = codebook_norm.detach().float()
        zr = rows(z, cbn.shape[0], normalize, norm_a, norm_b)
        return torch.softmax(-distance(zr, cbn) / temperature, dim=2).reshape(zr.shape[0], -1)

    def pq_accumulate(z, idx, num_codebook, use_norm=False, normalize="l2", norm_a=None, norm_b=None, out=None):
        M, K = idx.shape[0], int(num_codebook)
        zr = rows(z, M, normalize if use_norm else "none", norm_a, norm_b)
        packed = torch.zeros(M, K, zr.shape[2] + 1) if out is None else out
        for m in range(M):
            packed[m, :, :-1].index_add_(0, idx[m].long(), zr[:, m])
            packed[m, :, -1] += torch.bincount(idx[m].long(), minlength=K).float()
        return packed

    def ema_update(packed, decay, eps, vq_count, weight_avg, weight, exact_count=None):
        count, total = packed[:, :, -1], packed[:, :, :-1]
        K = count.shape[1]
        if exact_count is not None:
            exact_count += count
        vq_count.mul_(decay).add_(count, alpha=1 - decay)                                    # model/quantizer.py:242
        weight_avg.mul_(decay).add_(total, alpha=1 - decay)                                  # :245
        n = vq_count.sum(dim=1, keepdim=True)
        weight.copy_(weight_avg / ((vq_count + eps) / (n + K * eps) * n).unsqueeze(-1))      # :248-254
        return (count == 0).sum(dim=1).to(torch.int32)

    def usage_percentiles(count):
        count = count.float()
        M, K = count.shape
        prob = count / (count.sum(dim=1, keepdim=True) + 1)                                  # model/quantizer.py:16
        csum = torch.cumsum(torch.sort(prob, dim=1, descending=True)[0], dim=1)
        out = torch.full((M, 3), float("nan"))
        for m in range(M):
            for t, level in enumerate((0.1, 0.5, 0.9)):
                hit = torch.nonzero(csum[m] >= level)
                if hit.numel():
                    out[m, t] = float(hit[0, 0]) / K
        return out

    def pq_train_tail(packed, decay, eps, vq_count, weight_avg, weight, exact_count, sqerr, n_pixels, beta,
                      peers=None, zero_next=None):
        assert peers is None
        M, K, d1 = packed.shape
        if K > 1024:
            return None
        count = packed[:, :, -1].clone()
        unused = ema_update(packed, decay, eps, vq_count, weight_avg, weight, exact_count)
        stats = torch.zeros(10)
        stats[0:3] = usage_percentiles(exact_count).mean(dim=0)
        stats[3:6] = usage_percentiles(count).mean(dim=0)
        stats[6] = ((K - unused.float()) / K).mean()
        stats[7] = weight.abs().sum() / M
        if sqerr is not None:
            stats[8] = (sqerr / max(n_pixels * (d1 - 1), 1)).float().mean()
            stats[9] = beta * stats[8]
        return stats

    def channel_moments(z):
        zr = core._rows(z.detach().float(), 1)[:, 0]
        return torch.stack([zr.mean(dim=0), (zr * zr).mean(dim=0)])

    def pq_soft_stats(z, codebook_norm, cnorm2=None, normalize="l2", norm_a=None, norm_b=None, temperature=1.0):
        cbn = codebook_norm.detach().float()
        prob = pq_distance_prob(z, cbn, None, normalize, norm_a, norm_b, temperature)
        return core.soft_assignment_stats(prob, cbn.shape[0], cbn.shape[1])

    for name, fn in list(locals().items()):
        if callable(fn) and hasattr(ops, name):
            monkeypatch.setattr(ops, name, fn)
    return ops
