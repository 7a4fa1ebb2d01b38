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

    # ---- evaluation: probes at token resolution, interpolation + argmax + histogram at label resolution ------------
    import torch.nn.functional as F
    from equss_b200 import _native as N

    def probe_pack(wmat):
        wmat = wmat.detach().float()
        Ct, D = wmat.shape
        cpad = int(N.lib().equss_probe_cpad(Ct))                     # host helper of the library: channel padding rule
        wmat_t = torch.zeros(D, cpad)
        wmat_t[:, :Ct] = wmat.t()
        return wmat_t, Ct, None

    def probe_logits(feat, wmat, bias=None, algo=0):
        wmat_t, Ct, _ = wmat if isinstance(wmat, tuple) else probe_pack(wmat)
        B, D, h, w = feat.shape
        logits = feat.detach().float().permute(0, 2, 3, 1).reshape(-1, D) @ wmat_t
        if bias is not None:
            logits[:, :bias.numel()] += bias.detach().float().reshape(-1)
        return logits

    def upsampled(logits, B, h, w, H, W):
        maps = logits.reshape(B, h, w, -1).permute(0, 3, 1, 2)
        return maps if (h, w) == (H, W) else F.interpolate(maps, (H, W), mode="bilinear", align_corners=False)

    def confusion_update(preds, label, num_classes, confusion):
        p, l = preds.reshape(-1).long(), label.reshape(-1).long()
        keep = (l >= 0) & (l < num_classes) & (p >= 0) & (p < num_classes)               # model/metric.py:49
        confusion += torch.bincount(p[keep] * num_classes + l[keep], minlength=confusion.shape[0] * num_classes
                                    ).reshape(confusion.shape[0], num_classes)

    def probe_argmax_confusion(logits, B, h, w, c_total, label, num_classes, heads, want_preds=True, confusions=None):
        H, W = label.shape[1:]
        up = upsampled(logits, B, h, w, H, W)
        preds = []
        for i, (off, cnt) in enumerate(heads):
            p = up[:, off:off + cnt].argmax(dim=1)
            if confusions is not None and confusions[i] is not None:
                confusion_update(p, label, num_classes, confusions[i])
            preds.append(p if want_preds else None)
        return preds

    def token_gram(feat):
        x = feat.detach().float()
        B, D, h, w = x.shape
        g = torch.zeros(B, h, w, 5)
        g[..., 0] = (x * x).sum(1)
        g[:, :, :-1, 1] = (x[..., :, :-1] * x[..., :, 1:]).sum(1)                         # <x, right>
        g[:, :-1, :, 2] = (x[..., :-1, :] * x[..., 1:, :]).sum(1)                         # <x, down>
        g[:, :-1, :-1, 3] = (x[..., :-1, :-1] * x[..., 1:, 1:]).sum(1)                    # <x, down-right>
        g[:, :-1, 1:, 4] = (x[..., :-1, 1:] * x[..., 1:, :-1]).sum(1)                     # <x, down-left>
        return g.reshape(B * h * w, 5)

    def taps(out_size, in_size):
        src = ((torch.arange(out_size, dtype=torch.float32) + 0.5) * (float(in_size) / float(out_size)) - 0.5).clamp_min(0.0)
        i0 = src.floor().long().clamp_(max=in_size - 1)
        i1 = (i0 + 1).clamp(max=in_size - 1)
        return i0, i1, src - i0.float()

    def norm_from_gram(gram, B, h, w, H, W):
        """|sum_t w_t x_t| over the (up to) four taps of every label pixel, from the pairwise inner products."""
        g = gram.reshape(B, h, w, 5)
        y0, y1, ly = taps(H, h)
        x0, x1, lx = taps(W, w)
        Y0, Y1, X0, X1 = y0[:, None], y1[:, None], x0[None, :], x1[None, :]
        wts = {(0, 0): (1 - ly)[:, None] * (1 - lx)[None, :], (0, 1): (1 - ly)[:, None] * lx[None, :],
               (1, 0): ly[:, None] * (1 - lx)[None, :], (1, 1): ly[:, None] * lx[None, :]}
        pos = {(0, 0): (Y0, X0), (0, 1): (Y0, X1), (1, 0): (Y1, X0), (1, 1): (Y1, X1)}

        def inner(a, b):                                             # <x[pos a], x[pos b]> for every label pixel
            (ya, xa), (yb, xb) = pos[a], pos[b]
            ya, xa, yb, xb = (t.expand(H, W) for t in (ya, xa, yb, xb))
            dy, dx = yb - ya, xb - xa                                # in {0, 1} x {-1, 0, 1} after ordering by row
            swap = (dy < 0) | ((dy == 0) & (dx < 0))
            ya2, xa2 = torch.where(swap, yb, ya), torch.where(swap, xb, xa)
            dy, dx = torch.where(swap, -dy, dy), torch.where(swap, -dx, dx)
            term = torch.where(dy == 0, torch.where(dx == 0, 0, 1), torch.where(dx == 0, 2, torch.where(dx > 0, 3, 4)))
            return g[:, ya2, xa2, :].gather(-1, term.expand(B, H, W).unsqueeze(-1)).squeeze(-1)

        n2 = torch.zeros(B, H, W)
        keys = list(wts)
        for a in keys:
            for b in keys:
                n2 = n2 + wts[a] * wts[b] * inner(a, b)
        return n2.clamp_min(0).sqrt()

    def probe_losses(logits, gram, B, h, w, c_total, label, num_classes, cluster_head, linear_head, want_grad=False):
        """include/equss_b200.h K8b: sums[0] = sum over labelled pixels of logsumexp(v_lin) - v_lin[label],
        sums[1] = sum over all pixels of max_c v_clu / |upsampled feature|; grad_logits = d sums / d token logits."""
        H, W = label.shape[1:]
        (oc, cc), (ol, cl) = cluster_head, linear_head
        with torch.enable_grad():
            lg = logits.detach().clone().requires_grad_(True)
            up = upsampled(lg, B, h, w, H, W)
            lin = up[:, ol:ol + cl].permute(0, 2, 3, 1).reshape(-1, cl)
            lab = label.reshape(-1)
            valid = (lab >= 0) & (lab < num_classes)
            s_lin = (torch.logsumexp(lin[valid], dim=1) - lin[valid].gather(1, lab[valid].unsqueeze(1)).squeeze(1)).double().sum()
            norm = norm_from_gram(gram, B, h, w, H, W).clamp_min(1e-12)
            s_clu = (up[:, oc:oc + cc].max(dim=1).values / norm).double().sum()
            g = torch.autograd.grad(s_lin + s_clu, lg)[0] if want_grad else None
        return torch.stack([s_lin.detach(), s_clu.detach()]), valid.sum().reshape(1), g

    def head_gemm(a1, w, bias=None, a2=None, relu=False, out=None):
        """out[r, o] = act(sum_k [a1 | a2][r, k] w[o, k] + bias[o]); a1 NCHW (B, C1, h, w) or flat rows."""
        a = a1.detach().float()
        if a.dim() == 4:
            a = a.permute(0, 2, 3, 1).reshape(-1, a.shape[1])
        if a2 is not None:
            a = torch.cat([a, a2.detach().float()], dim=1)
        y = a @ w.detach().float().reshape(w.shape[0], -1).t()
        if bias is not None:
            y = y + bias.detach().float().reshape(1, -1)
        y = torch.relu(y) if relu else y
        if out is not None:
            out.copy_(y)
            return out
        return y

    def stego_feature_corr(f1, f2, pointwise=True):
        """model/loss.py:679-687: cosine correlation of the sampled positions, rows centred, global mean restored."""
        a, b = F.normalize(f1.detach().float(), dim=1, eps=1e-10), F.normalize(f2.detach().float(), dim=1, eps=1e-10)
        fd = torch.einsum("nchw,ncij->nhwij", a, b)
        if pointwise:
            old_mean = fd.mean()
            fd = fd - fd.mean(dim=[3, 4], keepdim=True)
            fd = fd - fd.mean() + old_mean
        return fd

    for name, fn in list(locals().items()):
        if callable(fn) and hasattr(ops, name):
            monkeypatch.setattr(ops, name, fn)
    return ops
