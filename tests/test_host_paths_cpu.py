"""Host-side PyTorch pieces of the research flags on CPU tensors, against the fixtures produced by the
unmodified reference (oracle/make_golden_flags.py): the Gumbel index draw and the weighted sum are plain torch code and
need no device; the kernels around them are covered by tests/test_gpu_flags.py."""
import os

import numpy as np
import torch
import torch.nn.functional as F


def _fixed_gumbel(noises):
    calls = [0]

    def fake(logits, tau=1.0, hard=True, dim=1):
        y = logits + noises[calls[0]]
        calls[0] += 1
        return F.one_hot(torch.argmax(y, dim=1), logits.shape[1]).to(logits.dtype)
    return fake, calls


def test_gumbel_indices_follow_the_reference_draw(golden_dir, monkeypatch):
    import equss_b200  # noqa: F401
    from equss_b200 import _pq_core as core
    from equss_b200._host_paths import gumbel_indices
    g = np.load(os.path.join(golden_dir, "pq_flag_ema_gumbel_l2.npz"))
    M = int(g["M"])
    z, w0 = torch.from_numpy(g["z0"]), torch.from_numpy(g["weight0"])
    fake, calls = _fixed_gumbel(list(torch.from_numpy(g["noise0"])))
    monkeypatch.setattr(torch.nn.functional, "gumbel_softmax", fake)
    z_norm = core._normalize_rows(core._rows(z, M), "l2", None, None)
    idx = gumbel_indices(z_norm, F.normalize(w0, dim=2), 0.01)                  # EMAVectorQuantizer: -distance / 0.01
    assert calls[0] == M and idx.dtype == torch.int32 and tuple(idx.shape) == (M, z.shape[0])
    assert np.array_equal(idx.numpy(), g["idx0"])
    # VectorQuantizer: plain -distance, one subspace, NCHW rows
    g = np.load(os.path.join(golden_dir, "pq_flag_param_gumbel.npz"))
    z = torch.from_numpy(g["z"])
    fake, calls = _fixed_gumbel([torch.from_numpy(g["noise"])])
    monkeypatch.setattr(torch.nn.functional, "gumbel_softmax", fake)
    z_norm = core._normalize_rows(core._rows(z, 1), "l2", None, None)
    idx = gumbel_indices(z_norm, F.normalize(torch.from_numpy(g["weight"]), dim=1).unsqueeze(0), None)
    assert calls[0] == 1 and np.array_equal(idx[0].numpy(), g["idx"])


def test_weighted_sum_matches_reference_values_and_gradients(golden_dir):
    import equss_b200  # noqa: F401
    from equss_b200.codebooks import _weighted_sum
    for variant, book in (("new_vq", 1.0), ("pqgo", 1.0)):
        g = np.load(os.path.join(golden_dir, f"pq_flag_inline_{variant}_weighted.npz"))
        K, ts = int(g["K"]), float(g["jsd_ts"])
        z = torch.from_numpy(g["z"]).requires_grad_(True)
        w = torch.from_numpy(g["weight"]).clone().requires_grad_(True)
        zf = z.permute(0, 2, 3, 1).reshape(-1, z.shape[1])
        dist = (zf ** 2).sum(1, keepdim=True) + (w ** 2).sum(1) - 2 * zf @ w.t()
        prob = F.softmax(-dist / ts, dim=1)                    # what the kernel's DistanceProb returns on the device
        out, commit, cb_loss = _weighted_sum(z, prob, w.unsqueeze(0), "none", None, None, 1, K)
        np.testing.assert_allclose(out.detach().numpy(), g["zq"], rtol=1e-5, atol=1e-7)
        loss = (book * cb_loss + 0.25 * commit).mean()
        assert abs(float(loss.detach()) - float(g["out/vq-loss"])) <= 1e-6 * abs(float(g["out/vq-loss"])) + 1e-9
        ((out * torch.from_numpy(g["go"])).sum() + loss).backward()
        np.testing.assert_allclose(z.grad.numpy(), g["grad_z"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(w.grad.numpy(), g["grad_w"], rtol=1e-4, atol=1e-6)
