"""Host-side data-parallel logic on CPU with gloo, world_size 2 (no GPU needed).

The kernels themselves need a B200; what is covered here is the plumbing around them: pixel sharding,
the clone-and-return contract of all_reduce_tensor, and that ONE all-reduce of the packed [M, K, d+1]
statistics buffer reproduces the single-process statistics (counts exactly, sums up to fp32 add order),
so every rank applies an identical EMA update; the query sharding + all_gather of the kNN table
(ragged shards), the single int64 all-reduce inside UnSegMetrics.compute, and ProductQuantizerWrapper.train() itself on two
ranks (kernel entry points replaced by their torch definitions): replicas bit-identical, equal to the oracle on the whole batch."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import equss_oracle as O
        from equss_b200 import dist_utils as DU
        torch.manual_seed(0)                       # same data on every rank, each takes its shard
        M, K, d, n = 3, 16, 8, 203                 # ragged: 203 pixels over 2 ranks
        z = torch.randn(n, M * d)
        w0 = torch.randn(M, K, d)
        lo, hi = DU.shard_range(n)
        assert (lo, hi) == ((0, 102) if rank == 0 else (102, 203))
        # all_reduce_tensor returns a reduced clone and leaves its input alone (utils/dist_utils.py:98-113)
        t = torch.full((4,), float(rank + 1))
        r = DU.all_reduce_tensor(t, op="sum")
        assert torch.equal(t, torch.full((4,), float(rank + 1))) and torch.equal(r, torch.full((4,), 3.0))
        assert torch.equal(DU.all_reduce_tensor(t, op="mean"), torch.full((4,), 1.5))
        with pytest.raises(RuntimeError):
            DU.all_reduce_tensor(t, op="max")
        # the in-kernel peer exchange needs NCCL + CUDA symmetric memory: with gloo / CPU tensors the step falls back to ONE
        # all-reduce of the packed buffer, silently (no warning: nothing failed, the backend simply has no peer memory)
        assert DU.packed_peer_exchange((M, K, d + 1), torch.device("cpu")) is None
        assert DU.packed_peer_exchange((M, K, d + 1), torch.device("cuda", 0)) is None      # backend is gloo
        # per-rank packed statistics on the shard (oracle arithmetic stands in for K4), one all-reduce (K5)
        packed = torch.zeros(M, K, d + 1)
        idx_full = []
        for m in range(M):
            zn, cn = O.normalize_pair(z[:, m * d:(m + 1) * d], w0[m], "l2")
            idx = torch.argmin(O.sq_distance(zn, cn), dim=1)
            idx_full.append(idx)
            oh = torch.nn.functional.one_hot(idx[lo:hi], K).float()
            packed[m, :, :d] = oh.t() @ z[lo:hi, m * d:(m + 1) * d]
            packed[m, :, d] = oh.sum(0)
        DU.all_reduce_packed_(packed)
        for m in range(M):
            oh = torch.nn.functional.one_hot(idx_full[m], K).float()
            assert torch.equal(packed[m, :, d], oh.sum(0))                                     # counts: exact
            torch.testing.assert_close(packed[m, :, :d], oh.t() @ z[:, m * d:(m + 1) * d], rtol=1e-5, atol=1e-5)
        # identical EMA update on every rank -> replicas stay bit-identical
        st = O.EmaState(w0[0])
        st.update(packed[0, :, d], packed[0, :, :d])
        gathered = [torch.empty_like(st.weight) for _ in range(world)]
        dist.all_gather(gathered, st.weight)
        assert torch.equal(gathered[0], gathered[1])
        # kNN: queries sharded over the ranks (ragged: 203 rows over 2 ranks), database replicated, table all-gathered
        # (data/precompute_knns.py:305-319).  The shard kernel is replaced by its torch definition; what runs here is
        # the sharding, the -1 padding of the last shard and the all_gather.
        from equss_b200 import knn as KN
        from equss_b200 import ops
        seen = []

        def knn_stub(q, db, k):
            seen.append(tuple(q.shape))
            return torch.topk(q @ db.t(), k, dim=1).indices

        real_knn, ops.knn_topk = ops.knn_topk, knn_stub
        try:
            feats = torch.nn.functional.normalize(torch.randn(n, 24), dim=1)
            table = KN.precompute_knns(feats, k=5)
            assert seen == [(hi - lo, 24)]
            assert table.shape == (n, 5) and table.dtype == torch.int64
            assert torch.equal(table, O.knn(feats, k=5)[0]) and torch.equal(table[:, 0], torch.arange(n))
            assert KN.precompute_knns(feats, k=5, sharded=False).shape == (n, 5) and seen[-1] == (n, 24)
        finally:
            ops.knn_topk = real_knn
        # UnSegMetrics.compute: ONE int64 all-reduce of the confusion matrix (model/metric.py:63), every rank then
        # runs the Hungarian match on the same matrix
        import tempfile
        from equss_b200.metric import UnSegMetrics
        os.chdir(tempfile.mkdtemp())                   # compute() writes ./class_matrix/*.csv like the reference
        C = 5
        lab = torch.randint(-1, C, (n,))
        pred = (lab.clamp_min(0) + (torch.arange(n) % 7 == 0).long()) % C
        met = UnSegMetrics(C, 0, True, torch.device("cpu"))
        met.confusion_matrix += O.confusion_update(torch.zeros(C, C, dtype=torch.long), pred[lo:hi], lab[lo:hi], C)
        res = met.compute("gloo")
        full = O.confusion_update(torch.zeros(C, C, dtype=torch.long), pred, lab, C)
        assert torch.equal(met.confusion_matrix.cpu(), full)
        want = O.metrics_compute(full, True)
        for k, v in want.items():
            assert abs(float(res[k]) - float(v)) < 1e-6, (k, float(res[k]), float(v))
        # The MODULE on two ranks: ProductQuantizerWrapper.train() on each rank's shard (kernel entry points replaced by
        # their torch definitions, tests/kernel_standins.py), packed statistics all-reduced by the module itself
        # (one call for all subspaces), z_trainable moments by one [2, D] all-reduce.  Replicas must stay bit-identical
        # and equal the single-process oracle on the whole batch.
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import kernel_standins

        class _Patch:
            def setattr(self, obj, name, value):
                setattr(obj, name, value)
        kernel_standins.install(_Patch())
        from equss_b200.quantizer import EMAVectorQuantizer, ProductQuantizerWrapper
        for mode in ("l2", "z_trainable"):
            torch.manual_seed(5)
            Mq, Kq, dq, nq = 2, 8, 4, 60
            pq = ProductQuantizerWrapper(Mq, Kq, Mq * dq, beta=0.25, normalize=mode, quantizer_cls=EMAVectorQuantizer)
            wq = torch.randn(Mq, Kq, dq)
            with torch.no_grad():
                for i, q in enumerate(pq.quantizers):
                    q.codebook.weight.copy_(wq[i]); q.codebook.weight_avg.copy_(wq[i])
            states = [O.EmaState(wq[i]) for i in range(Mq)]
            exact = [torch.zeros(Kq) for _ in range(Mq)]
            zstat = ([torch.zeros(dq) for _ in range(Mq)], [torch.zeros(dq) for _ in range(Mq)])
            pq.train()
            for step in range(3):
                zfull = torch.randn(nq, Mq * dq) + 0.2 * step
                a, b = DU.shard_range(nq)
                with torch.no_grad():
                    zq, out, _ = pq(zfull[a:b])
                ref_rows = []
                for i in range(Mq):
                    kw = {"z_mean": zstat[0][i], "z_log_var": zstat[1][i]} if mode == "z_trainable" else {}
                    qi, oi, _, _ = O.ema_vq_forward(zfull[:, i * dq:(i + 1) * dq], states[i], exact[i], normalize=mode,
                                                    beta=0.25, training=True, **kw)
                    ref_rows.append(qi)
                torch.testing.assert_close(zq, torch.cat(ref_rows, dim=1)[a:b], rtol=1e-5, atol=1e-6)
                for i, q in enumerate(pq.quantizers):
                    torch.testing.assert_close(q.codebook.weight, states[i].weight, rtol=1e-5, atol=1e-6)
                    assert torch.equal(q.vq_count, exact[i])                     # exact counts of the WHOLE batch
                    both = [torch.empty_like(q.codebook.weight) for _ in range(world)]
                    dist.all_gather(both, q.codebook.weight.contiguous())
                    assert torch.equal(both[0], both[1])                         # replicas bit-identical
        ret[rank] = "ok"
    except BaseException as e:   # noqa
        ret[rank] = f"{type(e).__name__}: {e}"
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}, dict(ret)


def test_single_process_helpers_are_identity():
    sys.path.insert(0, ROOT)
    from equss_b200 import dist_utils as DU
    t = torch.arange(4.0)
    assert DU.all_reduce_tensor(t) is t                    # reference returns the input itself (dist_utils.py:99-100)
    assert DU.all_reduce_packed_(t) is t
    assert DU.get_world_size() == 1 and DU.get_rank() == 0
    assert DU.packed_peer_exchange((2, 4, 5), torch.device("cpu")) is None      # no process group: no exchange at all
    assert DU.shard_range(10, 1, 4) == (3, 6) and DU.shard_range(10, 3, 4) == (9, 10)
