#!/usr/bin/env python
"""bench.py -- EQUSS product-quantization hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl equss|reference] [--no-extras]

Metric (BASELINE.json): PQ-quantized pixels/sec.  Headline workload at every N is BASELINE configs[1]
"cocostuff27 eval shape": per rank, features (32, 1024, 40, 40) fp32 NCHW -> PQ head (M=64 subspaces x
K=256 codewords, d=16, l2) assign + gather -> cluster + linear probe argmax at 320x320 label resolution
-> two 27x27 confusion histograms.  One "step" = one such batch (51 200 PQ-quantized pixels per rank).

value : device-resident throughput (inputs already in HBM), CUDA events, max over ranks.
e2e   : same step through the public module API (PQGOProductQuantizerWrapper / UnSegEvaluator) with pinned
        HOST inputs copied in, and the confusion matrices copied out, inside the timed region; also reported through
        the reference's unchanged call sequence (evaluator.forward + UnSegMetrics.update x2, train.py:268-281).
The same JSON line carries one block per remaining BASELINE config, measured the same way at the same N:
  train : configs[2], the EMA training step through the eager module API incl. the packed NCCL all-reduce (weak scaling)
  c4    : configs[3], cityscapes high-res, 50 176 pixels sharded over the ranks (strong scaling)
  c5    : configs[4], 50k x 768 kNN, queries sharded over the ranks + all_gather (strong scaling)
  c1    : configs[0], pq_baseline forward at 3 136 pixels (latency)
torch_eager_gpu (N = 1): the oracle port with its tensors on the GPU (PyTorch eager, library kernels) -- the "same-box
PyTorch" comparator of SURVEY 8d, a reported baseline like cpu_baseline.
`--impl reference` times the CPU oracle port of the reference path (oracle/equss_oracle.py, torch CPU ops
with all host threads) on a bounded sample of the same workload; the reference itself is Python and
/root/reference does not exist on the GPU box.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(B=32, D=1024, h=40, w=40, H=320, W=320, M=64, K=256, C=27)
METRIC = "PQ-quantized pixels/sec"
UNIT = "pixels/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------
# CPU oracle pipeline (reference arm + cpu_baseline)
# ------------------------------------------------------------------------------------------------------
def cpu_pipeline_factory(n_images, device="cpu"):
    """The oracle port of the reference path on `n_images` images of the headline workload.  device="cpu" is the
    cpu_baseline / reference arm; a CUDA device runs the SAME torch code with library kernels (ATen / cuBLAS) as the
    "same-box PyTorch eager" comparator of SURVEY 8d -- a reported baseline, never the product path."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import equss_oracle as O
    c = CFG
    g = torch.Generator().manual_seed(0)
    d = c["D"] // c["M"]
    feat = torch.randn(n_images, c["D"], c["h"], c["w"], generator=g).to(device)
    label = torch.randint(-1, c["C"], (n_images, c["H"], c["W"]), generator=g).to(device)
    cb = torch.randn(c["M"], c["K"], d, generator=g).to(device)
    clusters = torch.randn(c["C"], c["D"], generator=g).to(device)
    lin_w = (torch.randn(c["C"], c["D"], generator=g) * 0.03).to(device)
    lin_b = torch.zeros(c["C"], device=device)

    def step():
        qs = []
        for m in range(c["M"]):   # the reference's per-subspace Python loop (model/dino_pqgo.py:757-770)
            q, _, _, _ = O.param_vq_forward(feat[:, m * d:(m + 1) * d], cb[m], normalize="l2")
            qs.append(q)
        zq = torch.cat(qs, dim=1)
        _, lp, _, cp = O.evaluator_forward(zq, label, clusters, lin_w, lin_b, c["C"])
        conf_c = O.confusion_update(torch.zeros(c["C"], c["C"], dtype=torch.long, device=device), cp, label, c["C"])
        conf_l = O.confusion_update(torch.zeros(c["C"], c["C"], dtype=torch.long, device=device), lp, label, c["C"])
        return conf_c, conf_l

    return step, n_images * c["h"] * c["w"]


def time_cpu(step, n_warm, n_timed):
    for _ in range(n_warm):
        step()
    ts = []
    for _ in range(n_timed):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return ts


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    n_img = 4
    step, px = cpu_pipeline_factory(n_img)
    ts = time_cpu(step, args.warmup, args.steps)
    total = sum(ts)
    value = px * len(ts) / total
    sample = f"{n_img} of {CFG['B']} images per step ({px} pixels), oracle port of the reference path, torch CPU ops"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cocostuff27_eval (BASELINE configs[1]), bounded CPU sample", **CFG,
                   "normalize": "l2", "sample": sample, "same_config": False,
                   "sampling": f"{n_img} of {CFG['B']} images per step at unchanged per-image shapes; the metric is a per-pixel rate"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled every 100 ms from before the warm-up to after the last timed region.  The timed regions are
    milliseconds long, so the figure reported is the median SM clock over the samples taken while this process kept
    the GPU busy (warm-up + every timed block), not a single sample."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def wait_first(self, timeout=8.0):
        """nvidia-smi takes a while to start when eight ranks launch it at once: wait for its first line."""
        t_end = time.time() + timeout
        while self.proc is not None and not self.lines and time.time() < t_end:
            time.sleep(0.05)

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [ln.split(", ") for ts, ln in self.lines if t0 - 0.05 <= ts <= t1 + 0.15] or [ln.split(", ") for _, ln in self.lines]
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for nm, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank (and therefore the pinned host buffers it allocates afterwards: first-touch) to the CPU cores of the
    NUMA node its GPU hangs off.  torchrun starts ranks unbound; with eight of them streaming 236 MB per step from
    pinned memory, half of the traffic otherwise crosses the socket interconnect."""
    try:
        props = torch.cuda.get_device_properties(local_rank)
        if all(hasattr(props, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        else:
            bus = subprocess.run(["nvidia-smi", f"--id={local_rank}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=10).stdout.strip()
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node_file = f"/sys/bus/pci/devices/{bus}/numa_node"
        node = int(open(node_file).read().strip())
        if node < 0:
            return {"numa_node": None, "note": "platform reports no NUMA affinity for the GPU"}
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
        return {"numa_node": node, "cpus": len(ids)}
    except Exception as e:   # affinity is an optimisation, never a requirement
        return {"numa_node": None, "note": f"{type(e).__name__}: {e}"}


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def _extra_peaks():
    p = os.path.join(ROOT, "profiles", "r2_measured_peaks_extra.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def run_equss(args):
    import torch.distributed as dist
    import torch.nn.functional as F
    import equss_b200
    from equss_b200 import ops
    from equss_b200.codebooks import PQGOProductQuantizerWrapper
    from equss_b200.evaluator import UnSegEvaluator
    from equss_b200.knn import precompute_knns
    from equss_b200.metric import UnSegMetrics
    from equss_b200.quantizer import ProductQuantizerWrapper

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: equss_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else {"numa_node": None, "note": "single rank: not bound"}
    json_fd = None
    if world > 1:
        # keep stdout for the one JSON line: NCCL prints its version banner on fd 1 from C code, so everything that
        # writes to fd 1 during the run is sent to stderr and the JSON line goes to a duplicate of the real stdout
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    c = CFG
    B, D, h, w, H, W, M, K, C = (c[k] for k in ("B", "D", "h", "w", "H", "W", "M", "K", "C"))
    d = D // M
    N = B * h * w
    P = B * H * W
    torch.manual_seed(1234 + rank)
    NBUF = 3   # rotate inputs so every step reads buffers last touched ~0.7 GB ago (L2 is 126 MB)
    sampler = ClockSampler(local)
    sampler.start()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(step_fn, n_steps, n_warm, per_step_events=0):
        """n_warm untimed steps, then n_steps timed with CUDA events on the current stream, barrier + synchronize on both
        sides, MAX over ranks.  Returns (total ms, [per-step event lists])."""
        for i in range(n_warm):
            step_fn(i, None)
        sync_all()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(per_step_events)] for _ in range(n_steps)] if per_step_events else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_steps):
            step_fn(i, evs[i] if evs else None)
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), evs

    peak, peak_src = _peaks()
    xpk = _extra_peaks()
    sampler.wait_first()
    t_wall0 = time.time()
    l0 = ops.launch_count()

    # ================= headline: cocostuff27 eval step (BASELINE configs[1]) ==================================
    zs = [torch.randn(B, D, h, w, device=dev) for _ in range(NBUF)]
    labels = [torch.randint(-1, C, (B, H, W), device=dev) for _ in range(NBUF)]
    codebook = torch.randn(M, K, d, device=dev)
    clusters = torch.randn(C, D, device=dev)
    lin_w = torch.randn(C, D, device=dev) * 0.03
    Cp = (C + 3) // 4 * 4                      # each probe head starts at a multiple of four channels
    wmat = torch.zeros(Cp + C, D, device=dev)
    wmat[:C] = F.normalize(clusters, dim=1)
    wmat[Cp:] = lin_w
    bias = torch.zeros(Cp + C, device=dev)
    wpack = ops.probe_pack(wmat)
    conf_c = torch.zeros(C, C, dtype=torch.long, device=dev)
    conf_l = torch.zeros(C, C, dtype=torch.long, device=dev)
    stages = ["pq_assign_gather", "probe_logits", "probe_argmax_confusion"]

    def eval_step(i, ev):
        z, lab = zs[i % NBUF], labels[i % NBUF]
        if ev: ev[0].record()
        cbn, cn2 = ops.pq_prepare_codebook(codebook, "l2")                 # per step, like the module does
        idx, zq, sqerr = ops.pq_assign_gather(z, cbn, None, cn2, "l2")     # K1 + K3 fused: z is read once
        if ev: ev[1].record()
        logits = ops.probe_logits(zq, wpack, bias)
        if ev: ev[2].record()
        ops.probe_argmax_confusion(logits, B, h, w, Cp + C, lab, C, [(0, C), (Cp, C)], want_preds=False,
                                   confusions=[conf_c, conf_l])
        if ev: ev[3].record()

    alg_bytes = {
        "pq_assign_gather": 8 * N * D + 4 * N * M,        # read z, write z_q, write int32 indices
        "probe_logits": 4 * N * D + 4 * N * 56,
        "probe_argmax_confusion": 8 * P + 4 * N * 56,
    }
    lc0 = ops.launch_count()
    ms_total, evs = timed(eval_step, args.steps, args.warmup, per_step_events=4)
    launches_timed = ops.launch_count() - lc0
    if world > 1:
        dist.all_reduce(conf_c); dist.all_reduce(conf_l)                  # K10: one int64 all-reduce per compute()
    stage_ms = {name: statistics.mean(evs[i][si].elapsed_time(evs[i][si + 1]) for i in range(args.steps))
                for si, name in enumerate(stages)}
    value = world * N * args.steps / (ms_total / 1e3)
    gpu_launches = launches_timed * args.steps // (args.steps + args.warmup)
    del zs

    # ================= end-to-end: public module API, host buffers in, metrics out ============================
    torch.manual_seed(99 + rank)
    pqm = PQGOProductQuantizerWrapper(M, K, D, normalize="l2").to(dev).eval()
    pqm.materialize_prob = False
    with torch.no_grad():
        for q in pqm.quantizers:
            q.embedding.weight.copy_(torch.randn(K, d, device=dev))
    evalr = UnSegEvaluator(D, C).to(dev).eval()
    cm, lm = UnSegMetrics(C, 0, True, dev), UnSegMetrics(C, 0, False, dev)
    hz = [torch.randn(B, D, h, w).pin_memory() for _ in range(2)]
    hl = [torch.randint(-1, C, (B, H, W)).pin_memory() for _ in range(2)]
    out_host = torch.empty(2, C, C, dtype=torch.long).pin_memory()
    # Host->device copies run on a side stream one batch ahead of the compute stream (what a prefetching
    # data loader does); every step still pays its own H2D copy and D2H read inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    dz = [torch.empty(B, D, h, w, device=dev) for _ in range(2)]
    dl = [torch.empty(B, H, W, dtype=torch.long, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        k = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[k])
            dz[k].copy_(hz[k], non_blocking=True)
            dl[k].copy_(hl[k], non_blocking=True)
            ready[k].record(copy_stream)

    def e2e_run(n, reference_calls):
        for k in range(2):
            freed[k].record(torch.cuda.current_stream())
        for i in range(min(2, n)):
            prefetch(i)
        for i in range(n):
            k = i % 2
            cur = torch.cuda.current_stream()
            cur.wait_event(ready[k])
            with torch.no_grad():
                zq, _, _, _ = pqm(dz[k])
                if reference_calls:
                    # the unchanged call sequence of train.py:268-281: evaluator.forward, then two metric updates
                    _, lp, _, cp = evalr(zq, None, dl[k])
                    cm.update(cp, dl[k]); lm.update(lp, dl[k])
                else:
                    evalr.predict(zq, dl[k], cm.confusion_matrix, lm.confusion_matrix, want_preds=False)
            freed[k].record(cur)
            if i + 2 < n:
                prefetch(i + 2)
            out_host[0].copy_(cm.confusion_matrix, non_blocking=True)
            out_host[1].copy_(lm.confusion_matrix, non_blocking=True)
            cur.synchronize()                              # the step's result is on the host

    n_e2e = max(4, min(args.steps, 20))
    e2e_vals = {}
    for name, refcalls in (("fused", False), ("reference_calls", True)):
        evalr.compute_losses = refcalls
        e2e_run(3, refcalls)
        sync_all()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        e2e_run(n_e2e, refcalls)
        s1.record()
        sync_all()
        t = torch.tensor([s0.elapsed_time(s1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_vals[name] = world * N * n_e2e / (float(t.item()) / 1e3)
    e2e = {"value": e2e_vals["fused"], "unit": UNIT, "h2d_bytes_per_step": 4 * N * D + 8 * P,
           "d2h_bytes_per_step": 2 * C * C * 8, "steps": n_e2e,
           "api": "PQGOProductQuantizerWrapper.forward + UnSegEvaluator.predict (confusion accumulated inside the probe "
                  "kernel); H2D of batch i+1 overlaps compute of batch i on a copy stream",
           "reference_call_sequence": {
               "value": e2e_vals["reference_calls"], "unit": UNIT,
               "api": "the unchanged sequence of train.py:268-281 -- PQGOProductQuantizerWrapper.forward, "
                      "UnSegEvaluator.forward (both predictions + both losses), UnSegMetrics.update x2; same H2D / D2H"},
           "host_binding": numa}
    del hz, hl, dz, dl

    # ================= the other BASELINE configs, each as its own block of the same line =====================
    n_x = max(5, min(args.steps, 30))
    extras = {}
    if not args.no_extras:
        # ---- config 3: PQ training step with EMA update, data parallel, through the eager module API ------------
        pq = ProductQuantizerWrapper(M, K, D, normalize="l2").to(dev)
        pq.materialize_prob = False
        pq.train()
        with torch.no_grad():
            for q in pq.quantizers:
                q.codebook.weight.copy_(torch.randn(K, d, device=dev)); q.codebook.weight_avg.copy_(q.codebook.weight)
        ztr = [torch.randn(N, D, device=dev) for _ in range(NBUF)]

        def train_step(i, ev):
            with torch.no_grad():
                pq(ztr[i % NBUF])

        lc = ops.launch_count()
        t_tr, _ = timed(train_step, n_x, 5)
        tr_launch = (ops.launch_count() - lc) // (n_x + 5)
        packed_like = torch.zeros(M, K, d + 1, device=dev)
        t_ar = 0.0
        if world > 1:
            t_ar, _ = timed(lambda i, ev: dist.all_reduce(packed_like), 20, 5)
            t_ar /= 20
        tr_bytes = 3 * 4 * N * D + 3 * 4 * N * M
        from equss_b200 import dist_utils as _du
        peer_used = any(v is not None for v in _du._peer_exchanges.values())
        exchange = ("none (one rank)" if world == 1 else
                    "in-kernel sum of the ranks' symmetric-memory buffers over NVLink inside the EMA tail kernel (one device barrier, no collective call)"
                    if peer_used else "one NCCL all-reduce of the packed statistics")
        extras["train"] = {
            "workload": "pq_train (BASELINE configs[2]): assign + gather/loss + scatter-add + exchange of the packed EMA "
                        "statistics + EMA update + statistics, eager ProductQuantizerWrapper.forward in train() mode, flat (51200, 1024) per rank",
            "ms_per_step": t_tr / n_x, "value": world * N * n_x / (t_tr / 1e3), "unit": UNIT, "scaling": "weak",
            "launch": "eager module API (no CUDA graph)", "library_kernels_per_step": int(tr_launch),
            "exchange": exchange, "exchange_payload_bytes": 4 * M * K * (d + 1), "nccl_allreduce_ms_alone": round(t_ar, 4),
            "roofline": {"bound": "hbm", "achieved": round(tr_bytes / (t_tr / n_x * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(tr_bytes / (t_tr / n_x * 1e-3) / 1e9 / peak, 4), "algorithmic_MB": round(tr_bytes / 1e6, 1)}}
        del ztr, pq

        # ---- config 4: cityscapes high-res, 16 x 56 x 56 pixels sharded over the ranks ---------------------------
        N4, M4, K4, D4 = 16 * 56 * 56, 16, 512, 1024
        d4 = D4 // M4
        lo4 = N4 * rank // world
        hi4 = N4 * (rank + 1) // world
        n4 = hi4 - lo4
        nb4 = max(3, -(-400_000_000 // (n4 * D4 * 4)))
        z4 = [torch.randn(n4, D4, device=dev) for _ in range(nb4)]
        cb4 = F.normalize(torch.randn(M4, K4, d4, device=dev), dim=2).contiguous()
        cn24 = ops.pq_cnorm2(cb4)

        def c4_step(i, ev):
            z = z4[i % nb4]
            if ev: ev[0].record()
            idx, zq, sq = ops.pq_assign_gather(z, cb4, None, cn24, "l2")
            if ev: ev[1].record()
            ops.pq_accumulate(z, idx, K4)
            if ev: ev[2].record()

        t_c4, ev4 = timed(c4_step, n_x, 5, per_step_events=3)
        ms4 = t_c4 / n_x
        ag4 = statistics.mean(e[0].elapsed_time(e[1]) for e in ev4)
        flops4 = 2.0 * n4 * K4 * D4
        tens_pk = xpk.get("fp16_tflops_sustained") or 1405.3
        extras["c4"] = {
            "workload": "cityscapes high-res (BASELINE configs[3]): 16 x 56 x 56 = 50176 pixels, M=16 x K=512 (d=64), l2; pixels "
                        f"sharded contiguously over {world} rank(s); assign + gather/loss + scatter-add per shard (the EMA exchange is "
                        "the train block's)",
            "pixels_per_rank": n4, "ms_per_step": ms4, "value": N4 * n_x / (t_c4 / 1e3), "unit": UNIT, "scaling": "strong",
            "assign_gather_ms": round(ag4, 4), "input_buffers": nb4,
            "roofline": {"bound": "tensor", "achieved": round(flops4 / (ag4 * 1e-3) / 1e12, 1), "peak": tens_pk, "unit": "TFLOP/s",
                         "frac": round(flops4 / (ag4 * 1e-3) / 1e12 / tens_pk, 4),
                         "note": "useful flops 2*N*K*D of the distance GEMM over the assign(+gather) time; the fp16-split kernel issues 3.25x "
                                 "that on the kind::f16 pipe; peak = measured fp16 dense (profiles/r2_measured_peaks_extra.json) else bf16 sustained",
                         "hbm_frac": round((8.0 * n4 * D4 + 4 * n4 * M4) / (ag4 * 1e-3) / 1e9 / peak, 4)}}
        del z4

        # ---- config 5: global-feature kNN, queries sharded, database replicated ----------------------------------
        n5, F5, k5 = 50000, 768, 8
        g5 = torch.Generator(device=dev).manual_seed(7)
        db = F.normalize(torch.randn(n5, F5, device=dev, generator=g5), dim=1)

        def c5_step(i, ev):
            precompute_knns(db, k=k5)                           # shards the queries over the ranks + all_gather of the table

        t_c5, _ = timed(c5_step, max(3, n_x // 3), 2)
        ms5 = t_c5 / max(3, n_x // 3)
        flops5 = 2.0 * (n5 / world) * n5 * F5
        extras["c5"] = {
            "workload": f"precompute_knns (BASELINE configs[4]): 50000 x 768 unit-norm features, top-8 (self + 7), queries sharded over "
                        f"{world} rank(s), database replicated, index table all-gathered",
            "queries_per_rank": -(-n5 // world), "ms_per_step": ms5, "value": n5 / (ms5 / 1e3), "unit": "queries/s", "scaling": "strong",
            "roofline": {"bound": "tensor", "achieved": round(flops5 / (ms5 * 1e-3) / 1e12, 1), "peak": tens_pk, "unit": "TFLOP/s",
                         "frac": round(flops5 / (ms5 * 1e-3) / 1e12 / tens_pk, 4),
                         "note": "useful flops 2*nq*n*F per rank over the whole call (fp16 screening GEMM with the top-k in its "
                                 "epilogue + exact fp32 decision on the survivors + all_gather); peak = measured fp16 dense"}}
        del db

        # ---- config 1: pq_baseline ProductQuantizer forward (the reference's own CPU-runnable case) ----------------
        N1, D1, M1, K1 = 4 * 28 * 28, 512, 8, 256
        pq1 = ProductQuantizerWrapper(M1, K1, D1, normalize="l2").to(dev)
        pq1.materialize_prob = False
        with torch.no_grad():
            for q in pq1.quantizers:
                q.codebook.weight.copy_(torch.randn(K1, D1 // M1, device=dev)); q.codebook.weight_avg.copy_(q.codebook.weight)
        z1 = [torch.randn(N1, D1, device=dev) for _ in range(64)]
        res1 = {}
        for mode_name in ("train", "eval"):
            pq1.train(mode_name == "train")

            def c1_step(i, ev):
                with torch.no_grad():
                    pq1(z1[i % 64])

            t1, _ = timed(c1_step, 50, 10)
            res1[mode_name + "_us"] = round(1e3 * t1 / 50, 1)
        extras["c1"] = {"workload": "pq_baseline forward (BASELINE configs[0]): 4 x 28 x 28 = 3136 pixels, D=512, M=8 x K=256 (d=64), "
                                    "eager ProductQuantizerWrapper.forward per rank (launch-latency bound at this size)",
                        **res1, "value": world * N1 / (res1["train_us"] * 1e-6), "unit": UNIT}
        del z1

    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    if world > 1:
        sync_all()
        dist.destroy_process_group()
    if rank != 0:
        return

    dom = max(stage_ms, key=stage_ms.get)
    kern = {}
    for name in stages:
        gbs = alg_bytes[name] / (stage_ms[name] * 1e-3) / 1e9
        kern[name] = {"ms": round(stage_ms[name], 4), "algorithmic_MB": round(alg_bytes[name] / 1e6, 1),
                      "achieved_GBps": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 4)}
    roof = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_GBps"], "peak": peak, "unit": "GB/s",
            "frac": kern[dom]["frac_of_hbm_peak"], "traffic": None, "peak_source": peak_src,
            "share_of_step": round(stage_ms[dom] / sum(stage_ms.values()), 3)}
    traffic_file = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(traffic_file):
        try:
            roof["traffic"] = json.load(open(traffic_file)).get(dom)
        except Exception:
            pass

    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        n_img = 8
        cstep, px = cpu_pipeline_factory(n_img)
        ts = time_cpu(cstep, 1, 8)
        cpu_base = {"value": px * len(ts) / sum(ts), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                    "sample": f"{n_img} of {B} images ({px} pixels) x {len(ts)} timed runs of the oracle port "
                              f"(PQ loop + evaluator + 2 confusion updates), {sum(ts):.1f} s of CPU work"}
        # the same port on ONE host thread, on a smaller sample so that it stays within ~10 s
        torch.set_num_threads(1)
        cstep1, px1 = cpu_pipeline_factory(1)
        ts1 = time_cpu(cstep1, 1, 2)
        cpu_base["one_thread"] = {"value": px1 * len(ts1) / sum(ts1), "unit": UNIT, "cores": 1,
                                  "sample": f"1 of {B} images ({px1} pixels) x {len(ts1)} timed runs, {sum(ts1):.1f} s of CPU work"}
        torch.set_num_threads(os.cpu_count() or 1)

    # The same oracle code with its tensors on the GPU (library kernels, the reference's op sequence): what a user gets
    # from `.cuda()` on the reference today.  Reported next to cpu_baseline, one rank only, after every measurement of
    # the product path; a failure here (e.g. out of memory) is recorded and does not touch the rest of the line.
    eager_gpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            n_img = 8
            gstep, px = cpu_pipeline_factory(n_img, device=dev)
            with torch.no_grad():
                for _ in range(2):
                    gstep()
                torch.cuda.synchronize()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n_g = 5
                g0.record()
                for _ in range(n_g):
                    gstep()
                g1.record()
                torch.cuda.synchronize()
            ms_g = g0.elapsed_time(g1) / n_g
            eager_gpu = {"value": px / (ms_g * 1e-3), "unit": UNIT, "ms_per_step": round(ms_g, 3), "kind": "port",
                         "sample": f"{n_img} of {B} images ({px} pixels) x {n_g} timed runs of the oracle port with its tensors on "
                                   "cuda:0 (PyTorch eager, ATen / cuBLAS kernels, the reference's per-subspace loop and "
                                   "label-resolution evaluator), device-resident inputs"}
            del gstep
        except Exception as e:                                   # noqa: BLE001 -- a comparator must never cost the line
            eager_gpu = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
        torch.cuda.empty_cache()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (fp16-split / split-tf32 tensor-core contractions with exact fp32 re-score)", "data": "synthetic",
        "config": {"workload": "cocostuff27_eval (BASELINE configs[1]): PQ assign+gather, cluster+linear probe argmax, 2x 27x27 confusion",
                   **CFG, "normalize": "l2", "pixels_per_step_per_gpu": N, "parallelism": f"dp{world}",
                   "materialize_distance_prob": False,
                   "l2": f"{NBUF} rotating input sets of {(4 * N * D + 8 * P) / 1e6:.0f} MB each (> 126 MB L2), no flush kernel"},
        "roofline": roof, "kernels": kern, "cpu_baseline": cpu_base, "torch_eager_gpu": eager_gpu, "e2e": e2e,
        "gpu_launches": int(gpu_launches), "clocks": clocks, **extras,
    }
    if json_fd is not None:
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    else:
        print(json.dumps(line))
        sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="equss", choices=["equss", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the train / c4 / c5 / c1 blocks (BASELINE configs 0,2,3,4)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "equss" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_equss(args)


if __name__ == "__main__":
    main()
