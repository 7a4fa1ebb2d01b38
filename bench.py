#!/usr/bin/env python
"""bench.py -- EQUSS product-quantization hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl equss|reference] [--workload NAME]

Metric (BASELINE.json): PQ-quantized pixels/sec.  Default workload at every N is BASELINE configs[1]
"cocostuff27 eval shape": per rank, features (32, 1024, 40, 40) fp32 NCHW -> PQ head (M=64 subspaces x
K=256 codewords, d=16, l2) assign + gather -> cluster + linear probe argmax at 320x320 label resolution
-> two 27x27 confusion histograms.  One "step" = one such batch (51 200 PQ-quantized pixels per rank).
`--workload pq_train` times the config-3 EMA training step instead (assign + gather/loss + scatter-add +
packed NCCL all-reduce + EMA update on a flat (51200, 1024) batch per rank).

value : device-resident throughput (inputs already in HBM), CUDA events, max over ranks.
e2e   : same step through the public module API (ProductQuantizerWrapper / UnSegEvaluator) with pinned
        HOST inputs copied in, and the confusion matrices copied out, inside the timed region.
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
def cpu_pipeline_factory(n_images):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import equss_oracle as O
    c = CFG
    g = torch.Generator().manual_seed(0)
    d = c["D"] // c["M"]
    feat = torch.randn(n_images, c["D"], c["h"], c["w"], generator=g)
    label = torch.randint(-1, c["C"], (n_images, c["H"], c["W"]), generator=g)
    cb = torch.randn(c["M"], c["K"], d, generator=g)
    clusters = torch.randn(c["C"], c["D"], generator=g)
    lin_w = torch.randn(c["C"], c["D"], generator=g) * 0.03
    lin_b = torch.zeros(c["C"])

    def step():
        qs = []
        for m in range(c["M"]):   # the reference's per-subspace Python loop (model/dino_pqgo.py:757-770)
            q, _, _, _ = O.param_vq_forward(feat[:, m * d:(m + 1) * d], cb[m], normalize="l2")
            qs.append(q)
        zq = torch.cat(qs, dim=1)
        _, lp, _, cp = O.evaluator_forward(zq, label, clusters, lin_w, lin_b, c["C"])
        conf_c = O.confusion_update(torch.zeros(c["C"], c["C"], dtype=torch.long), cp, label, c["C"])
        conf_l = O.confusion_update(torch.zeros(c["C"], c["C"], dtype=torch.long), lp, label, c["C"])
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
                   "normalize": "l2", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
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

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
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


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def run_equss(args):
    import torch.distributed as dist
    import torch.nn.functional as F
    import equss_b200
    from equss_b200 import ops
    from equss_b200.codebooks import PQGOProductQuantizerWrapper
    from equss_b200.evaluator import UnSegEvaluator
    from equss_b200.metric import UnSegMetrics
    from equss_b200.quantizer import ProductQuantizerWrapper

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: equss_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
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
    train = args.workload == "pq_train"
    if train:
        zs = [torch.randn(N, D, device=dev) for _ in range(NBUF)]
    else:
        zs = [torch.randn(B, D, h, w, device=dev) for _ in range(NBUF)]
    labels = [torch.randint(-1, C, (B, H, W), device=dev) for _ in range(NBUF)]
    sampler = ClockSampler(local)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    stages, stage_ms = [], {}
    if not train:
        # ---------------- device-resident arm: the hot path as the sequence of its kernels ----------------
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
        cbn = F.normalize(codebook, dim=2).contiguous()
        cn2 = ops.pq_cnorm2(cbn)
        stages = ["pq_assign_gather", "probe_logits", "probe_argmax_confusion"]

        def step(i, ev=None):
            z, lab = zs[i % NBUF], labels[i % NBUF]
            if ev: ev[0].record()
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
        finish = (lambda: dist.all_reduce(conf_c) or dist.all_reduce(conf_l)) if world > 1 else (lambda: None)
    else:
        pq = ProductQuantizerWrapper(M, K, D, normalize="l2").to(dev)
        pq.materialize_prob = False
        pq.train()
        with torch.no_grad():
            for q in pq.quantizers:
                q.codebook.weight.copy_(torch.randn(K, d, device=dev)); q.codebook.weight_avg.copy_(q.codebook.weight)
        stages = ["pq_train_step"]
        graphs = None
        launches_per_graph = 0

        def eager_step(i):
            with torch.no_grad():
                return pq(zs[i % NBUF])

        if not args.no_graphs:
            # The step is ~10 library kernels plus a few dozen tiny torch ops (statistics on [64, 256] tensors): the
            # launches, not the GPU work, bound an eager loop.  One CUDA graph per rotating input buffer (including
            # the NCCL all-reduce of the packed EMA statistics) removes that; the module code is unchanged.
            try:
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for i in range(3):
                        eager_step(i)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                graphs = []
                lc0 = ops.launch_count()
                for k in range(NBUF):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        eager_step(k)
                    graphs.append(g)
                launches_per_graph = (ops.launch_count() - lc0) // NBUF     # library kernels each replay launches
            except Exception as e:      # capture is an optimisation, never a requirement
                print(f"bench.py: CUDA graph capture failed ({type(e).__name__}: {e}); timing the eager loop", file=sys.stderr)
                graphs = None
                torch.cuda.synchronize()

        def step(i, ev=None):
            if ev: ev[0].record()
            if graphs is not None:
                graphs[i % NBUF].replay()
            else:
                eager_step(i)
            if ev: ev[1].record()

        alg_bytes = {"pq_train_step": 3 * 4 * N * D + 3 * 4 * N * M}
        finish = lambda: None  # noqa: E731

    for i in range(args.warmup):
        step(i)
    sync_all()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    time.sleep(0.25)
    l0 = ops.launch_count()
    t_wall0 = time.time()
    sync_all()
    e0.record()
    for i in range(args.steps):
        step(i, evs[i])
    finish()
    e1.record()
    sync_all()
    t_wall1 = time.time()
    l1 = ops.launch_count()
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_total = e0.elapsed_time(e1)
    for si, name in enumerate(stages):
        stage_ms[name] = statistics.mean(evs[i][si].elapsed_time(evs[i][si + 1]) for i in range(args.steps))
    tmax = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    value = world * N * args.steps / (ms_total / 1e3)

    # ---------------- end-to-end arm: public module API, host buffers in, metrics out -------------------
    e2e = None
    if not train:
        torch.manual_seed(99 + rank)
        pqm = PQGOProductQuantizerWrapper(M, K, D, normalize="l2").to(dev).eval()
        pqm.materialize_prob = False
        with torch.no_grad():
            for q in pqm.quantizers:
                q.embedding.weight.copy_(torch.randn(K, d, device=dev))
        evalr = UnSegEvaluator(D, C).to(dev).eval()
        evalr.compute_losses = False
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

        def e2e_step(i, n_total):
            k = i % 2
            cur = torch.cuda.current_stream()
            cur.wait_event(ready[k])
            with torch.no_grad():
                zq, _, _, _ = pqm(dz[k])
                evalr.predict(zq, dl[k], cm.confusion_matrix, lm.confusion_matrix, want_preds=False)
            freed[k].record(cur)
            if i + 2 < n_total:
                prefetch(i + 2)
            out_host[0].copy_(cm.confusion_matrix, non_blocking=True)
            out_host[1].copy_(lm.confusion_matrix, non_blocking=True)
            cur.synchronize()                              # the step's result is on the host

        def e2e_run(n):
            for k in range(2):
                freed[k].record(torch.cuda.current_stream())
            for i in range(min(2, n)):
                prefetch(i)
            for i in range(n):
                e2e_step(i, n)

        n_e2e = max(4, min(args.steps, 20))
        e2e_run(3)
        sync_all()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        e2e_run(n_e2e)
        s1.record()
        sync_all()
        t = torch.tensor([s0.elapsed_time(s1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * N * n_e2e / (float(t.item()) / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": 4 * N * D + 8 * P, "d2h_bytes_per_step": 2 * C * C * 8, "steps": n_e2e,
               "api": "PQGOProductQuantizerWrapper.forward + UnSegEvaluator.predict (fused UnSegMetrics buffers); H2D of batch i+1 overlaps compute of batch i on a copy stream"}

    def finish_process():
        """Multi-rank teardown.  CUDA graphs that captured NCCL collectives keep the communicator busy: destroying the
        process group then deadlocked at 8 ranks (the JSON line was already out; torchrun never returned).  The train
        workload therefore drops the graphs, meets at a barrier and leaves without running the NCCL destructor."""
        if world <= 1:
            return
        if train:
            nonlocal graphs
            graphs = None
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush(); sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()

    if rank != 0:
        finish_process()
        return

    peak, peak_src = _peaks()
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
    if world == 1 and not args.no_cpu_baseline and not train:
        torch.set_num_threads(os.cpu_count() or 1)
        n_img = 8
        cstep, px = cpu_pipeline_factory(n_img)
        ts = time_cpu(cstep, 1, 8)
        cpu_base = {"value": px * len(ts) / sum(ts), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                    "sample": f"{n_img} of {B} images ({px} pixels) x {len(ts)} timed runs of the oracle port "
                              f"(PQ loop + evaluator + 2 confusion updates), {sum(ts):.1f} s of CPU work"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (fp16-split / split-tf32 tensor-core contractions with exact fp32 re-score)", "data": "synthetic",
        "config": {"workload": ("cocostuff27_eval (BASELINE configs[1]): PQ assign+gather, cluster+linear probe argmax, "
                                "2x 27x27 confusion" if not train else
                                "pq_train (BASELINE configs[2]): assign + gather/loss + scatter-add + packed all-reduce + EMA"),
                   **CFG, "normalize": "l2", "pixels_per_step_per_gpu": N, "parallelism": f"dp{world}",
                   "materialize_distance_prob": False,
                   **({"launch": "CUDA graph per input buffer" if graphs is not None else "eager"} if train else {}),
                   "l2": f"{NBUF} rotating input sets of {(4 * N * D + 8 * P) / 1e6:.0f} MB each (> 126 MB L2), no flush kernel"},
        "roofline": roof, "kernels": kern, "cpu_baseline": cpu_base, "e2e": e2e,
        "gpu_launches": int(l1 - l0) + (launches_per_graph * args.steps if (train and graphs is not None) else 0),
        "clocks": clocks,
    }
    if json_fd is not None:
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    else:
        print(json.dumps(line))
        sys.stdout.flush()
    finish_process()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="equss", choices=["equss", "reference"])
    ap.add_argument("--workload", default="cocostuff27_eval", choices=["cocostuff27_eval", "pq_train"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="pq_train: time the eager module loop instead of CUDA graphs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "equss" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_equss(args)


if __name__ == "__main__":
    main()
