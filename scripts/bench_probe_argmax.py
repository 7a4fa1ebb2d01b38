"""Time and cross-check the three argmax formulations of the label-resolution probe kernel on the cocostuff27 eval
shape (EQUSS_PROBE_ARGMAX_T = 0 sequential kernel, 1 = default: persistent tournament kernel).

    gpurun --timeout 300 -- 'python scripts/bench_probe_argmax.py > gpurun_out/probe_argmax_modes.txt 2>&1'

Every mode must produce bit-identical predictions and confusion matrices (random logits, tied logits, NaN rows)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from equss_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, D, h, w, H, W, C = 32, 1024, 40, 40, 320, 320, 27
CT = 28 + C
label = torch.randint(-1, C, (B, H, W), device=dev)
heads = [(0, C), (28, C)]


def run(mode, logits, want_preds, B_=B, h_=h, w_=w, label_=label, heads_=heads, ct=CT, C_=C):
    os.environ["EQUSS_PROBE_ARGMAX_T"] = str(mode)
    confs = [torch.zeros(C_, C_, dtype=torch.long, device=dev) for _ in heads_]
    preds = ops.probe_argmax_confusion(logits, B_, h_, w_, ct, label_, C_, heads_, want_preds=want_preds, confusions=confs)
    return preds, confs


def timeit(mode, logits, want_preds, iters=40):
    os.environ["EQUSS_PROBE_ARGMAX_T"] = str(mode)
    confs = [torch.zeros(C, C, dtype=torch.long, device=dev) for _ in heads]
    call = lambda: ops.probe_argmax_confusion(logits, B, h, w, CT, label, C, heads, want_preds=want_preds, confusions=confs)
    for _ in range(5):
        call()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    tot = 0.0
    for _ in range(iters):
        for _ in range(3):
            flush.zero_()          # > L2, and long enough for the host to run ahead of the device
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


feat = torch.randn(B, D, h, w, device=dev)
wmat = torch.randn(CT, D, device=dev) / 32
bias = torch.randn(CT, device=dev)
logits = ops.probe_logits(feat, ops.probe_pack(wmat), bias)
cases = {"random": logits}
tied = (torch.randint(-2, 3, logits.shape, device=dev)).float()          # many exact ties: first index must win
cases["tied"] = tied
nanl = logits.clone()
nanl.view(-1)[torch.randint(0, nanl.numel(), (200000,), device=dev)] = float("nan")
nanl.view(B * h * w, -1)[::97] = float("nan")                            # whole tokens NaN
nanl.view(B * h * w, -1)[5::101] = float("-inf")
cases["nan_inf"] = nanl
ok = True
for name, lg in cases.items():
    ref_p, ref_c = run(0, lg, True)
    for mode in (1,):
        p, c = run(mode, lg, True)
        same = all(torch.equal(a, b) for a, b in zip(p, ref_p)) and all(torch.equal(a, b) for a, b in zip(c, ref_c))
        print(f"case {name}: mode {mode} identical to mode 0: {same}")
        ok &= same
# other channel counts / ragged shapes (every instantiation of the kernel)
for cnt in (3, 4, 7, 10, 16, 19, 21, 27, 30, 32):
    cm = (cnt + 3) & ~3
    ct = 2 * cm
    Bs, hs, ws, Hs, Ws = 3, 9, 11, 70, 83
    lg = torch.randn(Bs * hs * ws, ct, device=dev).round(decimals=1)
    lab = torch.randint(-1, cnt, (Bs, Hs, Ws), device=dev)
    hd = [(0, cnt), (cm, cnt)]
    ref_p, ref_c = run(0, lg, True, Bs, hs, ws, lab, hd, ct, cnt)
    for mode in (1,):
        p, c = run(mode, lg, True, Bs, hs, ws, lab, hd, ct, cnt)
        same = all(torch.equal(a, b) for a, b in zip(p, ref_p)) and all(torch.equal(a, b) for a, b in zip(c, ref_c))
        if not same:
            print(f"channels {cnt}: mode {mode} DIFFERS")
        ok &= same
print("all identical:", ok)
for want_preds in (False, True):
    for mode in (0, 1, 0, 1):
        print(f"want_preds={want_preds} mode {mode}: {timeit(mode, logits, want_preds):.1f} us")
