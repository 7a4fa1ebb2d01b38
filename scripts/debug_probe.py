import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from equss_b200 import ops
dev = torch.device("cuda:0")
B, D, h, w, Ct = 1, 32, 8, 16, 8
wmat = torch.zeros(Ct, D, device=dev)
for j in range(Ct):
    wmat[j] = torch.arange(D, device=dev).float() + 100 * (j + 1)      # w[j][c] = 100(j+1) + c
pack = ops.probe_pack(wmat)
for (c0, s0) in [(0, 0), (1, 0), (0, 1), (5, 37), (9, 64), (31, 127), (8, 3), (16, 33)]:
    feat = torch.zeros(B, D, h, w, device=dev)
    feat.view(B, D, -1)[0, c0, s0] = 1.0
    lt = ops.probe_logits(feat, pack, None, algo=0)
    nz = lt.nonzero()
    rows = sorted(set(nz[:, 0].tolist()))
    print(f"c0={c0} s0={s0}: nonzero rows {rows[:8]} vals {lt[rows[0], :4].tolist() if rows else None}  expect row {s0} vals {[100*(j+1)+c0 for j in range(4)]}")
