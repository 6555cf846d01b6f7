#!/usr/bin/env python
"""Run one full-ranking eval configuration (for ncu / timing): python scripts/prof_eval.py R nI D precision [reps [k]]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisprrec_b200 import _lib  # noqa: E402

R, nI, D, prec = (int(x) for x in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
k = int(sys.argv[6]) if len(sys.argv) > 6 else 0
dev = torch.device('cuda')
g = torch.Generator(device=dev)
g.manual_seed(3407)
nU = max(2, R // 2)
U = torch.randn((nU, D), device=dev, generator=g) / D ** 0.5
I = torch.randn((nI, D), device=dev, generator=g)
user = torch.randint(0, nU, (R,), device=dev, generator=g)
pos = torch.randint(0, nI, (R,), device=dev, generator=g)
hl = 50
hp = torch.arange(0, (nU + 1) * hl, hl, device=dev, dtype=torch.int64)
hi = torch.sort(torch.randint(0, nI, (nU, hl), device=dev, generator=g), dim=1).values.to(torch.int32).reshape(-1).contiguous()
ws = _lib.Workspace(dev)
ms = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = _lib.eval_rank_topk(U, I, user, pos, hp, hi, ws, precision=prec, k=k)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
tf = 2.0 * R * nI * D / (min(ms) * 1e-3) / 1e12
print(f'R={R} nI={nI} D={D} precision={prec} k={k}: best {min(ms):.3f} ms  {R / (min(ms) * 1e-3):.3e} rows/s  {tf:.1f} TFLOP/s  mean rank {out[0].float().mean().item():.1f}')
