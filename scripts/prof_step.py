#!/usr/bin/env python
"""Time the single-launch BPRMF step on the bench shape, next to the event-timing floor (a 1-CTA kernel)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisprrec_b200 import _lib  # noqa: E402

nU, nI, D, B = 6040, 3706, 64, 2048
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
flush_on = (sys.argv[2] != 'noflush') if len(sys.argv) > 2 else True
dev = torch.device('cuda')
P = torch.randn((nU + nI, D), device=dev) * 0.1
M, V, G = torch.zeros_like(P), torch.zeros_like(P), torch.zeros_like(P)
ws, loss = _lib.Workspace(dev), torch.zeros(1, device=dev)
u = torch.randint(0, nU, (B,), device=dev)
p = torch.randint(0, nI, (B,), device=dev)
n = torch.randint(1, nI, (B,), device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
tiny = [torch.zeros(4, device=dev) for _ in range(4)]


def timed(fn):
    ms = []
    for _ in range(reps):
        if flush_on:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return np.median(ms) * 1e3, np.mean(ms) * 1e3


k = [0]
def step():
    k[0] += 1
    _lib.bprmf_step(P, M, V, G, u, p, n, nU, k[0], 1e-3, 1e-6, loss, ws)
def two():
    k[0] += 1
    _lib.bpr_fwd_bwd(P[:nU], P[nU:], u, p, n, G[:nU], G[nU:], loss, ws)
    _lib.adam_l2_sweep(P, M, V, G, k[0], 1e-3, 1e-6)
for name, fn in [('floor (4-element adam)', lambda: _lib.adam_l2_sweep(*tiny, 1, 1e-3, 0.0)),
                 ('bprmf_step (1 launch)', step), ('bpr_fwd_bwd + adam_l2_sweep', two),
                 ('bpr_fwd_bwd', lambda: _lib.bpr_fwd_bwd(P[:nU], P[nU:], u, p, n, G[:nU], G[nU:], loss, ws)),
                 ('adam_l2_sweep', lambda: _lib.adam_l2_sweep(P, M, V, G, 5, 1e-3, 1e-6))]:
    fn(); torch.cuda.synchronize()
    med, mean = timed(fn)
    print(f'{name:32s} median {med:7.2f} us   mean {mean:7.2f} us   flush={flush_on}')
