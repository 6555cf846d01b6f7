#!/usr/bin/env python
"""How much do duplicate ids cost the gradient scatter of wr_bpr_fwd_bwd (one red.global.add.v4.f32 per 16 bytes, no
pre-aggregation)?  Same kernel, same batch size, four id distributions."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisprrec_b200 import _lib  # noqa: E402

dev = torch.device('cuda')
D = 64
ws, loss = _lib.Workspace(dev), torch.zeros(1, device=dev)
g = torch.Generator(device=dev); g.manual_seed(1)


def run(name, nU, nI, B, user, pos, neg):
    P = torch.randn((nU + nI, D), device=dev) * 0.1
    G = torch.zeros_like(P)
    fn = lambda: _lib.bpr_fwd_bwd(P[:nU], P[nU:], user, pos, neg, G[:nU], G[nU:], loss, ws)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    uniq = len(torch.unique(torch.cat([pos, neg])))
    print(f'{name:46s} B={B:6d} distinct users {len(torch.unique(user)):6d} items {uniq:6d}  median {np.median(ms) * 1e3:8.1f} us')


for B in (2048, 65536):
    nU, nI = 1_000_000, 1_000_000
    run('large tables, uniform ids (almost no duplicates)', nU, nI, B, torch.randint(0, nU, (B,), device=dev, generator=g),
        torch.randint(0, nI, (B,), device=dev, generator=g), torch.randint(1, nI, (B,), device=dev, generator=g))
    nU, nI = 6040, 3706
    u = torch.randint(0, nU, (B,), device=dev, generator=g)
    run('ml-1m-sized tables, uniform ids', nU, nI, B, u, torch.randint(0, nI, (B,), device=dev, generator=g),
        torch.randint(1, nI, (B,), device=dev, generator=g))
    z = (torch.rand(B, device=dev, generator=g) ** 4 * nI).long().clamp(0, nI - 1)      # heavy head: popular items
    run('ml-1m-sized tables, skewed positives (u^4)', nU, nI, B, u, z, torch.randint(1, nI, (B,), device=dev, generator=g))
    run('one user, one positive for the whole batch', nU, nI, B, torch.zeros(B, dtype=torch.int64, device=dev),
        torch.full((B,), 7, dtype=torch.int64, device=dev), torch.randint(1, nI, (B,), device=dev, generator=g))
