#!/usr/bin/env python
"""SpMM at BASELINE configs[3] scale (10M x 2M x 494M edges, D=128): time of one pass against the size of the hot-row set
pinned in L2 (plan.hot_bits, evict_last) -- 0 = plain loads."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisprrec_b200 import _lib  # noqa: E402
from whisprrec_b200.models.general.LightGCN import build_norm_adj_device  # noqa: E402
from whisprrec_b200.utils import synthetic  # noqa: E402

dev = torch.device('cuda')
U, I, E, D = 10_000_000, 2_000_000, 500_000_000, 128
ws = _lib.Workspace(dev)
users, items = synthetic.power_law_pairs(U, I, E, device=dev)
rowptr, col, val, dinv = build_norm_adj_device(U, I, users, items, ws)
del users, items
torch.cuda.empty_cache()
X = torch.randn((U + I, D), device=dev) * 0.1
Y = torch.empty_like(X)
h = rowptr.cpu().numpy()
deg = np.diff(h)
out = {'nnz': int(col.numel()), 'deg_share_top': {}}
order = np.sort(deg)[::-1].astype(np.float64)
cs = np.cumsum(order) / order.sum()
for mb in (24, 40, 56, 80):
    k = (mb << 20) // (4 * D)
    out['deg_share_top'][str(mb)] = float(cs[k - 1])
for mb in (0, 24, 40, 56, 80):
    plan = _lib.SpmmPlan(h, D, dev, hot_budget_bytes=mb << 20)
    ms = []
    for r in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.csr_spmm(rowptr, col, val, X, Y=Y, plan=plan)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    out['hot_%dMB_ms' % mb] = float(np.median(ms[1:]))
    del plan
print(json.dumps(out))
json.dump(out, open(sys.argv[1], 'w'), indent=1) if len(sys.argv) > 1 else None
