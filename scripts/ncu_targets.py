#!/usr/bin/env python
"""Launch one kernel of interest at the shape its numbers are quoted on, for `ncu --set full` captures.

    python scripts/ncu_targets.py step|epoch|adam_big|adam_marked|eval_tc64|eval_tc128|eval_x64|eval_x128|spmm_cfg4|spmm_cfg4_masked|spmm_ml1m

Each target does a couple of un-profiled warm-up launches of everything it needs and then the launches to capture
(the ncu command line selects them with -k / -s / -c; see profiles/README.md).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisprrec_b200 import _lib  # noqa: E402

dev = torch.device('cuda')
what = sys.argv[1]
g = torch.Generator(device=dev)
g.manual_seed(3407)
ws = _lib.Workspace(dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

if what in ('step', 'epoch'):
    nU, nI, D, B, N = 6040, 3706, 64, 2048, 668862
    P = torch.randn((nU + nI, D), device=dev, generator=g) * 0.1
    M, V, G = torch.zeros_like(P), torch.zeros_like(P), torch.zeros_like(P)
    ids = torch.stack([torch.randint(0, nU, (N,), device=dev, generator=g), torch.randint(0, nI, (N,), device=dev, generator=g),
                       torch.randint(1, nI, (N,), device=dev, generator=g)])
    loss = torch.zeros(400, device=dev)
    if what == 'step':
        for k in range(6):
            flush.zero_()
            _lib.bprmf_step(P, M, V, G, ids[0, k * B:(k + 1) * B], ids[1, k * B:(k + 1) * B], ids[2, k * B:(k + 1) * B], nU,
                            k + 1, 1e-3, 1e-6, loss[:1], ws)
    else:
        for k in range(3):
            flush.zero_()
            _lib.bprmf_epoch(P, M, V, G, ids, B, nU, 327 * k, 1e-3, 1e-6, loss, ws)
elif what == 'adam_big':
    nU, nI, d, b = 10_000_000, 2_000_000, 128, 65536
    P = torch.empty((nU + nI, d), device=dev).normal_(0, 0.01)
    M, V, G = torch.zeros_like(P), torch.zeros_like(P), torch.zeros_like(P)
    loss = torch.zeros(1, device=dev)
    u = torch.randint(0, nU, (b,), device=dev)
    p = torch.randint(0, nI, (b,), device=dev)
    n = torch.randint(1, nI, (b,), device=dev)
    for k in range(3):
        _lib.bpr_fwd_bwd(P[:nU], P[nU:], u, p, n, G[:nU], G[nU:], loss, ws)
        _lib.adam_l2_sweep(P, M, V, G, k + 1, 1e-3, 1e-6)
elif what == 'adam_marked':       # the row-marked step on tables far beyond L2 (wr_bprmf_step_marked)
    nU, nI, d, b = 10_000_000, 2_000_000, 128, 65536
    P = torch.empty((nU + nI, d), device=dev).normal_(0, 0.01)
    M, V, G = torch.zeros_like(P), torch.zeros_like(P), torch.zeros_like(P)
    loss = torch.zeros(1, device=dev)
    touched = _lib.row_map(nU + nI, dev)
    u = torch.randint(0, nU, (b,), device=dev)
    p = torch.randint(0, nI, (b,), device=dev)
    n = torch.randint(1, nI, (b,), device=dev)
    for k in range(3):
        _lib.bprmf_step(P, M, V, G, u, p, n, nU, k + 1, 1e-3, 1e-6, loss, ws, touched=touched)
elif what.startswith('eval_'):
    d = 128 if what.endswith('128') else 64
    nUs, nIs, Rs = 200_000, 1_000_000, 262_144
    precision = {'eval_tc64': 1, 'eval_tc128': 1, 'eval_x64': 2, 'eval_x128': 2}[what]
    Ub = torch.randn((nUs, d), device=dev, generator=g) / d ** 0.5
    Ib = torch.randn((nIs, d), device=dev, generator=g)
    us = torch.randint(0, nUs, (Rs,), device=dev, generator=g)
    ps = torch.randint(0, nIs, (Rs,), device=dev, generator=g)
    hp_ = torch.arange(0, (nUs + 1) * 50, 50, device=dev, dtype=torch.int64)
    hi_ = torch.sort(torch.randint(0, nIs, (nUs, 50), device=dev, generator=g), dim=1).values.to(torch.int32).reshape(-1).contiguous()
    for k in range(2):
        _lib.eval_rank_topk(Ub, Ib, us, ps, hp_, hi_, ws, precision=precision)
elif what in ('spmm_cfg4', 'spmm_ml1m', 'spmm_cfg4_masked'):
    from whisprrec_b200.models.general.LightGCN import build_norm_adj_device
    from whisprrec_b200.utils import synthetic
    if what in ('spmm_cfg4', 'spmm_cfg4_masked'):
        U, I, E, D = 10_000_000, 2_000_000, 500_000_000, 128
    else:
        U, I, E, D = 6040, 3706, 669_000, 64
    users, items = synthetic.power_law_pairs(U, I, E, device=dev)
    rowptr, col, val, dinv = build_norm_adj_device(U, I, users, items, ws)
    users_n = users.numel()
    if what == 'spmm_cfg4_masked':
        users_keep, items_keep = users[:4_000_000].clone(), items[:4_000_000].clone()
        users_n = 4_000_000
    del users, items
    torch.cuda.empty_cache()
    X = torch.randn((U + I, D), device=dev, generator=g) * 0.1
    Y, pool = torch.empty_like(X), torch.empty_like(X)
    plan = _lib.SpmmPlan(rowptr.cpu().numpy(), D, dev)
    if what == 'spmm_cfg4_masked':     # the first adjoint pass: X = pooled gradient of a 65,536-row batch, unmarked rows skipped
        b = 65536
        sel = torch.randint(0, users_n, (b,), device=dev, generator=g)
        bu, bp = users_keep[sel].contiguous(), items_keep[sel].contiguous()
        bn = torch.randint(1, I, (b,), device=dev, generator=g)
        x_rows = _lib.row_map(U + I, dev)
        _lib.mark_rows(bu, bp, bn, U, I, x_rows)
        X.zero_()
        X[bu] = 0.01
        X[U + bp] = 0.01
        X[U + bn] = -0.01
        for k in range(2):
            _lib.csr_spmm(rowptr, col, val, X, Y=Y, add=pool, plan=plan, x_rows=x_rows)
    else:
        for k in range(2):
            _lib.csr_spmm(rowptr, col, val, X, Y=Y, acc_in=X, acc_out=pool, plan=plan)
else:
    raise SystemExit('unknown target ' + what)
torch.cuda.synchronize()
assert ws.status() == 0
print('ok', what)
