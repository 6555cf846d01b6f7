#!/usr/bin/env python
"""Phase times of the two large sharded training steps (max over ranks, CUDA events), under torchrun:

    torchrun --nproc-per-node N --master-addr 127.0.0.1 scripts/prof_sharded.py [out.json] [bprmf|lightgcn|both]

BPRMF 10M x 2M, D=128, 65,536 rows per GPU: staged fwd+bwd (remote gathers + inbox writes) / barrier / inbox
reduction / Adam / barrier, plus the gather-only and local-only variants of the first kernel.
LightGCN L=3, D=128 on the 10M x 2M x 494M-edge graph: all-gather and SpMM of every layer, the batch kernels, Adam.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from whisprrec_b200 import _lib, sharded as S  # noqa: E402

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev)
peers = S.PeerGroup(dev)
what = sys.argv[2] if len(sys.argv) > 2 else 'both'
out = {'world': world}


class Phases:
    def __init__(self):
        self.names, self.ev = [], []

    def mark(self, name=None):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.ev.append(e)
        if name is not None:
            self.names.append(name)

    def ms(self):
        torch.cuda.synchronize()
        return {n: self.ev[i].elapsed_time(self.ev[i + 1]) for i, n in enumerate(self.names)}


def reduce_max(d):
    keys = sorted(d)
    t = torch.tensor([d[k] for k in keys], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {k: float(v) for k, v in zip(keys, t.tolist())}


def median_runs(fn, reps=4):
    runs = []
    for _ in range(reps):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        runs.append(reduce_max(fn()))
    return {k: float(np.median([r[k] for r in runs[1:]])) for k in runs[0]}


if what in ('both', 'bprmf'):
    nU, nI, D, B = 10_000_000, 2_000_000, 128, 65536
    lay = S.ShardLayout(nU, nI, world, rank)
    tabs = S.ShardedTables(peers, lay, D)
    tabs.P.normal_(0, 0.01)
    peers.host_sync()
    g = torch.Generator(device=dev); g.manual_seed(100 + rank)
    u = torch.randint(0, nU, (B,), device=dev, generator=g)
    p = torch.randint(0, nI, (B,), device=dev, generator=g)
    n = torch.randint(1, nI, (B,), device=dev, generator=g)
    GB = B * world
    S.bprmf_step(tabs, u, p, n, GB, 1e-3, 1e-6)

    rec = []

    def timed(name, fn):
        def w(*a_, **k_):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r_ = fn(*a_, **k_)
            e1.record()
            rec.append((name, e0, e1))
            return r_
        return w
    WRAPPED = ('allgather_shards', 'csr_spmm_sharded', 'bpr_fwd_bwd_sharded', 'bpr_fwd_bwd_sharded_staged', 'inbox_scatter',
               'embloss_sumsq_sharded', 'embloss_scatter_sharded', 'adam_l2_sweep', 'peer_barrier', 'xchg_request', 'xchg_serve',
               'bpr_fwd_bwd_exchanged', 'embloss_owner_sumsq', 'embloss_owner_scatter', 'csr_spmm_sharded_dma', 'push_shard_dma',
               'push_marked_rows', 'adam_l2_sweep_marked', 'mark_rows')
    for name in WRAPPED:
        setattr(_lib, name, timed(name, getattr(_lib, name)))

    def step_phases():
        del rec[:]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        S.bprmf_step(tabs, u, p, n, GB, 1e-3, 1e-6)
        e1.record()
        torch.cuda.synchronize()
        d = {'step': e0.elapsed_time(e1)}
        for name, a_, b_ in rec:
            d['sum_' + name] = d.get('sum_' + name, 0.0) + a_.elapsed_time(b_)
            d['n_' + name] = d.get('n_' + name, 0) + 1
        return d
    out['bprmf_10Mx2M_d128_b65536'] = median_runs(step_phases)
    # the first kernel with ids that are all LOCAL to this rank (no NVLink traffic): what the arithmetic itself costs
    ul = (u // world) * world + rank
    pl_ = (p // world) * world + rank
    nl = (n // world) * world + rank
    ul, pl_, nl = ul.clamp(max=nU - 1), pl_.clamp(max=nI - 1), nl.clamp(max=nI - 1)
    ul = torch.where(ul % world == rank, ul, torch.full_like(ul, rank))
    pl_ = torch.where(pl_ % world == rank, pl_, torch.full_like(pl_, rank))
    nl = torch.where(nl % world == rank, nl, torch.full_like(nl, rank))

    def local_only():
        ph = Phases()
        ph.mark()
        _lib.bpr_fwd_bwd_sharded(tabs.T, tabs.Gd, ul, pl_, nl, GB, D, tabs.loss_part, tabs.ws)
        ph.mark('fwd_bwd_all_rows_local')
        peers.barrier()
        tabs.G.zero_()
        ph.mark('x')
        return ph.ms()
    out['bprmf_fwd_bwd_local_ids'] = median_runs(local_only)['fwd_bwd_all_rows_local']

    def gathers_only():       # remote gathers without the gradient traffic: the row fetch of the evaluation path
        ph = Phases()
        ph.mark()
        _lib.gather_rows_sharded(tabs.T, 0, u, D, tabs.ws)
        ph.mark('gather_users_65536')
        _lib.gather_rows_sharded(tabs.T, 1, p, D, tabs.ws)
        ph.mark('gather_items_65536')
        return ph.ms()
    out['gather_rows_sharded'] = median_runs(gathers_only)
    row_bytes = B * D * 4 * (world - 1) / world
    out['gather_rows_sharded']['remote_gbs_users'] = row_bytes / (out['gather_rows_sharded']['gather_users_65536'] * 1e-3) / 1e9
    del tabs
    torch.cuda.empty_cache()

if what in ('both', 'lightgcn'):
    from whisprrec_b200.models.general.LightGCN import build_norm_adj_device
    from whisprrec_b200.utils import synthetic
    U, I, E, D, L, B = 10_000_000, 2_000_000, 500_000_000, 128, 3, 65536
    users, items = synthetic.power_law_pairs(U, I, E, device=dev)
    rowptr, col, val, dinv = build_norm_adj_device(U, I, users, items)
    nE = users.numel()
    gg = torch.Generator(device=dev); gg.manual_seed(1000 * rank)
    sel = torch.randint(0, nE, (B,), device=dev, generator=gg)
    bu, bp = users[sel].contiguous(), items[sel].contiguous()
    bn = torch.randint(1, I, (B,), device=dev, generator=gg)
    del users, items, val
    lay = S.ShardLayout(U, I, world, rank)
    tabs = S.ShardedTables(peers, lay, D)
    tabs.P.uniform_(-0.02, 0.02)
    peers.host_sync()
    lg = S.ShardedLightGCN(tabs, rowptr, col, dinv, L, 1e-5)
    del rowptr, col
    torch.cuda.empty_cache()
    lg.step(bu, bp, bn, B * world, 1e-3, 0.0)

    if 'rec' not in globals():
        rec = []

        def timed(name, fn):
            def w(*a_, **k_):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r_ = fn(*a_, **k_)
                e1.record()
                rec.append((name, e0, e1))
                return r_
            return w
        for name in ('allgather_shards', 'csr_spmm_sharded', 'bpr_fwd_bwd_sharded', 'bpr_fwd_bwd_sharded_staged', 'inbox_scatter',
                     'embloss_sumsq_sharded', 'embloss_scatter_sharded', 'adam_l2_sweep', 'peer_barrier', 'xchg_request',
                     'xchg_serve', 'bpr_fwd_bwd_exchanged', 'embloss_owner_sumsq', 'embloss_owner_scatter', 'csr_spmm_sharded_dma',
                     'push_shard_dma', 'push_marked_rows', 'adam_l2_sweep_marked', 'mark_rows'):
            setattr(_lib, name, timed(name, getattr(_lib, name)))

    def lg_phases():
        del rec[:]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lg.step(bu, bp, bn, B * world, 1e-3, 0.0)
        e1.record()
        torch.cuda.synchronize()
        d = {'step': e0.elapsed_time(e1)}
        for name, a_, b_ in rec:
            d['sum_' + name] = d.get('sum_' + name, 0.0) + a_.elapsed_time(b_)
            d['n_' + name] = d.get('n_' + name, 0) + 1
        return d
    r = median_runs(lg_phases, 3)
    r['allgather_bytes_per_rank'] = (world - 1) * lay.n_local * D * 4
    r['nnz_local'] = int(lg.rowptr[-1].item())
    out['lightgcn_cfg4'] = r

if rank == 0:
    print(json.dumps(out, indent=1))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], 'w'), indent=1)
peers.close()
dist.destroy_process_group()
