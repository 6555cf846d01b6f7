#!/usr/bin/env python
"""LightGCN training step on a synthetic power-law bipartite graph (BASELINE.json configs[3] shape), 1..8 GPUs.

    python scripts/bench_lightgcn_scale.py --users 10000000 --items 2000000 --edges 500000000 --dim 128 --layers 3
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_lightgcn_scale.py ...

Prints one JSON line (rank 0): step time, interactions/s, algorithmic bytes and the fraction of
the measured HBM peak, per GPU and for the job.  The graph, tables and batches are generated on the device (Philox,
seed 3407); the per-step cost does not depend on the batch size, so B is a parameter.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from whisprrec_b200 import _lib  # noqa: E402
from whisprrec_b200.models.BaseModel import FusedTables  # noqa: E402
from whisprrec_b200.models.general.LightGCN import PropagationEngine, build_norm_adj_device  # noqa: E402
from whisprrec_b200.utils import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--users', type=int, default=1_000_000)
    ap.add_argument('--items', type=int, default=200_000)
    ap.add_argument('--edges', type=int, default=50_000_000)
    ap.add_argument('--dim', type=int, default=128)
    ap.add_argument('--layers', type=int, default=3)
    ap.add_argument('--batch', type=int, default=65536, help='interactions per GPU per step')
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    a = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
    torch.cuda.set_device(dev)
    peers = None
    if world > 1:
        import torch.distributed as dist
        from whisprrec_b200 import sharded as S
        dist.init_process_group('nccl', device_id=dev)
        peers = S.PeerGroup(dev)
    line = run(a, rank, world, dev, peers)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        peers.close()
        dist.destroy_process_group()


def run(a, rank, world, dev, peers=None):
    """One measurement; `a` carries users / items / edges / dim / layers / batch / steps / warmup.  world > 1 needs an
    initialised process group and a PeerGroup.  Returns the result line (every rank; rank 0 prints it)."""
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(
        os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0}
    U, I, D, L, B = a.users, a.items, a.dim, a.layers, a.batch
    t0 = time.time()
    users, items = synthetic.power_law_pairs(U, I, a.edges, device=dev)       # same seed on every rank -> same graph
    E = users.numel()
    rowptr, col, val, dinv = build_norm_adj_device(U, I, users, items)
    torch.cuda.synchronize()
    t_graph = time.time() - t0
    N, nnz = U + I, col.numel()
    g = torch.Generator(device=dev)
    g.manual_seed(3407)
    bound = (6.0 / (U + D)) ** 0.5
    lr, reg = 1e-3, 1e-5

    def batch(seed):
        gg = torch.Generator(device=dev)
        gg.manual_seed(seed)
        sel = torch.randint(0, E, (B,), device=dev, generator=gg)
        return users[sel].contiguous(), items[sel].contiguous(), torch.randint(1, I, (B,), device=dev, generator=gg)

    if world == 1:
        Uw = (torch.rand((U, D), device=dev, generator=g) * 2 - 1) * bound
        Iw = (torch.rand((I, D), device=dev, generator=g) * 2 - 1) * (6.0 / (I + D)) ** 0.5
        tabs = FusedTables(Uw, Iw)
        del Uw, Iw
        eng = PropagationEngine(tabs, rowptr, col, val, rowptr.cpu().numpy(), L, reg)
        k = [0]

        def step(b):
            k[0] += 1
            eng.fwd_bwd(b[0], b[1], b[2], tabs.loss)
            _lib.adam_l2_sweep(tabs.P, tabs.M, tabs.V, tabs.G, k[0], lr, 0.0)
        sync = torch.cuda.synchronize
        n_local, nnz_local = N, nnz
    else:
        import torch.distributed as dist
        from whisprrec_b200 import sharded as S
        lay = S.ShardLayout(U, I, world, rank)
        tabs = S.ShardedTables(peers, lay, D)
        tabs.P.copy_((torch.rand((lay.n_local, D), device=dev, generator=g) * 2 - 1) * bound)
        peers.host_sync()
        lg = S.ShardedLightGCN(tabs, rowptr, col, dinv, L, reg)
        del rowptr, col, val
        torch.cuda.empty_cache()

        def step(b):
            lg.step(b[0], b[1], b[2], B * world, lr, 0.0)

        def sync():
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
        n_local, nnz_local = lay.n_local, int(lg.rowptr[-1].item())
    batches = [batch(1000 * rank + s) for s in range(a.warmup + a.steps)]
    for s in range(a.warmup):
        step(batches[s])
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(a.steps):
        step(batches[a.warmup + s])
    e1.record()
    sync()
    ms = e0.elapsed_time(e1) / a.steps
    if world > 1:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    # one SpMM pass P = nnz (4 + 4 + 4D) + N (4 + 4D); step = 2L P + 2 (L+2) N 4D + 32 D N + 48 B D + 12 B
    # The first adjoint pass reads the pooled gradient of the batch, which is zero outside <= 3 B world rows: when the
    # step skips those rows (12 B world <= N, wr_spmm_plan.x_rows) its row reads are counted for the marked share only.
    sparse_first = L > 0 and 12 * B * world <= N
    marked_share = min(1.0, 3.0 * B * world / N) if sparse_first else 1.0
    def step_bytes(n_rows, n_nnz):
        P = n_nnz * (8 + 4 * D) + n_rows * (4 + 4 * D)
        P1 = n_nnz * (8 + 4 * D * marked_share) + n_rows * (4 + 4 * D)
        return (2 * L - 1) * P + P1 + 2 * (L + 2) * n_rows * 4 * D + 32 * D * n_rows + 48 * B * D + 12 * B
    per_gpu = step_bytes(n_local, nnz_local)
    gbs = per_gpu / (ms * 1e-3) / 1e9
    return {
        'workload': 'LightGCN L=%d D=%d on synthetic power-law graph' % (L, D), 'users': U, 'items': I, 'edges': E,
        'nodes': N, 'nnz': nnz, 'n_gpus': world, 'batch_per_gpu': B, 'ms_per_step': ms,
        'interactions_per_s': world * B / (ms * 1e-3), 'steps_per_s': 1e3 / ms,
        'algorithmic_bytes_per_gpu_step': per_gpu, 'algorithmic_gbs_per_gpu': gbs,
        'frac_of_hbm_peak': gbs / peaks['hbm_gbs'], 'hbm_peak_gbs': peaks['hbm_gbs'],
        'first_adjoint_pass': ('rows of the pooled gradient outside the batch are not fetched (marked share %.4f)' % marked_share)
        if sparse_first else 'dense',
        'compulsory_bytes_per_pass': nnz_local * 8 + n_local * (4 + 8 * D),
        'compulsory_bytes_per_gpu_step': 2 * L * (nnz_local * 8 + n_local * (4 + 8 * D)) + 2 * (L + 2) * n_local * 4 * D + 32 * D * n_local,
        'graph_build_s': t_graph, 'mem_gb': torch.cuda.max_memory_allocated() / 1e9}


if __name__ == '__main__':
    main()
