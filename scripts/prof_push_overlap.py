#!/usr/bin/env python
"""Does a copy-engine push of a finished layer overlap a running SpMM?  (torchrun, cfg4 graph)

Times on every rank (max over ranks): the SpMM alone, the SpMM with the fused store-push epilogue, the 7 peer copies of
one shard by cudaMemcpyAsync alone, and the SpMM with those copies running beside it on a second stream.
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from whisprrec_b200 import _lib, sharded as S  # noqa: E402
from whisprrec_b200.models.general.LightGCN import build_norm_adj_device  # noqa: E402
from whisprrec_b200.utils import synthetic  # noqa: E402

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev)
peers = S.PeerGroup(dev)
cudart = ctypes.CDLL('libcudart.so.12')
cudart.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]

scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
U, I, E, D, L = int(10_000_000 * scale), int(2_000_000 * scale), int(500_000_000 * scale), 128, 3
users, items = synthetic.power_law_pairs(U, I, E, device=dev)
rowptr, col, val, dinv = build_norm_adj_device(U, I, users, items)
del users, items, val
lay = S.ShardLayout(U, I, world, rank)
tabs = S.ShardedTables(peers, lay, D)
tabs.P.uniform_(-0.02, 0.02)
peers.host_sync()
lg = S.ShardedLightGCN(tabs, rowptr, col, dinv, L, 1e-5)
del rowptr, col
torch.cuda.empty_cache()
assert lg.gather_first
_lib.allgather_shards(tabs.T, lg.gathered[0], D)
peers.barrier()
X = lg._local_view(tabs.T, 0)
y, ys = lg.layer[0]
other = lg.layer[1][0]
step = lg.gathered[0][0].numel() * 4
push = [None if g == rank else lg.gathered_ptrs[1][g] + rank * step for g in range(world)]
side = torch.cuda.Stream()
n_slices = int(sys.argv[3]) if len(sys.argv) > 3 else 1


def spmm(push_ptrs=None):
    _lib.csr_spmm_sharded(lg.rowptr, lg.col, lg.val, lay.n_local, D, X, plan=lg.plan, push_ptrs=push_ptrs, Y=y)


def dma(src, stream):
    nb = src.numel() * 4 // n_slices // 16 * 16
    for s in range(n_slices):
        for k in range(1, world):
            g = (rank + k) % world
            rc = cudart.cudaMemcpyAsync(push[g] + s * nb, src.data_ptr() + s * nb, nb, 4, stream.cuda_stream)
            assert rc == 0, rc


def timed(fn, reps=4):
    out = []
    for _ in range(reps):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out.append(float(t))
    return float(np.median(out[1:]))


def overlapped():
    cur = torch.cuda.current_stream()
    side.wait_stream(cur)
    spmm()
    dma(other, side)
    cur.wait_stream(side)


def dma_only():
    cur = torch.cuda.current_stream()
    side.wait_stream(cur)
    dma(other, side)
    cur.wait_stream(side)


epoch = [0]


def spmm_dma():
    epoch[0] += 1
    _lib.csr_spmm_sharded_dma(lg.rowptr, lg.col, lg.val, lay.n_local, D, X, y, push, lg._progress, lg._block_rows, epoch[0],
                              side, nnz=lg.nnz, plan=lg.plan)


lg._progress.zero_()
res = {'world': world, 'n_chunks': int(lg.plan.n_chunks), 'nnz_local': int(lg.nnz), 'block_rows': lg._block_rows,
       'spmm_copy_engine_push_ms': timed(spmm_dma), 'n_local': lay.n_local, 'shard_mb': lay.n_local * D * 4 / 1e6,
       'spmm_ms': timed(spmm), 'spmm_fused_push_ms': timed(lambda: spmm(push)),
       'dma_push_only_ms': timed(dma_only), 'spmm_beside_dma_push_ms': timed(overlapped), 'dma_slices': n_slices}
res['dma_push_gbs_out'] = (world - 1) * lay.n_local * D * 4 / (res['dma_push_only_ms'] * 1e-3) / 1e9
if rank == 0:
    print(json.dumps(res, indent=1))
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], 'w'), indent=1)
peers.close()
dist.destroy_process_group()
