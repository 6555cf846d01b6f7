#!/usr/bin/env python
"""SpMM time on the ml-1m-shaped adjacency for different long-row slice sizes (wr_spmm_plan)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from whisprrec_b200 import _lib  # noqa: E402
from whisprrec_b200.utils import synthetic  # noqa: E402

dev = torch.device('cuda')
corpus = synthetic.ml1m_shaped_corpus(cache_dir='/tmp/wr_cache/r0')
for D in (64, 128):
    bench.D = D
    lg, _, _ = bench.make_model(corpus, dev, 'LightGCN', gcn_layers=2)
    rowptr = lg._adj_host[0]
    N, nnz = lg.tables.P.shape[0], lg.adj_col.numel()
    Pb = nnz * (8 + 4 * D) + N * (4 + 4 * D)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for chunk in (None, 32, 64, 128, 256, 512, 1024):
        plan = None if chunk is None else _lib.SpmmPlan(rowptr, D, dev, threshold=chunk, chunk=chunk)
        fn = lambda: _lib.csr_spmm(lg.adj_rowptr, lg.adj_col, lg.adj_val, lg.tables.P, Y=lg.layer[0], plan=plan)
        fn(); torch.cuda.synchronize()
        ms = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        med = float(np.median(ms))
        print(f'D={D} slice={chunk}: {med * 1e3:7.1f} us  {Pb / (med * 1e-3) / 1e9:7.0f} GB/s algorithmic'
              f'  slices={0 if plan is None else plan.n_chunks}')
