#!/usr/bin/env python
"""Where the host-side preparation of an epoch goes (ml-1m-shaped corpus): sampler, permutation, upload."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from whisprrec_b200 import _lib  # noqa: E402
from whisprrec_b200.helpers.BaseRunner import dataloader_draws  # noqa: E402
from whisprrec_b200.utils import synthetic  # noqa: E402

dev = torch.device('cuda')
corpus = synthetic.ml1m_shaped_corpus(cache_dir='/tmp/wr_cache/r0')
model, runner, data = bench.make_model(corpus, dev)
ds = data['train']
T = lambda: (torch.cuda.synchronize(), time.perf_counter())[1]
for it in range(3):
    t0 = T(); st = np.random.get_state(); t1 = T()
    cols = ds._device_cols(dev)
    neg = _lib.neg_sample_numpy_stream(cols[0], int(corpus.n_users), int(corpus.n_items), cols[2], cols[3], model.tables.ws)
    t2 = T(); host_neg = neg.cpu().numpy(); t3 = T()
    perm = dataloader_draws(len(ds), shuffle=True); t4 = T()
    p = torch.from_numpy(perm).to(dev); b = torch.stack([cols[0][p], cols[1][p], neg[p]]); t5 = T()
    model.device_sampler = False
    np.random.set_state(st); ds.actions_before_epoch(); t6 = T()
    model.device_sampler = True
    print(f'get_state {1e3 * (t1 - t0):6.2f} ms | device sampler (incl. state round trip) {1e3 * (t2 - t1):6.2f} ms | '
          f'neg D2H {1e3 * (t3 - t2):5.2f} ms | torch randperm + draws {1e3 * (t4 - t3):6.2f} ms | perm H2D + gathers '
          f'{1e3 * (t5 - t4):5.2f} ms | host sampler (NumPy path) {1e3 * (t6 - t5):7.2f} ms')
    assert (ds.data['neg_items'] == host_neg).all()
