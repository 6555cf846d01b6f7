#!/usr/bin/env python
"""Train BPRMF / LightGCN to convergence on ml-100k through the reference-facing runner (BaseRunner.train: early stop,
checkpointing, final test metrics) and print what the reference prints, for comparison with its README / BASELINE.md.

    python scripts/train_ml100k.py BPRMF        # reference: test HR@10 0.2247 NDCG@10 0.1110 (lr 1e-3, l2 1e-6), 73 epochs
    python scripts/train_ml100k.py LightGCN     # reference: test HR@10 0.2292 NDCG@10 0.1174 (lr 2e-3, L=2), 114 epochs

The corpus is the reference reader's own split of ml-100k, frozen in tests/golden/ml100k_corpus.npz.
"""
import json
import logging
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests.helpers import ml100k_corpus, model_args  # noqa: E402
from whisprrec_b200.helpers.BaseRunner import BaseRunner  # noqa: E402
from whisprrec_b200.models.general.BPRMF import BPRMF  # noqa: E402
from whisprrec_b200.models.general.LightGCN import LightGCN  # noqa: E402
from whisprrec_b200.utils import utils  # noqa: E402


def train(name, eval_precision=0, verbose=False):
    cls, over = {'BPRMF': (BPRMF, dict(lr=1e-3, l2=1e-6)), 'LightGCN': (LightGCN, dict(lr=2e-3, gcn_layers=2))}[name]
    if verbose:
        logging.basicConfig(level=logging.INFO, stream=sys.stderr)
    corpus = ml100k_corpus()
    args = model_args(cls, eval_precision=eval_precision, **over)
    args.device = torch.device('cuda')
    args.model_path = os.path.join(tempfile.gettempdir(), 'wr_ml100k_%s.pt' % name)
    utils.init_seed(3407)
    model = cls(args, corpus).to(args.device)
    data = {ph: cls.Dataset(model, corpus, ph) for ph in ('train', 'dev', 'test')}
    runner = BaseRunner(args)
    t0 = time.time()
    runner.train(data)
    torch.cuda.synchronize()
    wall = time.time() - t0
    res = runner.evaluate(data['test'], [10, 20], ['NDCG', 'HR'])
    return {'model': name, 'test': {k: float(v) for k, v in res.items()}, 'train_wall_s': wall,
            'last_epoch': runner.last_epoch_stats}


if __name__ == '__main__':
    print(json.dumps(train(sys.argv[1] if len(sys.argv) > 1 else 'BPRMF', verbose=True)))
