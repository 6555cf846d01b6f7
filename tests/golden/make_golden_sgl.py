#!/usr/bin/env python
"""Golden vectors for SGL from the UNMODIFIED reference -- groundwork for the next row of the
hot-path table; the product does not implement SGL yet.

Same mechanics as make_golden.py (scratch copy, NumPy alias, the reference's own classes).  Records, for a tiny problem:
the two edge-dropout sub-graphs the reference draws with Python's `random.sample` (SGL.py:67-79, augmentor.py:77-111)
after `init_seed`, three optimiser steps (loss, gradients, parameters) and the pooled tables of the three graphs.

Usage:  python tests/golden/make_golden_sgl.py [--ref /root/reference]
"""
import argparse
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402


def dense(t):
    return (t.to_dense() if t.is_sparse else t).numpy().copy()


def sgl_cases(ref_dst, out):
    import torch
    from models.general import SGL
    from utils import utils
    store = {}
    for tag, D, extra in (('sgl_d16_l2', 16, dict(lr=5e-3, l2=0.0, gcn_layers=2, reg_weight=1e-3, ssl_tau=0.2, ssl_weight=0.1,
                                                  drop_ratio=0.2)),
                          ('sgl_d32_l3', 32, dict(lr=1e-3, l2=1e-5, gcn_layers=3, reg_weight=1e-4, ssl_tau=0.1, ssl_weight=0.05,
                                                  drop_ratio=0.1))):
        args = G.make_args(SGL.SGL, dict(embedding_size=D, batch_size=96, type='ED',
                                         model_path=os.path.join(ref_dst, tag + '.pt'), **extra))
        utils.init_seed(1234)                                   # seeds python's `random` too (utils.py:13-20)
        corpus, parts = G.fake_corpus(37, 53, 400, seed=7)
        model = SGL.SGL(args, corpus).to(args.device)
        model.graph_construction()                              # what Dataset.actions_before_epoch does (SGL.py:261-262)
        store[tag + '/train'] = parts['train'].astype(np.int32)
        store[tag + '/hp'] = np.array([args.lr, args.l2, args.reg_weight, args.gcn_layers, D, args.ssl_tau, args.ssl_weight,
                                       args.drop_ratio], dtype=np.float64)
        store[tag + '/U0'] = model.user_embedding.weight.detach().numpy().copy()
        store[tag + '/I0'] = model.item_embedding.weight.detach().numpy().copy()
        store[tag + '/graph'] = dense(model.train_graph)
        store[tag + '/sub1'] = dense(model.sub_graph1)
        store[tag + '/sub2'] = dense(model.sub_graph2)
        with torch.no_grad():
            for name, g in (('main', model.train_graph), ('sub1', model.sub_graph1), ('sub2', model.sub_graph2)):
                pu, pi = model.forward(g)
                store[f'{tag}/pooled_{name}_user'] = pu.numpy().copy()
                store[f'{tag}/pooled_{name}_item'] = pi.numpy().copy()
        opt = torch.optim.Adam(model.parameters(), lr=args.lr, weight_decay=args.l2)
        rng = np.random.RandomState(99)
        tr = parts['train']
        for step in range(3):
            sel = rng.randint(0, len(tr), size=96 if step < 2 else 41)
            user, pos = tr[sel, 0].astype(np.int64), tr[sel, 1].astype(np.int64)
            neg = rng.randint(1, 53, size=len(sel)).astype(np.int64)
            batch = {'user_id': torch.from_numpy(user), 'pos_item': torch.from_numpy(pos),
                     'neg_items': torch.from_numpy(neg), 'batch_size': len(sel), 'phase': 'train'}
            opt.zero_grad()
            loss = model.predict(batch)
            loss.backward()
            store[f'{tag}/s{step}/user'], store[f'{tag}/s{step}/pos'] = user.astype(np.int32), pos.astype(np.int32)
            store[f'{tag}/s{step}/neg'] = neg.astype(np.int32)
            store[f'{tag}/s{step}/loss'] = np.float32(loss.detach().reshape(-1)[0].item())
            store[f'{tag}/s{step}/gU'] = model.user_embedding.weight.grad.numpy().copy()
            store[f'{tag}/s{step}/gI'] = model.item_embedding.weight.grad.numpy().copy()
            opt.step()
            store[f'{tag}/s{step}/U'] = model.user_embedding.weight.detach().numpy().copy()
            store[f'{tag}/s{step}/I'] = model.item_embedding.weight.detach().numpy().copy()
        # the next epoch's sub-graphs continue python's random stream
        model.graph_construction()
        store[tag + '/sub1_epoch2_nnz'] = np.int64(np.count_nonzero(dense(model.sub_graph1)))
        store[tag + '/sub1_epoch2_sha'] = np.array(G.sha(dense(model.sub_graph1)))
        print(tag, 'losses', [float(store[f'{tag}/s{s}/loss']) for s in range(3)])
    np.savez_compressed(os.path.join(out, 'sgl_cases.npz'), **store)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ref', default='/root/reference')
    ap.add_argument('--out', default=HERE)
    a = ap.parse_args()
    out = os.path.abspath(a.out)
    dst = G.import_reference(a.ref)
    import torch
    torch.set_num_threads(1)
    sgl_cases(dst, out)
    shutil.rmtree(os.path.dirname(dst), ignore_errors=True)


if __name__ == '__main__':
    main()
