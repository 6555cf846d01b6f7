#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Runs only in the authoring container (needs /root/reference).  The reference is
Python, so it cannot travel to the GPU box; its outputs do, as the small .npz
files this script writes.  Nothing here is product code and nothing in the
product imports it.

How the reference is driven:
  * a writable scratch copy (the reader/runner write ../data, ../log, ../model)
  * `np.float_ = np.float64` (reference src/utils/utils.py:66 predates NumPy 2)
  * `sys.path.insert(0, <copy>/src)` then the reference's own classes are
    imported and called: BaseReader, BPRMF, LightGCN, BaseRunner.fit /
    interface / evaluate_method.  No reference source is edited or copied
    into this repository.

Usage:  python tests/golden/make_golden.py [--ref /root/reference]
"""
import argparse
import hashlib
import os
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def import_reference(ref_root):
    scratch = tempfile.mkdtemp(prefix='wr_ref_')
    dst = os.path.join(scratch, 'ref')
    shutil.copytree(ref_root, dst)
    for r, ds, fs in os.walk(dst):
        for n in ds + fs:
            os.chmod(os.path.join(r, n), 0o755)
    if not hasattr(np, 'float_'):
        np.float_ = np.float64
    src = os.path.join(dst, 'src')
    sys.path.insert(0, src)
    os.chdir(src)
    return dst


def make_args(model_mod, overrides):
    from helpers import BaseReader, BaseRunner
    p = argparse.ArgumentParser()
    p = BaseReader.BaseReader.parse_reader_args(p)
    p = BaseRunner.BaseRunner.parse_runner_args(p)
    p = model_mod.parse_model_args(p)
    args, _ = p.parse_known_args([])
    import torch
    args.device = torch.device('cpu')
    args.num_workers = 0          # batch order does not depend on workers
    for k, v in overrides.items():
        setattr(args, k, v)
    return args


def record_fit(runner, model, ds, n_keep_batches=3):
    """Run the reference's own BaseRunner.fit once, recording what predict saw."""
    losses, batches = [], []
    orig = model.predict

    def spy(batch):
        if len(batches) < n_keep_batches:
            batches.append({k: v.clone().numpy() for k, v in batch.items() if hasattr(v, 'clone')})
        loss = orig(batch)
        losses.append(float(loss.detach().reshape(-1)[0]))
        return loss

    model.predict = spy
    mean_loss = runner.fit(ds, epoch=1)
    model.predict = orig
    return mean_loss, np.array(losses, dtype=np.float64), batches


def ranks_from_predictions(pred):
    sort_idx = (-pred).argsort(axis=1)
    return (np.argwhere(sort_idx == 0)[:, 1] + 1).astype(np.int32)


def ml100k(ref_dst, out):
    import torch
    from helpers import BaseReader, BaseRunner
    from models.general import BPRMF, LightGCN
    from utils import utils

    # ---------------- BPRMF, configs[0]: emb 64, lr 1e-3, l2 1e-6 -----------------
    args = make_args(BPRMF.BPRMF, dict(lr=1e-3, l2=1e-6, model_path=os.path.join(ref_dst, 'bprmf.pt')))
    utils.init_seed(3407)                                   # main.py:44
    corpus = BaseReader.BaseReader(args)                    # main.py:61
    model = BPRMF.BPRMF(args, corpus).to(args.device)       # main.py:66
    data = {ph: BPRMF.BPRMF.Dataset(model, corpus, ph) for ph in ('train', 'dev', 'test')}
    runner = BaseRunner.BaseRunner(args)

    cor = {}
    for ph in ('train', 'dev', 'test'):
        df = corpus.data_df[ph]
        cor[ph + '_user'] = df['user_id'].to_numpy().astype(np.int16)
        cor[ph + '_item'] = df['item_id'].to_numpy().astype(np.int16)
    np.savez_compressed(os.path.join(out, 'ml100k_corpus.npz'),
                        n_users=np.int64(corpus.n_users), n_items=np.int64(corpus.n_items), **cor)

    init_u = model.user_embeddings.weight.detach().numpy().copy()
    init_i = model.item_embeddings.weight.detach().numpy().copy()

    mean_loss, losses, batches = record_fit(runner, model, data['train'])
    neg1 = data['train'].data['neg_items'].copy()
    after_u = model.user_embeddings.weight.detach().numpy().copy()
    after_i = model.item_embeddings.weight.detach().numpy().copy()
    # dev eval after epoch 1 through the reference's interface()/evaluate_method()
    pred = runner.interface(data['dev'])
    dev_rank = ranks_from_predictions(pred)
    dev_res = BaseRunner.BaseRunner.evaluate_method(pred, [10, 20], ['NDCG', 'HR'])
    dev_target = pred[:, 0].astype(np.float32)
    # epoch 2 negatives (stream continues) -- only their hash
    data['train'].actions_before_epoch()
    neg2 = data['train'].data['neg_items'].copy()

    np.savez_compressed(
        os.path.join(out, 'ml100k_bprmf.npz'),
        init_user_head=init_u[:8], init_item_head=init_i[:8],
        init_user_sha=sha(init_u), init_item_sha=sha(init_i),
        neg_epoch1=neg1.astype(np.int16), neg_epoch2_sha=sha(neg2.astype(np.int64)),
        neg_epoch2_head=neg2[:64].astype(np.int16),
        batch0_user=batches[0]['user_id'].astype(np.int16), batch0_pos=batches[0]['pos_item'].astype(np.int16),
        batch0_neg=batches[0]['neg_items'].astype(np.int16),
        batch1_user=batches[1]['user_id'].astype(np.int16), batch1_pos=batches[1]['pos_item'].astype(np.int16),
        batch1_neg=batches[1]['neg_items'].astype(np.int16),
        step_losses=losses, epoch_mean_loss=np.float64(mean_loss),
        after_user_rows=after_u[:64], after_item_rows=after_i[:64],
        after_user_norm=np.float64(np.linalg.norm(after_u.astype(np.float64))),
        after_item_norm=np.float64(np.linalg.norm(after_i.astype(np.float64))),
        dev_rank=dev_rank.astype(np.int16), dev_target=dev_target,
        dev_metric_keys=np.array(sorted(dev_res.keys())),
        dev_metric_vals=np.array([dev_res[k] for k in sorted(dev_res.keys())], dtype=np.float64),
    )
    print('BPRMF epoch1 mean loss', mean_loss, 'first losses', losses[:3], 'dev', dev_res)

    # ---------------- LightGCN, lr 2e-3, gcn_layers 2 (README.md:41 setting) --------
    args = make_args(LightGCN.LightGCN, dict(lr=2e-3, gcn_layers=2, model_path=os.path.join(ref_dst, 'lgcn.pt')))
    utils.init_seed(3407)
    corpus = BaseReader.BaseReader(args)
    model = LightGCN.LightGCN(args, corpus).to(args.device)
    data = {ph: LightGCN.LightGCN.Dataset(model, corpus, ph) for ph in ('train', 'dev', 'test')}
    runner = BaseRunner.BaseRunner(args)
    adj = model.norm_adj.numpy()                            # dense [N,N] as shipped (LightGCN.py:119-121)
    r, c = np.nonzero(adj)
    vals = adj[r, c].astype(np.float32)
    init_u = model.user_embedding.weight.detach().numpy().copy()
    init_i = model.item_embedding.weight.detach().numpy().copy()
    with torch.no_grad():
        pu, pi = model.forward()
    mean_loss, losses, batches = record_fit(runner, model, data['train'])
    after_u = model.user_embedding.weight.detach().numpy().copy()
    after_i = model.item_embedding.weight.detach().numpy().copy()
    pred = runner.interface(data['dev'])
    dev_rank = ranks_from_predictions(pred)
    dev_res = BaseRunner.BaseRunner.evaluate_method(pred, [10, 20], ['NDCG', 'HR'])
    np.savez_compressed(
        os.path.join(out, 'ml100k_lightgcn.npz'),
        adj_nnz=np.int64(len(vals)), adj_row_sha=sha(r.astype(np.int64)), adj_col_sha=sha(c.astype(np.int64)),
        adj_val_sha=sha(vals), adj_row_head=r[:512].astype(np.int32), adj_col_head=c[:512].astype(np.int32),
        adj_val_head=vals[:512],
        init_user_head=init_u[:8], init_item_head=init_i[:8], init_user_sha=sha(init_u), init_item_sha=sha(init_i),
        pooled_user_rows=pu.numpy()[:64].copy(), pooled_item_rows=pi.numpy()[:64].copy(),
        pooled_user_norm=np.float64(np.linalg.norm(pu.numpy().astype(np.float64))),
        pooled_item_norm=np.float64(np.linalg.norm(pi.numpy().astype(np.float64))),
        step_losses=losses, epoch_mean_loss=np.float64(mean_loss),
        after_user_rows=after_u[:64], after_item_rows=after_i[:64],
        after_user_norm=np.float64(np.linalg.norm(after_u.astype(np.float64))),
        after_item_norm=np.float64(np.linalg.norm(after_i.astype(np.float64))),
        dev_rank=dev_rank.astype(np.int16),
        dev_metric_keys=np.array(sorted(dev_res.keys())),
        dev_metric_vals=np.array([dev_res[k] for k in sorted(dev_res.keys())], dtype=np.float64),
    )
    print('LightGCN epoch1 mean loss', mean_loss, 'first losses', losses[:3], 'dev', dev_res)


def fake_corpus(n_users, n_items, n_inter, seed):
    """A tiny corpus object with the attributes the reference models read."""
    import pandas as pd
    rng = np.random.RandomState(seed)
    pairs = set()
    while len(pairs) < n_inter:
        # skewed so that batches contain many duplicate users and items
        u = int(n_users * rng.rand() ** 2)
        i = int(n_items * rng.rand() ** 1.5)
        pairs.add((u, i))
    pairs = np.array(sorted(pairs))
    rng.shuffle(pairs)
    n_tr = int(0.8 * n_inter)
    n_dev = (n_inter - n_tr) // 2
    parts = {'train': pairs[:n_tr], 'dev': pairs[n_tr:n_tr + n_dev], 'test': pairs[n_tr + n_dev:]}
    c = types.SimpleNamespace()
    c.n_users, c.n_items = np.int64(n_users), np.int64(n_items)
    c.data_df = {k: pd.DataFrame({'user_id': v[:, 0], 'item_id': v[:, 1], 'timestamp': np.arange(len(v))})
                 for k, v in parts.items()}
    c.train_clicked_set = {u: set() for u in range(n_users)}
    c.residual_clicked_set = {u: set() for u in range(n_users)}
    for k, v in parts.items():
        for u, i in v:
            (c.train_clicked_set if k == 'train' else c.residual_clicked_set)[int(u)].add(int(i))
    return c, parts


def small_cases(ref_dst, out):
    """Every tensor of a few reference steps on tiny problems: pins the oracle element-wise."""
    import torch
    from helpers import BaseRunner
    from models.general import BPRMF, LightGCN
    from utils import utils
    store = {}
    for tag, mod, cls, D, extra in (
            ('bprmf_d16', BPRMF, 'BPRMF', 16, dict(lr=1e-2, l2=1e-3)),
            ('bprmf_d64', BPRMF, 'BPRMF', 64, dict(lr=1e-3, l2=1e-6)),
            ('lgcn_d16_l2', LightGCN, 'LightGCN', 16, dict(lr=5e-3, l2=1e-4, gcn_layers=2, reg_weight=1e-3)),
            ('lgcn_d64_l3', LightGCN, 'LightGCN', 64, dict(lr=2e-3, l2=0.0, gcn_layers=3, reg_weight=1e-5)),
    ):
        klass = getattr(mod, cls)
        args = make_args(klass, dict(embedding_size=D, batch_size=96, model_path=os.path.join(ref_dst, tag + '.pt'),
                                     **extra))
        utils.init_seed(1234)
        corpus, parts = fake_corpus(37, 53, 400, seed=7)
        model = klass(args, corpus).to(args.device)
        is_gcn = cls == 'LightGCN'
        ue = model.user_embedding if is_gcn else model.user_embeddings
        ie = model.item_embedding if is_gcn else model.item_embeddings
        store[tag + '/train'] = parts['train'].astype(np.int32)
        store[tag + '/dev'] = parts['dev'].astype(np.int32)
        store[tag + '/test'] = parts['test'].astype(np.int32)
        store[tag + '/hp'] = np.array([args.lr, args.l2, getattr(args, 'reg_weight', 0.0),
                                       getattr(args, 'gcn_layers', 0), D], dtype=np.float64)
        store[tag + '/U0'] = ue.weight.detach().numpy().copy()
        store[tag + '/I0'] = ie.weight.detach().numpy().copy()
        if is_gcn:
            adj = model.norm_adj.numpy()
            store[tag + '/adj_dense'] = adj.copy()
            with torch.no_grad():
                pu, pi = model.forward()
            store[tag + '/pooled_user0'] = pu.numpy().copy()
            store[tag + '/pooled_item0'] = pi.numpy().copy()
        opt = torch.optim.Adam(model.parameters(), lr=args.lr, weight_decay=args.l2)   # BaseRunner.py:120-124
        rng = np.random.RandomState(99)
        tr = parts['train']
        for step in range(3):
            sel = rng.randint(0, len(tr), size=96 if step < 2 else 41)   # duplicates on purpose; ragged last batch
            user = tr[sel, 0].astype(np.int64)
            pos = tr[sel, 1].astype(np.int64)
            neg = rng.randint(1, 53, size=len(sel)).astype(np.int64)
            batch = {'user_id': torch.from_numpy(user), 'pos_item': torch.from_numpy(pos),
                     'neg_items': torch.from_numpy(neg), 'batch_size': len(sel), 'phase': 'train'}
            opt.zero_grad()
            loss = model.predict(batch)
            loss.backward()
            store[f'{tag}/s{step}/user'] = user.astype(np.int32)
            store[f'{tag}/s{step}/pos'] = pos.astype(np.int32)
            store[f'{tag}/s{step}/neg'] = neg.astype(np.int32)
            store[f'{tag}/s{step}/loss'] = np.float32(loss.detach().reshape(-1)[0].item())
            store[f'{tag}/s{step}/gU'] = ue.weight.grad.numpy().copy()
            store[f'{tag}/s{step}/gI'] = ie.weight.grad.numpy().copy()
            opt.step()
            store[f'{tag}/s{step}/U'] = ue.weight.detach().numpy().copy()
            store[f'{tag}/s{step}/I'] = ie.weight.detach().numpy().copy()
        st = opt.state[ue.weight]
        store[tag + '/mU'] = st['exp_avg'].numpy().copy()
        store[tag + '/vU'] = st['exp_avg_sq'].numpy().copy()
        # full-ranking eval through the reference's interface()/evaluate_method()
        ds = klass.Dataset(model, corpus, 'test')
        args.eval_batch_size = 16
        runner = BaseRunner.BaseRunner(args)
        pred = runner.interface(ds)
        store[tag + '/eval_pred'] = pred.astype(np.float32)
        store[tag + '/eval_rank'] = ranks_from_predictions(pred)
        res = BaseRunner.BaseRunner.evaluate_method(pred, [5, 10, 20], ['NDCG', 'HR', 'RECALL', 'PRECISION'])
        keys = sorted(res.keys())
        store[tag + '/eval_keys'] = np.array(keys)
        store[tag + '/eval_vals'] = np.array([res[k] for k in keys], dtype=np.float64)
        print(tag, 'losses', [float(store[f'{tag}/s{s}/loss']) for s in range(3)], 'HR@10', res['HR@10'])
    np.savez_compressed(os.path.join(out, 'small_cases.npz'), **store)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ref', default='/root/reference')
    ap.add_argument('--out', default=HERE)
    a = ap.parse_args()
    out = os.path.abspath(a.out)
    dst = import_reference(a.ref)
    import torch
    torch.set_num_threads(1)      # fixed reduction order inside ATen for the recorded numbers
    small_cases(dst, out)
    ml100k(dst, out)
    shutil.rmtree(os.path.dirname(dst), ignore_errors=True)


if __name__ == '__main__':
    main()
