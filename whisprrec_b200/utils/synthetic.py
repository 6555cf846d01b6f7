"""Synthetic interaction data in the reference's `.inter` schema.

`data/ml-1m/ml-1m.inter` is absent from the reference checkout (its .MISSING_LARGE_BLOBS lists it) and there is
no network, so the ml-1m configurations run on an ml-1m-SHAPED stand-in: same user / item / row counts, same
rating histogram, log-normal user activity, Zipf item popularity.  It is written in the 4-column TSV format the
reference reader parses, so the reference and this package can consume the same file.  Results on it are
labelled "ml-1m-shaped synthetic", never "ml-1m".
"""
import os

import numpy as np
import pandas as pd

ML1M = dict(n_users=6040, n_items=3706, n_rows=1_000_209, rating_p=(.056, .108, .261, .349, .226))


def ml1m_shaped(seed=3407, n_users=ML1M['n_users'], n_items=ML1M['n_items'], n_rows=ML1M['n_rows']):
    """DataFrame with the raw `.inter` columns (before the reader's filtering / id remap / split)."""
    rng = np.random.default_rng(seed)
    # rows per user: >= 20, log-normal tail (median ~96), capped by the item count, rescaled to n_rows in total
    raw = np.exp(rng.normal(np.log(96.0), 1.0, size=n_users))
    cnt = np.clip(raw, 20, min(2314, n_items)).astype(np.int64)
    for _ in range(50):
        diff = n_rows - int(cnt.sum())
        if diff == 0:
            break
        room = (cnt < min(2314, n_items)) if diff > 0 else (cnt > 20)
        idx = rng.choice(np.nonzero(room)[0], size=min(abs(diff), int(room.sum())), replace=False)
        cnt[idx] += 1 if diff > 0 else -1
    # distinct items per user with Zipf(0.9) popularity: Gumbel top-k == weighted sampling without replacement
    logp = -0.9 * np.log(np.arange(1, n_items + 1, dtype=np.float64))
    logp = logp[rng.permutation(n_items)]
    users, items = [], []
    for lo in range(0, n_users, 512):
        hi = min(n_users, lo + 512)
        key = logp[None, :] + rng.gumbel(size=(hi - lo, n_items))
        order = np.argsort(-key, axis=1)
        for r in range(hi - lo):
            c = cnt[lo + r]
            users.append(np.full(c, lo + r + 1, dtype=np.int64))      # raw ids are 1-based tokens
            items.append(order[r, :c] + 1)
    users, items = np.concatenate(users), np.concatenate(items)
    rating = rng.choice(np.arange(1, 6), size=len(users), p=ML1M['rating_p'])
    ts = 956_703_932 + np.sort(rng.integers(0, 90_000_000, size=len(users)))
    shuffle = rng.permutation(len(users))                              # file order is not grouped by user
    return pd.DataFrame({'user_id:token': users[shuffle], 'item_id:token': items[shuffle],
                         'rating:float': rating[shuffle], 'timestamp:float': ts})


def write_inter(df, path):
    """Write the 4-column TSV (`user_id:token item_id:token rating:float timestamp:float`)."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    df.to_csv(path, sep='\t', index=False)


def ml1m_shaped_corpus(seed=3407, cache_dir=None):
    """The stand-in pushed through the reader's own pipeline (filter, remap, 80/10/10 split) -> BaseReader."""
    from ..helpers import BaseReader as R
    cache = None if cache_dir is None else os.path.join(cache_dir, 'ml1m_shaped_%d.npz' % seed)
    if cache and os.path.exists(cache):
        z = np.load(cache)
        parts = [pd.DataFrame({'user_id': z[p + '_u'], 'item_id': z[p + '_i'], 'timestamp': z[p + '_t']})
                 for p in ('train', 'dev', 'test')]
    else:
        df = R.count_statics(ml1m_shaped(seed), 'ml-1m')
        parts = list(R.random_split(df))
        if cache:
            os.makedirs(cache_dir, exist_ok=True)
            np.savez(cache, **{p + s: parts[k][c].to_numpy() for k, p in enumerate(('train', 'dev', 'test'))
                               for s, c in (('_u', 'user_id'), ('_i', 'item_id'), ('_t', 'timestamp'))})
    return R.BaseReader.from_frames(*parts)


def corpus_from_npz(path):
    """BaseReader from an .npz of (user, item) pair arrays per split (`train_user`, `train_item`, `dev_*`, `test_*`,
    ids already remapped) -- the layout of tests/golden/ml100k_corpus.npz."""
    from ..helpers.BaseReader import BaseReader
    c = np.load(path, allow_pickle=False)
    frames = []
    for ph in ('train', 'dev', 'test'):
        u, i = np.asarray(c[ph + '_user'], dtype=np.int64), np.asarray(c[ph + '_item'], dtype=np.int64)
        frames.append(pd.DataFrame({'user_id': u, 'item_id': i, 'timestamp': np.zeros(len(u), dtype=np.int64)}))
    return BaseReader.from_frames(*frames)


def power_law_pairs(n_users, n_items, n_edges, seed=3407, alpha=1.8, zipf=0.8, device='cpu'):
    """Distinct (user, item) pairs of a power-law bipartite graph, as int64 torch tensors on `device`.

    User activity ~ truncated Pareto(alpha) rescaled to n_edges in total; items by Zipf(zipf) inverse CDF; ids
    are randomly permuted so that contiguous row shards are balanced.  Duplicates are dropped, so the result
    has slightly fewer than n_edges pairs.
    """
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    u01 = torch.rand(n_users, generator=g, device=device, dtype=torch.float64)
    act = (1.0 - u01).pow(-1.0 / alpha)                               # Pareto, min 1
    act = act.clamp(max=float(n_items) / 8)
    deg = torch.clamp((act * (n_edges / act.sum())).round().long(), min=1)
    users = torch.repeat_interleave(torch.arange(n_users, device=device), deg)
    # Zipf inverse CDF (continuous approximation): rank = ((1-u) * (I^(1-s) - 1) + 1)^(1/(1-s))
    u = torch.rand(users.numel(), generator=g, device=device, dtype=torch.float64)
    s = zipf
    rank = ((u * (float(n_items) ** (1 - s) - 1.0)) + 1.0).pow(1.0 / (1 - s))
    items = (rank.long() - 1).clamp(0, n_items - 1)
    items = torch.randperm(n_items, generator=g, device=device)[items]
    users = torch.randperm(n_users, generator=g, device=device)[users]
    code = torch.unique(users * n_items + items)
    return code // n_items, code % n_items
