"""Shared helpers for the tests: golden loading and the tolerance the parity bar uses."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# north_star: losses / embeddings / metrics within 1e-5 relative in fp32.  Element-wise relative
# error is meaningless on near-zero elements (hard part 4), so the bound is
# |a-b| <= RTOL*|b| + ATOL_SCALE*max|b|.
RTOL = 1e-5
ATOL_SCALE = 1e-6


def close(a, b, rtol=RTOL, atol_scale=ATOL_SCALE):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.max(np.abs(b)) if b.size else 0.0
    return np.abs(a - b) <= rtol * np.abs(b) + atol_scale * scale + 1e-30


def assert_close(a, b, what='', rtol=RTOL, atol_scale=ATOL_SCALE):
    ok = close(a, b, rtol, atol_scale)
    if not np.all(ok):
        a64, b64 = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        err = np.abs(a64 - b64)
        raise AssertionError(f'{what}: {np.count_nonzero(~ok)} of {ok.size} elements out of tolerance; '
                             f'max abs err {err.max():.3e}, max |ref| {np.abs(b64).max():.3e}')


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def small_case(store, tag):
    """dict view of one small case in small_cases.npz"""
    pre = tag + '/'
    return {k[len(pre):]: store[k] for k in store.files if k.startswith(pre)}


def clicked_sets(n_users, pairs):
    s = {u: set() for u in range(int(n_users))}
    for u, i in pairs:
        s[int(u)].add(int(i))
    return s


# ---- corpora / args for the host-side mirror ----------------------------------------------------------

def frames_corpus(train, dev, test, n_users=None, n_items=None):
    """BaseReader built from (user, item) pair arrays."""
    import pandas as pd
    from whisprrec_b200.helpers.BaseReader import BaseReader
    frames = []
    for pairs in (train, dev, test):
        pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
        frames.append(pd.DataFrame({'user_id': pairs[:, 0], 'item_id': pairs[:, 1],
                                    'timestamp': np.zeros(len(pairs), dtype=np.int64)}))
    return BaseReader.from_frames(*frames, n_users=n_users, n_items=n_items)


def ml100k_corpus():
    c = load('ml100k_corpus.npz')
    parts = [np.stack([c[ph + '_user'], c[ph + '_item']], 1) for ph in ('train', 'dev', 'test')]
    return frames_corpus(*parts)


def model_args(cls, **over):
    from whisprrec_b200 import main as wr_main
    return wr_main.default_args(cls, **over)
