"""Corpus reader with the reference's interface (reference src/helpers/BaseReader.py, src/utils/sample.py).

One-time host preprocessing, outside the accelerated path; restated so both sides see identical ids and
splits: same column renames, same >=20-row user filter *before* the rating filter (sample.py:33-40), ids
remapped by first appearance (:47-50), sklearn `train_test_split(random_state=42)` twice (:139-140).
On top of the reference's attributes the corpus also exposes the CSR views the device kernels consume.
"""
import logging
import os
from datetime import datetime, timezone

import numpy as np
import pandas as pd


_RENAMES = {'user_id:token': 'user_id', 'item_id:token': 'item_id', 'rating:float': 'rating',
            'timestamp:float': 'timestamp'}
_RATING_FLOOR = {'ml-1m': 3, 'ml-100k': 3, 'ml-10m': 3, 'yelp': 4, 'food': 4}


def count_statics(data_df, dataset):
    """sample.py:17-70."""
    data_df = data_df.rename(columns=_RENAMES)
    per_user = data_df['user_id'].value_counts()
    data_df = data_df[data_df['user_id'].isin(per_user[per_user >= 20].index)]
    if dataset in _RATING_FLOOR:
        data_df = data_df.loc[data_df['rating'] >= _RATING_FLOOR[dataset]].copy()
        data_df.drop(columns=['rating'], inplace=True)
    else:
        data_df = data_df.copy()
    for column in ('user_id', 'item_id'):
        codes, _ = pd.factorize(data_df[column], sort=False)      # ids in order of first appearance
        data_df[column] = codes
    n_users, n_items, n_clicks = data_df['user_id'].nunique(), data_df['item_id'].nunique(), len(data_df)
    logging.info('# Users: %d', n_users)
    logging.info('# Items: %d', n_items)
    logging.info('# Interactions: %d', n_clicks)
    logging.info('# density: %.8f%%', n_clicks / (n_users * n_items))
    fmt = '%Y-%m-%d'
    t0 = datetime.fromtimestamp(float(data_df['timestamp'].min()), tz=timezone.utc).strftime(fmt)
    t1 = datetime.fromtimestamp(float(data_df['timestamp'].max()), tz=timezone.utc).strftime(fmt)
    logging.info('Time Span: {}/{}'.format(t0, t1))
    return data_df


def random_split(data_df, ratios=(0.8, 0.1, 0.1)):
    """sample.py:116-151."""
    from sklearn.model_selection import train_test_split
    assert sum(ratios) == 1.0, 'ratios should sum to 1'
    train_df, rest = train_test_split(data_df, train_size=ratios[0], random_state=42, shuffle=True)
    dev_df, test_df = train_test_split(rest, train_size=ratios[1] / (ratios[1] + ratios[2]), random_state=42,
                                       shuffle=True)
    logging.info('Dataset has been split. Train dataset length: %d, Dev dataset length: %d, Test dataset length: %d',
                 len(train_df), len(dev_df), len(test_df))
    return train_df, dev_df, test_df


def leave_one_out_split(data_df):
    """sample.py:73-113: first row per user stays in train, last two go to test / dev."""
    first = data_df.groupby('user_id').head(1)
    rest = data_df[~data_df.index.isin(first.index)]
    test_df = rest.groupby('user_id').tail(1).copy()
    rest = rest[~rest.index.isin(test_df.index)]
    dev_df = rest.groupby('user_id').tail(1).copy()
    rest = rest[~rest.index.isin(dev_df.index)]
    train_df = pd.concat([first, rest]).sort_index()
    logging.info('Dataset has been split. Train dataset length: %d, Dev dataset length: %d, Test dataset length: %d',
                 len(train_df), len(dev_df), len(test_df))
    return train_df, dev_df, test_df


class BaseReader(object):
    @staticmethod
    def parse_reader_args(parser):
        parser.add_argument('--path', type=str, default='../data/', help='Input data dir.')
        parser.add_argument('--dataset', type=str, default='ml-100k', help='Choose a dataset.')
        parser.add_argument('--sep', type=str, default='\t', help='sep of csv file.')
        parser.add_argument('--sample', type=str, default='random', help='random or leave one out')
        return parser

    def __init__(self, args):
        self.sep = args.sep
        self.prefix = args.path
        self.dataset = args.dataset
        self.sample = args.sample
        self._read_data(*self._check_file())
        self._build_clicked_sets()

    @classmethod
    def from_frames(cls, train_df, dev_df, test_df, n_users=None, n_items=None):
        """Build a corpus from ready-made splits (synthetic workloads, tests)."""
        self = cls.__new__(cls)
        self.sep, self.prefix, self.dataset, self.sample = '\t', '', 'frames', 'given'
        self.all_df = pd.concat([train_df, dev_df, test_df])
        self._read_data(train_df, dev_df, test_df)
        if n_users is not None:
            self.n_users = n_users
        if n_items is not None:
            self.n_items = n_items
        self._build_clicked_sets()
        return self

    def _check_file(self):
        """BaseReader.py:48-65."""
        logging.info('Generating train, val and test dataset now···')
        inter_path = os.path.join(self.prefix, self.dataset, self.dataset + '.inter')
        if not os.path.exists(inter_path):
            logging.error('Interactions file not found.')
            raise FileNotFoundError(inter_path)
        data_df = count_statics(pd.read_csv(inter_path, sep='\t', header=0), self.dataset)
        self.all_df = data_df
        return random_split(data_df) if self.sample == 'random' else leave_one_out_split(data_df)

    def _read_data(self, train_df, dev_df, test_df):
        """BaseReader.py:67-86."""
        logging.info('Reading data from %s, dataset = %s', self.prefix, self.dataset)
        self.data_df = {'train': train_df, 'dev': dev_df, 'test': test_df}
        self.n_users = self.all_df['user_id'].max() + 1
        self.n_items = self.all_df['item_id'].max() + 1

    def _build_clicked_sets(self):
        """BaseReader.py:33-46: per-user train set and residual (dev+test) set."""
        self.train_clicked_set, self.residual_clicked_set = dict(), dict()
        for key in ('train', 'dev', 'test'):
            df = self.data_df[key]
            target = self.train_clicked_set if key == 'train' else self.residual_clicked_set
            for uid, iid in zip(df['user_id'].to_numpy(), df['item_id'].to_numpy()):
                if uid not in self.train_clicked_set:
                    self.train_clicked_set[uid] = set()
                    self.residual_clicked_set[uid] = set()
                target[uid].add(iid)
        self._csr_cache = {}

    # ---- CSR views for the device kernels (not in the reference) ------------------------------------
    def _pairs(self, phases):
        u = np.concatenate([self.data_df[p]['user_id'].to_numpy(dtype=np.int64) for p in phases])
        i = np.concatenate([self.data_df[p]['item_id'].to_numpy(dtype=np.int64) for p in phases])
        return u, i

    def _user_csr(self, phases):
        """Sorted, de-duplicated item lists per user over the given phases: (ptr int64 [U+1], idx int32)."""
        key = tuple(phases)
        cache = self.__dict__.setdefault('_csr_cache', {})
        if key not in cache:
            u, i = self._pairs(phases)
            stride = int(self.n_items)
            code = np.unique(u * stride + i)
            uu, ii = code // stride, code % stride
            ptr = np.zeros(int(self.n_users) + 1, dtype=np.int64)
            np.cumsum(np.bincount(uu, minlength=int(self.n_users)), out=ptr[1:])
            cache[key] = (ptr, ii.astype(np.int32))
        return cache[key]

    def train_csr(self):
        """train_clicked_set as CSR (negative-sampling rejection, LightGCN adjacency)."""
        return self._user_csr(('train',))

    def history_csr(self):
        """train_clicked_set | residual_clicked_set as CSR: the eval mask of BaseRunner.py:246-255."""
        return self._user_csr(('train', 'dev', 'test'))
