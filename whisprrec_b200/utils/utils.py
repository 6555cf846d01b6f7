"""Host helpers with the reference's names and behaviour (reference src/utils/utils.py)."""
import datetime
import logging
import os
import random

import numpy as np
import torch


def init_seed(seed):
    """utils.py:13-20: python, numpy (legacy global RandomState) and torch generators, in the reference's order (the
    negative sampler and the DataLoader replay depend on exactly these two global streams)."""
    seeders = [random.seed, np.random.seed, torch.manual_seed]
    if torch.cuda.is_available():
        seeders.append(torch.cuda.manual_seed_all)
    for seed_fn in seeders:
        seed_fn(seed)
    cudnn = torch.backends.cudnn
    cudnn.benchmark, cudnn.deterministic = False, True


def df_to_dict(df):
    """utils.py:26-30: column -> numpy array."""
    return {name: np.array(values) for name, values in df.to_dict('list').items()}


def batch_to_gpu(batch, device):
    """utils.py:33-37: every tensor of the feed dict moves to `device`, in place; other entries stay."""
    batch.update({k: v.to(device) for k, v in batch.items() if type(v) is torch.Tensor})
    return batch


def check(check_list):
    """utils.py:40-47: dump the tensors a model put in `check_list`."""
    logging.info('')
    for name, tensor in check_list:
        arr = np.array(tensor.detach().cpu())
        logging.info(os.linesep.join([name + '\t' + str(arr.shape), np.array2string(arr, threshold=20)]) + os.linesep)


_FLOATS = (float, np.floating)
_INTS = (int, np.integer)


def format_metric(result_dict):
    """utils.py:57-70: 'HR@10:0.1234,NDCG@10:0.0567', cut-offs ascending, metric names sorted.

    (The reference line :66 names `np.float_`, which NumPy 2 removed; the type test here accepts any
    NumPy floating / integer scalar, which is what that line meant.)
    """
    assert type(result_dict) == dict
    names = sorted({key.split('@')[0] for key in result_dict})
    cuts = sorted({int(key.split('@')[1]) for key in result_dict})
    parts = []
    for k in cuts:
        for name in names:
            key = '{}@{}'.format(name, k)
            value = result_dict[key]
            if isinstance(value, _FLOATS):
                parts.append('{}:{:<.4f}'.format(key, value))
            elif isinstance(value, _INTS):
                parts.append('{}:{}'.format(key, value))
    return ','.join(parts)


def format_arg_str(args, exclude_lst, max_len=20):
    """utils.py:73-94: the argument table printed at start-up."""
    sep = os.linesep
    items = {k: v for k, v in vars(args).items() if k not in exclude_lst}
    head_k, head_v = 'Arguments', 'Values'
    wk = max(len(head_k), max(len(str(k)) for k in items))
    wv = max(len(head_v), min(max(len(str(v)) for v in items.values()), max_len))
    rule = '=' * (wk + wv + 5)
    out = sep + rule + sep + ' ' + head_k.ljust(wk) + ' | ' + head_v.ljust(wv) + ' ' + sep + rule + sep
    for key in sorted(items):
        value = items[key]
        if value is None:
            continue
        text = str(value).replace('\t', '\\t')
        if len(text) > max_len:
            text = text[:max_len - 3] + '...'
        out += ' ' + str(key).ljust(wk) + ' | ' + text.ljust(wv) + sep
    return out + rule


def check_dir(file_name):
    folder = os.path.dirname(file_name)
    if not os.path.exists(folder):
        print('make dirs:', folder)
        os.makedirs(folder)


def non_increasing(lst):
    return all(a >= b for a, b in zip(lst, lst[1:]))


def get_time():
    return datetime.datetime.now().strftime('%Y-%m-%d %H:%M:%S')
