"""Multi-rank paths: host-side shard logic under gloo on the CPU, the peer-memory kernels on
one B200 per rank."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from whisprrec_b200.sharded import ShardLayout, combine_shard_ranks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, 'tests', 'dist_worker.py')


def torchrun(mode, world, port, timeout):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(port), WORKER, mode]
    env = dict(os.environ, OMP_NUM_THREADS='1')
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


@pytest.mark.parametrize('world', [2, 3])
def test_shard_logic_under_gloo(world):
    """World-size-2/3 gloo run: ownership, batch slices, gradient sum, item-sharded ranks / top-k merge,
    row-partitioned SpMM -- the host logic of the sharded path with the oracle as the compute."""
    r = torchrun('cpu', world, 29531 + world, 600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'dist_worker cpu ok' in r.stdout


@pytest.mark.parametrize('world', [1, 2, 4, 8])
def test_layout_arithmetic(world):
    nU, nI = 103, 57
    seen_u, seen_i = np.zeros(nU, int), np.zeros(nI, int)
    for rank in range(world):
        lay = ShardLayout(nU, nI, world, rank)
        assert lay.rows_u_local == -(-nU // world) and lay.n_local == lay.rows_u_local + lay.rows_i_local
        seen_u[lay.local_users()] += 1
        seen_i[lay.local_items()] += 1
        items = np.arange(nI)
        loc = lay.item_local_index(items)
        assert (lay.item_global_index(loc[loc >= 0]) == items[loc >= 0]).all()
        t = lay.item_local_index(torch.arange(nI))
        assert (t.numpy() == loc).all()
        lo, hi = lay.batch_slice(1001)
        assert 0 <= lo <= hi <= 1001 and (hi - lo) in (1001 // world, 1001 // world + 1)
    assert (seen_u == 1).all() and (seen_i == 1).all()
    assert (combine_shard_ranks([np.array([1, 3]), np.array([2, 1])]) == np.array([2, 3])).all()


def test_localised_history_keeps_order_and_membership():
    rng = np.random.RandomState(0)
    nU, nI, world = 20, 50, 4
    ptr = np.zeros(nU + 1, dtype=np.int64)
    rows = [np.sort(rng.choice(nI, rng.randint(0, 12), replace=False)) for _ in range(nU)]
    ptr[1:] = np.cumsum([len(r) for r in rows])
    idx = np.concatenate(rows).astype(np.int32)
    for rank in range(world):
        lay = ShardLayout(nU, nI, world, rank)
        lp, li = lay.localise_history(ptr, idx)
        for u in range(nU):
            want = rows[u][rows[u] % world == rank] // world
            assert (li[lp[u]:lp[u + 1]] == want).all()


@pytest.mark.gpu
def test_sharded_paths_on_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs one GPU per rank (ranks that spin on each other must never share a GPU)')
    r = torchrun('gpu', 2, 29541, 900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count('dist_worker gpu ok') == 5
