"""The oracle against the reference's own outputs (tests/golden, made by make_golden.py). CPU only."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import whispr_oracle as O
from tests.helpers import assert_close, clicked_sets, load, small_case

SMALL = ['bprmf_d16', 'bprmf_d64', 'lgcn_d16_l2', 'lgcn_d64_l3']


def test_mt19937_matches_numpy_legacy_stream():
    mt = O.MT19937(3407)
    rs = np.random.RandomState(3407)
    assert (mt.raw(5000) == rs.randint(0, 2 ** 32, size=5000, dtype=np.uint32)).all()
    for hi in (2, 3, 53, 1574, 3706, 2_000_000):
        mt, rs = O.MT19937(11), np.random.RandomState(11)
        assert (mt.randint_fill(1, hi, 3000) == rs.randint(1, hi, size=(3000, 1)).reshape(-1)).all()
        assert mt.randint(1, hi) == rs.randint(1, hi)


def test_negative_sampler_bit_exact_on_ml100k():
    g, c = load('ml100k_bprmf.npz'), load('ml100k_corpus.npz')
    tu, ti = c['train_user'].astype(np.int64), c['train_item'].astype(np.int64)
    clicked = clicked_sets(c['n_users'], zip(tu, ti))
    mt = O.MT19937(3407)
    n1 = O.neg_sample_epoch(mt, tu, int(c['n_items']), clicked)
    assert n1.dtype == np.int64 and (n1 == g['neg_epoch1']).all()
    assert n1.min() >= 1                      # item 0 is never a negative
    n2 = O.neg_sample_epoch(mt, tu, int(c['n_items']), clicked)
    assert hashlib.sha256(n2.tobytes()).hexdigest() == str(g['neg_epoch2_sha'])


def test_batch_order_bit_exact_on_ml100k():
    g, c = load('ml100k_bprmf.npz'), load('ml100k_corpus.npz')
    # main.py order: seed, reader (no torch draws), model init (4 torch draws on CPU), then the fit loader
    torch.manual_seed(3407)
    U = torch.nn.Embedding(int(c['n_users']), 64)
    I = torch.nn.Embedding(int(c['n_items']), 64)
    torch.nn.init.xavier_normal_(U.weight.data)
    torch.nn.init.xavier_normal_(I.weight.data)
    assert hashlib.sha256(U.weight.detach().numpy().tobytes()).hexdigest() == str(g['init_user_sha'])
    assert hashlib.sha256(I.weight.detach().numpy().tobytes()).hexdigest() == str(g['init_item_sha'])
    perm = O.dataloader_draws(len(c['train_user']), shuffle=True)
    for b in (0, 1):
        sel = perm[b * 2048:(b + 1) * 2048]
        assert (c['train_user'][sel] == g[f'batch{b}_user']).all()
        assert (c['train_item'][sel] == g[f'batch{b}_pos']).all()
        assert (g['neg_epoch1'][sel] == g[f'batch{b}_neg']).all()


@pytest.mark.parametrize('tag', SMALL[:2])
def test_bprmf_steps_match_reference(tag):
    s = small_case(load('small_cases.npz'), tag)
    lr, l2 = float(s['hp'][0]), float(s['hp'][1])
    U, I = torch.from_numpy(s['U0'].copy()), torch.from_numpy(s['I0'].copy())
    mU, vU, mI, vI = (torch.zeros_like(x) for x in (U, U, I, I))
    for step in range(3):
        loss, gU, gI = O.bpr_fwd_bwd(U, I, s[f's{step}/user'], s[f's{step}/pos'], s[f's{step}/neg'])
        assert_close(loss.item(), s[f's{step}/loss'], f'{tag} loss step {step}')
        assert_close(gU.numpy(), s[f's{step}/gU'], f'{tag} gU step {step}')
        assert_close(gI.numpy(), s[f's{step}/gI'], f'{tag} gI step {step}')
        O.adam_l2_step(U, mU, vU, gU, step + 1, lr, l2)
        O.adam_l2_step(I, mI, vI, gI, step + 1, lr, l2)
        assert_close(U.numpy(), s[f's{step}/U'], f'{tag} U step {step}')
        assert_close(I.numpy(), s[f's{step}/I'], f'{tag} I step {step}')
    assert_close(mU.numpy(), s['mU'], f'{tag} exp_avg')
    assert_close(vU.numpy(), s['vU'], f'{tag} exp_avg_sq')


@pytest.mark.parametrize('tag', SMALL[2:])
def test_lightgcn_adjacency_bit_exact(tag):
    s = small_case(load('small_cases.npz'), tag)
    nU, nI = s['U0'].shape[0], s['I0'].shape[0]
    rowptr, col, val = O.build_norm_adj_csr(nU, nI, s['train'][:, 0], s['train'][:, 1])
    dense = np.zeros((nU + nI, nU + nI), dtype=np.float32)
    for r in range(nU + nI):
        dense[r, col[rowptr[r]:rowptr[r + 1]]] = val[rowptr[r]:rowptr[r + 1]]
    assert dense.tobytes() == s['adj_dense'].tobytes()
    assert (dense == dense.T).all()


def test_lightgcn_adjacency_ml100k():
    g, c = load('ml100k_lightgcn.npz'), load('ml100k_corpus.npz')
    rowptr, col, val = O.build_norm_adj_csr(c['n_users'], c['n_items'], c['train_user'], c['train_item'])
    assert len(val) == int(g['adj_nnz']) == 132032
    rows = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr))
    assert hashlib.sha256(rows.astype(np.int64).tobytes()).hexdigest() == str(g['adj_row_sha'])
    assert hashlib.sha256(col.astype(np.int64).tobytes()).hexdigest() == str(g['adj_col_sha'])
    assert hashlib.sha256(val.tobytes()).hexdigest() == str(g['adj_val_sha'])     # weights bit-equal on every edge


def test_deg_inv_sqrt_vs_correctly_rounded():
    """NumPy's fp32 power (LightGCN.py:92) is SIMD-dispatched and NOT correctly rounded, so its bits depend on
    the host CPU.  The product therefore takes d^-1/2 from the host (same NumPy call => same bits as the
    reference on the same machine); the device-side fallback fl32(1/sqrt(double(deg))) is within 1 ulp of it."""
    deg = np.arange(1, 1_000_001)
    a = O.deg_inv_sqrt(deg)
    b = (1.0 / np.sqrt(deg.astype(np.float64))).astype(np.float32)
    ulp = np.abs(a.view(np.int32) - b.view(np.int32))
    assert ulp.max() <= 1
    assert O.deg_inv_sqrt(np.array([0]))[0] == pytest.approx(1e5, rel=1e-6)   # isolated node: (0+1e-10)^-0.5


@pytest.mark.parametrize('tag', SMALL[2:])
def test_lightgcn_steps_match_reference(tag):
    s = small_case(load('small_cases.npz'), tag)
    lr, l2, reg_w, L = float(s['hp'][0]), float(s['hp'][1]), float(s['hp'][2]), int(s['hp'][3])
    nU, nI = s['U0'].shape[0], s['I0'].shape[0]
    rowptr, col, val = O.build_norm_adj_csr(nU, nI, s['train'][:, 0], s['train'][:, 1])
    A = O.csr_to_torch(rowptr, col, val, nU + nI)
    U, I = torch.from_numpy(s['U0'].copy()), torch.from_numpy(s['I0'].copy())
    P = O.lightgcn_propagate(A, torch.cat([U, I]), L)
    assert_close(P[:nU].numpy(), s['pooled_user0'], f'{tag} pooled users')
    assert_close(P[nU:].numpy(), s['pooled_item0'], f'{tag} pooled items')
    mU, vU, mI, vI = (torch.zeros_like(x) for x in (U, U, I, I))
    for step in range(3):
        loss, gU, gI = O.lightgcn_fwd_bwd(A, U, I, s[f's{step}/user'], s[f's{step}/pos'], s[f's{step}/neg'], L, reg_w)
        assert_close(loss.item(), s[f's{step}/loss'], f'{tag} loss step {step}')
        assert_close(gU.numpy(), s[f's{step}/gU'], f'{tag} gU step {step}')
        assert_close(gI.numpy(), s[f's{step}/gI'], f'{tag} gI step {step}')
        O.adam_l2_step(U, mU, vU, gU, step + 1, lr, l2)
        O.adam_l2_step(I, mI, vI, gI, step + 1, lr, l2)
        assert_close(U.numpy(), s[f's{step}/U'], f'{tag} U step {step}')
        assert_close(I.numpy(), s[f's{step}/I'], f'{tag} I step {step}')


@pytest.mark.parametrize('tag', SMALL)
def test_eval_rank_and_metrics_match_reference(tag):
    s = small_case(load('small_cases.npz'), tag)
    nU, nI = s['U0'].shape[0], s['I0'].shape[0]
    U, I = torch.from_numpy(s['s2/U']), torch.from_numpy(s['s2/I'])
    if tag.startswith('lgcn'):
        rowptr, col, val = O.build_norm_adj_csr(nU, nI, s['train'][:, 0], s['train'][:, 1])
        P = O.lightgcn_propagate(O.csr_to_torch(rowptr, col, val, nU + nI), torch.cat([U, I]), int(s['hp'][3]))
        U, I = P[:nU], P[nU:]
    user, pos = s['test'][:, 0], s['test'][:, 1]
    hp, hi = O.history_csr(nU, s['train'], np.concatenate([s['dev'], s['test']]))
    scores = O.full_scores(U, I, user).numpy()
    pred = O.predictions_matrix(scores, user, pos, hp, hi)
    assert_close(pred[np.isfinite(pred)], s['eval_pred'][np.isfinite(s['eval_pred'])], f'{tag} predictions')
    assert (np.isfinite(pred) == np.isfinite(s['eval_pred'])).all()          # same cells masked
    ranks, _ = O.ranks_count(scores, user, pos, hp, hi)
    assert (ranks == s['eval_rank']).all()
    assert (O.ranks_argsort(s['eval_pred']) == s['eval_rank']).all()
    res = O.evaluate_method(ranks, [5, 10, 20], ['NDCG', 'HR', 'RECALL', 'PRECISION'])
    for k, v in zip(s['eval_keys'], s['eval_vals']):
        assert res[str(k)] == pytest.approx(float(v), rel=1e-12), k
    idx, vals = O.topk_masked(scores, user, hp, hi, 10)
    for r in range(len(user)):
        assert not set(idx[r][np.isfinite(vals[r])]) & set(hi[hp[user[r]]:hp[user[r] + 1]])


def test_ml100k_epoch_trajectory_bprmf():
    """One full reference epoch (33 steps) re-run with the oracle: losses, parameters, dev ranks."""
    g, c = load('ml100k_bprmf.npz'), load('ml100k_corpus.npz')
    nU, nI = int(c['n_users']), int(c['n_items'])
    torch.manual_seed(3407)
    Ue, Ie = torch.nn.Embedding(nU, 64), torch.nn.Embedding(nI, 64)
    torch.nn.init.xavier_normal_(Ue.weight.data)
    torch.nn.init.xavier_normal_(Ie.weight.data)
    U, I = Ue.weight.detach().clone(), Ie.weight.detach().clone()
    perm = O.dataloader_draws(len(c['train_user']), shuffle=True)
    tu, ti, neg = c['train_user'].astype(np.int64), c['train_item'].astype(np.int64), g['neg_epoch1'].astype(np.int64)
    mU, vU, mI, vI = (torch.zeros_like(x) for x in (U, U, I, I))
    losses = []
    for step, lo in enumerate(range(0, len(perm), 2048)):
        sel = perm[lo:lo + 2048]
        loss, gU, gI = O.bpr_fwd_bwd(U, I, tu[sel], ti[sel], neg[sel])
        losses.append(loss.item())
        O.adam_l2_step(U, mU, vU, gU, step + 1, 1e-3, 1e-6)
        O.adam_l2_step(I, mI, vI, gI, step + 1, 1e-3, 1e-6)
    assert len(losses) == 33 and len(perm) - 32 * 2048 == 480
    assert_close(np.array(losses), g['step_losses'], 'step losses')
    assert np.mean(np.array(losses, dtype=np.float32)) == pytest.approx(float(g['epoch_mean_loss']), rel=1e-6)
    # Adam's m/sqrt(v) amplifies 1-ulp gradient differences early in training: parameters are compared
    # at 1e-4 of the table's largest element rather than at the per-op 1e-5 bound
    assert_close(U[:64].numpy(), g['after_user_rows'], 'U after epoch 1', rtol=1e-4, atol_scale=1e-4)
    assert_close(I[:64].numpy(), g['after_item_rows'], 'I after epoch 1', rtol=1e-4, atol_scale=1e-4)
    du, di = c['dev_user'].astype(np.int64), c['dev_item'].astype(np.int64)
    hp, hi = O.history_csr(nU, np.stack([tu, ti], 1),
                           np.concatenate([np.stack([du, di], 1),
                                           np.stack([c['test_user'], c['test_item']], 1).astype(np.int64)]))
    ranks, target = O.ranks_count(O.full_scores(U, I, du).numpy(), du, di, hp, hi)
    assert_close(target, g['dev_target'], 'dev target scores', rtol=1e-4, atol_scale=1e-4)
    assert np.mean(ranks != g['dev_rank']) < 0.01          # near-ties may swap after 33 Adam steps
    res = O.evaluate_method(ranks, [10, 20], ['NDCG', 'HR'])
    for k, v in zip(g['dev_metric_keys'], g['dev_metric_vals']):
        assert res[str(k)] == pytest.approx(float(v), abs=5e-4), k


# --------------------------------------------------------------------------------------------------------------
# SGL: the oracle of the next row, pinned to the reference before any kernel exists
# --------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize('tag', ['sgl_d16_l2', 'sgl_d32_l3'])
def test_sgl_oracle_matches_reference(tag):
    import random
    from oracle import sgl_oracle as S
    s = small_case(load('sgl_cases.npz'), tag)
    lr, l2, reg, L, D, tau, w_ssl, drop = (float(x) for x in s['hp'])
    L = int(L)
    nU, nI = s['U0'].shape[0], s['I0'].shape[0]
    N = nU + nI
    tr = s['train']
    # same seeding as utils.init_seed(1234); building the reference model consumes torch's generator only, so python's
    # `random` is still at its seed when graph_construction() draws the two views
    random.seed(1234)
    rows, cols = S.symmetric_edges(nU, nI, tr[:, 0], tr[:, 1])
    main = S.normalise(rows, cols, N)
    subs = [S.normalise(*S.edge_dropout(rows, cols, drop), N) for _ in range(2)]
    graphs = [O.csr_to_torch(*g, N) for g in (main, subs[0], subs[1])]
    for g, name in zip(graphs, ('graph', 'sub1', 'sub2')):
        assert (g.to_dense().numpy() == s[name]).all(), name          # views bit for bit (structure and fp32 weights)
    E0 = torch.from_numpy(np.concatenate([s['U0'], s['I0']]))
    for g, name in zip(graphs, ('main', 'sub1', 'sub2')):
        pooled = S.propagate(g, E0, L).numpy()
        assert_close(pooled[:nU], s[f'pooled_{name}_user'], name + ' users')
        assert_close(pooled[nU:], s[f'pooled_{name}_item'], name + ' items')
    U, I = torch.from_numpy(s['U0'].copy()), torch.from_numpy(s['I0'].copy())
    mom = [torch.zeros_like(U), torch.zeros_like(U), torch.zeros_like(I), torch.zeros_like(I)]
    for step in range(3):
        user, pos, neg = s[f's{step}/user'], s[f's{step}/pos'], s[f's{step}/neg']
        val, gU, gI = S.fwd_bwd(U, I, graphs, user, pos, neg, L, reg, tau, w_ssl)
        assert_close(val.item(), s[f's{step}/loss'], f'loss step {step}', rtol=2e-6)
        assert_close(gU.numpy(), s[f's{step}/gU'], f'gU step {step}', rtol=2e-5, atol_scale=2e-6)
        assert_close(gI.numpy(), s[f's{step}/gI'], f'gI step {step}', rtol=2e-5, atol_scale=2e-6)
        O.adam_l2_step(U, mom[0], mom[1], gU, step + 1, lr, l2)
        O.adam_l2_step(I, mom[2], mom[3], gI, step + 1, lr, l2)
        assert_close(U.numpy(), s[f's{step}/U'], f'U step {step}', rtol=2e-5, atol_scale=2e-6)
        assert_close(I.numpy(), s[f's{step}/I'], f'I step {step}', rtol=2e-5, atol_scale=2e-6)
    # the next epoch's views continue python's stream
    nxt = O.csr_to_torch(*S.normalise(*S.edge_dropout(rows, cols, drop), N), N).to_dense().numpy()
    assert np.count_nonzero(nxt) == int(s['sub1_epoch2_nnz'])
