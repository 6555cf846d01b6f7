"""CPU oracle for the WhisprRec general-recommender hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain restatement, on the CPU, of what the reference computes on
the BPRMF / LightGCN path.  It exists to CHECK the CUDA path; only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  Nothing under `whisprrec_b200/` imports it and the
product fails loudly without its CUDA library instead of falling back to this.

Parity status: PINNED.  The reference ships no tests, so
the pins are outputs of the unmodified reference classes run in the authoring
container by `tests/golden/make_golden.py` and committed under `tests/golden/`:
every tensor of three optimiser steps on four small problems (BPRMF D=16/64,
LightGCN L=2/3), and on ml-100k the epoch-1 negatives, batch order, step
losses, parameters after one epoch and the dev ranks / HR / NDCG.
`tests/test_oracle.py` checks every function below against them.

The arithmetic of the reference lives in PyTorch 2.x ATen (gather, mul/sum,
sigmoid/log/mean, autograd, torch.optim.Adam `_single_tensor_adam`), NumPy's
legacy `RandomState` (MT19937 + masked rejection) and SciPy sparse products.
Data movement here uses torch CPU tensor ops (index_select / index_add_ /
sparse CSR mm) so the oracle can serve as a multi-threaded CPU baseline; all
the *math* (loss, analytic backward, Adam, normalisation, pooling, ranks) is
written out explicitly rather than delegated to autograd or torch.optim.

Each function cites the reference lines it follows (paths relative to the
reference checkout).
"""
import math

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# a1. negative sampling -- src/models/BaseModel.py:167-177, NumPy legacy RandomState
# --------------------------------------------------------------------------------------


class MT19937:
    """MT19937 as seeded by `np.random.seed(int)` (init_genrand) with 32-bit outputs.

    NumPy's legacy `RandomState.randint(low, high)` for a range that fits 32 bits
    serves every attempt with one raw output r: v = r & mask, accept iff v <= high-1-low
    (numpy/random/_bounded_integers: buffered_bounded_masked_uint32).
    """
    N, M = 624, 397

    def __init__(self, seed):
        mt = np.empty(self.N, dtype=np.uint64)
        s = np.uint64(seed & 0xFFFFFFFF)
        for i in range(self.N):
            mt[i] = s
            s = (np.uint64(1812433253) * (s ^ (s >> np.uint64(30))) + np.uint64(i + 1)) & np.uint64(0xFFFFFFFF)
        self.mt = mt.astype(np.uint32)
        self.buf = None
        self.stream = np.empty(0, dtype=np.uint32)
        self.cursor = 0

    def _twist(self):
        mt, N, M = self.mt, self.N, self.M
        UP, LO, A = np.uint32(0x80000000), np.uint32(0x7FFFFFFF), np.uint32(0x9908B0DF)

        def mix(hi, lo, src):
            y = (hi & UP) | (lo & LO)
            return src ^ (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), A, np.uint32(0))
        # dependencies reach back M words, so regenerate in chunks no longer than N-M = 227
        k = N - M
        mt[0:k] = mix(mt[0:k], mt[1:k + 1], mt[M:N])
        mt[k:2 * k] = mix(mt[k:2 * k], mt[k + 1:2 * k + 1], mt[0:k])
        rest = N - 1 - 2 * k
        mt[2 * k:N - 1] = mix(mt[2 * k:N - 1], mt[2 * k + 1:N], mt[k:k + rest])
        mt[N - 1] = mix(mt[N - 1:N], mt[0:1], mt[M - 1:M])[0]
        y = mt.copy()
        y ^= y >> np.uint32(11)
        y ^= (y << np.uint32(7)) & np.uint32(0x9D2C5680)
        y ^= (y << np.uint32(15)) & np.uint32(0xEFC60000)
        y ^= y >> np.uint32(18)
        self.buf = y

    def _ensure(self, n):
        """make at least n unread outputs available in self.stream"""
        while len(self.stream) - self.cursor < n:
            self._twist()
            self.stream = np.concatenate([self.stream[self.cursor:], self.buf])
            self.cursor = 0

    def raw(self, n):
        """next n tempered 32-bit outputs"""
        self._ensure(n)
        out = self.stream[self.cursor:self.cursor + n].copy()
        self.cursor += n
        return out

    @staticmethod
    def _mask(rng):
        mask = rng
        for sh in (1, 2, 4, 8, 16):
            mask |= mask >> sh
        return mask

    def randint_fill(self, low, high, n):
        """`np.random.randint(low, high, size=n)` on this stream (int64 result)."""
        rng = high - 1 - low
        if rng == 0:
            return np.full(n, low, dtype=np.int64)
        mask = self._mask(rng)
        out = np.empty(n, dtype=np.int64)
        got = 0
        while got < n:
            want = n - got
            r = self.raw(max(64, int(want * 1.3)))
            v = r & np.uint32(mask)
            ok = np.nonzero(v <= rng)[0]
            if len(ok) >= want:
                self.cursor -= len(r) - (int(ok[want - 1]) + 1)      # hand back what was not consumed
                ok = ok[:want]
            out[got:got + len(ok)] = low + v[ok].astype(np.int64)
            got += len(ok)
        return out

    def randint(self, low, high):
        rng = high - 1 - low
        if rng == 0:
            return low
        mask = self._mask(rng)
        while True:
            v = int(self.raw(1)[0]) & mask
            if v <= rng:
                return low + v


def neg_sample_epoch(mt, user_ids, n_items, train_clicked_set, num_neg=1):
    """BaseModel.py:167-177: bulk draw, then per row redraw while the item is in the user's train set."""
    n = len(user_ids)
    neg = mt.randint_fill(1, n_items, n * num_neg).reshape(n, num_neg)
    for i, u in enumerate(user_ids):
        clicked = train_clicked_set[int(u)]
        for j in range(num_neg):
            while int(neg[i, j]) in clicked:
                neg[i, j] = mt.randint(1, n_items)
    return neg.reshape(-1)


# --------------------------------------------------------------------------------------
# a2. batch order -- src/helpers/BaseRunner.py:188-193 (torch DataLoader + RandomSampler)
# --------------------------------------------------------------------------------------

def dataloader_draws(n, shuffle, generator=None):
    """What one `iter(DataLoader(...))` takes from the global torch CPU generator.

    torch/utils/data/dataloader.py: `_base_seed` is one int64 `random_()` for every loader;
    a shuffling loader's RandomSampler then draws its own int64 seed, seeds a fresh
    generator with it and yields `torch.randperm(n, generator=g)`.
    Returns the permutation (or None for a sequential loader).
    """
    torch.empty((), dtype=torch.int64).random_(generator=generator)          # _base_seed
    if not shuffle:
        return None
    seed = int(torch.empty((), dtype=torch.int64).random_(generator=generator).item())
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g).numpy()


# --------------------------------------------------------------------------------------
# a4-a7. BPRMF predict + analytic backward -- BPRMF.py:69-80, loss.py:33-39
# --------------------------------------------------------------------------------------

def _t(x, dtype=None):
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    return t if dtype is None else t.to(dtype)


def bpr_loss_and_coeff(s_pos, s_neg, gamma=1e-10):
    """loss = -mean(log(gamma + sigmoid(s_pos - s_neg)));  coeff_b = dLoss/d(s_pos_b) = -dLoss/d(s_neg_b)."""
    x = s_pos - s_neg
    sig = torch.sigmoid(x)
    loss = -(torch.log(gamma + sig)).mean()
    B = x.numel()
    coeff = -(sig * (1.0 - sig)) / (gamma + sig) / B
    return loss, coeff


def bpr_fwd_bwd(U, I, user, pos, neg, gamma=1e-10):
    """Returns loss (fp32 scalar), dense gU [n_users,D], dense gI [n_items,D]."""
    U, I = _t(U), _t(I)
    user, pos, neg = _t(user, torch.int64), _t(pos, torch.int64), _t(neg, torch.int64)
    ue, pe, ne = U.index_select(0, user), I.index_select(0, pos), I.index_select(0, neg)
    s_pos = (ue * pe).sum(dim=1)
    s_neg = (ue * ne).sum(dim=1)
    loss, c = bpr_loss_and_coeff(s_pos, s_neg, gamma)
    c = c.unsqueeze(1)
    gU = torch.zeros_like(U).index_add_(0, user, c * pe - c * ne)
    gI = torch.zeros_like(I).index_add_(0, pos, c * ue).index_add_(0, neg, -c * ue)
    return loss, gU, gI


# --------------------------------------------------------------------------------------
# a8. dense Adam with coupled L2 -- BaseRunner.py:120-124,199; torch/optim/adam.py _single_tensor_adam
# --------------------------------------------------------------------------------------

def adam_l2_step(p, m, v, g, step, lr, l2, beta1=0.9, beta2=0.999, eps=1e-8):
    """In place on torch fp32 tensors p, m, v.  `step` is 1-based.  Every row moves every step."""
    if l2 != 0:
        g = g.add(p, alpha=l2)
    w = 1.0 - beta1
    m.add_((g - m) * w)                                   # lerp_(grad, 1-beta1), weight < 0.5 branch
    v.mul_(beta2).add_(g * g * (1.0 - beta2))             # mul_(beta2).addcmul_(g, g, value=1-beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.add_(m / denom * (-step_size))                      # addcdiv_(m, denom, value=-step_size)


def adam_scalars(step, lr, beta1=0.9, beta2=0.999):
    """The two Python-double scalars Adam feeds its fp32 kernels: (lr/bc1, sqrt(bc2))."""
    return lr / (1.0 - beta1 ** step), math.sqrt(1.0 - beta2 ** step)


# --------------------------------------------------------------------------------------
# a9. LightGCN adjacency -- LightGCN.py:54-121
# --------------------------------------------------------------------------------------

def deg_inv_sqrt(deg):
    """np.power(fp32(rowsum) + 1e-10, -0.5) in fp32 (LightGCN.py:89-93); isolated nodes give 1e5."""
    rowsum = deg.astype(np.float32) + 1e-10
    return np.power(rowsum, -0.5).astype(np.float32)


def build_norm_adj_csr(n_users, n_items, train_user, train_item):
    """CSR of D^-1/2 [[0,R],[R^T,0]] D^-1/2 over distinct train pairs; rows and columns ascending.

    Weight recipe: w = fl32(fl32(d[row] * 1) * d[col]) -- the two SciPy diag products of
    LightGCN.py:95-97 -- bit-equal to the reference on every edge.
    """
    U, I = int(n_users), int(n_items)
    u = np.asarray(train_user, dtype=np.int64)
    i = np.asarray(train_item, dtype=np.int64)
    key = np.unique(u * I + i)                              # R is binary: duplicates collapse (dok assignment)
    u, i = key // I, key % I
    rows = np.concatenate([u, U + i])
    cols = np.concatenate([U + i, u])
    order = np.lexsort((cols, rows))
    rows, cols = rows[order], cols[order]
    N = U + I
    deg = np.bincount(rows, minlength=N)
    rowptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(deg, out=rowptr[1:])
    d = deg_inv_sqrt(deg)
    val = ((d[rows] * np.float32(1.0)) * d[cols]).astype(np.float32)
    return rowptr, cols.astype(np.int32), val


def csr_to_torch(rowptr, col, val, N):
    return torch.sparse_csr_tensor(_t(rowptr, torch.int64), _t(col, torch.int64), _t(val), size=(N, N))


# --------------------------------------------------------------------------------------
# a10-a11. LightGCN propagation, loss and analytic backward -- LightGCN.py:123-175, loss.py:83-98
# --------------------------------------------------------------------------------------

def lightgcn_propagate(A, E0, n_layers):
    """E^{k+1} = A E^k; pooled = mean_k E^k (sum in layer order, then divide: ATen CPU mean = sum().div_())."""
    acc = E0.clone()
    E = E0
    for _ in range(n_layers):
        E = torch.mm(A, E) if A.layout == torch.strided else torch.sparse.mm(A, E)
        acc = acc + E
    return acc / float(n_layers + 1)


def lightgcn_fwd_bwd(A, U0, I0, user, pos, neg, n_layers, reg_weight, gamma=1e-10):
    """Returns loss ([1]-shaped in the reference; scalar here), dense gU, gI on the EGO tables."""
    U0, I0 = _t(U0), _t(I0)
    nU = U0.shape[0]
    user, pos, neg = _t(user, torch.int64), _t(pos, torch.int64), _t(neg, torch.int64)
    E0 = torch.cat([U0, I0], dim=0)
    P = lightgcn_propagate(A, E0, n_layers)
    ue, pe, ne = P.index_select(0, user), P.index_select(0, nU + pos), P.index_select(0, nU + neg)
    s_pos, s_neg = (ue * pe).sum(1), (ue * ne).sum(1)
    mf, c = bpr_loss_and_coeff(s_pos, s_neg, gamma)
    # EmbLoss(require_pow=False): (||U0_b||_F + ||P0_b||_F + ||N0_b||_F) / B on the ego rows
    B = user.numel()
    u0, p0, n0 = U0.index_select(0, user), I0.index_select(0, pos), I0.index_select(0, neg)
    nu, npos, nneg = torch.linalg.norm(u0), torch.linalg.norm(p0), torch.linalg.norm(n0)
    reg = (nu + npos + nneg) / B
    loss = mf + reg_weight * reg
    # backward of the BPR part w.r.t. the pooled table
    c = c.unsqueeze(1)
    G = torch.zeros_like(P)
    G.index_add_(0, user, c * pe - c * ne)
    G.index_add_(0, nU + pos, c * ue)
    G.index_add_(0, nU + neg, -c * ue)
    # pooled = (1/(L+1)) sum_k A^k E0, A symmetric  =>  dE0 = H_0, H_L = G', H_{k-1} = G' + A^T H_k
    Gs = G / float(n_layers + 1)
    H = Gs
    At = A.t() if A.layout == torch.strided else A          # CSR built here is symmetric
    for _ in range(n_layers):
        H = (torch.mm(At, H) if A.layout == torch.strided else torch.sparse.mm(At, H)) + Gs
    # backward of the regulariser: d||X||_F/dX = X/||X||_F per occurrence
    k = reg_weight / B
    H = H.clone()
    H.index_add_(0, user, u0 * (k / nu))
    H.index_add_(0, nU + pos, p0 * (k / npos))
    H.index_add_(0, nU + neg, n0 * (k / nneg))
    return loss, H[:nU], H[nU:]


# --------------------------------------------------------------------------------------
# a12-a14. full-ranking evaluation -- BPRMF.py:82-91, BaseRunner.py:218-258, 50-92
# --------------------------------------------------------------------------------------

def full_scores(Uemb, Iemb, user):
    return torch.matmul(_t(Uemb).index_select(0, _t(user, torch.int64)), _t(Iemb).t())


def bf16_round(x):
    """fp32 -> bf16 (round to nearest even) -> fp32: the operand rounding of the tensor-core scoring mode."""
    return _t(x).to(torch.float32).bfloat16().float()


def full_scores_bf16(Uemb, Iemb, user):
    """Scores of the bf16 scoring mode: bf16-rounded operands, products and sums in fp32."""
    return torch.matmul(bf16_round(Uemb).index_select(0, _t(user, torch.int64)), bf16_round(Iemb).t())


def history_csr(n_users, train_pairs, residual_pairs):
    """Sorted union of train and residual (dev+test) items per user as CSR (BaseRunner.py:246-255)."""
    allp = np.concatenate([np.asarray(train_pairs, dtype=np.int64).reshape(-1, 2),
                           np.asarray(residual_pairs, dtype=np.int64).reshape(-1, 2)])
    big = int(allp[:, 1].max()) + 1 if len(allp) else 1
    key = np.unique(allp[:, 0] * big + allp[:, 1])
    u, i = key // big, key % big
    ptr = np.zeros(int(n_users) + 1, dtype=np.int64)
    np.cumsum(np.bincount(u, minlength=int(n_users)), out=ptr[1:])
    return ptr, i.astype(np.int32)


def predictions_matrix(scores, user, pos, hist_ptr, hist_idx):
    """[target | masked scores] exactly as BaseRunner.interface builds it (BaseRunner.py:238-257)."""
    s = np.array(scores, dtype=np.float32, copy=True)
    rows = np.arange(len(user))
    target = s[rows, pos].copy()
    for r, u in enumerate(user):
        s[r, hist_idx[hist_ptr[u]:hist_ptr[u + 1]]] = -np.inf
    return np.concatenate([target[:, None], s], axis=1)


def ranks_argsort(pred):
    """BaseRunner.py:72-73."""
    sort_idx = (-pred).argsort(axis=1)
    return np.argwhere(sort_idx == 0)[:, 1] + 1


def ranks_count(scores, user, pos, hist_ptr, hist_idx):
    """rank = 1 + #{unmasked j : s_j > s_target}; equals ranks_argsort when there are no ties."""
    s = np.asarray(scores, dtype=np.float32)
    rows = np.arange(len(user))
    target = s[rows, pos]
    gt = s > target[:, None]
    for r, u in enumerate(user):
        gt[r, hist_idx[hist_ptr[u]:hist_ptr[u + 1]]] = False
    return 1 + gt.sum(axis=1), target


def topk_masked(scores, user, hist_ptr, hist_idx, k):
    """Indices of the k best unmasked items per row, ties to the lower item id."""
    s = np.array(scores, dtype=np.float32, copy=True)
    for r, u in enumerate(user):
        s[r, hist_idx[hist_ptr[u]:hist_ptr[u + 1]]] = -np.inf
    idx = np.argsort(-s, axis=1, kind='stable')[:, :k]
    return idx.astype(np.int32), np.take_along_axis(s, idx, axis=1)


def evaluate_method(gt_rank, topk, metrics):
    """BaseRunner.py:76-88 given the ranks; float64 means."""
    gt_rank = np.asarray(gt_rank)
    out = {}
    for k in topk:
        hit = gt_rank <= k
        for metric in metrics:
            key = f'{metric}@{k}'
            ml = metric.lower()
            if ml in ('hr', 'recall'):
                out[key] = hit.mean()
            elif ml == 'ndcg':
                out[key] = np.mean(hit / np.log2(gt_rank + 1))
            elif ml == 'precision':
                out[key] = hit.sum() / (hit.shape[0] * k)
            else:
                raise ValueError(f'Undefined evaluation metric: {metric}.')
    return out
