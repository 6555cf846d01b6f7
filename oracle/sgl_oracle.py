"""CPU oracle for SGL (reference src/models/general/SGL.py, src/utils/augmentor.py).  TEST INFRASTRUCTURE ONLY.

Groundwork for the next row of the hot-path table: the product does not implement SGL yet; this
file restates what the reference computes so that the kernels, when they come, have a pinned checker.

Parity status: PINNED to outputs of the unmodified reference (tests/golden/make_golden_sgl.py -> tests/golden/
sgl_cases.npz): the two edge-dropout sub-graphs bit for bit (they are drawn with Python's `random.sample`, which
`init_seed` seeds), pooled tables, three optimiser steps.  tests/test_oracle.py checks every function below.
The forward math is written out explicitly; its gradient is taken with torch autograd on that restatement (the
hand-written adjoint belongs to the kernels' design, not to the checker).
"""
import random

import numpy as np
import torch

from . import whispr_oracle as O


def symmetric_edges(n_users, n_items, train_user, train_item):
    """Non-zeros of [[0, R], [R^T, 0]] in the order `adj_matrix.nonzero()` yields them (SGL.py:81-103, augmentor.py:92):
    row-major, columns ascending -- the order edge_dropout's sampled indices refer to."""
    u = np.asarray(train_user, dtype=np.int64)
    i = np.asarray(train_item, dtype=np.int64)
    pairs = np.unique(np.stack([u, i], 1), axis=0)
    rows = np.concatenate([pairs[:, 0], n_users + pairs[:, 1]])
    cols = np.concatenate([n_users + pairs[:, 1], pairs[:, 0]])
    order = np.lexsort((cols, rows))
    return rows[order], cols[order]


def edge_dropout(rows, cols, drop_ratio, rnd=random):
    """augmentor.py:77-111: keep int(nnz * (1 - ratio)) edges chosen by `random.sample(range(nnz), keep)` (Python's
    global generator unless `rnd` is given).  Directions are dropped independently: the result is not symmetric."""
    n = len(rows)
    keep = rnd.sample(range(n), int(n * (1 - drop_ratio)))
    keep = np.asarray(keep, dtype=np.int64)
    return rows[keep], cols[keep]


def normalise(rows, cols, n_nodes):
    """SGL.py:105-132 (csr2tensor): D^-1/2 A D^-1/2 with D = row sums (+1e-10), all in fp32 as SciPy computes it.
    Returns CSR (rowptr int64, col int32, val fp32), columns ascending inside a row."""
    deg = np.bincount(rows, minlength=n_nodes).astype(np.float32)
    dinv = np.power(deg + 1e-10, -0.5).astype(np.float32)
    dinv[np.isinf(dinv)] = 0.
    order = np.lexsort((cols, rows))
    r, c = rows[order], cols[order]
    val = (dinv[r] * np.float32(1.0)) * dinv[c]
    rowptr = np.zeros(n_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(r, minlength=n_nodes), out=rowptr[1:])
    return rowptr, c.astype(np.int32), val.astype(np.float32)


def propagate(A, E0, n_layers):
    """SGL.py:148-163: mean over [E0, A E0, ..., A^L E0]."""
    return O.lightgcn_propagate(A, E0, n_layers)


def loss(U0, I0, graphs, user, pos, neg, n_layers, reg_weight, ssl_tau, ssl_weight):
    """SGL.py:165-246: sum-form BPR on the main graph + reg_weight * EmbLoss on the ego rows + InfoNCE between the two
    sub-graph views (users of the batch against all users, positive items against all items)."""
    nU = U0.shape[0]
    E0 = torch.cat([U0, I0])
    pooled = [propagate(A, E0, n_layers) for A in graphs]          # main, sub1, sub2
    ue, ie = pooled[0][:nU], pooled[0][nU:]
    u, p, n = (torch.as_tensor(np.asarray(x), dtype=torch.int64) for x in (user, pos, neg))
    x = (ue[u] * ie[p]).sum(1) - (ue[u] * ie[n]).sum(1)
    l1 = torch.sum(-torch.nn.functional.logsigmoid(x))
    B = u.numel()
    l2 = (torch.norm(U0[u], p=2) + torch.norm(I0[p], p=2) + torch.norm(I0[n], p=2)) / B        # EmbLoss, loss.py:83-98

    def info_nce(t1, t2, idx):
        a = torch.nn.functional.normalize(t1[idx], dim=1)
        b = torch.nn.functional.normalize(t2[idx], dim=1)
        allb = torch.nn.functional.normalize(t2, dim=1)
        pos_s = torch.exp((a * b).sum(1) / ssl_tau)
        tot_s = torch.exp(a.matmul(allb.T) / ssl_tau).sum(1)
        return -torch.sum(torch.log(pos_s / tot_s))
    ssl = info_nce(pooled[1][:nU], pooled[2][:nU], u) + info_nce(pooled[1][nU:], pooled[2][nU:], p)
    return l1 + l2 * reg_weight + ssl * ssl_weight


def fwd_bwd(U0, I0, graphs, user, pos, neg, n_layers, reg_weight, ssl_tau, ssl_weight):
    """(loss, dL/dU0, dL/dI0) of one SGL.predict + backward."""
    U = torch.as_tensor(np.asarray(U0), dtype=torch.float32).clone().requires_grad_(True)
    I = torch.as_tensor(np.asarray(I0), dtype=torch.float32).clone().requires_grad_(True)
    val = loss(U, I, graphs, user, pos, neg, n_layers, reg_weight, ssl_tau, ssl_weight)
    val.backward()
    return val.detach(), U.grad.detach(), I.grad.detach()
