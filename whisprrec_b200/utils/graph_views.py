"""Augmented graph views for SGL (reference src/models/general/SGL.py:67-79, src/utils/augmentor.py:77-111) -- host side of
the SGL row of the hot-path table (models/general/SGL.py is the device side).

`edge_dropout_view` draws exactly the edges the reference keeps (Python's global `random` stream, reproduced by
wr_pyrandom_sample in the library, ~20x faster than `random.sample` on a million edges) and normalises the result as
`SGL.csr2tensor` does (row sums only: the two directions of an edge are dropped independently, the view is not symmetric).
"""
import numpy as np

from .. import _lib


def edge_dropout_view(rowptr, col, drop_ratio, transpose=False):
    """rowptr / col: CSR structure of the full symmetric adjacency [[0, R], [R^T, 0]] (rows and columns ascending, as
    `build_norm_adj_csr` returns it; that is also the order `adj_matrix.nonzero()` has in the reference).
    Returns (rowptr int64 [N+1], col int32 [kept], val fp32 [kept]) of one augmented, normalised view; with
    transpose=True also the same three arrays of its TRANSPOSE (the view is not symmetric, and the backward pass of a
    propagation multiplies by A^T)."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    col = np.asarray(col)
    n_nodes, nnz = len(rowptr) - 1, int(rowptr[-1])
    keep = _lib.py_random_sample(nnz, int(nnz * (1 - drop_ratio)))          # augmentor.py:95-98
    row_of_edge = np.repeat(np.arange(n_nodes, dtype=np.int64), np.diff(rowptr))
    rows, cols = row_of_edge[keep], col[:nnz][keep].astype(np.int64)
    deg = np.bincount(rows, minlength=n_nodes).astype(np.float32)
    dinv = np.power(deg + 1e-10, -0.5).astype(np.float32)                   # SGL.py:113-117
    dinv[np.isinf(dinv)] = 0.

    def csr(r, c):
        order = np.lexsort((c, r))
        r, c = r[order], c[order]
        # (D^-1/2 A) D^-1/2, two fp32 roundings; the product is commutative, so the transposed entry gets the same bits
        val = (dinv[r] * np.float32(1.0)) * dinv[c]
        out_ptr = np.zeros(n_nodes + 1, dtype=np.int64)
        np.cumsum(np.bincount(r, minlength=n_nodes), out=out_ptr[1:])
        return out_ptr, c.astype(np.int32), val.astype(np.float32)
    fwd = csr(rows, cols)
    if not transpose:
        return fwd
    t_ptr, t_col, t_val = csr(cols, rows)
    return fwd, (t_ptr, t_col, t_val)


def edge_dropout_view_device(rowptr_d, col_d, drop_ratio, ws):
    """The same view built on the device (SGL.graph_construction): the kept edge numbers are drawn on the host exactly as
    above (Python's stream cannot move), everything else -- row / column of every kept edge, the sort into CSR order,
    the transposed structure, the fp32 weights -- is wr_subgraph_csr + wr_csr_norm_weights; only the node degrees visit
    the host, for the reference's own NumPy `power` call.  rowptr_d / col_d: device CSR of the full adjacency.
    Returns ((rowptr, col, val), (rowptr_T, col_T, val_T)) as device tensors."""
    import torch
    nnz = int(col_d.numel())
    keep = torch.from_numpy(_lib.py_random_sample(nnz, int(nnz * (1 - drop_ratio)))).to(col_d.device)
    out = []
    dinv_d = None
    for transpose in (False, True):
        ptr, col = _lib.subgraph_csr(rowptr_d, col_d, keep, ws, transpose=transpose)
        if dinv_d is None:
            deg = (ptr[1:] - ptr[:-1]).to(torch.int32).cpu().numpy().astype(np.float32)
            dinv = np.power(deg + 1e-10, -0.5).astype(np.float32)              # SGL.py:113-117, row sums of the VIEW
            dinv[np.isinf(dinv)] = 0.
            dinv_d = torch.from_numpy(dinv).to(col_d.device)
        val = torch.empty(col.numel(), dtype=torch.float32, device=col_d.device)
        _lib.csr_norm_weights(ptr, col, dinv_d, val)                            # a * b == b * a: the transpose gets the same bits
        out.append((ptr, col, val))
    return out[0], out[1]
