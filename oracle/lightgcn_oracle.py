"""CPU oracle (TEST INFRASTRUCTURE, not product) for the LightGCN hot path.

Restates, in plain torch fp32 on the CPU:

* ``SparseTensor(row, col, sparse_sizes)`` CSR/CSC construction
  (reference call site ``data/lightgcn_loader.py:65-79``; third-party
  torch_sparse, algorithm per SURVEY.md Appendix A1)            -> ``csr_from_coo``, ``csc_from_csr``
* ``gcn_norm(adj, add_self_loops=False)`` (``model/lightgcn.py:56``; third-party
  PyG, Appendix A2)                                               -> ``gcn_norm_values``
* ``matmul(adj_t, x)`` (``model/lightgcn.py:85-87``; torch_sparse, A3)
                                                                   -> ``spmm`` / ``spmm_rowloop``
* ``LightGCN.forward`` (``model/lightgcn.py:46-80``)            -> ``lightgcn_forward``
* ``bpr_loss`` (``utils/metrics_lightgcn.py:9-45``)             -> ``bpr_loss``
* one training iteration (``run_pipeline_lightgcn.py:120-158``) -> ``train_iteration``

Gradients are obtained by torch autograd THROUGH this restatement, which is
exactly how the reference obtains them.  "parity unpinned" for the torch_sparse /
PyG pieces (see ``oracle/__init__.py``).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------
# A1. COO -> CSR / CSC   (torch_sparse SparseStorage semantics)
# --------------------------------------------------------------------------
def csr_from_coo(row: Tensor, col: Tensor, n_rows: int, n_cols: int
                 ) -> Tuple[Tensor, Tensor, Tensor]:
    """``SparseTensor(row=, col=, sparse_sizes=)`` -> (rowptr, col_sorted, perm).

    key = row*n_cols + col; entries are permuted by ``argsort(key)`` only when
    the keys are not already non-decreasing; duplicates are kept;
    ``rowptr[i] = #{e : row[e] < i}``  (data/lightgcn_loader.py:65-79 -> A1).
    """
    row = row.to(torch.int64).cpu()
    col = col.to(torch.int64).cpu()
    key = row * n_cols + col
    if key.numel() > 1 and bool((key[1:] < key[:-1]).any()):
        perm = torch.argsort(key, stable=True)
        row, col = row[perm], col[perm]
    else:
        perm = torch.arange(key.numel(), dtype=torch.int64)
    counts = torch.bincount(row, minlength=n_rows)
    rowptr = torch.zeros(n_rows + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(counts, 0)
    return rowptr, col, perm


def rows_from_rowptr(rowptr: Tensor) -> Tensor:
    n = rowptr.numel() - 1
    return torch.repeat_interleave(torch.arange(n, dtype=torch.int64), rowptr[1:] - rowptr[:-1])


def csc_from_csr(rowptr: Tensor, col: Tensor, n_cols: int
                 ) -> Tuple[Tensor, Tensor, Tensor]:
    """CSR -> (colptr, row_sorted_by_col, csr2csc) with ``csr2csc = argsort(col*M + row)`` (A1)."""
    n_rows = rowptr.numel() - 1
    row = rows_from_rowptr(rowptr)
    key = col * n_rows + row
    csr2csc = torch.argsort(key, stable=True)
    counts = torch.bincount(col, minlength=n_cols)
    colptr = torch.zeros(n_cols + 1, dtype=torch.int64)
    colptr[1:] = torch.cumsum(counts, 0)
    return colptr, row[csr2csc], csr2csc


# --------------------------------------------------------------------------
# A2. gcn_norm on a value-less SparseTensor, add_self_loops=False
# --------------------------------------------------------------------------
def gcn_norm_values(rowptr: Tensor, col: Tensor) -> Tuple[Tensor, Tensor]:
    """(dinv, val): deg = row sums of the all-ones matrix; dinv = deg^-1/2 with inf->0;
    val[e] = (1*dinv[row[e]])*dinv[col[e]] -- row scaling first, then column (model/lightgcn.py:56 -> A2).
    The matrix is square in every reference call, so ``dinv[col]`` is the ROW degree of node ``col``.
    """
    n = rowptr.numel() - 1
    deg = (rowptr[1:] - rowptr[:-1]).to(torch.float32)
    dinv = deg.pow(-0.5)
    dinv.masked_fill_(dinv == float("inf"), 0.0)
    row = rows_from_rowptr(rowptr)
    val = torch.ones(col.numel(), dtype=torch.float32)
    val = val * dinv[row]
    val = val * dinv[col]
    assert n >= 0
    return dinv, val


# --------------------------------------------------------------------------
# A3. SpMM  out[r,:] = sum_e val[e] * X[col[e],:]
# --------------------------------------------------------------------------
def spmm_rowloop(rowptr: Tensor, col: Tensor, val: Optional[Tensor], X: Tensor) -> Tensor:
    """Literal per-row sequential restatement (small inputs only)."""
    n = rowptr.numel() - 1
    out = torch.zeros(n, X.shape[1], dtype=X.dtype)
    rp = rowptr.tolist()
    for r in range(n):
        acc = torch.zeros(X.shape[1], dtype=X.dtype)
        for e in range(rp[r], rp[r + 1]):
            w = val[e] if val is not None else 1.0
            acc = acc + w * X[col[e]]
        out[r] = acc
    return out


def spmm(rowptr: Tensor, col: Tensor, val: Optional[Tensor], X: Tensor) -> Tensor:
    """Differentiable (w.r.t. X) SpMM: gather, scale, ``index_add_`` in CSR order.

    ``index_add_`` on the CPU visits the entries sequentially, i.e. the same per-row
    summation order as torch_sparse's CPU ``spmm`` (model/lightgcn.py:85-87 -> A3).
    Backward through autograd is ``A^T @ grad`` with constant values, as in torch_sparse.
    """
    n = rowptr.numel() - 1
    row = rows_from_rowptr(rowptr)
    msg = X.index_select(0, col)
    if val is not None:
        msg = msg * val.unsqueeze(1)
    return torch.zeros(n, X.shape[1], dtype=X.dtype).index_add_(0, row, msg)


def spmm_mkl(rowptr: Tensor, col: Tensor, val: Tensor, n_cols: int):
    """Same operator as a ``torch.sparse_csr_tensor`` (MKL, multi-threaded) -- used for the timed
    CPU baseline at sizes where ``spmm``'s nnz*d message buffer does not fit."""
    n = rowptr.numel() - 1
    return torch.sparse_csr_tensor(rowptr, col, val, size=(n, n_cols))


class _CsrMatmul(torch.autograd.Function):
    """y = A @ x with backward A^T @ g served by a pre-built CSR of the transpose -- exactly how
    torch_sparse's SPMMSum::backward re-runs its kernel on the cached (colptr, row[csr2csc], val[csr2csc])."""

    @staticmethod
    def forward(ctx, x, A, At):
        ctx.At = At
        return torch.sparse.mm(A, x)

    @staticmethod
    def backward(ctx, g):
        return torch.sparse.mm(ctx.At, g.contiguous()), None, None


def make_cpu_matmul(rowptr: Tensor, col: Tensor, val: Tensor, n_cols: int):
    """Multi-threaded CPU ``matmul(adj_t, x)`` closure with the torch_sparse-style transposed backward
    (the timed CPU baseline of bench.py)."""
    A = spmm_mkl(rowptr, col, val, n_cols)
    colptr, row_t, csr2csc = csc_from_csr(rowptr, col, n_cols)
    At = spmm_mkl(colptr, row_t, val[csr2csc], rowptr.numel() - 1)
    return lambda x: _CsrMatmul.apply(x, A, At)


# --------------------------------------------------------------------------
# LightGCN.forward  (model/lightgcn.py:46-80)
# --------------------------------------------------------------------------
def lightgcn_forward(users_w: Tensor, items_w: Tensor, rowptr: Tensor, col: Tensor,
                     num_iterations: int, val: Optional[Tensor] = None, matmul=None):
    """gcn_norm -> E0 = cat(Wu, Wi) -> K x (A_hat @ E) -> stack -> mean(dim=1) -> split.

    Returns (u_final, users_w, i_final, items_w) exactly like the reference (elements 1 and 3 are
    the layer-0 tables themselves).  ``matmul(E)`` may be injected (e.g. the MKL operator).
    """
    if val is None:
        _, val = gcn_norm_values(rowptr, col)                      # :56
    emb_0 = torch.cat([users_w, items_w])                           # :58
    embs = [emb_0]
    emb_k = emb_0
    for _ in range(num_iterations):                                 # :63-65
        emb_k = matmul(emb_k) if matmul is not None else spmm(rowptr, col, val, emb_k)
        embs.append(emb_k)
    emb_final = torch.mean(torch.stack(embs, dim=1), dim=1)         # :67-68
    u_final, i_final = torch.split(emb_final, [users_w.shape[0], items_w.shape[0]])  # :70-72
    return u_final, users_w, i_final, items_w


# --------------------------------------------------------------------------
# bpr_loss  (utils/metrics_lightgcn.py:9-45)
# --------------------------------------------------------------------------
def bpr_loss(u_f: Tensor, u_0: Tensor, p_f: Tensor, p_0: Tensor, n_f: Tensor, n_0: Tensor,
             lambda_val: float) -> Tensor:
    """reg = lambda*(|u0|_F^2 + |p0|_F^2 + |n0|_F^2);  loss = -mean(softplus(<u,p> - <u,n>)) + reg.

    Sign and the un-normalised regulariser are the reference's (utils/metrics_lightgcn.py:32-43).
    """
    reg = lambda_val * (u_0.norm(2).pow(2) + p_0.norm(2).pow(2) + n_0.norm(2).pow(2))   # :32-36
    pos = (u_f * p_f).sum(dim=-1)                                                       # :38-39
    neg = (u_f * n_f).sum(dim=-1)                                                       # :40-41
    return -torch.mean(torch.nn.functional.softplus(pos - neg)) + reg                   # :43


# --------------------------------------------------------------------------
# One training iteration without the optimizer (run_pipeline_lightgcn.py:120-158)
# --------------------------------------------------------------------------
def train_iteration(users_w: Tensor, items_w: Tensor, rowptr: Tensor, col: Tensor, num_iterations: int,
                    user_idx: Tensor, pos_idx: Tensor, neg_idx: Tensor, lambda_val: float,
                    val: Optional[Tensor] = None, matmul=None):
    """forward -> six row gathers -> bpr_loss -> backward.  Returns (loss, dWu, dWi, u_final, i_final)."""
    Wu = users_w.detach().clone().requires_grad_(True)
    Wi = items_w.detach().clone().requires_grad_(True)
    u_f, u_0, i_f, i_0 = lightgcn_forward(Wu, Wi, rowptr, col, num_iterations, val=val, matmul=matmul)  # :120
    loss = bpr_loss(u_f[user_idx], u_0[user_idx],                                                   # :133-155
                    i_f[pos_idx], i_0[pos_idx],
                    i_f[neg_idx], i_0[neg_idx], lambda_val)
    loss.backward()                                                                                 # :158
    return loss.detach(), Wu.grad, Wi.grad, u_f.detach(), i_f.detach()


# --------------------------------------------------------------------------
# Graph wirings used by the tests and the bench
# --------------------------------------------------------------------------
def wiring_reference(users: Tensor, items: Tensor, num_users: int, num_items: int):
    """(R): rows = user ids, cols = item ids in [0, I), matrix (U+I)^2 -- exactly what
    data/lightgcn_loader.py:61-69 hands to ``SparseTensor`` (non-symmetric; ids collide)."""
    n = num_users + num_items
    return users.to(torch.int64), items.to(torch.int64), n


def wiring_symmetric(users: Tensor, items: Tensor, num_users: int, num_items: int):
    """(S): the proper bipartite block matrix [[0, R], [R^T, 0]], nnz = 2E."""
    n = num_users + num_items
    u = users.to(torch.int64)
    i = items.to(torch.int64) + num_users
    return torch.cat([u, i]), torch.cat([i, u]), n
