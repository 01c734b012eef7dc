"""Integer plan of the sliced-ELLPACK copy of the duplicate-free CSR (host side, once per mesh, on demand).

The Krylov solvers of folax_b200/solvers multiply with the Jacobian many times per Newton step; one thread per
row on a column-major slice layout reads the matrix coalesced and needs no cross-lane reduction, so the product
is deterministic (csrc/krylov_threads.cuh).  Pure integer work.
"""
import numpy as np

SLICE = 32


def build(indptr, indices, dofs_per_node=1):
    """CSR structure -> dict(slice_ptr int64 (nslices+1), cols int32 (total), src int32 (total; CSR position of each
    SELL entry, -1 for padding), diag_src int32 (nrows; CSR position of the diagonal, -1 if absent), nrows, total).
    With dofs_per_node = d in (2, 3) and rows made of runs of d consecutive dofs of one node (FE Jacobians), also
    node_cols int32 (total/d): one node index per run, for fol_sell_spmv_block (None if the structure does not fit)."""
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    n = indptr.size - 1
    nnz = int(indptr[-1])
    row_len = np.diff(indptr)
    nsl = (n + SLICE - 1) // SLICE
    padded = np.zeros(nsl * SLICE, dtype=np.int64)
    padded[:n] = row_len
    width = padded.reshape(nsl, SLICE).max(axis=1) if nsl else np.zeros(0, np.int64)
    slice_ptr = np.concatenate([[0], np.cumsum(width * SLICE)]).astype(np.int64)
    total = int(slice_ptr[-1])
    if total >= 2 ** 31 or nnz >= 2 ** 31:
        raise ValueError("SELL plan exceeds int32 value indexing")
    rows = np.repeat(np.arange(n, dtype=np.int64), row_len)
    k = np.arange(nnz, dtype=np.int64) - indptr[rows]
    pos = slice_ptr[rows // SLICE] + k * SLICE + rows % SLICE
    cols = np.zeros(total, dtype=np.int32)
    src = np.full(total, -1, dtype=np.int32)
    cols[pos] = indices
    src[pos] = np.arange(nnz, dtype=np.int64)
    diag_src = np.full(n, -1, dtype=np.int32)
    on_diag = np.flatnonzero(indices == rows)
    diag_src[rows[on_diag]] = on_diag
    node_cols = None
    d = int(dofs_per_node)
    if d in (2, 3) and nnz and nnz % d == 0 and not np.any(row_len % d):
        runs = indices.reshape(-1, d)                      # rows hold whole runs, so this never straddles two rows
        if np.all(runs[:, 0] % d == 0) and np.all(runs == runs[:, :1] + np.arange(d)):
            node_cols = np.zeros(total // d, dtype=np.int32)
            first = np.arange(0, nnz, d)
            r = rows[first]
            q = k[first] // d
            node_cols[slice_ptr[r // SLICE] // d + q * SLICE + r % SLICE] = runs[:, 0] // d
    return {"slice_ptr": slice_ptr, "cols": cols, "src": src, "diag_src": diag_src, "nrows": n, "total": total,
            "nnz": nnz, "node_cols": node_cols, "dofs_per_node": d}
