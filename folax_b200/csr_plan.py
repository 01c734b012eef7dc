"""Integer plan of the duplicate-free CSR hand-off (host side, built once per mesh, on demand).

The BCOO of the reference keeps one entry per (element, i, j) (fe_loss.py:313-316); its consumers
sum duplicates on the host (scipy.sparse.csr_array, fe_solver.py:71-72).  This plan lets the GPU do
that sum in a fixed order.  Pure integer work, deterministic.
"""
import numpy as np


def build(conn, nn, d):
    """conn (ne, A) -> dict with CSR structure (indptr (ndof+1), indices (nnz)) and the value plan."""
    conn = np.asarray(conn, dtype=np.int64)
    ne, A = conn.shape
    rows = np.repeat(conn, A, axis=1).reshape(-1)            # node of local a for every (e, a, b)
    cols = np.tile(conn, (1, A)).reshape(-1)                 # node of local b
    key = rows * nn + cols
    order = np.argsort(key, kind="stable")                   # contributors grouped by pair, ascending (e, a, b)
    skey = key[order]
    first = np.concatenate([[True], skey[1:] != skey[:-1]])
    pair_start = np.flatnonzero(first)
    pair_ptr = np.concatenate([pair_start, [len(skey)]])
    pkey = skey[pair_start]
    pn, pm = pkey // nn, pkey % nn                           # pairs sorted by (n, m)
    npairs = len(pkey)
    deg = np.bincount(pn, minlength=nn)                      # neighbours per node
    node_ptr = np.concatenate([[0], np.cumsum(deg)])
    q = np.arange(npairs) - node_ptr[pn]                     # index of the pair within its node row
    row_stride = d * deg[pn]
    out_base = d * d * node_ptr[pn] + q * d
    # scalar CSR structure
    ndof = d * nn
    row_nnz = np.repeat(d * deg, d)
    indptr = np.concatenate([[0], np.cumsum(row_nnz)])
    nnz = int(indptr[-1])
    if nnz >= 2 ** 31 or len(order) >= 2 ** 31:
        raise ValueError("CSR plan exceeds int32 indexing")
    indices = np.empty(nnz, dtype=np.int32)
    ar = np.arange(d)
    for p0 in range(0, npairs, 1 << 22):          # in chunks of pairs: the d*d-fold expansion is the memory peak otherwise
        sl = slice(p0, min(npairs, p0 + (1 << 22)))
        base = out_base[sl, None, None] + ar[None, :, None] * row_stride[sl, None, None] + ar[None, None, :]
        indices[base.reshape(-1)] = np.broadcast_to(d * pm[sl, None, None] + ar[None, None, :],
                                                    base.shape).reshape(-1).astype(np.int32)
    i32 = lambda x: np.ascontiguousarray(x, dtype=np.int32)
    return {"indptr": i32(indptr), "indices": indices, "pair_ptr": i32(pair_ptr), "contrib": i32(order),
            "out_base": i32(out_base), "row_stride": i32(row_stride), "npairs": npairs, "nnz": nnz, "ndof": ndof}
