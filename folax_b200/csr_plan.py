"""Integer plan of the duplicate-free CSR hand-off (host side, built once per mesh, on demand).

The BCOO of the reference keeps one entry per (element, i, j) (fe_loss.py:313-316); its consumers
sum duplicates on the host (scipy.sparse.csr_array, fe_solver.py:71-72).  This plan lets the GPU do
that sum in a fixed order.  Pure integer work, deterministic.

`build` runs in the library (csrc/plan_host.cu: node rows are independent, all host threads; 128^3 Hex8 in seconds);
`build_numpy` is the same plan written with NumPy sorts -- the readable restatement, which the tests hold `build`
against array for array.
"""
import ctypes as C

import numpy as np


def build(conn, nn, d):
    """conn (ne, A) -> dict with CSR structure (indptr (ndof+1), indices (nnz)) and the value plan."""
    from . import _lib
    lib = _lib.load()
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    ne, A = conn.shape
    nn, d = int(nn), int(d)
    if ne * A * A >= 2 ** 31:
        raise ValueError("CSR plan exceeds int32 indexing")
    p32 = lambda a: a.ctypes.data_as(C.c_void_p)
    adj_ptr, adj, deg = np.empty(nn + 1, np.int32), np.empty(ne * A, np.int32), np.empty(nn, np.int32)
    _lib.check(lib.fol_csr_plan_count_host(p32(conn), ne, A, nn, p32(adj_ptr), p32(adj), p32(deg)))
    node_ptr = np.zeros(nn + 1, np.int64)
    np.cumsum(deg, out=node_ptr[1:])
    npairs = int(node_ptr[-1])
    nnz, ndof = d * d * npairs, d * nn
    if nnz >= 2 ** 31:
        raise ValueError("CSR plan exceeds int32 indexing")
    out = {"indptr": np.empty(ndof + 1, np.int32), "indices": np.empty(nnz, np.int32),
           "pair_ptr": np.empty(npairs + 1, np.int32), "contrib": np.empty(ne * A * A, np.int32),
           "out_base": np.empty(npairs, np.int32), "row_stride": np.empty(npairs, np.int32)}
    _lib.check(lib.fol_csr_plan_fill_host(p32(conn), ne, A, nn, d, p32(adj_ptr), p32(adj), p32(node_ptr),
                                          p32(out["pair_ptr"]), p32(out["contrib"]), p32(out["out_base"]),
                                          p32(out["row_stride"]), p32(out["indptr"]), p32(out["indices"])))
    out.update(npairs=npairs, nnz=nnz, ndof=ndof)
    return out


def build_numpy(conn, nn, d):
    """The same plan with NumPy (one stable sort of all (e, a, b) by node pair); see the module docstring."""
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
