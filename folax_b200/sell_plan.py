"""Integer plan of the sliced-ELLPACK copy of the duplicate-free CSR (host side, once per mesh, on demand).

The Krylov solvers of folax_b200/solvers multiply with the Jacobian many times per Newton step; one thread per
row on a column-major slice layout reads the matrix coalesced and needs no cross-lane reduction, so the product
is deterministic (csrc/krylov_threads.cuh).  Pure integer work.
"""
import numpy as np

SLICE = 32


def build(indptr, indices, dofs_per_node=1, chunk_rows=1 << 17):
    """CSR structure -> dict(slice_ptr int64 (nslices+1), cols int32 (total), src int32 (total; CSR position of each
    SELL entry, -1 for padding), diag_src int32 (nrows; CSR position of the diagonal, -1 if absent), nrows, total).
    With dofs_per_node = d in (2, 3) and rows made of runs of d consecutive dofs of one node (FE Jacobians), also
    node_cols int32 (total/d): one node index per run, for fol_sell_spmv_block (None if the structure does not fit).
    The fill runs in the library (csrc/plan_host.cu, slices in parallel on the host threads); `build_numpy` below is the
    NumPy restatement the tests compare it with (`chunk_rows` only matters there)."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
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
    d = int(dofs_per_node)
    blocked = d in (2, 3) and nnz > 0 and nnz % d == 0 and not np.any(row_len % d)
    cols = np.zeros(total, dtype=np.int32)
    src = np.full(total, -1, dtype=np.int32)
    diag_src = np.full(n, -1, dtype=np.int32)
    node_cols = np.zeros(total // d, dtype=np.int32) if blocked else None
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    fits = C.c_int(0)
    _lib.check(lib.fol_sell_plan_fill_host(p(indptr), p(indices), n, d, SLICE, p(slice_ptr), p(cols), p(src),
                                           p(diag_src), p(node_cols), C.byref(fits)))
    if not fits.value:
        node_cols = None
    return {"slice_ptr": slice_ptr, "cols": cols, "src": src, "diag_src": diag_src, "nrows": n, "total": total,
            "nnz": nnz, "node_cols": node_cols, "dofs_per_node": d}


def build_numpy(indptr, indices, dofs_per_node=1, chunk_rows=1 << 17):
    """CSR structure -> dict(slice_ptr int64 (nslices+1), cols int32 (total), src int32 (total; CSR position of each
    SELL entry, -1 for padding), diag_src int32 (nrows; CSR position of the diagonal, -1 if absent), nrows, total).
    With dofs_per_node = d in (2, 3) and rows made of runs of d consecutive dofs of one node (FE Jacobians), also
    node_cols int32 (total/d): one node index per run, for fol_sell_spmv_block (None if the structure does not fit).
    Rows are processed in chunks (a multiple of the slice height), so the temporaries stay small next to the
    outputs: 128^3 Hex8 elasticity (514 M entries) needs the outputs (4.1 GB of int32) plus well under 1 GB."""
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices)
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
    d = int(dofs_per_node)
    blocked = d in (2, 3) and nnz > 0 and nnz % d == 0 and not np.any(row_len % d)
    cols = np.zeros(total, dtype=np.int32)
    src = np.full(total, -1, dtype=np.int32)
    diag_src = np.full(n, -1, dtype=np.int32)
    node_cols = np.zeros(total // d, dtype=np.int32) if blocked else None
    chunk_rows = max(SLICE, (int(chunk_rows) // SLICE) * SLICE)
    for r0 in range(0, n, chunk_rows):
        r1 = min(n, r0 + chunk_rows)
        e0, e1 = int(indptr[r0]), int(indptr[r1])
        if e1 == e0:
            continue
        rows = np.repeat(np.arange(r0, r1, dtype=np.int64), row_len[r0:r1])
        k = np.arange(e0, e1, dtype=np.int64) - indptr[rows]
        pos = slice_ptr[rows // SLICE] + k * SLICE + rows % SLICE
        ind = indices[e0:e1].astype(np.int64)
        cols[pos] = ind
        src[pos] = np.arange(e0, e1, dtype=np.int64)
        on_diag = np.flatnonzero(ind == rows)
        diag_src[rows[on_diag]] = e0 + on_diag
        if blocked:
            runs = ind.reshape(-1, d)                     # rows hold whole runs, so this never straddles two rows
            if np.all(runs[:, 0] % d == 0) and np.all(runs == runs[:, :1] + np.arange(d)):
                rr, qq = rows[::d], k[::d] // d
                node_cols[slice_ptr[rr // SLICE] // d + qq * SLICE + rr % SLICE] = runs[:, 0] // d
            else:
                blocked, node_cols = False, None
    return {"slice_ptr": slice_ptr, "cols": cols, "src": src, "diag_src": diag_src, "nrows": n, "total": total,
            "nnz": nnz, "node_cols": node_cols, "dofs_per_node": d}
