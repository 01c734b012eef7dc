"""Device-resident sparse operator and BiCGSTAB: the GPU side of fol/solvers/fe_solver.py:60-103.

The reference converts the duplicate-keeping BCOO to a SciPy CSR on the host (fe_solver.py:71-72) or hands it to
`jax.scipy.sparse.linalg.bicgstab` (:62-67).  Here the Jacobian never leaves the device: `fol_csr_values` sums the
duplicates in a fixed order, `fol_gather_values` permutes the values into the sliced-ELLPACK layout, and the Krylov
iteration runs on `fol_sell_spmv` / `fol_vec_op` / `fol_dot` (csrc/krylov.cu).  PyTorch only owns the vectors.
The host reads the BiCGSTAB coefficients back four times per iteration (six scalars); everything else stays on the
stream.
"""
import math

import numpy as np
import torch

from . import _lib


class SellOperator:
    """y = J x for a Jacobian returned by `ComputeJacobianMatrixAndResidualVector` (duplicates summed)."""

    use_block_kernel = True      # False: always the scalar-column kernel (A/B checks; results are bit-identical)

    def __init__(self, loss, jacobian):
        self.loss = loss
        lib = _lib.load()
        self.indptr, self.indices, self.csr_values = loss.JacobianToCSR(jacobian)
        sp = loss._sell_plan()
        self.plan = sp
        self.n = sp["nrows"]
        self.values = torch.empty(max(sp["total"], 1), dtype=loss.dtype, device=loss.device)
        _lib.check(lib.fol_gather_values(_lib.stream_ptr(), loss._dt, sp["total"], _lib.ptr(sp["src"]),
                                         _lib.ptr(self.csr_values), _lib.ptr(self.values)))

    def matvec(self, x, out=None):
        L = self.loss
        if out is None:
            out = torch.empty_like(x)
        if self.plan.get("node_cols") is not None and self.use_block_kernel:
            # runs of d dofs per neighbour node: one column index per run (8 + 4/d bytes per entry instead of 12)
            _lib.check(_lib.load().fol_sell_spmv_block(_lib.stream_ptr(), L._dt, self.plan["dofs_per_node"], self.n,
                                                       _lib.ptr(self.plan["slice_ptr"]), _lib.ptr(self.plan["node_cols"]),
                                                       _lib.ptr(self.values), _lib.ptr(x), _lib.ptr(out)))
            return out
        _lib.check(_lib.load().fol_sell_spmv(_lib.stream_ptr(), L._dt, self.n, _lib.ptr(self.plan["slice_ptr"]),
                                             _lib.ptr(self.plan["cols"]), _lib.ptr(self.values), _lib.ptr(x),
                                             _lib.ptr(out)))
        return out

    def diagonal(self):
        L = self.loss
        d = torch.empty(max(self.n, 1), dtype=L.dtype, device=L.device)
        _lib.check(_lib.load().fol_gather_values(_lib.stream_ptr(), L._dt, self.n, _lib.ptr(self.plan["diag_src"]),
                                                 _lib.ptr(self.csr_values), _lib.ptr(d)))
        return d

    def to_scipy_csr(self):
        import scipy.sparse as sp
        return sp.csr_array((self.csr_values.cpu().numpy(), self.indices.cpu().numpy(), self.indptr.cpu().numpy()),
                            shape=(self.n, self.n))


class SlabOperator:
    """The Jacobian of a slab-partitioned mesh (distributed.SlabPartition: one process per GPU, contiguous element
    slabs along z) as ONE operator: every rank holds the SELL matrix of its own elements, so its local product
    carries partial sums on the two interface node planes; the neighbour exchange that completes the residual
    (`SlabPartition.halo_sum`, or the NVLink peer path) completes the product too.  Vectors are stored per rank with
    both interface planes, identical on the two ranks that share them; dot products count each shared dof once
    (the rank below owns the plane) and are summed with one all-reduce per read.  The Dirichlet treatment (rows
    zeroed, diagonal kept per element, fe_loss.py:191-207) is element-local, so the summed rows equal those of
    the undivided mesh."""

    def __init__(self, loss, jacobian, part, group=None):
        self.local = SellOperator(loss, jacobian)
        self.loss, self.part, self.group, self.n = loss, part, group, self.local.n
        d = loss.number_dofs_per_node
        w = torch.ones(self.n, dtype=loss.dtype, device=loss.device)
        if part.world > 1 and part.rank > 0:
            w[:part.plane_nodes * d] = 0                 # the lower plane is owned by the rank below
        self.weights = w

    def matvec(self, x, out=None):
        out = self.local.matvec(x, out)
        self.part.halo_sum(out, self.loss.number_dofs_per_node, self.group)
        return out

    def diagonal(self):
        diag = self.local.diagonal()
        self.part.halo_sum(diag, self.loss.number_dofs_per_node, self.group)
        return diag

    def vectors(self, n):
        return _Vectors(self.loss._dt, n, self.loss.dtype, self.loss.device, weights=self.weights, group=self.group)


class _Vectors:
    """Vector updates and dot products of one solve, through the C ABI."""

    def __init__(self, dt, n, dtype, device, weights=None, group=None):
        """weights / group: slab-partitioned vectors (SlabOperator) -- every interface dof lives on two ranks, the
        weights (1 on owned dofs, 0 on the copies) make each dof count once, and `read` sums the ranks' partial
        dot products with one all-reduce."""
        self.lib, self.dt, self.n = _lib.load(), dt, n
        self.work = torch.empty(int(self.lib.fol_dot_work_size()), dtype=dtype, device=device)
        self.slots = torch.zeros(4, dtype=dtype, device=device)
        self.weights, self.group = weights, group
        self.tmp = torch.empty(n, dtype=dtype, device=device) if weights is not None else None

    def axpby(self, a, x, b, y, out):
        _lib.check(self.lib.fol_vec_op(_lib.stream_ptr(), self.dt, 0, self.n, float(a), _lib.ptr(x), float(b),
                                       _lib.ptr(y) if y is not None else None, _lib.ptr(out)))
        return out

    def divide(self, x, y, out):
        _lib.check(self.lib.fol_vec_op(_lib.stream_ptr(), self.dt, 2, self.n, 1.0, _lib.ptr(x), 0.0, _lib.ptr(y),
                                       _lib.ptr(out)))
        return out

    def dot_into(self, x, y, slot):
        if self.weights is not None:                    # count every shared dof once
            _lib.check(self.lib.fol_vec_op(_lib.stream_ptr(), self.dt, 1, self.n, 1.0, _lib.ptr(x), 0.0,
                                           _lib.ptr(self.weights), _lib.ptr(self.tmp)))
            x = self.tmp
        _lib.check(self.lib.fol_dot(_lib.stream_ptr(), self.dt, self.n, _lib.ptr(x), _lib.ptr(y), _lib.ptr(self.work),
                                    self.slots.data_ptr() + slot * self.slots.element_size()))

    def read(self, count=1):
        if self.weights is not None:
            import torch.distributed as dist
            if dist.is_initialized() and dist.get_world_size(self.group) > 1:
                dist.all_reduce(self.slots[:count], op=dist.ReduceOp.SUM, group=self.group)   # only what was just written
        vals = self.slots[:count].tolist()             # one device -> host read for `count` scalars
        return vals[0] if count == 1 else vals

    def dot(self, x, y):
        self.dot_into(x, y, 0)
        return self.read()


def norm(loss, x):
    v = _Vectors(loss._dt, x.numel(), loss.dtype, loss.device)
    return math.sqrt(max(v.dot(x, x), 0.0))


def bicgstab(A, b, x0=None, tol=1e-5, atol=0.0, maxiter=None, M_diagonal=None):
    """BiCGSTAB with the recurrences, start (rho = alpha = omega = 1, p = q = 0), stopping rule
    (|r|^2 <= max(tol^2 |b|^2, atol^2), also tested on the half step) and break-down codes of
    `jax.scipy.sparse.linalg.bicgstab`, which fe_solver.py:62-67 calls (JAX 0.8; restated from its documented
    algorithm -- the JAX sources are not in the reference tree).  A: SellOperator; b, x0: device vectors;
    M_diagonal: optional Jacobi preconditioner (the diagonal; the reference passes no preconditioner).
    Returns (x, info): info = number of iterations, or -10 / -11 on a rho / (alpha, omega) break-down."""
    L = A.loss
    n = b.numel()
    v = A.vectors(n) if hasattr(A, "vectors") else _Vectors(L._dt, n, L.dtype, L.device)
    if maxiter is None:
        maxiter = 10 * n
    new = lambda: torch.empty_like(b)
    x = b.new_zeros(n) if x0 is None else _lib.to_device(x0, L.dtype).reshape(-1).clone()
    q, r = A.matvec(x), new()
    v.axpby(1.0, b, -1.0, q, r)                        # r0 = b - A x0
    rhat = r.clone()
    p, phat, s, shat, t, tmp = b.new_zeros(n), new(), new(), new(), new(), new()
    q.zero_()
    v.dot_into(b, b, 0)
    v.dot_into(r, r, 1)
    v.dot_into(rhat, r, 2)
    bb, rs, rho_new = v.read(3)                          # rho of the next iteration rides along with |r|^2: one read
    atol2 = max(tol * tol * bb, atol * atol)
    rho = alpha = omega = 1.0
    k = 0
    while rs > atol2 and 0 <= k < maxiter:
        if rho_new == 0.0:
            k = -10
            break
        beta = rho_new / rho * alpha / omega
        v.axpby(1.0, p, -omega, q, tmp)                # p = r + beta (p - omega q)
        v.axpby(1.0, r, beta, tmp, p)
        if M_diagonal is not None:
            v.divide(p, M_diagonal, phat)
        else:
            phat = p
        A.matvec(phat, q)
        rq = v.dot(rhat, q)
        if rq == 0.0:
            k = -11
            break
        alpha = rho_new / rq
        v.axpby(1.0, r, -alpha, q, s)
        ss = v.dot(s, s)
        if ss < atol2:                                  # converged on the half step
            v.axpby(1.0, x, alpha, phat, x)
            r, s = s, r
            rs, rho, k = ss, rho_new, k + 1
            break                                       # rs = ss < atol2: the loop condition is false anyway
        if M_diagonal is not None:
            v.divide(s, M_diagonal, shat)
        else:
            shat = s
        A.matvec(shat, t)
        v.dot_into(t, s, 0)
        v.dot_into(t, t, 1)
        ts, tt = v.read(2)
        omega = ts / tt if tt != 0.0 else 0.0
        v.axpby(1.0, x, alpha, phat, x)
        v.axpby(1.0, x, omega, shat, x)
        v.axpby(1.0, s, -omega, t, r)
        rho = rho_new
        v.dot_into(r, r, 0)
        v.dot_into(rhat, r, 1)
        rs, rho_new = v.read(2)
        if omega == 0.0 or alpha == 0.0:
            k = -11
            break
        k += 1
    return x, k


def bicgstab_fused(A, b, x0=None, tol=1e-5, atol=0.0, maxiter=None, M_diagonal=None):
    """`bicgstab` as ONE persistent launch (csrc/krylov_fused.cu): same recurrences, stopping rule and break-down codes,
    dot products summed per CTA instead of per fixed block (iterates equal to rounding, deterministic run to run).  For
    the sizes where the multi-launch loop is launch-latency-bound; one host read when the solve is over.  Not for
    SlabOperator (its dot products need a collective)."""
    L = A.loss
    lib = _lib.load()
    if hasattr(A, "vectors"):
        raise ValueError("bicgstab_fused: slab-partitioned operators use bicgstab")
    n = b.numel()
    if maxiter is None:
        maxiter = 10 * n
    x = b.new_zeros(n) if x0 is None else _lib.to_device(x0, L.dtype).reshape(-1).clone()
    work = torch.empty(int(lib.fol_bicgstab_fused_work_size(n)), dtype=L.dtype, device=L.device)
    block = A.plan.get("node_cols") is not None and A.use_block_kernel
    d = int(A.plan["dofs_per_node"]) if block else 0
    cols = A.plan["node_cols"] if block else A.plan["cols"]
    _lib.check(lib.fol_bicgstab_fused(_lib.stream_ptr(), L._dt, d, n, _lib.ptr(A.plan["slice_ptr"]), _lib.ptr(cols),
                                      _lib.ptr(A.values), _lib.ptr(b.contiguous()), _lib.ptr(x),
                                      _lib.ptr(M_diagonal) if M_diagonal is not None else None, float(tol), float(atol),
                                      int(maxiter), _lib.ptr(work)))
    k, _, lost = work[8 * n + 16384: 8 * n + 16384 + 3].tolist()          # the one host read
    if lost:
        raise _lib.FolaxError("bicgstab_fused: a grid barrier timed out (the persistent grid was not fully resident)")
    return x, int(k)


# slots of the device scalar array (csrc/krylov_threads.cuh, enum BS_*)
(_BB, _RS, _RHO, _ALPHA, _OMEGA, _RHO_NEW, _BETA, _RQ, _SS, _TS, _TT, _ATOL2, _STATE, _K, _RS_NEXT, _RHO_NEXT,
 _MAXITER) = range(17)
_RUN, _HALF, _DONE, _BROKEN = 1, 2, 4, 8                # state masks (bit s = state s)


def bicgstab_device(A, b, x0=None, tol=1e-5, atol=0.0, maxiter=None, M_diagonal=None, check_every=8, use_graph=False):
    """`bicgstab` with the recurrence scalars kept on the device: the host enqueues `check_every` whole iterations
    (two products, the vector updates gated by the device-side state, five dot products, five one-thread scalar
    stages) without reading anything back, then looks at (state, k) once.  Same iterates, same stopping iteration
    and break-down codes as `bicgstab` (tests/test_solvers_glue_cpu.py compares them); not for SlabOperator (its dot
    products need a collective between the dot and the scalar stage).
    use_graph: the batch of `check_every` iterations contains no host decision, no allocation and no synchronisation,
    so it is captured once in a CUDA graph (after one eager batch) and replayed -- one launch per batch instead of
    ~25 per iteration.  (Written without GPU time; tests/test_zzz_graph_bicgstab_gpu.py is its first run.)"""
    L = A.loss
    lib = _lib.load()
    n = b.numel()
    if hasattr(A, "vectors"):
        raise ValueError("bicgstab_device: slab-partitioned operators use bicgstab")
    if maxiter is None:
        maxiter = 10 * n
    s_ptr = _lib.stream_ptr
    v = _Vectors(L._dt, n, L.dtype, L.device)
    nsc = int(lib.fol_bicg_scalar_count())
    sc = torch.zeros(nsc, dtype=L.dtype, device=L.device)
    esz = sc.element_size()
    slot = lambda i: sc.data_ptr() + i * esz

    def dot(x, y, i):
        _lib.check(lib.fol_dot(s_ptr(), L._dt, n, _lib.ptr(x), _lib.ptr(y), _lib.ptr(v.work), slot(i)))

    def vec(mask, ia, sa, x, ib, sb, y, out):
        _lib.check(lib.fol_vec_op_dev(s_ptr(), L._dt, n, _lib.ptr(sc), mask, ia, float(sa), _lib.ptr(x), ib, float(sb),
                                      _lib.ptr(y) if y is not None else None, _lib.ptr(out)))

    def stage(k):
        _lib.check(lib.fol_bicg_scalars(s_ptr(), L._dt, k, _lib.ptr(sc)))

    new = lambda: torch.zeros_like(b)      # the un-gated kernels (products, dots) read these even after the stop
    x = b.new_zeros(n) if x0 is None else _lib.to_device(x0, L.dtype).reshape(-1).clone()
    q, r = A.matvec(x), new()
    v.axpby(1.0, b, -1.0, q, r)
    rhat = r.clone()
    p, phat, s, shat, t, tmp = b.new_zeros(n), new(), new(), new(), new(), new()
    q.zero_()
    dot(b, b, _BB)
    dot(r, r, _RS)
    dot(rhat, r, _RHO_NEW)
    bb = float(sc[_BB])                                  # the only read before the loop: atol2 = max(tol^2 |b|^2, atol^2)
    init = torch.zeros(nsc, dtype=torch.float64)
    init[_RHO] = init[_ALPHA] = init[_OMEGA] = 1.0
    init[_ATOL2] = max(tol * tol * bb, atol * atol)
    init[_MAXITER] = float(maxiter)
    keep = torch.zeros(nsc, dtype=torch.bool)
    keep[[_BB, _RS, _RHO_NEW]] = True
    sc.copy_(torch.where(keep.to(sc.device), sc, init.to(device=sc.device, dtype=sc.dtype)))
    done = 0
    graph = None

    def run_batch():
        nonlocal phat, shat
        for _ in range(check_every):
            stage(0)                                                     # beta | stop test | rho break-down
            vec(_RUN, -1, 1.0, p, _OMEGA, -1.0, q, tmp)                  # tmp = p - omega q
            vec(_RUN, -1, 1.0, r, _BETA, 1.0, tmp, p)                    # p = r + beta tmp
            if M_diagonal is not None:
                v.divide(p, M_diagonal, phat)
            else:
                phat = p
            A.matvec(phat, q)
            dot(rhat, q, _RQ)
            stage(1)                                                     # alpha | break-down
            vec(_RUN, -1, 1.0, r, _ALPHA, -1.0, q, s)                    # s = r - alpha q
            dot(s, s, _SS)
            stage(2)                                                     # converged on the half step?
            if M_diagonal is not None:
                v.divide(s, M_diagonal, shat)
            else:
                shat = s
            A.matvec(shat, t)
            dot(t, s, _TS)
            dot(t, t, _TT)
            stage(3)                                                     # omega
            vec(_RUN | _HALF, -1, 1.0, x, _ALPHA, 1.0, phat, x)          # x += alpha phat
            vec(_RUN, -1, 1.0, x, _OMEGA, 1.0, shat, x)                  # x += omega shat
            vec(_RUN, -1, 1.0, s, _OMEGA, -1.0, t, r)                    # r = s - omega t
            dot(r, r, _RS_NEXT)
            dot(rhat, r, _RHO_NEXT)
            stage(4)                                                     # commit the iteration | (alpha, omega) break-down

    while True:
        if graph is not None:
            graph.replay()
        else:
            run_batch()
            if use_graph and done == 0:                                  # the first batch ran eagerly: capture the next ones
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    run_batch()
        done += check_every
        state, k = sc[[_STATE, _K]].tolist()                             # one host read per batch
        if int(state) != 0 or done >= maxiter + check_every:
            return x, int(k)
