"""CPU stand-ins used ONLY by the `-m "not gpu"` glue tests (test_responses_glue_cpu.py, test_solvers_glue_cpu.py).

There is no GPU in the build container, so the host-side classes (responses, solvers, BiCGSTAB) are exercised
against a stand-in of the C ABI that routes every entry point whose kernel body is a __host__ __device__ function
to tests/host_shim (the kernel's own per-thread code compiled for the CPU) and does the plain reductions in NumPy;
the loss object is a stand-in built from the oracle.  None of this is importable from the product and the product
has no such path (tests/test_cabi.py::test_no_cpu_fallback_without_cuda)."""
import ctypes as C
import os
import subprocess
import types

import numpy as np
import scipy.sparse as sp
import torch

from folax_b200 import _lib, sell_plan
from folax_b200.sparse import BCOO
from oracle import assembly

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_shim(out_dir=None):
    """nvcc -shared of the host shims (host code only is ever called).  The library is cached next to the sources
    (tests/host_shim/_build, git-ignored like every .so), keyed by a hash of everything it is compiled from."""
    import concurrent.futures as cf
    import hashlib
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    shim_dir = os.path.join(ROOT, "tests", "host_shim")
    srcs = [os.path.join(shim_dir, f) for f in ("adjoint_host.cu", "krylov_host.cu", "assemble_ad_host.cu", "j2_host.cu")]
    csrc = os.path.join(ROOT, "folax_b200", "csrc")
    deps = srcs + [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f.endswith((".cuh", ".h"))] + \
        [os.path.join(ROOT, "include", "folax_b200.h")]
    h = hashlib.sha256()
    for f in deps:
        with open(f, "rb") as fh:
            h.update(f.encode() + fh.read())
    tag = h.hexdigest()[:16]
    build_dir = str(out_dir) if out_dir is not None else os.path.join(shim_dir, "_build")
    os.makedirs(build_dir, exist_ok=True)
    out = os.path.join(build_dir, f"libfolax_host_shim_{tag}.so")
    if not os.path.exists(out):
        def compile_one(src):
            obj = os.path.join(build_dir, f"{os.path.basename(src)[:-3]}_{tag}.o")
            r = subprocess.run([nvcc, "-O1", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets",
                                "-Xcompiler", "-fPIC", "-c", src, "-o", obj], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            return obj
        with cf.ThreadPoolExecutor(max_workers=len(srcs)) as ex:
            objs = list(ex.map(compile_one, srcs))
        tmp = f"{out}.{os.getpid()}.tmp"
        r = subprocess.run([nvcc, "-shared", "-Wno-deprecated-gpu-targets"] + objs + ["-o", tmp],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        os.replace(tmp, out)
    return C.CDLL(out)


def arr(ptr, n, ctype=C.c_double):
    return np.ctypeslib.as_array((ctype * n).from_address(ptr)) if n else np.zeros(0)


_REAL_LOAD = _lib.load     # captured before any test swaps _lib.load for a FakeLib


class FakeLib:
    """fol_* entry points used by the response / solver classes, on host pointers (float64 only)."""

    def __init__(self, shim_lib, ne, nnode):
        self.shim, self.ne, self.nnode = shim_lib, ne, nnode
        self.calls = {}

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    def fol_gauss_interpolate(self, s, dt, element, num_gp, d, ne, conn, ctrl, u, kg, ug):
        self._count("gauss_interpolate")
        return self.shim.host_gauss_interpolate(element, num_gp, d, C.c_longlong(ne),
                                                *map(C.c_void_p, (conn, ctrl, u, kg, ug)))

    def fol_response_elements(self, s, dt, element, num_gp, d, ne, xyz, conn, f, fk, fu, val, du, dk, dx):
        self._count("response_elements")
        return self.shim.host_response_elements(element, num_gp, d, C.c_longlong(ne),
                                                *map(C.c_void_p, (xyz, conn, f, fk, fu, val, du, dk, dx)))

    def fol_residual_adjoint_elements(self, s, dt, physics, element, num_gp, acc, ne, xyz, conn, ctrl, u, lam, aux,
                                      params, dk, dx):
        self._count("residual_adjoint_elements")
        return self.shim.host_residual_adjoint_elements(physics, element, num_gp, acc, C.c_longlong(ne),
                                                        *map(C.c_void_p, (xyz, conn, ctrl, u, lam, aux)), params,
                                                        C.c_void_p(dk), C.c_void_p(dx))

    def fol_residual_gather(self, s, dt, nn, nnode, width, adj_ptr, adj, elem, out):
        ap, ad = arr(adj_ptr, nn + 1, C.c_int32), arr(adj, self.ne * nnode, C.c_int32)
        ev, o = arr(elem, self.ne * nnode * width), arr(out, nn * width)
        for n in range(nn):
            for k in range(width):
                o[n * width + k] = sum(ev[int(x) * width + k] for x in ad[ap[n]:ap[n + 1]])
        return 0

    # ---- host-side integer plans (csrc/plan_host.cu): pure host code, served by the real library
    def fol_csr_plan_count_host(self, *a):
        return _REAL_LOAD().fol_csr_plan_count_host(*a)

    def fol_csr_plan_fill_host(self, *a):
        return _REAL_LOAD().fol_csr_plan_fill_host(*a)

    def fol_sell_plan_fill_host(self, *a):
        return _REAL_LOAD().fol_sell_plan_fill_host(*a)

    def fol_sum(self, s, dt, n, x, out):
        arr(out, 1)[0] = arr(x, n).sum()
        return 0

    # ---- linear algebra (csrc/krylov.cu)
    def fol_gather_values(self, s, dt, n, src_index, src, dst):
        self._count("gather_values")
        return self.shim.host_gather_values(C.c_longlong(n), *map(C.c_void_p, (src_index, src, dst)))

    def fol_sell_spmv(self, s, dt, nrows, slice_ptr, cols, vals, x, y):
        self._count("sell_spmv")
        assert x != y
        return self.shim.host_sell_spmv(C.c_longlong(nrows), *map(C.c_void_p, (slice_ptr, cols, vals, x, y)))

    def fol_sell_spmv_block(self, s, dt, d, nrows, slice_ptr, node_cols, vals, x, y):
        self._count("sell_spmv")
        self._count("sell_spmv_block")
        assert x != y
        return self.shim.host_sell_spmv_block(d, C.c_longlong(nrows), *map(C.c_void_p, (slice_ptr, node_cols, vals, x, y)))

    def fol_vec_op(self, s, dt, op, n, a, x, b, y, out):
        self._count("vec_op")
        return self.shim.host_vec_op(op, C.c_longlong(n), C.c_double(a), C.c_void_p(x), C.c_double(b), C.c_void_p(y),
                                     C.c_void_p(out))

    def fol_bicg_scalar_count(self):
        return self.shim.host_bicg_scalar_count()

    def fol_bicg_scalars(self, s, dt, stage, sc):
        self._count("bicg_scalars")
        return self.shim.host_bicg_scalars(stage, C.c_void_p(sc))

    def fol_vec_op_dev(self, s, dt, n, sc, mask, ia, sa, x, ib, sb, y, out):
        self._count("vec_op_dev")
        return self.shim.host_vec_op_dev(C.c_longlong(n), C.c_void_p(sc), mask, ia, C.c_double(sa), C.c_void_p(x), ib,
                                         C.c_double(sb), C.c_void_p(y), C.c_void_p(out))

    def fol_dot_work_size(self):
        return 592

    def fol_dot(self, s, dt, n, x, y, work, out):
        self._count("dot")
        arr(out, 1)[0] = float(np.dot(arr(x, n), arr(y, n)))
        return 0


def fake_loss(physics, element_type, num_gp, coords, conn, node_sets, ordered_dofs, bc, params, params_array):
    """A loss object with the surface the response / solver classes use, computed by the oracle on CPU tensors."""
    didx, dval = assembly.dirichlet_vectors(ordered_dofs, bc, node_sets)
    ne, nn = conn.shape[0], coords.shape[0]
    a, d = conn.shape[1], len(ordered_dofs)
    ndof = d * nn
    order = np.argsort(conn.reshape(-1), kind="stable")            # entries e*a + local, ascending per node
    counts = np.bincount(conn.reshape(-1), minlength=nn)
    from oracle.geometry import ELEMENTS
    L = types.SimpleNamespace(
        physics=physics, dofs=list(ordered_dofs), dtype=torch.float64, device=torch.device("cpu"), _dt=_lib.F64,
        _ne=ne, _nn=nn, _nnode=a, _ngauss=len(ELEMENTS[element_type].gauss(num_gp)[1]), num_gp=num_gp,
        number_dofs_per_node=d, total_number_of_dofs=ndof, dim=ELEMENTS[element_type].dim,
        element_type=element_type, fe_element=types.SimpleNamespace(code=_lib.ELEMENTS[element_type]),
        fe_mesh=types.SimpleNamespace(GetNumberOfNodes=lambda: nn),
        _xyz=torch.as_tensor(np.ascontiguousarray(coords, dtype=np.float64)), _conn=torch.as_tensor(conn),
        _adj_ptr=torch.as_tensor(np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)),
        _adj=torch.as_tensor(order.astype(np.int32)), _dir_idx=torch.as_tensor(didx.astype(np.int32)),
        _params=_lib.params_array(params_array), Initialize=lambda reinitialize=False: None,
        dirichlet_indices=didx, dirichlet_values=dval, oracle_params=params, coords=coords, conn=conn)
    L.GetName = lambda: "oracle_backed_loss"

    def jac_and_res(ctrl, u, transpose=False):
        data, idx, R = assembly.assemble(physics, element_type, num_gp, coords, conn, np.asarray(ctrl, float).reshape(-1),
                                         np.asarray(u, float).reshape(-1), didx, params, transpose=transpose)
        return BCOO((torch.as_tensor(data), torch.as_tensor(idx)), shape=(ndof, ndof)), torch.as_tensor(R)

    def apply_bc(u, load_increment=1.0):
        out = torch.as_tensor(np.array(np.asarray(u, float).reshape(-1), copy=True))
        out[torch.as_tensor(didx.astype(np.int64))] = torch.as_tensor(load_increment * dval)
        return out

    def to_csr(jac):
        idx = jac.indices.numpy()
        A = sp.csr_array((jac.data.numpy(), (idx[:, 0], idx[:, 1])), shape=jac.shape)
        A.sum_duplicates()
        A.sort_indices()
        L._csr_structure = (A.indptr.copy(), A.indices.copy())
        return (torch.as_tensor(A.indptr.astype(np.int32)), torch.as_tensor(A.indices.astype(np.int32)),
                torch.as_tensor(A.data.copy()))

    def splan():
        if getattr(L, "_splan", None) is None:
            plan = sell_plan.build(*L._csr_structure, L.number_dofs_per_node)
            L._splan = {k: (torch.as_tensor(v) if isinstance(v, np.ndarray) else v) for k, v in plan.items()}
        return L._splan

    L.ComputeJacobianMatrixAndResidualVector = jac_and_res
    L.ApplyDirichletBCOnDofVector = apply_bc
    L.JacobianToCSR = to_csr
    L._sell_plan = splan
    return L


def install_globally(loss, shim_lib):
    """The same routing without pytest's monkeypatch -- for spawned worker processes (gloo tests), which end with
    the test and need no clean-up."""
    fake = FakeLib(shim_lib, loss._ne, loss._nnode)
    _lib.load = lambda: fake
    _lib.check = lambda rc: (_ for _ in ()).throw(RuntimeError(rc)) if rc else None
    _lib.stream_ptr = lambda: 0
    _lib.ptr = lambda t: None if t is None else t.data_ptr()
    _lib.to_device = (lambda x, dtype, device=None: torch.as_tensor(np.asarray(x)).to(dtype).contiguous()
                      if not isinstance(x, torch.Tensor) else x.to(dtype).contiguous())
    return fake


def make_cpu_backend(monkeypatch, shim):
    """install(loss) -> FakeLib: routes folax_b200._lib to the stand-ins for the duration of one test
    (the `cpu_backend` fixture of tests/conftest.py)."""
    def install(loss):
        fake = FakeLib(shim, loss._ne, loss._nnode)
        monkeypatch.setattr(_lib, "load", lambda: fake)
        monkeypatch.setattr(_lib, "check", lambda rc: (_ for _ in ()).throw(RuntimeError(rc)) if rc else None)
        monkeypatch.setattr(_lib, "stream_ptr", lambda: 0)
        monkeypatch.setattr(_lib, "ptr", lambda t: None if t is None else t.data_ptr())
        monkeypatch.setattr(_lib, "to_device",
                            lambda x, dtype, device=None: torch.as_tensor(np.asarray(x)).to(dtype).contiguous()
                            if not isinstance(x, torch.Tensor) else x.to(dtype).contiguous())
        return fake
    return install
