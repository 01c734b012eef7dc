"""CPU check of the host glue of folax_b200.responses.FiniteElementResponse against the reference's own
known answers (tests/unit/test_sensitivity_analysis.py:54-80, tests/integration/test_mechanical_2D_sa.py:81-113).

No GPU here, so the C ABI is replaced -- IN THIS TEST ONLY -- by a stand-in that routes the three adjoint entry
points to tests/host_shim (the kernels' own per-thread code compiled for the CPU) and does the two trivial
reductions in NumPy; the loss object is a stand-in built from the oracle.  What is exercised for real is the
response class: formula handling, autograd partials, array layouts, call order, signs, Dirichlet handling.
The product itself has no such path (tests/test_cabi.py::test_no_cpu_fallback_without_cuda)."""
import ctypes as C
import json
import os
import types

import numpy as np
import pytest
import torch

from folax_b200 import _lib
from folax_b200.responses import FiniteElementResponse, NodalControl
from folax_b200.sparse import BCOO
from oracle import assembly
from tests.test_adjoint_host_shim import shim  # noqa: F401  (fixture)
from tests.test_oracle_golden import _square_mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _arr(ptr, n, ctype=C.c_double):
    return np.ctypeslib.as_array((ctype * n).from_address(ptr)) if n else np.zeros(0)


class _FakeLib:
    """fol_* entry points used by FiniteElementResponse, on host pointers (float64 only)."""

    def __init__(self, shim_lib, ne, nnode):
        self.shim, self.ne, self.nnode = shim_lib, ne, nnode

    def fol_gauss_interpolate(self, s, dt, element, num_gp, d, ne, conn, ctrl, u, kg, ug):
        return self.shim.host_gauss_interpolate(element, num_gp, d, C.c_longlong(ne), *map(C.c_void_p, (conn, ctrl, u, kg, ug)))

    def fol_response_elements(self, s, dt, element, num_gp, d, ne, xyz, conn, f, fk, fu, val, du, dk, dx):
        return self.shim.host_response_elements(element, num_gp, d, C.c_longlong(ne),
                                                *map(C.c_void_p, (xyz, conn, f, fk, fu, val, du, dk, dx)))

    def fol_residual_adjoint_elements(self, s, dt, physics, element, num_gp, acc, ne, xyz, conn, ctrl, u, lam, aux,
                                      params, dk, dx):
        return self.shim.host_residual_adjoint_elements(physics, element, num_gp, acc, C.c_longlong(ne),
                                                        *map(C.c_void_p, (xyz, conn, ctrl, u, lam, aux)), params,
                                                        C.c_void_p(dk), C.c_void_p(dx))

    def fol_residual_gather(self, s, dt, nn, nnode, width, adj_ptr, adj, elem, out):
        ap, ad = _arr(adj_ptr, nn + 1, C.c_int32), _arr(adj, self.ne * nnode, C.c_int32)
        ev, o = _arr(elem, self.ne * nnode * width), _arr(out, nn * width)
        for n in range(nn):
            for k in range(width):
                o[n * width + k] = sum(ev[int(x) * width + k] for x in ad[ap[n]:ap[n + 1]])
        return 0

    def fol_sum(self, s, dt, n, x, out):
        _arr(out, 1)[0] = _arr(x, n).sum()
        return 0


def _fake_loss(N, K_unused=None):
    coords, conn, sets = _square_mesh(N)
    bc = {"Ux": {"left": 0.0, "right": 0.05}, "Uy": {"left": 0.0, "right": 0.05}}
    didx, dval = assembly.dirichlet_vectors(["Ux", "Uy"], bc, sets)
    params = {"young_modulus": 1.0, "poisson_ratio": 0.3}
    ne, nn = conn.shape[0], coords.shape[0]
    order = np.argsort(conn.reshape(-1), kind="stable")            # entries e*a + local, ascending per node
    counts = np.bincount(conn.reshape(-1), minlength=nn)
    L = types.SimpleNamespace(
        physics="mechanical", dofs=["Ux", "Uy"], dtype=torch.float64, device=torch.device("cpu"), _dt=_lib.F64,
        _ne=ne, _nn=nn, _nnode=4, _ngauss=4, num_gp=2, number_dofs_per_node=2, total_number_of_dofs=2 * nn, dim=2,
        fe_element=types.SimpleNamespace(code=_lib.ELEMENTS["quad"]),
        fe_mesh=types.SimpleNamespace(GetNumberOfNodes=lambda: nn),
        _xyz=torch.as_tensor(coords), _conn=torch.as_tensor(conn),
        _adj_ptr=torch.as_tensor(np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)),
        _adj=torch.as_tensor(order.astype(np.int32)), _dir_idx=torch.as_tensor(didx.astype(np.int32)),
        _params=_lib.params_array([1.0, 0.3] + [0.0] * 10), Initialize=lambda reinitialize=False: None,
        dirichlet_indices=didx, dirichlet_values=dval)

    def jac_and_res(ctrl, u, transpose=False):
        data, idx, R = assembly.assemble("mechanical", "quad", 2, coords, conn, ctrl.numpy(), u.numpy(), didx, params,
                                         transpose=transpose)
        return BCOO((torch.as_tensor(data), torch.as_tensor(idx)), shape=(2 * nn, 2 * nn)), torch.as_tensor(R)

    L.ComputeJacobianMatrixAndResidualVector = jac_and_res
    return L, coords, conn, didx, dval, params


@pytest.fixture()
def cpu_backend(monkeypatch, shim):  # noqa: F811
    def install(loss):
        fake = _FakeLib(shim, loss._ne, loss._nnode)
        monkeypatch.setattr(_lib, "load", lambda: fake)
        monkeypatch.setattr(_lib, "check", lambda rc: (_ for _ in ()).throw(RuntimeError(rc)) if rc else None)
        monkeypatch.setattr(_lib, "stream_ptr", lambda: 0)
        monkeypatch.setattr(_lib, "ptr", lambda t: None if t is None else t.data_ptr())
        monkeypatch.setattr(_lib, "to_device",
                            lambda x, dtype, device=None: torch.as_tensor(np.asarray(x)).to(dtype).contiguous()
                            if not isinstance(x, torch.Tensor) else x.to(dtype).contiguous())
    return install


@pytest.fixture(scope="module")
def goldens():
    with open(os.path.join(ROOT, "tests", "golden", "reference_unit_goldens.json")) as fh:
        return json.load(fh)


def test_unit_golden_through_the_response_class(cpu_backend, goldens):
    rec = goldens["tests/unit/test_sensitivity_analysis.py"]["test_quad"]
    L, coords, conn, didx, _, _ = _fake_loss(3)
    cpu_backend(L)
    resp = FiniteElementResponse("test_response", "(E**2)*U[0]", L, NodalControl("E", L.fe_mesh))
    resp.Initialize()
    u = np.array(rec["assign"]["random_FE_UV"])
    K = np.array(rec["assign"]["random_K"])
    lam = np.array(rec["assign"]["random_adj_FE_UV"])
    jac, rhs = resp.ComputeAdjointJacobianMatrixAndRHSVector(K, u)
    a = rec["asserts"]
    np.testing.assert_allclose(jac.todense()[8, :].numpy(), a[0]["value"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(rhs.numpy(), a[1]["value"], rtol=1e-5, atol=1e-5)
    cd = resp.ComputeAdjointNodalControlDerivatives(K, u, lam)
    sd = resp.ComputeAdjointNodalShapeDerivatives(K, u, lam)
    np.testing.assert_allclose(cd.numpy(), a[2]["value"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(sd.numpy(), a[3]["value"], rtol=1e-5, atol=1e-5)
    # value: against the oracle's quadrature
    from oracle import responses
    f = responses.response_function("(E**2)*U[0]", "E", "Ux")
    ref = responses.compute_value(f, "mechanical", "quad", 2, coords, conn, K, u)
    assert abs(float(resp.ComputeValue(K, u)) - ref) <= 1e-13 * abs(ref)


def test_integration_golden_through_the_response_class(cpu_backend, goldens):
    """test_mechanical_2D_sa.py: FE solve, adjoint solve, control and shape derivatives on a 5x5-node mesh."""
    rec = goldens["tests/integration/test_mechanical_2D_sa.py"]
    K = np.array(rec["setUp"]["assign"]["random_K"])
    L, coords, conn, didx, dval, params = _fake_loss(5)
    cpu_backend(L)
    resp = FiniteElementResponse("test_response", "(E**2)*U[0]", L, NodalControl("E", L.fe_mesh))
    resp.Initialize()
    ndof = L.total_number_of_dofs
    u0 = assembly.full_dof_vector(np.zeros((1, ndof)), didx, dval)[0]
    jac, R = L.ComputeJacobianMatrixAndResidualVector(torch.as_tensor(K), torch.as_tensor(u0))
    u = u0 + np.linalg.solve(jac.todense().numpy(), -R.numpy())                # fe_linear_residual_based_solver.py:15-24
    adj_jac, adj_rhs = resp.ComputeAdjointJacobianMatrixAndRHSVector(K, u)
    lam = np.linalg.solve(adj_jac.todense().numpy(), adj_rhs.numpy())          # adjoint_fe_solver.py:19-24
    a = rec["test_sensitivites"]["asserts"]
    np.testing.assert_allclose(resp.ComputeAdjointNodalControlDerivatives(K, u, lam).numpy(), a[0]["value"],
                               rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(resp.ComputeAdjointNodalShapeDerivatives(K, u, lam).numpy(), a[1]["value"],
                               rtol=1e-5, atol=1e-5)


def test_adjoint_gradient_matches_finite_differences(cpu_backend):
    """The point of the adjoint: d value(K, u(K)) / dK from ONE extra solve equals finite differences of the
    re-solved problem (what ComputeFDNodalControlDerivatives does in the reference, fe_response.py:527-567)."""
    L, coords, conn, didx, dval, params = _fake_loss(4)
    cpu_backend(L)
    resp = FiniteElementResponse("r", "jnp.sin(E)*U[0]**2 + E*U[1]", L, NodalControl("E", L.fe_mesh))
    resp.Initialize()
    rng = np.random.default_rng(5)
    K = rng.uniform(0.3, 1.0, L._nn)
    ndof = L.total_number_of_dofs

    class Solver:
        def Solve(self, Kp, dofs):
            u0 = assembly.full_dof_vector(np.asarray(dofs, float).reshape(1, -1), didx, dval)[0]
            jac, R = L.ComputeJacobianMatrixAndResidualVector(torch.as_tensor(np.asarray(Kp, float)), torch.as_tensor(u0))
            return u0 + np.linalg.solve(jac.todense().numpy(), -R.numpy())

    u = Solver().Solve(K, np.zeros(ndof))
    adj_jac, adj_rhs = resp.ComputeAdjointJacobianMatrixAndRHSVector(K, u)
    lam = np.linalg.solve(adj_jac.todense().numpy(), adj_rhs.numpy())
    grad = resp.ComputeAdjointNodalControlDerivatives(K, u, lam).numpy()
    fd = resp.ComputeFDNodalControlDerivatives(K, Solver(), fd_step_size=1e-6, fd_mode="CD")
    assert np.abs(grad - fd).max() <= 1e-7 * np.abs(fd).max()
