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
from tests.cpu_backend import fake_loss
from tests.test_oracle_golden import _square_mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_loss(N):
    coords, conn, sets = _square_mesh(N)
    bc = {"Ux": {"left": 0.0, "right": 0.05}, "Uy": {"left": 0.0, "right": 0.05}}
    params = {"young_modulus": 1.0, "poisson_ratio": 0.3}
    L = fake_loss("mechanical", "quad", 2, coords, conn, sets, ["Ux", "Uy"], bc, params, [1.0, 0.3] + [0.0] * 10)
    return L, coords, conn, L.dirichlet_indices, L.dirichlet_values, params


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
