"""GPU runs of folax_b200.solvers / folax_b200.linalg: SELL SpMV, BiCGSTAB and the solver classes
of fol/solvers on the device-resident Jacobian, against SciPy, the oracle and the reference's integration golden."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

import folax_b200
from folax_b200 import linalg
from folax_b200.loss_functions import (ElastoplasticityLoss2DQuad, MechanicalLoss2DQuad, MechanicalLoss3DHexa,
                                       NeoHookeMechanicalLoss3DTetra)
from folax_b200.responses import FiniteElementResponse, NodalControl
from folax_b200.solvers import (AdjointFiniteElementSolver, FiniteElementLinearResidualBasedSolver,
                                FiniteElementNonLinearResidualBasedSolver,
                                FiniteElementNonLinearResidualBasedSolverWithStateUpdate)
from oracle import assembly
from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BC2 = {"Ux": {"left": 0.0, "right": 0.05}, "Uy": {"left": 0.0, "right": 0.05}}


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-13), ("float32", 2e-5)])
def test_sell_spmv_matches_scipy(dtype, tol):
    mesh = gh.make_mesh("hexahedron", 5, perturb=0.2, seed=1)
    loss = gh.make_loss("mechanical", "hexahedron", mesh, num_gp=2, dtype=dtype)
    K, u = gh.fields("mechanical", mesh, loss, seed=2)
    jac, _ = loss.ComputeJacobianMatrixAndResidualVector(K, u)
    A = linalg.SellOperator(loss, jac)
    ref = A.to_scipy_csr().astype(np.float64)
    x = np.random.default_rng(3).standard_normal(loss.total_number_of_dofs)
    xt = torch.as_tensor(x, device="cuda").to(loss.dtype)
    y = A.matvec(xt)
    want = ref @ xt.cpu().numpy().astype(np.float64)
    assert np.abs(y.cpu().numpy() - want).max() <= tol * np.abs(want).max()
    assert torch.equal(y, A.matvec(xt))                                   # deterministic
    assert A.plan["node_cols"] is not None                                # 3 dofs per node: the block-column kernel ran
    A.use_block_kernel = False
    assert torch.equal(y, A.matvec(xt))                                   # scalar-column kernel: bit-identical
    A.use_block_kernel = True
    assert np.array_equal(A.diagonal().cpu().numpy(), ref.diagonal().astype(y.cpu().numpy().dtype))
    # the same product without the matrix (ApplyJacobian) agrees
    y_mf = loss.ApplyJacobian(K, u, xt)
    assert np.abs(y_mf.cpu().numpy() - want).max() <= max(tol, 1e-12) * np.abs(want).max()


def test_dot_and_vector_kernels():
    from folax_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for n in (1, 1000, 592 * 256 * 3 + 17):
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        xt, yt = torch.as_tensor(x, device="cuda"), torch.as_tensor(y, device="cuda")
        v = linalg._Vectors(_lib.F64, n, torch.float64, xt.device)
        got = v.dot(xt, yt)
        assert abs(got - float(np.dot(x, y))) <= 1e-12 * np.sqrt(n) * max(1.0, abs(np.dot(x, y)))
        assert got == v.dot(xt, yt)                                        # fixed reduction tree
        out = torch.empty_like(xt)
        v.axpby(2.0, xt, -0.5, yt, out)
        assert np.allclose(out.cpu().numpy(), 2.0 * x - 0.5 * y, rtol=1e-15, atol=1e-15)
        v.axpby(3.0, xt, 0.0, None, out)
        assert np.array_equal(out.cpu().numpy(), 3.0 * x)
        v.divide(xt, yt, out)
        assert np.allclose(out.cpu().numpy(), x / y, rtol=1e-15)
        v.axpby(1.0, xt, 1.0, yt, xt)                                       # in place
        assert np.array_equal(xt.cpu().numpy(), x + y)


@pytest.mark.parametrize("precond", ["ilu", "jacobi"])
def test_config0_linear_solve_with_device_bicgstab(precond):
    """BASELINE.json configs[0] (examples/mechanical_square, 50x50 quads) through the solver class with the
    reference's default linear solver, on the GPU end to end; checked against SciPy's direct solve of the oracle's
    system."""
    mesh = folax_b200.create_2D_square_mesh(1.0, 51)
    loss = MechanicalLoss2DQuad("mechanical_loss_2d", {"dirichlet_bc_dict": BC2, "num_gp": 2,
                                                       "material_dict": dict(gh.MATERIAL)}, mesh)
    solver = FiniteElementLinearResidualBasedSolver("lin", loss, {"linear_solver_settings": {
        "solver": "JAX-bicgstab", "tol": 1e-11, "atol": 1e-14, "maxiter": 5000, "pre-conditioner": precond}})
    loss.Initialize()
    solver.Initialize()
    nn, ndof = mesh.GetNumberOfNodes(), loss.GetTotalNumberOfDOFs()
    K = np.random.default_rng(25).uniform(0.1, 1.0, nn)
    u = solver.Solve(K, np.zeros(ndof)).cpu().numpy()
    assert solver.last_linear_solve_info > 0
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("quad")
    u0 = assembly.full_dof_vector(np.zeros((1, ndof)), loss.dirichlet_indices, loss.dirichlet_values)[0]
    data, idx, R = assembly.assemble("mechanical", "quad", 2, coords, conn, K, u0, loss.dirichlet_indices, gh.MATERIAL)
    A = sp.csr_array((data, (idx[:, 0], idx[:, 1])), shape=(ndof, ndof))
    ref = u0 + spla.spsolve(A.tocsc(), -R)
    assert np.abs(u - ref).max() <= 1e-7 * np.abs(ref).max()


@pytest.mark.parametrize("precond", [None, "jacobi"])
def test_device_scalar_bicgstab_equals_the_host_scalar_loop(precond):
    """linalg.bicgstab_device (recurrence scalars on the device, gated vector kernels, one host read per batch) against
    linalg.bicgstab: same stopping iteration, bit-identical solution."""
    mesh = gh.make_mesh("hexahedron", 4, perturb=0.2, seed=3)
    loss = gh.make_loss("mechanical", "hexahedron", mesh, num_gp=2)
    K, _ = gh.fields("mechanical", mesh, loss, seed=1)
    u0 = loss.ApplyDirichletBCOnDofVector(np.zeros(loss.total_number_of_dofs))
    jac, R = loss.ComputeJacobianMatrixAndResidualVector(K, u0)
    A = linalg.SellOperator(loss, jac)
    diag = A.diagonal() if precond else None
    rhs = -R
    for tol, maxiter in ((1e-10, 3000), (1e-10, 5), (1e-3, 3000)):
        x_h, k_h = linalg.bicgstab(A, rhs, x0=u0, tol=tol, atol=0.0, maxiter=maxiter, M_diagonal=diag)
        for every in (1, 8):
            x_d, k_d = linalg.bicgstab_device(A, rhs, x0=u0, tol=tol, atol=0.0, maxiter=maxiter, M_diagonal=diag,
                                              check_every=every)
            assert k_d == k_h, (tol, maxiter, every, k_d, k_h)
            assert torch.equal(x_d, x_h)


def test_reference_integration_golden_through_the_solver_classes():
    """tests/integration/test_mechanical_2D_sa.py:21-113, same objects, same settings."""
    with open(os.path.join(ROOT, "tests", "golden", "reference_unit_goldens.json")) as fh:
        rec = json.load(fh)["tests/integration/test_mechanical_2D_sa.py"]
    K = np.array(rec["setUp"]["assign"]["random_K"])
    mesh = folax_b200.create_2D_square_mesh(1.0, 5)
    loss = MechanicalLoss2DQuad("mechanical_loss_2d", {"dirichlet_bc_dict": BC2, "num_gp": 2,
                                                       "material_dict": {"young_modulus": 1, "poisson_ratio": 0.3}}, mesh)
    resp = FiniteElementResponse("test_response", "(E**2)*U[0]", loss, NodalControl("E", mesh))
    fe_setting = {"linear_solver_settings": {"solver": "JAX-direct", "tol": 1e-6, "atol": 1e-6, "maxiter": 1000,
                                             "pre-conditioner": "ilu"},
                  "nonlinear_solver_settings": {"rel_tol": 1e-5, "abs_tol": 1e-5, "maxiter": 10, "load_incr": 5}}
    lin = FiniteElementLinearResidualBasedSolver("linear_fe_solver", loss, fe_setting)
    adj = AdjointFiniteElementSolver("first_adj_fe_solver", resp, {"linear_solver_settings": {"solver": "JAX-direct"}})
    loss.Initialize()
    resp.Initialize()
    lin.Initialize()
    adj.Initialize()
    ndof = 2 * mesh.GetNumberOfNodes()
    FE_UV = lin.Solve(K, np.zeros(ndof))
    FE_adj_UV = adj.Solve(K, FE_UV, np.ones(ndof))
    a = rec["test_sensitivites"]["asserts"]
    np.testing.assert_allclose(resp.ComputeAdjointNodalControlDerivatives(K, FE_UV, FE_adj_UV).cpu().numpy(),
                               a[0]["value"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(resp.ComputeAdjointNodalShapeDerivatives(K, FE_UV, FE_adj_UV).cpu().numpy(),
                               a[1]["value"], rtol=1e-5, atol=1e-5)


def test_config3_newton_solver_neo_hooke_tetra():
    """BASELINE.json configs[3] at test size through FiniteElementNonLinearResidualBasedSolver with the device
    BiCGSTAB: assembly, de-duplication, SELL conversion and the Krylov iteration all on the GPU, Jacobian
    re-assembled every Newton iteration; against the same loop on the oracle with SciPy's direct solver."""
    mesh = folax_b200.create_3D_tetra_box_mesh(5, 5, 5, 1.0, 1.0, 1.0)
    folax_b200.perturb_interior_nodes(mesh, 0.15, 1)
    bc = {"Ux": {"left": 0.0, "right": 0.2}, "Uy": {"left": 0.0, "right": 0.05}, "Uz": {"left": 0.0, "right": -0.05}}
    mat = {"young_modulus": 1.0, "poisson_ratio": 0.3}
    loss = NeoHookeMechanicalLoss3DTetra("nh", {"dirichlet_bc_dict": bc, "material_dict": dict(mat)}, mesh)
    settings = {"linear_solver_settings": {"solver": "JAX-bicgstab", "tol": 1e-12, "atol": 1e-15, "maxiter": 3000,
                                           "pre-conditioner": "jacobi"},
                "nonlinear_solver_settings": {"rel_tol": 1e-10, "abs_tol": 1e-10, "maxiter": 10, "load_incr": 3}}
    solver = FiniteElementNonLinearResidualBasedSolver("nl", loss, settings)
    loss.Initialize()
    solver.Initialize()
    ndof = loss.GetTotalNumberOfDOFs()
    K = np.random.default_rng(0).uniform(0.5, 1.0, mesh.GetNumberOfNodes())
    u = solver.Solve(K, np.zeros(ndof)).cpu().numpy()
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("tetra")
    ref = np.zeros(ndof)
    for step in range(1, 4):
        ref[loss.dirichlet_indices] = step / 3 * loss.dirichlet_values
        for i in range(1, 11):
            data, idx, R = assembly.assemble("neohooke", "tetra", 1, coords, conn, K, ref, loss.dirichlet_indices, mat)
            A = sp.csr_array((data, (idx[:, 0], idx[:, 1])), shape=(ndof, ndof))
            du = spla.spsolve(A.tocsc(), -R)
            if np.linalg.norm(R) < 1e-10 or np.linalg.norm(du) < 1e-10 or i == 10:
                break
            ref = ref + du
    assert np.abs(u - ref).max() <= 1e-7 * np.abs(ref).max()
    hist = solver.convergence_history
    assert all(len(hist[s]["res_norm"]) >= 4 and hist[s]["res_norm"][-1] < 1e-8 for s in (1, 2, 3))


def test_newton_with_state_update_elastoplastic_quad():
    """fe_nonlinear_residual_based_solver_with_history_update.py on a J2 quad mesh: the solver's final dofs and
    committed Gauss-point state equal the same loop on the oracle."""
    mat = {"young_modulus": 3.0, "poisson_ratio": 0.3, "iso_hardening_parameter_1": 0.4,
           "iso_hardening_param_2": 10.0, "yield_limit": 0.2}
    mesh = folax_b200.create_2D_square_mesh(1.0, 7)
    bc = {"Ux": {"left": 0.0, "right": 0.15}, "Uy": {"left": 0.0, "right": 0.02}}
    loss = ElastoplasticityLoss2DQuad("ep", {"dirichlet_bc_dict": bc, "material_dict": dict(mat)}, mesh)
    settings = {"linear_solver_settings": {"solver": "JAX-direct"},
                "nonlinear_solver_settings": {"rel_tol": 1e-10, "abs_tol": 1e-10, "maxiter": 12, "load_incr": 3}}
    solver = FiniteElementNonLinearResidualBasedSolverWithStateUpdate("ep_solver", loss, settings)
    loss.Initialize()
    solver.Initialize()
    ndof, nn = loss.GetTotalNumberOfDOFs(), mesh.GetNumberOfNodes()
    u, state, hist = solver.Solve(np.ones(nn), np.zeros(ndof))
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("quad")
    ref, st = np.zeros(ndof), np.zeros(loss.GetStateShape())
    for step in range(1, 4):
        ref[loss.dirichlet_indices] = step / 3 * loss.dirichlet_values
        for i in range(1, 13):
            new_st, data, idx, R = assembly.assemble_j2("quad", 2, coords, conn, ref, st, loss.dirichlet_indices, mat)
            A = sp.csr_array((data, (idx[:, 0], idx[:, 1])), shape=(ndof, ndof))
            du = spla.spsolve(A.tocsc(), -R)
            if np.linalg.norm(R) < 1e-10 or np.linalg.norm(du) < 1e-10 or i == 12:
                break
            ref, st = ref + du, new_st
    assert np.abs(u.cpu().numpy() - ref).max() <= 1e-8 * np.abs(ref).max()
    assert np.abs(state.cpu().numpy() - st).max() <= 1e-8 * max(np.abs(st).max(), 1e-300)
    assert st[..., -1].max() > 0.0, "the test must reach the plastic range"
    assert set(hist) == {1, 2, 3}
