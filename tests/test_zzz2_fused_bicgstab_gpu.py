"""One-launch BiCGSTAB (csrc/krylov_fused.cu, linalg.bicgstab_fused) against the multi-launch host-scalar loop
(linalg.bicgstab) and SciPy's direct solve: same recurrences and stopping rule, dot products summed per CTA instead of
per fixed block, so iterates agree to rounding (not bit for bit) and the stopping iteration may move by one or two when
|r|^2 sits at the threshold.  Block columns (3 dofs per node), scalar columns (thermal, 1 dof per node), both
precisions, Jacobi or no preconditioner, converged / maxiter-stopped runs, run-to-run determinism."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla
import torch

from folax_b200 import linalg
from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu


def _system(physics, element, n, dtype="float64"):
    mesh = gh.make_mesh(element, n, perturb=0.2, seed=3)
    loss = gh.make_loss(physics, element, mesh, num_gp=2, dtype=dtype)
    K, _ = gh.fields(physics, mesh, loss, seed=1)
    u0 = loss.ApplyDirichletBCOnDofVector(np.zeros(loss.total_number_of_dofs))
    jac, R = loss.ComputeJacobianMatrixAndResidualVector(K, u0)
    return loss, linalg.SellOperator(loss, jac), u0, -R


@pytest.mark.parametrize("physics,element,n", [("mechanical", "hexahedron", 6), ("thermal", "quad", 24),
                                                ("mechanical", "quad", 20)])
@pytest.mark.parametrize("precond", [None, "jacobi"])
def test_fused_solve_equals_the_multi_launch_loop(physics, element, n, precond):
    loss, A, u0, rhs = _system(physics, element, n)
    diag = A.diagonal() if precond else None
    x_h, k_h = linalg.bicgstab(A, rhs, x0=u0, tol=1e-11, atol=0.0, maxiter=5000, M_diagonal=diag)
    x_f, k_f = linalg.bicgstab_fused(A, rhs, x0=u0, tol=1e-11, atol=0.0, maxiter=5000, M_diagonal=diag)
    assert k_h > 0 and k_f > 0 and abs(k_f - k_h) <= 5 + k_h // 4, (k_f, k_h)
    scale = x_h.abs().max().item()
    assert (x_f - x_h).abs().max().item() <= 1e-8 * scale
    # and both solve the system: direct solve of the same de-duplicated matrix
    ref = spla.spsolve(A.to_scipy_csr().tocsc(), rhs.cpu().numpy())
    assert np.abs(x_f.cpu().numpy() - ref).max() <= 1e-7 * max(np.abs(ref).max(), 1e-30)
    # deterministic run to run (same grid, same per-CTA partial sums)
    x_f2, k_f2 = linalg.bicgstab_fused(A, rhs, x0=u0, tol=1e-11, atol=0.0, maxiter=5000, M_diagonal=diag)
    assert k_f2 == k_f and torch.equal(x_f2, x_f)


def test_fused_solve_stops_at_maxiter_and_on_a_converged_start():
    loss, A, u0, rhs = _system("mechanical", "hexahedron", 5)
    x_f, k_f = linalg.bicgstab_fused(A, rhs, x0=u0, tol=1e-14, atol=0.0, maxiter=7)
    x_h, k_h = linalg.bicgstab(A, rhs, x0=u0, tol=1e-14, atol=0.0, maxiter=7)
    assert k_f == 7 and k_h == 7
    assert (x_f - x_h).abs().max().item() <= 1e-9 * x_h.abs().max().item()
    # start from the solution: zero iterations, x returned unchanged
    x_s, _ = linalg.bicgstab(A, rhs, x0=u0, tol=1e-13, atol=0.0, maxiter=5000)
    x_0, k_0 = linalg.bicgstab_fused(A, rhs, x0=x_s, tol=1e-6, atol=0.0, maxiter=50)
    assert k_0 == 0 and torch.equal(x_0, x_s)


def test_fused_solve_float32():
    loss, A, u0, rhs = _system("mechanical", "hexahedron", 5, dtype="float32")
    diag = A.diagonal()
    x_f, k_f = linalg.bicgstab_fused(A, rhs, x0=u0, tol=1e-5, atol=0.0, maxiter=2000, M_diagonal=diag)
    assert 0 < k_f < 2000
    ref = spla.spsolve(A.to_scipy_csr().astype(np.float64).tocsc(), rhs.cpu().numpy().astype(np.float64))
    assert np.abs(x_f.cpu().numpy() - ref).max() <= 2e-3 * np.abs(ref).max()


def test_solver_class_setting_fused():
    """FiniteElementLinearResidualBasedSolver with the extra setting "fused": same solution as the default loop."""
    import folax_b200
    from folax_b200.loss_functions import MechanicalLoss2DQuad
    from folax_b200.solvers import FiniteElementLinearResidualBasedSolver
    mesh = folax_b200.create_2D_square_mesh(1.0, 31)
    bc = {"Ux": {"left": 0.0, "right": 0.05}, "Uy": {"left": 0.0, "right": 0.05}}
    sols = []
    for fused in (False, True):
        loss = MechanicalLoss2DQuad("m2d", {"dirichlet_bc_dict": bc, "num_gp": 2, "material_dict": dict(gh.MATERIAL)}, mesh)
        solver = FiniteElementLinearResidualBasedSolver("lin", loss, {"linear_solver_settings": {
            "solver": "JAX-bicgstab", "tol": 1e-11, "atol": 1e-14, "maxiter": 5000, "fused": fused}})
        loss.Initialize()
        solver.Initialize()
        K = np.random.default_rng(25).uniform(0.1, 1.0, mesh.GetNumberOfNodes())
        sols.append(solver.Solve(K, np.zeros(loss.GetTotalNumberOfDOFs())).cpu().numpy())
        assert solver.last_linear_solve_info > 0
    assert np.abs(sols[0] - sols[1]).max() <= 1e-8 * np.abs(sols[0]).max()
