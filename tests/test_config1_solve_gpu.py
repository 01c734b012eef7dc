"""configs[0]: examples/mechanical_square -- 2-D linear elasticity on a generated 50x50 quad mesh, FE residual +
Jacobian assembly and the linear solve of FiniteElementLinearResidualBasedSolver.Solve
(fe_linear_residual_based_solver.py:15-24, fe_solver.py:70-80: BCOO -> scipy CSR -> sparse direct solve with -R).
The solver itself stays on the host (out of scope); what is checked is that the assembled system handed to it gives
the oracle's solution, through the BCOO, through the duplicate-free CSR hand-off and matrix-free."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import folax_b200
from folax_b200.loss_functions import MechanicalLoss2DQuad
from oracle import assembly

pytestmark = pytest.mark.gpu


def test_mechanical_square_assembly_and_linear_solve():
    mesh = folax_b200.create_2D_square_mesh(1.0, 51)                        # 2 500 quads, 5 202 dofs
    bc = {"Ux": {"left": 0.0, "right": 0.05}, "Uy": {"left": 0.0, "right": 0.05}}
    mat = {"young_modulus": 1.0, "poisson_ratio": 0.3}
    loss = MechanicalLoss2DQuad("mechanical_loss_2d", {"dirichlet_bc_dict": bc, "num_gp": 2, "material_dict": mat}, mesh)
    loss.Initialize()
    nn, ndof = mesh.GetNumberOfNodes(), loss.GetTotalNumberOfDOFs()
    assert (mesh.GetNumberOfElements("quad"), nn, ndof) == (2500, 2601, 5202)
    rng = np.random.default_rng(25)
    K = rng.uniform(0.1, 1.0, nn)                                            # a heterogeneous stiffness field
    u_bc = loss.ApplyDirichletBCOnDofVector(np.zeros(ndof))
    jac, R = loss.ComputeJacobianMatrixAndResidualVector(K, u_bc)
    assert jac.data.numel() == 160000 and tuple(jac.shape) == (ndof, ndof)

    # the reference's own solve on the oracle's system
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("quad")
    u0 = assembly.full_dof_vector(np.zeros((1, ndof)), loss.dirichlet_indices, loss.dirichlet_values)[0]
    data, idx, Rref = assembly.assemble("mechanical", "quad", 2, coords, conn, K, u0, loss.dirichlet_indices, mat)
    A_ref = sp.csr_array((data, (idx[:, 0], idx[:, 1])), shape=(ndof, ndof))
    u_ref = u0 + spla.spsolve(A_ref.tocsc(), -Rref)

    # 1. through the BCOO exactly as fe_solver.py:71-80 does
    A = sp.csr_array((jac.data.cpu().numpy(), (jac.indices[:, 0].cpu().numpy(), jac.indices[:, 1].cpu().numpy())),
                     shape=(ndof, ndof))
    u1 = u_bc.cpu().numpy() + spla.spsolve(A.tocsc(), -R.cpu().numpy())
    scale = np.abs(u_ref).max()
    assert np.abs(u1 - u_ref).max() <= 1e-9 * scale
    # 2. through the duplicate-free CSR built on the GPU
    indptr, indices, values = loss.JacobianToCSR(jac)
    A2 = sp.csr_array((values.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(ndof, ndof))
    u2 = u_bc.cpu().numpy() + spla.spsolve(A2.tocsc(), -R.cpu().numpy())
    assert np.abs(u2 - u_ref).max() <= 1e-9 * scale
    # 3. matrix-free: a Krylov solve whose operator is the GPU product J v (no matrix assembled)
    op = spla.LinearOperator((ndof, ndof), dtype=np.float64,
                             matvec=lambda v: loss.ApplyJacobian(K, u_bc, v).cpu().numpy())
    du, info = spla.gmres(op, -R.cpu().numpy(), rtol=1e-12, atol=0.0, restart=200, maxiter=20,
                          M=sp.diags(1.0 / A2.diagonal()))
    assert info == 0
    assert np.abs(u_bc.cpu().numpy() + du - u_ref).max() <= 1e-7 * scale
    # the solution satisfies the discrete equilibrium: free-dof residual vanishes, Dirichlet values are kept
    _, R_end = loss.ComputeJacobianMatrixAndResidualVector(K, u1)
    free = loss.non_dirichlet_indices
    assert np.abs(R_end.cpu().numpy()[free]).max() <= 1e-10 * np.abs(Rref).max()
    assert np.array_equal(u1[loss.dirichlet_indices], u0[loss.dirichlet_indices])
