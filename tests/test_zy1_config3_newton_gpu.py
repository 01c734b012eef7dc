"""configs[3] of BASELINE.json at test size: 3-D Neo-Hookean hyperelasticity on a tetra box, incremental
Newton-Raphson re-assembling the Jacobian every iteration -- the loop of FiniteElementNonLinearResidualBasedSolver.Solve
(fe_nonlinear_residual_based_solver.py:107-140: ApplyDirichletBCOnDofVector(load fraction) ->
ComputeJacobianMatrixAndResidualVector -> LinearSolve -> update).  The linear solve is the caller's (host SciPy on the
GPU-built duplicate-free CSR, as fe_solver.py:70-80 does with the BCOO); assembly runs on the GPU each iteration and
the whole trajectory is compared with the same loop on the oracle."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import folax_b200
from folax_b200.loss_functions import NeoHookeMechanicalLoss3DTetra
from oracle import assembly

pytestmark = pytest.mark.gpu

BC = {"Ux": {"left": 0.0, "right": 0.2}, "Uy": {"left": 0.0, "right": 0.05}, "Uz": {"left": 0.0, "right": -0.05}}
MAT = {"young_modulus": 1.0, "poisson_ratio": 0.3}


def test_neo_hooke_tetra_newton_iterations():
    mesh = folax_b200.create_3D_tetra_box_mesh(5, 5, 5, 1.0, 1.0, 1.0)
    folax_b200.perturb_interior_nodes(mesh, 0.15, 1)
    loss = NeoHookeMechanicalLoss3DTetra("nh", {"dirichlet_bc_dict": BC, "material_dict": dict(MAT)}, mesh)
    loss.Initialize()
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("tetra")
    ndof = loss.GetTotalNumberOfDOFs()
    K = np.random.default_rng(0).uniform(0.5, 1.0, mesh.GetNumberOfNodes())
    load_steps, maxiter = 3, 10

    u_gpu = np.zeros(ndof)
    u_ref = np.zeros(ndof)
    iters = 0
    for step in range(1, load_steps + 1):
        u_gpu = loss.ApplyDirichletBCOnDofVector(u_gpu, step / load_steps).cpu().numpy()
        u_ref[loss.dirichlet_indices] = step / load_steps * loss.dirichlet_values
        assert np.array_equal(u_gpu[loss.dirichlet_indices], u_ref[loss.dirichlet_indices])
        norms = []
        for _ in range(maxiter):
            jac, R = loss.ComputeJacobianMatrixAndResidualVector(K, u_gpu)
            data, idx, R_ref = assembly.assemble("neohooke", "tetra", 1, coords, conn, K, u_ref,
                                                 loss.dirichlet_indices, MAT)
            rn, rn_ref = float(np.linalg.norm(R.cpu().numpy())), float(np.linalg.norm(R_ref))
            norms.append(rn)
            assert abs(rn - rn_ref) <= 1e-9 * max(norms[0], 1e-300), (step, len(norms), rn, rn_ref)
            if rn < 1e-10:
                break
            indptr, indices, values = loss.JacobianToCSR(jac)
            A = sp.csr_array((values.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(ndof, ndof))
            u_gpu = u_gpu + spla.spsolve(A.tocsc(), -R.cpu().numpy())
            A_ref = sp.csr_array((data, (idx[:, 0], idx[:, 1])), shape=(ndof, ndof))
            u_ref = u_ref + spla.spsolve(A_ref.tocsc(), -R_ref)
            iters += 1
        assert norms[-1] < 1e-10, f"Newton did not converge in load step {step}: {norms}"
        # a consistent tangent converges quadratically: the last reduction is far better than linear
        assert len(norms) <= 6 and norms[-2] < 1e-5 and norms[-1] < 1e-3 * norms[-2]
    assert iters >= 9                                 # the Jacobian really was re-assembled every iteration
    assert np.abs(u_gpu - u_ref).max() <= 1e-9 * np.abs(u_ref).max()
