"""Adjoint sensitivity analysis of a 2-D elasticity problem, end to end on the GPU -- the flow of the reference's
tests/integration/test_mechanical_2D_sa.py (and examples/sensitivity_analysis) with folax_b200's classes:

    FE solve  ->  response value  ->  adjoint solve  ->  d(response)/d(control), d(response)/d(node positions)
    (+ a finite-difference check of a few control derivatives)

    python examples/sensitivity_analysis_2D.py [N]          (N x N nodes, default 41; needs a CUDA device)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import folax_b200
from folax_b200.loss_functions import MechanicalLoss2DQuad
from folax_b200.responses import FiniteElementResponse, NodalControl
from folax_b200.solvers import AdjointFiniteElementSolver, FiniteElementLinearResidualBasedSolver


def main(N=41):
    fe_mesh = folax_b200.create_2D_square_mesh(L=1, N=N)
    bc_dict = {"Ux": {"left": 0.0, "right": 0.05}, "Uy": {"left": 0.0, "right": 0.05}}
    loss = MechanicalLoss2DQuad("mechanical_loss_2d", loss_settings={
        "dirichlet_bc_dict": bc_dict, "num_gp": 2, "material_dict": {"young_modulus": 1, "poisson_ratio": 0.3}},
        fe_mesh=fe_mesh)
    response = FiniteElementResponse("test_response", response_formula="(E**2)*U[0]", fe_loss=loss,
                                     control=NodalControl("E", fe_mesh))
    settings = {"linear_solver_settings": {"solver": "JAX-bicgstab", "tol": 1e-12, "atol": 1e-14, "maxiter": 5000,
                                           "pre-conditioner": "jacobi"}}
    fe_solver = FiniteElementLinearResidualBasedSolver("linear_fe_solver", loss, settings)
    adj_solver = AdjointFiniteElementSolver("adjoint_fe_solver", response, settings)
    for obj in (loss, response, fe_solver, adj_solver):
        obj.Initialize()

    nn, ndof = fe_mesh.GetNumberOfNodes(), loss.GetTotalNumberOfDOFs()
    x, y = fe_mesh.GetNodesX(), fe_mesh.GetNodesY()
    K = 0.55 + 0.45 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y)          # a smooth stiffness field in [0.1, 1]

    U = fe_solver.Solve(K, np.zeros(ndof))
    value = float(response.ComputeValue(K, U))
    adjoint = adj_solver.Solve(K, U, np.zeros(ndof))
    dJ_dK = response.ComputeAdjointNodalControlDerivatives(K, U, adjoint).cpu().numpy()
    dJ_dx = response.ComputeAdjointNodalShapeDerivatives(K, U, adjoint).cpu().numpy().reshape(-1, 3)
    print(f"{nn} nodes, {ndof} dofs: response = {value:.8e}, BiCGSTAB iterations (last solve) = "
          f"{adj_solver.last_linear_solve_info}")
    print(f"|dJ/dK|_max = {np.abs(dJ_dK).max():.4e},  |dJ/dx|_max = {np.abs(dJ_dx).max():.4e}")

    # central differences of the re-solved problem on a few controls (what ComputeFDNodalControlDerivatives does for all)
    h = 1e-5
    for i in np.random.default_rng(0).choice(nn, size=4, replace=False):
        Kp, Km = K.copy(), K.copy()
        Kp[i] += h
        Km[i] -= h
        fd = (float(response.ComputeValue(Kp, fe_solver.Solve(Kp, np.zeros(ndof)))) -
              float(response.ComputeValue(Km, fe_solver.Solve(Km, np.zeros(ndof))))) / (2 * h)
        print(f"  control {i:5d}: adjoint {dJ_dK[i]: .8e}   central difference {fd: .8e}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 41)
