"""Finite-strain Neo-Hooke block under incremental load, Newton-Raphson with the Jacobian re-assembled every iteration
(BASELINE.json configs[3]; the reference's FiniteElementNonLinearResidualBasedSolver on NeoHookeMechanicalLoss3DTetra),
everything on the GPU: assembly, de-duplication, SELL SpMV, Jacobi-BiCGSTAB.

    python examples/neo_hooke_newton_3D.py [cells_per_side]      (default 12 -> 10 368 Tet4; needs a CUDA device)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import folax_b200
from folax_b200.loss_functions import NeoHookeMechanicalLoss3DTetra
from folax_b200.solvers import FiniteElementNonLinearResidualBasedSolver


def main(n=12):
    fe_mesh = folax_b200.create_3D_tetra_box_mesh(n, n, n, 1.0, 1.0, 1.0)
    bc_dict = {"Ux": {"left": 0.0, "right": 0.25}, "Uy": {"left": 0.0, "right": 0.05}, "Uz": {"left": 0.0, "right": -0.05}}
    loss = NeoHookeMechanicalLoss3DTetra("neo_hooke_3d", {"dirichlet_bc_dict": bc_dict,
                                                          "material_dict": {"young_modulus": 1.0, "poisson_ratio": 0.3}},
                                         fe_mesh=fe_mesh)
    fe_setting = {"linear_solver_settings": {"solver": "JAX-bicgstab", "tol": 1e-10, "atol": 1e-14, "maxiter": 5000,
                                             "pre-conditioner": "jacobi"},
                  "nonlinear_solver_settings": {"rel_tol": 1e-8, "abs_tol": 1e-8, "maxiter": 10, "load_incr": 5}}
    solver = FiniteElementNonLinearResidualBasedSolver("nonlin_fe_solver", loss, fe_setting)
    loss.Initialize()
    solver.Initialize()
    K = np.ones(fe_mesh.GetNumberOfNodes())
    t0 = time.time()
    U = solver.Solve(K, np.zeros(loss.GetTotalNumberOfDOFs()))
    print(f"{fe_mesh.GetNumberOfElements('tetra')} Tet4, {loss.GetTotalNumberOfDOFs()} dofs, "
          f"{sum(len(h['res_norm']) for h in solver.convergence_history.values())} Newton iterations in "
          f"{time.time() - t0:.2f} s; max |u| = {float(U.abs().max()):.4f}")
    for step, h in solver.convergence_history.items():
        print(f"  load step {step}: residual norms " + " ".join(f"{r:.2e}" for r in h["res_norm"]))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 12)
