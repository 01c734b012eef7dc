"""Incremental Newton-Raphson of fol/solvers/fe_nonlinear_residual_based_solver.py:73-170: per load step the
Dirichlet values are scaled (ApplyDirichletBCOnDofVector), then Jacobian and residual are RE-ASSEMBLED every
iteration (BASELINE.json configs[3]) -- on the device, with the linear solve on the device-resident matrix."""
import math

from .. import _lib, linalg
from ..tools import fol_info
from .fe_linear_residual_based_solver import FiniteElementLinearResidualBasedSolver, add_vectors


class FiniteElementNonLinearResidualBasedSolver(FiniteElementLinearResidualBasedSolver):
    def __init__(self, fe_solver_name: str, fe_loss_function, fe_solver_settings: dict = {},
                 history_plot_settings: dict = {}) -> None:
        super().__init__(fe_solver_name, fe_loss_function, fe_solver_settings)
        self.nonlinear_solver_settings = {"rel_tol": 1e-8, "abs_tol": 1e-8, "maxiter": 20, "load_incr": 5}
        self.history_plot_settings = history_plot_settings          # plotting is the caller's business here
        self.convergence_history = {}

    def Initialize(self) -> None:
        super().Initialize()
        if "nonlinear_solver_settings" in self.fe_solver_settings.keys():
            self.nonlinear_solver_settings = {**self.nonlinear_solver_settings,
                                              **self.fe_solver_settings["nonlinear_solver_settings"]}

    def _newton_report(self, load_step, i, res_norm, delta_norm, converged):
        s = self.nonlinear_solver_settings
        fol_info(f"load step {load_step}  Newton iteration {i} (max {s['maxiter']})  residual norm {res_norm:.3e} "
                 f"(abs_tol {s['abs_tol']:.3e})  delta dofs norm {delta_norm:.3e} (rel_tol {s['rel_tol']:.3e})  "
                 f"converged {converged}")

    def Solve(self, current_control_vars, current_dofs_np):
        L = self.fe_loss_function
        s = self.nonlinear_solver_settings
        current_dofs = _lib.to_device(current_dofs_np, L.dtype).reshape(-1)
        num_load_steps = s["load_incr"]
        convergence_history = {}
        for load_step in range(1, num_load_steps + 1):
            current_dofs = L.ApplyDirichletBCOnDofVector(current_dofs, load_step / num_load_steps)
            convergence_history[load_step] = {"res_norm": [], "delta_dofs_norm": []}
            for i in range(1, s["maxiter"] + 1):
                BC_applied_jac, BC_applied_r = L.ComputeJacobianMatrixAndResidualVector(current_control_vars,
                                                                                        current_dofs)
                res_norm = linalg.norm(L, BC_applied_r)
                if math.isnan(res_norm):
                    raise ValueError("Residual norm contains NaN values.")
                delta_dofs = self.LinearSolve(BC_applied_jac, BC_applied_r, current_dofs)
                delta_norm = linalg.norm(L, delta_dofs)
                newton_converged = (res_norm < s["abs_tol"] or delta_norm < s["rel_tol"] or i == s["maxiter"])
                self._newton_report(load_step, i, res_norm, delta_norm, newton_converged)
                convergence_history[load_step]["res_norm"].append(res_norm)
                convergence_history[load_step]["delta_dofs_norm"].append(delta_norm)
                if newton_converged:            # the reference leaves the loop BEFORE applying the last update (:162-166)
                    break
                current_dofs = add_vectors(L, current_dofs, delta_dofs)
        self.convergence_history = convergence_history
        return current_dofs
