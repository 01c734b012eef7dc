"""Newton-Raphson with Gauss-point history (elastoplasticity):
fol/solvers/fe_nonlinear_residual_based_solver_with_history_update.py:37-164.  The state returned by an assembly is
kept only when the iteration goes on (:143-145) -- exactly as the reference commits it."""
import math

import torch

from .. import _lib, linalg
from .fe_linear_residual_based_solver import add_vectors
from .fe_nonlinear_residual_based_solver import FiniteElementNonLinearResidualBasedSolver


class FiniteElementNonLinearResidualBasedSolverWithStateUpdate(FiniteElementNonLinearResidualBasedSolver):
    def Solve(self, current_control_vars, current_dofs, current_state=None, return_all_steps: bool = False):
        L = self.fe_loss_function
        s = self.nonlinear_solver_settings
        current_dofs = _lib.to_device(current_dofs, L.dtype).reshape(-1)
        num_load_steps = s["load_incr"]
        if current_state is None:
            current_state = torch.zeros(L.GetStateShape(), dtype=L.dtype, device=L.device)
        else:
            current_state = _lib.to_device(current_state, L.dtype)
        solution_history_dict = {}
        load_steps_solutions, load_steps_states = None, None
        for load_step in range(1, num_load_steps + 1):
            current_dofs = L.ApplyDirichletBCOnDofVector(current_dofs, load_step / num_load_steps)
            solution_history_dict[load_step] = {"res_norm": [], "delta_dofs_norm": []}
            for i in range(1, s["maxiter"] + 1):
                new_state, BC_applied_jac, BC_applied_r = L.ComputeJacobianMatrixAndResidualVector(
                    current_control_vars, current_dofs, old_state_gps=current_state)
                res_norm = linalg.norm(L, BC_applied_r)
                if math.isnan(res_norm):
                    raise ValueError("Residual norm contains NaN values.")
                delta_dofs = self.LinearSolve(BC_applied_jac, BC_applied_r, current_dofs)
                delta_norm = linalg.norm(L, delta_dofs)
                newton_converged = (res_norm < s["abs_tol"] or delta_norm < s["rel_tol"] or i == s["maxiter"])
                self._newton_report(load_step, i, res_norm, delta_norm, newton_converged)
                solution_history_dict[load_step]["res_norm"].append(res_norm)
                solution_history_dict[load_step]["delta_dofs_norm"].append(delta_norm)
                if newton_converged:
                    break
                current_dofs = add_vectors(L, current_dofs, delta_dofs)
                current_state = new_state
            if return_all_steps:
                if load_step == 1:
                    load_steps_solutions = current_dofs.clone()
                    load_steps_states = current_state.clone()[None, ...]
                else:
                    load_steps_solutions = torch.vstack([load_steps_solutions, current_dofs])
                    load_steps_states = torch.vstack([load_steps_states, current_state[None, ...]])
            else:
                load_steps_solutions = current_dofs.clone()
                load_steps_states = current_state.clone()
        return load_steps_solutions, load_steps_states, solution_history_dict
