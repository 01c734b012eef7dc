"""fol/solvers/adjoint_fe_solver.py:11-24: J^T lam = rhs of a FiniteElementResponse."""
import torch

from .. import _lib
from .fe_solver import FiniteElementSolver


class AdjointFiniteElementSolver(FiniteElementSolver):
    def __init__(self, adj_fe_solver_name: str, fe_response, adj_fe_solver_settings: dict = {}) -> None:
        super().__init__(adj_fe_solver_name, fe_response.fe_loss, adj_fe_solver_settings)
        self.fe_response = fe_response

    def Solve(self, current_control_vars, current_dofs, current_adjoint_dofs):
        L = self.fe_loss_function
        BC_applied_jac, BC_applied_rhs = self.fe_response.ComputeAdjointJacobianMatrixAndRHSVector(
            current_control_vars, current_dofs)
        # the linear solvers negate their right-hand side (:21-23)
        neg = torch.empty_like(BC_applied_rhs)
        _lib.check(_lib.load().fol_vec_op(_lib.stream_ptr(), L._dt, 0, neg.numel(), -1.0, _lib.ptr(BC_applied_rhs),
                                          0.0, None, _lib.ptr(neg)))
        return self.LinearSolve(BC_applied_jac, neg, _lib.to_device(current_adjoint_dofs, L.dtype).reshape(-1))
