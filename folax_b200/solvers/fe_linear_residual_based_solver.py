"""fol/solvers/fe_linear_residual_based_solver.py:11-24."""
import torch

from .. import _lib
from .fe_solver import FiniteElementSolver


def add_vectors(loss, x, y):
    """x + y on the device (the solvers' dof updates), through the C ABI."""
    out = torch.empty_like(x)
    _lib.check(_lib.load().fol_vec_op(_lib.stream_ptr(), loss._dt, 0, x.numel(), 1.0, _lib.ptr(x), 1.0, _lib.ptr(y),
                                      _lib.ptr(out)))
    return out


class FiniteElementLinearResidualBasedSolver(FiniteElementSolver):
    def Solve(self, current_control_vars, current_dofs):
        L = self.fe_loss_function
        BC_applied_dofs = L.ApplyDirichletBCOnDofVector(current_dofs)
        BC_applied_jac, BC_applied_r = L.ComputeJacobianMatrixAndResidualVector(current_control_vars, BC_applied_dofs)
        delta_dofs = self.LinearSolve(BC_applied_jac, BC_applied_r, BC_applied_dofs)
        return add_vectors(L, BC_applied_dofs, delta_dofs)
