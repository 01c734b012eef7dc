"""Finite-element solvers with the API of fol/solvers, running on the device-resident Jacobian."""
from .solver import Solver
from .fe_solver import FiniteElementSolver
from .fe_linear_residual_based_solver import FiniteElementLinearResidualBasedSolver
from .fe_nonlinear_residual_based_solver import FiniteElementNonLinearResidualBasedSolver
from .fe_nonlinear_residual_based_solver_with_history_update import \
    FiniteElementNonLinearResidualBasedSolverWithStateUpdate
from .adjoint_fe_solver import AdjointFiniteElementSolver

__all__ = ["Solver", "FiniteElementSolver", "FiniteElementLinearResidualBasedSolver",
           "FiniteElementNonLinearResidualBasedSolver", "FiniteElementNonLinearResidualBasedSolverWithStateUpdate",
           "AdjointFiniteElementSolver"]
