"""FiniteElementSolver: linear-solver selection of fol/solvers/fe_solver.py:22-103 on the device-resident Jacobian.

  "JAX-bicgstab"  (default)  BiCGSTAB on the GPU (folax_b200/linalg.py: SELL SpMV + vector kernels), same arguments
                             as the reference's call `bicgstab(A, -r, x0=dofs, tol, atol, maxiter)` (:62-67)
  "JAX-direct"               sparse direct solve: the duplicate-free CSR is BUILT on the GPU and factorised by SciPy
                             on the host -- the reference's spsolve path is a host factorisation as well (:70-80);
                             what changes is that 1/2.4 of the bytes cross PCIe (no duplicates) and no host sort
  "PETSc-*"                  petsc4py is not in this image: falls back to BiCGSTAB with a warning, as :47-49 does

Extra settings (not in the reference): "pre-conditioner": "jacobi" applies the diagonal to BiCGSTAB;
"device_scalars": True keeps the BiCGSTAB recurrence scalars on the device (linalg.bicgstab_device: same iterates,
one host read per "check_every" iterations instead of four per iteration); "fused": True / False runs the whole solve as
ONE persistent cooperative launch (linalg.bicgstab_fused: same recurrences, iterates equal to rounding) or forces the
multi-launch loop -- default: fused up to 3 M dofs, where the multi-launch loop is latency-bound; with "use_graph": True a batch of
"check_every" iterations is captured once in a CUDA graph and replayed (one launch per batch).
"""
import numpy as np
import torch

from .. import _lib, linalg
from ..tools import fol_error, fol_info
from .solver import Solver


class FiniteElementSolver(Solver):
    def __init__(self, fe_solver_name: str, fe_loss_function, fe_solver_settings: dict = {}) -> None:
        super().__init__(fe_solver_name)
        self.fe_loss_function = fe_loss_function
        self.fe_solver_settings = fe_solver_settings
        self.linear_solver_settings = {"solver": "JAX-bicgstab", "tol": 1e-6, "atol": 1e-6, "maxiter": 1000,
                                       "pre-conditioner": "ilu"}
        self.last_linear_solve_info = None

    def Initialize(self) -> None:
        if "linear_solver_settings" in self.fe_solver_settings.keys():
            self.linear_solver_settings = {**self.linear_solver_settings,
                                           **self.fe_solver_settings["linear_solver_settings"]}
        linear_solver = self.linear_solver_settings["solver"]
        available_linear_solver = ["PETSc-bcgsl", "PETSc-tfqmr", "PETSc-minres", "PETSc-gmres", "JAX-direct",
                                   "JAX-bicgstab"]
        if linear_solver == "JAX-direct":
            self.LinearSolve = self.JaxDirectLinearSolver
        elif linear_solver == "JAX-bicgstab":
            self.LinearSolve = self.JaxBicgstabLinearSolver
        elif linear_solver in available_linear_solver:
            fol_info("petsc4py is not available, falling back to the default iterative solver: JAX-bicgstab ")
            self.LinearSolve = self.JaxBicgstabLinearSolver
        else:
            fol_error(f"linear solver {linear_solver} does exist, available options are {available_linear_solver}")

    # names kept from the reference so that subclasses / callers that pick a method by name keep working
    def JaxBicgstabLinearSolver(self, tangent_matrix, residual_vector, dofs_vector):
        L = self.fe_loss_function
        A = linalg.SellOperator(L, tangent_matrix)
        r = _lib.to_device(residual_vector, L.dtype).reshape(-1)
        rhs = torch.empty_like(r)
        _lib.check(_lib.load().fol_vec_op(_lib.stream_ptr(), L._dt, 0, r.numel(), -1.0, _lib.ptr(r), 0.0, None,
                                          _lib.ptr(rhs)))
        s = self.linear_solver_settings
        diag = A.diagonal() if str(s.get("pre-conditioner", "")).lower() == "jacobi" else None
        fused = s.get("fused")
        if fused is None:                 # default: one launch where the multi-launch loop is latency-bound (measured:
            # 0.028 vs 0.22 ms per iteration at 53 k dofs, 0.26 vs 0.39 at 1.07 M, 2.1 vs 2.0 at 6.4 M dofs)
            fused = (L.device.type == "cuda" and rhs.numel() <= 3_000_000 and not s.get("device_scalars"))
        if fused:                         # the whole solve as one persistent launch (csrc/krylov_fused.cu)
            x, info = linalg.bicgstab_fused(A, rhs, x0=dofs_vector, tol=s["tol"], atol=s["atol"], maxiter=s["maxiter"],
                                            M_diagonal=diag)
        elif s.get("device_scalars"):     # extra setting: recurrence scalars on the device, one host read per batch
            x, info = linalg.bicgstab_device(A, rhs, x0=dofs_vector, tol=s["tol"], atol=s["atol"], maxiter=s["maxiter"],
                                             M_diagonal=diag, check_every=int(s.get("check_every", 8)),
                                             use_graph=bool(s.get("use_graph", False)))
        else:
            x, info = linalg.bicgstab(A, rhs, x0=dofs_vector, tol=s["tol"], atol=s["atol"], maxiter=s["maxiter"],
                                      M_diagonal=diag)
        self.last_linear_solve_info = info
        return x

    def JaxDirectLinearSolver(self, tangent_matrix, residual_vector, dofs_vector):
        import scipy.sparse.linalg as spla
        L = self.fe_loss_function
        A = linalg.SellOperator.__new__(linalg.SellOperator)           # CSR only: no SELL copy for a host solve
        A.loss, A.n = L, L.total_number_of_dofs
        A.indptr, A.indices, A.csr_values = L.JacobianToCSR(tangent_matrix)
        r = _lib.to_device(residual_vector, L.dtype).reshape(-1)
        delta = spla.spsolve(A.to_scipy_csr().tocsc(), -r.cpu().numpy().astype(np.float64))
        self.last_linear_solve_info = 0
        return _lib.to_device(delta, L.dtype)

    def Solve(self, current_control_vars, current_dofs):
        raise NotImplementedError

    def Finalize(self) -> None:
        pass
