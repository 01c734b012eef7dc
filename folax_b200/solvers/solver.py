"""The contract of a solver (fol/solvers/solver.py:10-47): a named object with `Initialize`, `Solve(control_vars,
dofs)` and `Finalize`; `BatchSolve` maps `Solve` over the leading axis of both arguments."""
import torch


class Solver:
    _required = ("Initialize", "Solve", "Finalize")

    def __init__(self, solver_name: str) -> None:
        self._solver_name = solver_name

    def GetName(self) -> str:
        return self._solver_name

    def BatchSolve(self, batch_control_vars, batch_dofs):
        """solver.py:38-39 (`jax.vmap(self.Solve)`): one solve per sample, stacked."""
        return torch.stack([self.Solve(c, d) for c, d in zip(batch_control_vars, batch_dofs)]).squeeze()

    def __new__(cls, *args, **kwargs):
        missing = [m for m in Solver._required if not callable(getattr(cls, m, None))]
        if missing:
            raise TypeError(f"Can't instantiate {cls.__name__}: it does not define {', '.join(missing)}")
        return super().__new__(cls)
