"""Solver ABC: the same contract as fol/solvers/solver.py:10-47."""
from abc import ABC, abstractmethod


class Solver(ABC):
    def __init__(self, solver_name: str) -> None:
        self.__name = solver_name

    def GetName(self) -> str:
        return self.__name

    @abstractmethod
    def Initialize(self) -> None:
        pass

    @abstractmethod
    def Solve(self) -> None:
        pass

    @abstractmethod
    def Finalize(self) -> None:
        pass
