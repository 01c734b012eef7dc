"""Loss ABC: the same contract as fol/loss_functions/loss.py:9-62."""
from abc import ABC, abstractmethod


class Loss(ABC):
    def __init__(self, loss_name: str) -> None:
        self.__name = loss_name
        self.initialized = False

    def GetName(self) -> str:
        return self.__name

    @abstractmethod
    def Initialize(self) -> None:
        pass

    @abstractmethod
    def GetFullDofVector(self, known_dofs, unknown_dofs):
        pass

    @abstractmethod
    def GetNumberOfUnknowns(self) -> int:
        pass

    @abstractmethod
    def ComputeBatchLoss(self) -> None:
        pass

    @abstractmethod
    def Finalize(self) -> None:
        pass
