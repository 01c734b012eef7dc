"""What every loss provides (the contract of fol/loss_functions/loss.py:9-62): a name, `Initialize` / `Finalize`,
`GetFullDofVector(known_dofs, unknown_dofs)`, `GetNumberOfUnknowns()` and `ComputeBatchLoss(batch_params, batch_dofs)`."""


class Loss:
    _required = ("Initialize", "GetFullDofVector", "GetNumberOfUnknowns", "ComputeBatchLoss", "Finalize")

    def __init__(self, loss_name: str) -> None:
        self.initialized = False
        self._loss_name = loss_name

    def GetName(self) -> str:
        return self._loss_name

    def __new__(cls, *args, **kwargs):
        missing = [m for m in Loss._required if not callable(getattr(cls, m, None))]
        if missing:
            raise TypeError(f"Can't instantiate {cls.__name__}: it does not define {', '.join(missing)}")
        return super().__new__(cls)
