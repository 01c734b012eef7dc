"""The contract every response fulfils -- what fol/responses/response.py:9-105 declares: a named object with
`Initialize` / `Finalize`, a value, the adjoint system of that value, its adjoint-based control and shape derivatives
and their finite-difference checks."""

class Response:
    _required = ("Initialize", "ComputeValue", "ComputeAdjointJacobianMatrixAndRHSVector",
                 "ComputeAdjointNodalControlDerivatives", "ComputeAdjointNodalShapeDerivatives",
                 "ComputeFDNodalControlDerivatives", "ComputeFDNodalShapeDerivatives", "Finalize")

    def __init__(self, response_name: str) -> None:
        self.initialized = False
        self._response_name = response_name

    def GetName(self) -> str:
        return self._response_name

    def __new__(cls, *args, **kwargs):
        missing = [m for m in Response._required if not callable(getattr(cls, m, None))]
        if missing:
            raise TypeError(f"Can't instantiate {cls.__name__}: it does not define {', '.join(missing)}")
        return super().__new__(cls)
