"""Adjoint sensitivities of finite-element responses: the API of fol/responses."""
from .response import Response
from .fe_response import FiniteElementResponse, NodalControl

__all__ = ["Response", "FiniteElementResponse", "NodalControl"]
