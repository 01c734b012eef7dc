"""Small-strain linear elasticity with a nodal Young's-modulus control field.
Same classes and settings as fol/loss_functions/mechanical.py:14-145."""
import numpy as np

from ..tools import fol_error
from .fe_loss import FiniteElementLoss


class MechanicalLoss(FiniteElementLoss):
    physics = "mechanical"
    _second_order = "zero"
    _has_control_gradient = False  # Se sits inside stop_gradient (mechanical.py:116): dE/dK = 0

    def Initialize(self, reinitialize=False) -> None:
        if "material_dict" not in self.loss_settings.keys():
            fol_error("material_dict should provided in the loss settings !", self.GetName())
        super().Initialize(reinitialize)

    def _material_params(self):
        md = self.loss_settings["material_dict"]
        p = super()._material_params()
        p[0], p[1] = float(md["young_modulus"]), float(md["poisson_ratio"])
        body = np.zeros(3)
        if "body_foce" in self.loss_settings:  # (sic) mechanical.py:26
            b = np.asarray(self.loss_settings["body_foce"], dtype=float).reshape(-1)
            body[: b.size] = b
        self.body_force = body[: self.loss_settings["compute_dims"]].reshape(-1, 1)
        p[2:5] = body.tolist()
        return p

    def _energy_and_grads(self, batch_params, batch_dofs, **kw):
        energy, grad_u, _ = super()._energy_and_grads(batch_params, batch_dofs, **kw)
        return energy, grad_u, None


class MechanicalLoss3DTetra(MechanicalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        super().__init__(name, {**loss_settings, "compute_dims": 3, "ordered_dofs": ["Ux", "Uy", "Uz"],
                                "element_type": "tetra"}, fe_mesh)


class MechanicalLoss3DHexa(MechanicalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        if "num_gp" not in loss_settings.keys():
            loss_settings["num_gp"] = 2
        super().__init__(name, {**loss_settings, "compute_dims": 3, "ordered_dofs": ["Ux", "Uy", "Uz"],
                                "element_type": "hexahedron"}, fe_mesh)


class MechanicalLoss2DTri(MechanicalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        super().__init__(name, {**loss_settings, "compute_dims": 2, "ordered_dofs": ["Ux", "Uy"],
                                "element_type": "triangle"}, fe_mesh)


class MechanicalLoss2DQuad(MechanicalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        if "num_gp" not in loss_settings.keys():
            loss_settings["num_gp"] = 2
        super().__init__(name, {**loss_settings, "compute_dims": 2, "ordered_dofs": ["Ux", "Uy"],
                                "element_type": "quad"}, fe_mesh)
