"""(Non)linear heat conduction, kappa = K (1 + beta T^c).
Same classes and settings as fol/loss_functions/thermal.py:15-78."""
from .fe_loss import FiniteElementLoss


class ThermalLoss(FiniteElementLoss):
    physics = "thermal"
    _second_order = "zero_in_u"

    def Initialize(self, reinitialize=False) -> None:
        if self.initialized and not reinitialize:
            return
        self.thermal_loss_settings = {"beta": 0, "c": 1}
        for key in ("beta", "c"):
            if key in self.loss_settings.keys():
                self.thermal_loss_settings[key] = self.loss_settings[key]
        super().Initialize(reinitialize)

    def _material_params(self):
        p = super()._material_params()
        p[5], p[6] = float(self.thermal_loss_settings["beta"]), float(self.thermal_loss_settings["c"])
        return p


class ThermalLoss3DTetra(ThermalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        super().__init__(name, {**loss_settings, "compute_dims": 3, "ordered_dofs": ["T"],
                                "element_type": "tetra"}, fe_mesh)


class ThermalLoss3DHexa(ThermalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        if "num_gp" not in loss_settings.keys():
            loss_settings["num_gp"] = 2
        super().__init__(name, {**loss_settings, "compute_dims": 3, "ordered_dofs": ["T"],
                                "element_type": "hexahedron"}, fe_mesh)


class ThermalLoss2DQuad(ThermalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        if "num_gp" not in loss_settings.keys():
            loss_settings["num_gp"] = 2
        super().__init__(name, {**loss_settings, "compute_dims": 2, "ordered_dofs": ["T"],
                                "element_type": "quad"}, fe_mesh)


class ThermalLoss2DTri(ThermalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        super().__init__(name, {**loss_settings, "compute_dims": 2, "ordered_dofs": ["T"],
                                "element_type": "triangle"}, fe_mesh)
