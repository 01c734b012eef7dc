"""Finite-strain Saint-Venant-Kirchhoff elasticity (analytic stress / tangent + geometric stiffness).
Same classes and settings as fol/loss_functions/mechanical_saint_venant.py:18-332, including the
`parametric_boundary_learning` mode, where the batch parameters are the Dirichlet values of every sample
and the control field is the mesh's heterogeneity field (:59-66)."""
import numpy as np

from .mechanical_neohooke import NeoHookeMechanicalLoss


class SaintVenantMechanicalLoss(NeoHookeMechanicalLoss):
    physics = "stvenant"
    _second_order = None       # the reference tangent (2 mu shear entries) is not the Hessian of psi
    default_material_settings = {"young_modulus": 1.0, "poisson_ratio": 0.3, "heterogeneity_field_name": "K",
                                 "heterogeneity_default_value": 1.0}

    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        md = dict(self.default_material_settings)
        md.update({k: v for k, v in loss_settings.get("material_dict", {}).items() if k in md})
        md["poisson_ratio"] = float(md["poisson_ratio"])
        loss_settings["material_dict"] = md
        super().__init__(name, loss_settings, fe_mesh)

    def Initialize(self, reinitialize=False) -> None:
        super().Initialize(reinitialize)
        if self.loss_settings.get("parametric_boundary_learning"):
            md = self.loss_settings["material_dict"]
            name = md["heterogeneity_field_name"]
            if not self.fe_mesh.HasPointData(name):
                self.fe_mesh[name] = md["heterogeneity_default_value"] * np.ones(self.fe_mesh.GetNumberOfNodes())
            self.heterogeneity_field = np.asarray(self.fe_mesh[name], dtype=float)


class SaintVenantMechanicalLoss2DQuad(SaintVenantMechanicalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        if "num_gp" not in loss_settings.keys():
            loss_settings["num_gp"] = 2
        super().__init__(name, {**loss_settings, "compute_dims": 2, "ordered_dofs": ["Ux", "Uy"],
                                "element_type": "quad"}, fe_mesh)


class SaintVenantMechanicalLoss2DTri(SaintVenantMechanicalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        super().__init__(name, {**loss_settings, "compute_dims": 2, "ordered_dofs": ["Ux", "Uy"],
                                "element_type": "triangle"}, fe_mesh)


class SaintVenantMechanicalLoss3DTetra(SaintVenantMechanicalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        super().__init__(name, {**loss_settings, "compute_dims": 3, "ordered_dofs": ["Ux", "Uy", "Uz"],
                                "element_type": "tetra"}, fe_mesh)


class SaintVenantMechanicalLoss3DHexa(SaintVenantMechanicalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        if "num_gp" not in loss_settings.keys():
            loss_settings["num_gp"] = 2
        super().__init__(name, {**loss_settings, "compute_dims": 3, "ordered_dofs": ["Ux", "Uy", "Uz"],
                                "element_type": "hexahedron"}, fe_mesh)
