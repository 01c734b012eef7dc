"""The AD variant of the Saint-Venant-Kirchhoff loss: same class names and settings as
fol/loss_functions/mechanical_saint_venant_AD.py:18-337.

`SaintVenantAD` takes the stress as the gradient of lam/2 tr(E)^2 + mu tr(E E) w.r.t. the Voigt vector of E, which
doubles the Voigt shear stresses (saint_venant.py:36-64); the stiffness is `jax.jacfwd(residual)`.  It coincides with
the analytic class only at F = I, which is what the reference's test compares
(test_saint_venant_mechanical_loss.py:41-42).  Element stage: csrc/assemble_ad_threads.cuh (dual numbers).  The
ENERGY is the same function as in the analytic class and the batched loss differentiates the energy, so
ComputeBatchLoss and its gradient are those of mechanical_saint_venant."""
from .mechanical_neohooke_AD import _ADElementStage
from .mechanical_saint_venant import SaintVenantMechanicalLoss as _AnalyticSaintVenant


class SaintVenantMechanicalLoss(_ADElementStage, _AnalyticSaintVenant):
    physics = "stvenant_ad"

    def _batch_physics(self):
        return "stvenant"


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
