"""KratosSmallDisplacement3DTetra: the reference's own FFI loss class
(fol/loss_functions/kratos_small_displacement.py:28-173), served by the sm_100a kernels instead of Kratos on the host.

In the reference this class is the precedent for a native plugin: `compute_elements` / `compute_nodal_residuals` are
XLA-FFI custom calls into Kratos Multiphysics (`SmallDisplacementElement3D4N` + `LinearElastic3DLaw`,
ffi_functions/kr_small_displacement_element.cc:137-159) that copy the buffers to the HOST, loop over the elements with
OpenMP and copy back (:193-214).  The element is the constant-strain Tet4 with one integration point and a uniform
(E, nu): exactly `MechanicalLoss3DTetra` with a unit control field and no body force -- the reference's dense-Jacobian
golden of this class (tests/unit/test_kratos_ffi_mechanical_loss.py:72) is reproduced by the oracle and by the kernels.
Semantics kept: the control vector is ignored (:92-100, 104-117), `ComputeElement` is an error (:89-90), so the
batched loss of the base class is not available either ("differentiation").
Kratos itself is third-party and un-vendored (version unpinned in the reference); no Kratos code is involved here."""
import torch

from .. import _lib
from ..tools import fol_error
from .mechanical import MechanicalLoss


class KratosSmallDisplacement3DTetra(MechanicalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        super().__init__(name, {**loss_settings, "compute_dims": 3, "ordered_dofs": ["Ux", "Uy", "Uz"],
                                "element_type": "tetra"}, fe_mesh)

    def Initialize(self, reinitialize=False) -> None:
        if self.initialized and not reinitialize:
            return
        self.material_settings = {"poisson_ratio": 0.3, "young_modulus": 1.0}          # :83-85
        if "material_dict" in self.loss_settings.keys():
            self.material_settings = self.loss_settings["material_dict"]
        self.loss_settings["material_dict"] = self.material_settings
        self.loss_settings["num_gp"] = 1                 # SmallDisplacementElement3D4N: one integration point
        super().Initialize(reinitialize)
        self._unit_control = torch.ones(self._nn, dtype=self.dtype, device=self.device)

    def _material_params(self):
        p = super()._material_params()
        p[2:5] = [0.0, 0.0, 0.0]                         # the FFI call passes no body force to Kratos (:96-99, 117-121)
        return p

    def ComputeElement(self, xyze, de, te, body_force=0):
        fol_error(" is not implemented for KratosSmallDisplacement3DTetra!")

    def ComputeBatchLoss(self, batch_params, batch_dofs):
        fol_error(" is not implemented for KratosSmallDisplacement3DTetra! (the reference's base-class batched loss "
                  "routes through ComputeElement, kratos_small_displacement.py:89-90)")

    def ComputeTotalEnergy(self, total_control_vars, total_primal_vars):
        """u . R(u), R the assembled nodal residual (:92-100); the control vector is ignored."""
        dofs = self._as_batch(total_primal_vars, self.total_number_of_dofs)
        energy, _, _ = self._energy_and_grads(self._unit_control.reshape(1, -1), dofs)
        return energy[0]

    def ComputeJacobianMatrixAndResidualVector(self, total_control_vars, total_primal_vars,
                                               transpose_jacobian: bool = False):
        """:102-173 -> (BCOO, residual); the control vector is ignored."""
        return super().ComputeJacobianMatrixAndResidualVector(self._unit_control, total_primal_vars, transpose_jacobian)

    def ApplyJacobian(self, total_control_vars, total_primal_vars, vector, transpose_jacobian: bool = False,
                      state_in=None):
        return super().ApplyJacobian(self._unit_control, total_primal_vars, vector, transpose_jacobian)
