"""Implicit-Euler Allen-Cahn phase-field loss.  Same classes, settings and call signatures as
fol/loss_functions/phase_field.py:15-92: the (control, dof) slots of the generic assembly carry the
(current, next) nodal phase field."""
import torch

from .. import _lib
from ..sparse import BCOO
from ..tools import fol_error
from .fe_loss import FiniteElementLoss


class AllenCahnLoss(FiniteElementLoss):
    physics = "allen_cahn"

    def Initialize(self, reinitialize=False) -> None:
        if "material_dict" not in self.loss_settings.keys():
            fol_error("material_dict should provided in the loss settings !", self.GetName())
        md = self.loss_settings["material_dict"]
        self.rho, self.cp, self.dt, self.epsilon = md["rho"], md["cp"], md["dt"], md["epsilon"]
        super().Initialize(reinitialize)

    def _material_params(self):
        p = [0.0] * _lib.NUM_PARAMS
        p[8], p[9], p[10], p[11] = float(self.rho), float(self.cp), float(self.dt), float(self.epsilon)
        return p

    def ComputeElement(self, xyze, phi_e_c, phi_e_n, body_force=0):
        """(energy, residual (a,1), tangent (a,a)) of one element -- phase_field.py:38-70."""
        lib, A = _lib.load(), self._nnode
        xyz = _lib.to_device(xyze, self.dtype).reshape(A, 3)
        pc = _lib.to_device(phi_e_c, self.dtype).reshape(A)
        pn = _lib.to_device(phi_e_n, self.dtype).reshape(A)
        conn = torch.arange(A, dtype=torch.int32, device=self.device).reshape(1, A)
        flags = torch.zeros(A, dtype=torch.uint8, device=self.device)
        ke = torch.empty(A * A, dtype=self.dtype, device=self.device)
        re = torch.empty(A, dtype=self.dtype, device=self.device)
        en = torch.empty(1, dtype=self.dtype, device=self.device)
        _lib.check(lib.fol_assemble_elements(_lib.stream_ptr(), self._dt, _lib.PHYSICS[self.physics],
                                             self.fe_element.code, self.num_gp, 0, 1, A, _lib.ptr(xyz), _lib.ptr(conn),
                                             _lib.ptr(pc), _lib.ptr(pn), _lib.ptr(flags), self._params, _lib.ptr(ke),
                                             _lib.ptr(re), None, _lib.ptr(en)))
        return en[0], re.reshape(A, 1), ke.reshape(A, A)

    def ComputeJacobianMatrixAndResidualVector(self, nodal_current_phi, nodal_next_phi, transpose_jacobian: bool = False):
        data, R = self._assemble(nodal_current_phi, nodal_next_phi, transpose_jacobian)
        jac = BCOO((data, self._bcoo_indices()), shape=(self.total_number_of_dofs, self.total_number_of_dofs))
        return jac, R

    def ComputeTotalEnergy(self, nodal_current_phi, nodal_next_phi):
        en = torch.empty(self._ne, dtype=self.dtype, device=self.device)
        self._assemble(nodal_current_phi, nodal_next_phi, False, state_out=en)
        return en.sum()



class AllenCahnLoss2DQuad(AllenCahnLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        if "num_gp" not in loss_settings.keys():
            loss_settings["num_gp"] = 2
        super().__init__(name, {**loss_settings, "compute_dims": 2, "ordered_dofs": ["Phi"],
                                "element_type": "quad"}, fe_mesh)


class AllenCahnLoss2DTri(AllenCahnLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        super().__init__(name, {**loss_settings, "compute_dims": 2, "ordered_dofs": ["Phi"],
                                "element_type": "triangle"}, fe_mesh)


class AllenCahnLoss3DHexa(AllenCahnLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        if "num_gp" not in loss_settings.keys():
            loss_settings["num_gp"] = 2
        super().__init__(name, {**loss_settings, "compute_dims": 3, "ordered_dofs": ["Phi"],
                                "element_type": "hexahedron"}, fe_mesh)
