"""Implicit-Euler transient (non)linear heat conduction.  Same classes, settings and call signatures as
fol/loss_functions/transient_thermal.py:16-172: the (control, dof) slots of the generic assembly carry the
(current, next) nodal temperatures, the conductivity heterogeneity k0 comes from material_dict."""
import numpy as np
import torch

from .. import _lib
from ..sparse import BCOO
from ..tools import fol_error
from .thermal import ThermalLoss


class TransientThermalLoss(ThermalLoss):
    physics = "transient_thermal"
    _second_order = None       # a true potential (transient_thermal.py:42-73)

    def Initialize(self, reinitialize=False) -> None:
        if self.initialized and not reinitialize:
            return
        nn = self.fe_mesh.GetNumberOfNodes()
        self.default_material_settings = {"rho": 1.0, "cp": 1.0, "k0": np.ones(nn), "beta": 0.0, "c": 1.0}
        self.material_settings = dict(self.default_material_settings)
        md = self.loss_settings.get("material_dict", {})
        self.material_settings.update({k: v for k, v in md.items() if k in self.material_settings})
        k0 = np.asarray(self.material_settings["k0"], dtype=float)
        if k0.shape != (nn,):
            fol_error(f"provided k0({k0.shape}) in the material_dict does not match the mesh with {nn} nodes !",
                      self.GetName())
        self.time_integration_settings = {"method": "implicit-euler", "time_step": None}
        self.time_integration_settings.update(self.loss_settings.get("time_integration_dict", {}))
        if self.time_integration_settings["time_step"] is None:
            fol_error("time step should be provided in the time_integration_dict ", self.GetName())
        super().Initialize(reinitialize)
        self._k0 = _lib.to_device(k0, self.dtype)
        self._geom_aux = self._k0      # interpolated into the geometry cache of the batched loss

    def _material_params(self):
        p = [0.0] * _lib.NUM_PARAMS
        # note: the reference reads beta from material_dict and c from the thermal loss settings
        # (transient_thermal.py:52-53); both are honoured here with material_dict taking precedence for c
        p[5] = float(self.material_settings["beta"])
        p[6] = float(self.thermal_loss_settings["c"])
        p[8], p[9] = float(self.material_settings["rho"]), float(self.material_settings["cp"])
        p[10] = float(self.time_integration_settings["time_step"])
        return p

    def ComputeElement(self, xyze, Te_c, Te_n, Ke):
        """(energy, residual (a,1), tangent (a,a)) of one element -- transient_thermal.py:42-73."""
        lib, A = _lib.load(), self._nnode
        xyz = _lib.to_device(xyze, self.dtype).reshape(A, 3)
        tc = _lib.to_device(Te_c, self.dtype).reshape(A)
        tn = _lib.to_device(Te_n, self.dtype).reshape(A)
        k0 = _lib.to_device(Ke, self.dtype).reshape(A)
        conn = torch.arange(A, dtype=torch.int32, device=self.device).reshape(1, A)
        flags = torch.zeros(A, dtype=torch.uint8, device=self.device)
        ke = torch.empty(A * A, dtype=self.dtype, device=self.device)
        re = torch.empty(A, dtype=self.dtype, device=self.device)
        en = torch.empty(1, dtype=self.dtype, device=self.device)
        _lib.check(lib.fol_assemble_elements(_lib.stream_ptr(), self._dt, _lib.PHYSICS[self.physics],
                                             self.fe_element.code, self.num_gp, 0, 1, A, _lib.ptr(xyz), _lib.ptr(conn),
                                             _lib.ptr(tc), _lib.ptr(tn), _lib.ptr(flags), self._params, _lib.ptr(ke),
                                             _lib.ptr(re), _lib.ptr(k0), _lib.ptr(en)))
        return en[0], re.reshape(A, 1), ke.reshape(A, A)

    def ComputeJacobianMatrixAndResidualVector(self, nodal_current_temps, nodal_next_temps,
                                               transpose_jacobian: bool = False):
        data, R = self._assemble(nodal_current_temps, nodal_next_temps, transpose_jacobian, state_in=self._k0)
        jac = BCOO((data, self._bcoo_indices()), shape=(self.total_number_of_dofs, self.total_number_of_dofs))
        return jac, R

    def ComputeTotalEnergy(self, nodal_current_temps, nodal_next_temps):
        en = torch.empty(self._ne, dtype=self.dtype, device=self.device)
        self._assemble(nodal_current_temps, nodal_next_temps, False, state_in=self._k0, state_out=en)
        return en.sum()



class TransientThermalLoss3DTetra(TransientThermalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        super(ThermalLoss, self).__init__(name, {**loss_settings, "compute_dims": 3, "ordered_dofs": ["T"],
                                                 "element_type": "tetra"}, fe_mesh)


class TransientThermalLoss3DHexa(TransientThermalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        if "num_gp" not in loss_settings.keys():
            loss_settings["num_gp"] = 2
        super(ThermalLoss, self).__init__(name, {**loss_settings, "compute_dims": 3, "ordered_dofs": ["T"],
                                                 "element_type": "hexahedron"}, fe_mesh)


class TransientThermalLoss2DQuad(TransientThermalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        if "num_gp" not in loss_settings.keys():
            loss_settings["num_gp"] = 2
        super(ThermalLoss, self).__init__(name, {**loss_settings, "compute_dims": 2, "ordered_dofs": ["T"],
                                                 "element_type": "quad"}, fe_mesh)


class TransientThermalLoss2DTri(TransientThermalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        super(ThermalLoss, self).__init__(name, {**loss_settings, "compute_dims": 2, "ordered_dofs": ["T"],
                                                 "element_type": "triangle"}, fe_mesh)
