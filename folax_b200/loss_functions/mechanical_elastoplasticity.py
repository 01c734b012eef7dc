"""Small-strain J2 elastoplasticity with per-Gauss-point history.
Same classes, settings and return tuples as fol/loss_functions/mechanical_elastoplasticity.py:16-248.
`ElastoplasticityLoss3DHexa` does not exist in the reference; it is the obvious subclass needed by
the 256^3 hex configuration of BASELINE.json."""
import torch

from .. import _lib
from ..sparse import BCOO
from .mechanical import MechanicalLoss


class ElastoplasticityLoss(MechanicalLoss):
    physics = "j2plasticity"
    _second_order = None

    def _material_params(self):
        p = super()._material_params()
        md = self.loss_settings["material_dict"]
        p[5] = float(md["yield_limit"])
        p[6] = float(md["iso_hardening_parameter_1"])
        p[7] = float(md["iso_hardening_param_2"])
        return p

    def GetStateShape(self):
        """(number of elements, Gauss points per element, 7 | 4): eps_p (6|3) + cumulative plastic
        strain, zeros initially (plasticity.py:122-130; solver default width, ..._with_history_update.py:81-85)."""
        return (self._ne, self._ngauss, 7 if self.dim == 3 else 4)

    def _state(self, x, ne):
        t = _lib.to_device(x, self.dtype).reshape(ne, self._ngauss, -1).contiguous()
        if t.shape[2] != (7 if self.dim == 3 else 4):
            raise ValueError(f"{self.GetName()}: Gauss-point state must have {7 if self.dim == 3 else 4} entries")
        return t

    def ComputeElement(self, xyze, de, uvwe, element_state_gps):
        """(energy, new Gauss-point state (g, 7|4), residual (nd,1), tangent (nd,nd)) -- :32-94."""
        lib = _lib.load()
        A, nd = self._nnode, self._nd
        xyz = _lib.to_device(xyze, self.dtype).reshape(A, 3)
        ctrl = torch.ones(A, dtype=self.dtype, device=self.device)      # controls are unused by this law
        u = _lib.to_device(uvwe, self.dtype).reshape(nd)
        st_in = self._state(element_state_gps, 1)
        st_out = torch.empty_like(st_in)
        conn = torch.arange(A, dtype=torch.int32, device=self.device).reshape(1, A)
        flags = torch.zeros(nd, dtype=torch.uint8, device=self.device)
        ke = torch.empty(nd * nd, dtype=self.dtype, device=self.device)
        re = torch.empty(nd, dtype=self.dtype, device=self.device)
        _lib.check(lib.fol_assemble_elements(_lib.stream_ptr(), self._dt, _lib.PHYSICS[self.physics],
                                             self.fe_element.code, self.num_gp, 0, 1, A, _lib.ptr(xyz),
                                             _lib.ptr(conn), _lib.ptr(ctrl), _lib.ptr(u), _lib.ptr(flags),
                                             self._params, _lib.ptr(ke), _lib.ptr(re), _lib.ptr(st_in),
                                             _lib.ptr(st_out)))
        return torch.dot(u, re), st_out[0], re.reshape(nd, 1), ke.reshape(nd, nd)

    def ComputeJacobianMatrixAndResidualVector(self, total_control_vars, total_primal_vars, old_state_gps,
                                               transpose_jacobian: bool = False):
        """:153-235 -> (new_state_gps, BCOO, residual)."""
        st_in = self._state(old_state_gps, self._ne)
        st_out = torch.empty_like(st_in)
        data, R = self._assemble(total_control_vars, total_primal_vars, transpose_jacobian, state_in=st_in,
                                 state_out=st_out)
        jac = BCOO((data, self._bcoo_indices()), shape=(self.total_number_of_dofs, self.total_number_of_dofs))
        return st_out, jac, R

    def ComputeBatchLoss(self, batch_params, batch_dofs):
        raise NotImplementedError("the reference's batched energy loss does not cover history-dependent "
                                  "elements (ComputeElement needs the Gauss-point state)")


class ElastoplasticityLoss2DQuad(ElastoplasticityLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        if "num_gp" not in loss_settings.keys():
            loss_settings["num_gp"] = 2
        super(MechanicalLoss, self).__init__(name, {**loss_settings, "compute_dims": 2, "ordered_dofs": ["Ux", "Uy"],
                                                    "element_type": "quad"}, fe_mesh)


class ElastoplasticityLoss3DTetra(ElastoplasticityLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        super(MechanicalLoss, self).__init__(name, {**loss_settings, "compute_dims": 3,
                                                    "ordered_dofs": ["Ux", "Uy", "Uz"],
                                                    "element_type": "tetra"}, fe_mesh)


class ElastoplasticityLoss3DHexa(ElastoplasticityLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        if "num_gp" not in loss_settings.keys():
            loss_settings["num_gp"] = 2
        super(MechanicalLoss, self).__init__(name, {**loss_settings, "compute_dims": 3,
                                                    "ordered_dofs": ["Ux", "Uy", "Uz"],
                                                    "element_type": "hexahedron"}, fe_mesh)
