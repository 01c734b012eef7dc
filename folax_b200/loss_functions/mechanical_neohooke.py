"""Finite-strain Neo-Hooke elasticity (analytic tangent + geometric stiffness).
Same classes and settings as fol/loss_functions/mechanical_neohooke.py:17-303."""
from .mechanical import MechanicalLoss


class NeoHookeMechanicalLoss(MechanicalLoss):
    physics = "neohooke"
    _second_order = "hessian"  # a true potential whose analytic tangent is its Hessian
    _has_control_gradient = True  # psi is differentiable in the control field

    def Initialize(self, reinitialize=False) -> None:
        super().Initialize(reinitialize)
        self.e = self.loss_settings["material_dict"]["young_modulus"]
        self.v = self.loss_settings["material_dict"]["poisson_ratio"]

    def _energy_and_grads(self, batch_params, batch_dofs, **kw):
        from .fe_loss import FiniteElementLoss
        return FiniteElementLoss._energy_and_grads(self, batch_params, batch_dofs, **kw)

    def _element_energy(self, xyz, conn, ctrl, u, re):
        # energy = sum_g w detJ psi (mechanical_neohooke.py:262, 271), not u . re
        import numpy as np
        import torch
        from .. import _lib, energy_plan
        lib = _lib.load()
        A, s = self._nnode, _lib.stream_ptr()
        width = A * self._edim + 1
        geom = torch.empty(self._ngauss * width, dtype=self.dtype, device=self.device)
        _lib.check(lib.fol_geometry_cache(s, self._dt, self.fe_element.code, self.num_gp, 1, _lib.ptr(xyz),
                                          _lib.ptr(conn), _lib.ptr(geom)))
        ep = {k: (torch.as_tensor(v, device=self.device) if isinstance(v, np.ndarray) else v)
              for k, v in energy_plan.build(xyz.cpu().numpy(), conn.cpu().numpy()).items()}
        gu, gk = torch.empty_like(u), torch.empty_like(ctrl)
        energy = torch.empty(1, dtype=self.dtype, device=self.device)
        work = torch.empty(lib.fol_energy_work_size(ep["ntiles"], 1), dtype=self.dtype, device=self.device)
        _lib.check(lib.fol_energy_and_grads(s, self._dt, _lib.PHYSICS[self.physics], self.fe_element.code,
                                            self.num_gp, 1, A, 1, _lib.ptr(geom), _lib.ptr(conn),
                                            _lib.ptr(ep["adj_ptr"]), _lib.ptr(ep["adj_local"]),
                                            _lib.ptr(ep["tile_node_ptr"]), _lib.ptr(ep["tile_nodes"]),
                                            _lib.ptr(ep["tile_elem_ptr"]), _lib.ptr(ep["tile_elems"]), _lib.ptr(ep["tile_conn"]),
                                            _lib.ptr(ep["tile_lnode_ptr"]), _lib.ptr(ep["tile_lnodes"]), ep["ntiles"],
                                            ep["ecap"], ep["lcap"], ep["ncap"], _lib.ptr(ctrl), _lib.ptr(u), None, None, 1.0, self._params, _lib.ptr(gu),
                                            _lib.ptr(gk), _lib.ptr(energy), _lib.ptr(work)))
        return energy[0]


class NeoHookeMechanicalLoss2DQuad(NeoHookeMechanicalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        if "num_gp" not in loss_settings.keys():
            loss_settings["num_gp"] = 2
        super().__init__(name, {**loss_settings, "compute_dims": 2, "ordered_dofs": ["Ux", "Uy"],
                                "element_type": "quad"}, fe_mesh)


class NeoHookeMechanicalLoss2DTri(NeoHookeMechanicalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        super().__init__(name, {**loss_settings, "compute_dims": 2, "ordered_dofs": ["Ux", "Uy"],
                                "element_type": "triangle"}, fe_mesh)


class NeoHookeMechanicalLoss3DTetra(NeoHookeMechanicalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        super().__init__(name, {**loss_settings, "compute_dims": 3, "ordered_dofs": ["Ux", "Uy", "Uz"],
                                "element_type": "tetra"}, fe_mesh)


class NeoHookeMechanicalLoss3DHexa(NeoHookeMechanicalLoss):
    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        if "num_gp" not in loss_settings.keys():
            loss_settings["num_gp"] = 2
        super().__init__(name, {**loss_settings, "compute_dims": 3, "ordered_dofs": ["Ux", "Uy", "Uz"],
                                "element_type": "hexahedron"}, fe_mesh)
