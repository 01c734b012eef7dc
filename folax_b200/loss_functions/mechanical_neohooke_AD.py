"""The AD variant of the Neo-Hooke loss: same class names and settings as
fol/loss_functions/mechanical_neohooke_AD.py:18-320 (import them from this module, as the reference's tests do).

Not an alias of mechanical_neohooke: `NeoHookianModelAD` differentiates mu/2 (J^-2/3 tr C - 3) - mu ln J + lam/2 ln^2 J
(2-D: mu/2 (tr C - 2) - mu ln J + lam/2 ln^2 J) w.r.t. the Voigt vector of C, so its Voigt shear stresses are twice
the tensor components and its 3-D stress does not vanish at F = I (neo_hooke.py:112-209); the stiffness is
`jax.jacfwd(residual)` (:275-283).  The element stage (csrc/assemble_ad_threads.cuh) does the same with dual numbers
and reproduces the reference's 19-digit goldens (tests/unit/test_neo_hooke_mechanical_loss_AD.py).  Like in the
reference these classes are cross-checks, not fast paths: one element per thread, nd forward sweeps.
`ComputeBatchLoss` is not available (the batched kernels do not carry this energy)."""
import torch

from .. import _lib
from .mechanical_neohooke import NeoHookeMechanicalLoss as _AnalyticNeoHooke


class _ADElementStage:
    """Mix-in: element energies come from the AD element stage itself (its `state_out` slot)."""

    _second_order = None

    def _element_energy(self, xyz, conn, ctrl, u, re):
        nd = self._nd
        flags = torch.zeros(nd, dtype=torch.uint8, device=self.device)
        ke = torch.empty(nd * nd, dtype=self.dtype, device=self.device)
        re2 = torch.empty(nd, dtype=self.dtype, device=self.device)
        en = torch.empty(1, dtype=self.dtype, device=self.device)
        _lib.check(_lib.load().fol_assemble_elements(
            _lib.stream_ptr(), self._dt, _lib.PHYSICS[self.physics], self.fe_element.code, self.num_gp, 0, 1,
            self._nnode, _lib.ptr(xyz), _lib.ptr(conn), _lib.ptr(ctrl), _lib.ptr(u), _lib.ptr(flags), self._params,
            _lib.ptr(ke), _lib.ptr(re2), None, _lib.ptr(en)))
        return en[0]

    def ComputeTotalEnergy(self, total_control_vars, total_primal_vars):
        """fe_loss.py:175-176: sum of the element energies (no Dirichlet overwrite)."""
        en = torch.empty(max(self._ne, 1), dtype=self.dtype, device=self.device)
        self._assemble(total_control_vars, total_primal_vars, False, state_out=en)
        out = torch.empty(1, dtype=self.dtype, device=self.device)
        _lib.check(_lib.load().fol_sum(_lib.stream_ptr(), self._dt, self._ne, _lib.ptr(en), _lib.ptr(out)))
        return out[0]


class NeoHookeMechanicalLoss(_ADElementStage, _AnalyticNeoHooke):
    physics = "neohooke_ad"

    def ComputeBatchLoss(self, batch_params, batch_dofs):
        raise NotImplementedError("the AD Neo-Hooke variant is an element-stage cross-check here: its energy "
                                  "(neo_hooke.py:128-134) is not in the batched loss kernels; use mechanical_neohooke")


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
