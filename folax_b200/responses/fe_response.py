"""FiniteElementResponse: host side of the adjoint sensitivities, same API as
fol/responses/fe_response.py:17-614.

    value = sum_e sum_g w detJ f(K_g, U_g),   f = the user's `response_formula` in the control's name and the
    first letter of the first dof (e.g. "(E**2)*U[0]", fe_response.py:63-66).

Work split.  The integration, the gathers, the geometric derivatives and the residual sensitivities
lam^T d re/dK, lam^T d re/dx run in the sm_100a kernels of csrc/adjoint.cu (C ABI: fol_gauss_interpolate,
fol_response_elements, fol_residual_adjoint_elements, fol_sum) and the node sums in fol_residual_gather (fixed
order, no atomics).  The formula is the caller's Python expression, exactly as in the reference: it is evaluated
pointwise over the (ne, g) Gauss-point arrays on the device, and its partials df/dK, df/dU come from one
torch.autograd.grad over those arrays (the reference uses jax.grad for the same purpose).  No CPU fallback.

Residual sensitivities: closed forms for the mechanical, thermal, transient-thermal and Allen-Cahn families,
closed-form geometry + dual-number derivatives of the point law for Neo-Hooke and St-Venant; the history-dependent J2
loss raises FolaxError.
"""
import math

import numpy as np
import torch

from .. import _lib
from ..tools import fol_error
from .response import Response


class _Linalg:
    @staticmethod
    def norm(x):
        return torch.sqrt((x * x).sum(0))


class _JnpOnTorch:
    """What `jnp` means inside a response formula.  The formula sees one Gauss point in the reference (control:
    scalar, dofs: (d,)); here the arguments carry trailing (ne, g) axes, so reductions over "the vector" are
    reductions over axis 0 only.  Everything else is the torch function of the same name."""

    pi, e, inf = math.pi, math.e, math.inf
    linalg = _Linalg()

    @staticmethod
    def sum(x, axis=None):
        return x.sum(0)

    @staticmethod
    def dot(a, b):
        return (a * b).sum(0)

    @staticmethod
    def power(x, p):
        return torch.pow(x, p)

    @staticmethod
    def array(x):
        return x if isinstance(x, torch.Tensor) else torch.as_tensor(x)

    def __getattr__(self, name):
        try:
            return getattr(torch, name)
        except AttributeError:
            raise AttributeError(f"jnp.{name} is not available in a response formula on this backend")


class NodalControl:
    """Smallest object with the surface FiniteElementResponse reads off a fol Control (control.py:18-42):
    a name, Initialize() and one controlled variable per node (identity_control.py:24-25)."""

    def __init__(self, control_name: str, fe_mesh):
        self._name, self.fe_mesh, self.initialized = control_name, fe_mesh, False
        self.num_control_vars = self.num_controlled_vars = None

    def GetName(self):
        return self._name

    def Initialize(self, reinitialize=False):
        self.num_control_vars = self.num_controlled_vars = self.fe_mesh.GetNumberOfNodes()
        self.initialized = True

    def ComputeControlledVariables(self, variable_vector):
        return variable_vector

    def Finalize(self):
        pass


class FiniteElementResponse(Response):
    def __init__(self, name: str, response_formula: str, fe_loss, control):
        super().__init__(name)
        self.response_formula = response_formula
        self.fe_loss = fe_loss
        self.control = control

    def Initialize(self, reinitialize=False) -> None:
        if self.initialized and not reinitialize:
            return
        self.fe_loss.Initialize()
        self.control.Initialize()
        variables_list = [self.control.GetName(), self.fe_loss.dofs[0][0]]          # fe_response.py:63
        func_str = f"lambda {', '.join(variables_list)}: {self.response_formula}"
        self.response_function = eval(func_str, {"jnp": _JnpOnTorch(), "torch": torch, "math": math})
        nn = self.fe_loss.fe_mesh.GetNumberOfNodes()
        ncv = getattr(self.control, "num_controlled_vars", None)
        if ncv is not None and int(ncv) != nn:
            fol_error(f"{self.GetName()}: one controlled variable per node is supported "
                      f"({ncv} controlled variables on {nn} nodes)")
        self.initialized = True

    # ------------------------------------------------------------------ plumbing
    def _fields(self, nodal_control_values, nodal_dof_values):
        L = self.fe_loss
        ctrl = _lib.to_device(nodal_control_values, L.dtype).reshape(-1)
        u = _lib.to_device(nodal_dof_values, L.dtype).reshape(-1)
        if ctrl.numel() != L._nn or u.numel() != L.total_number_of_dofs:
            raise ValueError(f"{self.GetName()}: controls must have {L._nn} entries and dofs {L.total_number_of_dofs}")
        return ctrl, u

    def _formula_at_gauss_points(self, ctrl, u, need_grads):
        """(f, df/dK, df/dU) over the Gauss points: kernel interpolation, caller's formula, autograd partials."""
        L = self.fe_loss
        d, g = L.number_dofs_per_node, L._ngauss
        kg = torch.empty((L._ne, g), dtype=L.dtype, device=L.device)
        ug = torch.empty((d, L._ne, g), dtype=L.dtype, device=L.device)
        _lib.check(_lib.load().fol_gauss_interpolate(_lib.stream_ptr(), L._dt, L.fe_element.code, L.num_gp, d, L._ne,
                                                     _lib.ptr(L._conn), _lib.ptr(ctrl), _lib.ptr(u), _lib.ptr(kg),
                                                     _lib.ptr(ug)))
        if not need_grads:
            with torch.no_grad():
                f = self.response_function(kg, ug)
            return (f * torch.ones_like(kg)).contiguous(), None, None
        kg.requires_grad_(True)
        ug.requires_grad_(True)
        with torch.enable_grad():
            f = self.response_function(kg, ug) * torch.ones_like(kg)       # a formula may ignore an argument
            fk, fu = torch.autograd.grad(f.sum(), (kg, ug), allow_unused=True)
        fk = torch.zeros_like(kg) if fk is None else fk.contiguous()
        fu = torch.zeros_like(ug) if fu is None else fu.contiguous()
        return f.detach().contiguous(), fk, fu

    def _response_elements(self, f, fk, fu, want):
        L = self.fe_loss
        A, d = L._nnode, L.number_dofs_per_node
        out = {"val": (L._ne,), "du": (L._ne, A * d), "dk": (L._ne, A), "dx": (L._ne, A * 3)}
        bufs = {k: (torch.empty(max(int(np.prod(s)), 1), dtype=L.dtype, device=L.device) if k in want else None)
                for k, s in out.items()}
        _lib.check(_lib.load().fol_response_elements(
            _lib.stream_ptr(), L._dt, L.fe_element.code, L.num_gp, d, L._ne, _lib.ptr(L._xyz), _lib.ptr(L._conn),
            _lib.ptr(f), _lib.ptr(fk), _lib.ptr(fu), _lib.ptr(bufs["val"]), _lib.ptr(bufs["du"]), _lib.ptr(bufs["dk"]),
            _lib.ptr(bufs["dx"])))
        return bufs

    def _node_sum(self, elem_values, width):
        L = self.fe_loss
        out = torch.empty(L._nn * width, dtype=L.dtype, device=L.device)
        _lib.check(_lib.load().fol_residual_gather(_lib.stream_ptr(), L._dt, L._nn, L._nnode, width,
                                                   _lib.ptr(L._adj_ptr), _lib.ptr(L._adj), _lib.ptr(elem_values),
                                                   _lib.ptr(out)))
        return out

    def _residual_adjoint(self, ctrl, u, lam, dk, dx):
        L = self.fe_loss
        _lib.check(_lib.load().fol_residual_adjoint_elements(
            _lib.stream_ptr(), L._dt, _lib.PHYSICS[L.physics], L.fe_element.code, L.num_gp, 1, L._ne,
            _lib.ptr(L._xyz), _lib.ptr(L._conn), _lib.ptr(ctrl), _lib.ptr(u), _lib.ptr(lam),
            _lib.ptr(getattr(L, "_geom_aux", None)), L._params, _lib.ptr(dk), _lib.ptr(dx)))

    # ------------------------------------------------------------------ API of fe_response.py
    def ComputeValue(self, nodal_control_values, nodal_dof_values):
        """fe_response.py:205-221."""
        ctrl, u = self._fields(nodal_control_values, nodal_dof_values)
        f, _, _ = self._formula_at_gauss_points(ctrl, u, need_grads=False)
        val = self._response_elements(f, None, None, ("val",))["val"]
        L = self.fe_loss
        out = torch.empty(1, dtype=L.dtype, device=L.device)
        _lib.check(_lib.load().fol_sum(_lib.stream_ptr(), L._dt, L._ne, _lib.ptr(val), _lib.ptr(out)))
        return out[0]

    def ComputeAdjointJacobianMatrixAndRHSVector(self, nodal_control_values, nodal_dof_values):
        """fe_response.py:245-283 -> (transposed, BC-applied Jacobian of the loss; rhs = -d value/d u, zero at
        the Dirichlet dofs)."""
        L = self.fe_loss
        ctrl, u = self._fields(nodal_control_values, nodal_dof_values)
        f, fk, fu = self._formula_at_gauss_points(ctrl, u, need_grads=True)
        du = self._response_elements(f, None, fu, ("du",))["du"]
        rhs = self._node_sum(du, L.number_dofs_per_node)
        rhs[L._dir_idx.to(torch.int64)] = 0.0
        rhs = -rhs
        sparse_jacobian, _ = L.ComputeJacobianMatrixAndResidualVector(ctrl, u, True)
        return sparse_jacobian, rhs

    def ComputeAdjointNodalControlDerivatives(self, nodal_control_values, nodal_dof_values, nodal_adj_dof_values):
        """fe_response.py:486-524: d value/d K + lam^T d R/d K per node."""
        L = self.fe_loss
        ctrl, u = self._fields(nodal_control_values, nodal_dof_values)
        lam = _lib.to_device(nodal_adj_dof_values, L.dtype).reshape(-1)
        if lam.numel() != u.numel():
            raise ValueError(f"{self.GetName()}: the adjoint vector must have {u.numel()} entries")
        f, fk, fu = self._formula_at_gauss_points(ctrl, u, need_grads=True)
        dk = self._response_elements(f, fk, None, ("dk",))["dk"]
        self._residual_adjoint(ctrl, u, lam, dk, None)
        return self._node_sum(dk, 1)

    def ComputeAdjointNodalShapeDerivatives(self, nodal_control_values, nodal_dof_values, nodal_adj_dof_values):
        """fe_response.py:358-394: d value/d x + lam^T d R/d x, three entries per node."""
        L = self.fe_loss
        ctrl, u = self._fields(nodal_control_values, nodal_dof_values)
        lam = _lib.to_device(nodal_adj_dof_values, L.dtype).reshape(-1)
        if lam.numel() != u.numel():
            raise ValueError(f"{self.GetName()}: the adjoint vector must have {u.numel()} entries")
        f, _, _ = self._formula_at_gauss_points(ctrl, u, need_grads=False)
        dx = self._response_elements(f, None, None, ("dx",))["dx"]
        self._residual_adjoint(ctrl, u, lam, None, dx)
        return self._node_sum(dx, 3)

    # ------------------------------------------------------------------ finite-difference checks (debug tools)
    def ComputeFDNodalControlDerivatives(self, nodal_control_values, fe_solver, fd_step_size: float = 1e-4,
                                         fd_mode="FWD"):
        """fe_response.py:527-567: one (or two) FE solves per control; `fe_solver.Solve(controls, dofs)` is the
        caller's solver."""
        if fd_mode not in ("FWD", "CD"):
            fol_error("only Forward (FWD), Central Difference (CD) methods are implemented !")
        L = self.fe_loss
        K = np.array(torch.as_tensor(nodal_control_values).detach().cpu().numpy(), dtype=float).reshape(-1)
        zeros = np.zeros(L.total_number_of_dofs)

        def value(Kp):
            return float(self.ComputeValue(Kp, fe_solver.Solve(Kp, zeros)))

        base = value(K)
        grad = np.zeros_like(K)
        for i in range(K.size):
            Kp = K.copy()
            Kp[i] += fd_step_size
            fw = value(Kp)
            if fd_mode == "FWD":
                grad[i] = (fw - base) / fd_step_size
            else:
                Kp[i] -= 2.0 * fd_step_size
                grad[i] = (fw - value(Kp)) / (2.0 * fd_step_size)
        return grad

    def ComputeFDNodalShapeDerivatives(self, nodal_control_values, fe_solver, fd_step_size: float = 1e-4,
                                       fd_mode="FWD"):
        """fe_response.py:569-611: perturbs the mesh coordinates node by node; the loss's device-resident plan is
        rebuilt for every perturbed mesh."""
        if fd_mode not in ("FWD", "CD"):
            fol_error("only Forward (FWD), Central Difference (CD) methods are implemented !")
        L = self.fe_loss
        mesh = L.fe_mesh
        zeros = np.zeros(L.total_number_of_dofs)
        coords0 = np.array(mesh.nodes_coordinates, dtype=float, copy=True)

        def value(coords):
            mesh.nodes_coordinates = coords
            L.Initialize(reinitialize=True)
            return float(self.ComputeValue(nodal_control_values, fe_solver.Solve(nodal_control_values, zeros)))

        base = value(coords0)
        grad = np.zeros((mesh.GetNumberOfNodes(), 3))
        for n in range(mesh.GetNumberOfNodes()):
            for c in range(L.dim):
                cp = coords0.copy()
                cp[n, c] += fd_step_size
                fw = value(cp)
                if fd_mode == "FWD":
                    grad[n, c] = (fw - base) / fd_step_size
                else:
                    cp[n, c] -= 2.0 * fd_step_size
                    grad[n, c] = (fw - value(cp)) / (2.0 * fd_step_size)
        value(coords0)
        return grad.reshape(-1)

    def Finalize(self) -> None:
        pass
