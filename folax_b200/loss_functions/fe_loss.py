"""FiniteElementLoss: host side of the hot path, same API as fol/loss_functions/fe_loss.py:19-318.

All arithmetic runs in the sm_100a kernels of libfolax_b200 (through the C ABI, include/
folax_b200.h); this class only owns the device-resident *plan* of a mesh (coordinates,
connectivity, Dirichlet flags, node->element adjacency, BCOO indices) and marshals pointers.
PyTorch supplies device memory, streams and the autograd hook (the analogue of jax.custom_vjp).
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from ..sparse import BCOO
from ..tools import fol_error, fol_info
from .loss import Loss


class ElementInfo:
    """What callers read off `loss.fe_element` (fol/geometries/geometry.py:12-47)."""

    def __init__(self, element_type):
        self.__name = element_type
        self.code = _lib.ELEMENTS[element_type]
        self.integration_method = "GI_GAUSS_1"
        self.num_gp = 1

    def GetName(self):
        return self.__name

    def SetGaussIntegrationMethod(self, gi_method: str):
        order = {"GI_GAUSS_1": 1, "GI_GAUSS_2": 2, "GI_GAUSS_3": 3}.get(gi_method)
        if order is None:
            fol_error(f"{gi_method} integration method is not implemented.")
        self.integration_method, self.num_gp = gi_method, order

    def info(self):
        a, d, g = C.c_int(), C.c_int(), C.c_int()
        _lib.check(_lib.load().fol_element_info(self.code, self.num_gp, C.byref(a), C.byref(d), C.byref(g)))
        return a.value, d.value, g.value


def _aligned16(t):
    """The bulk copies of csrc/energy_grid.cu read from 16-byte aligned arrays (fresh torch tensors always are)."""
    return t if t.data_ptr() % 16 == 0 else t.clone(memory_format=torch.contiguous_format)


class _BatchLossFn(torch.autograd.Function):
    """custom_vjp of ComputeBatchLoss: forward runs the fused energy+gradient kernel and the loss
    tail, backward only rescales the saved cotangents -- SURVEY.md A.7.
    fuse_dirichlet: `batch_dofs` are the raw dofs and the kernel overwrites the Dirichlet entries while it
    stages them (fe_loss.py:91-92, 255) -- no full-dof copy is made.  With exponent 1 the kernel also writes the
    cotangents with their final 1/B scale and zero Dirichlet entries, so backward is a no-op for an upstream
    cotangent of 1 (decided on the device: no host synchronisation)."""

    @staticmethod
    def forward(ctx, loss, batch_params, batch_dofs, mask_dirichlet=True, fuse_dirichlet=False):
        ctx.mask_dirichlet = mask_dirichlet
        ctx.set_materialize_grads(False)      # cotangents of unused outputs stay None: no zero-fill / add kernels per step
        nb = batch_dofs.shape[0]
        ctx.prescaled = float(loss.loss_function_exponent) == 1.0
        energy, grad_u, grad_k = loss._energy_and_grads(
            batch_params, batch_dofs, dir_values=loss._dir_full if fuse_dirichlet else None,
            dir_flag=loss._dir_flag if (mask_dirichlet and ctx.prescaled) else None,
            out_scale=(1.0 / nb) if ctx.prescaled else 1.0)
        out4 = torch.empty(4, dtype=loss.dtype, device=energy.device)
        scale = torch.empty_like(energy)
        _lib.check(_lib.load().fol_loss_reduce(_lib.stream_ptr(), loss._dt, energy.shape[0],
                                               float(loss.loss_function_exponent), _lib.ptr(energy),
                                               _lib.ptr(out4), _lib.ptr(scale)))
        ctx.loss, ctx.grads, ctx.scale, ctx.used = loss, (grad_u, grad_k), scale, False
        ctx.save_for_backward(batch_params, batch_dofs)       # second-order path only (inputs: no copy is made)
        ctx.fuse_dirichlet = fuse_dirichlet
        ctx.need = (ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        ctx.mark_non_differentiable(energy)
        return out4[0], out4[1], out4[2], out4[3], energy

    @staticmethod
    def backward(ctx, g_mean, g_min, g_max, g_mean2, _g_energy):
        if ctx.used:
            raise RuntimeError("ComputeBatchLoss: backward through the FFI loss a second time is not supported")
        ctx.used = True
        loss = ctx.loss
        if torch.is_grad_enabled():
            # create_graph=True: the caller differentiates THROUGH this gradient (the latent-code steps of
            # meta_implicit_parametric_operator_learning.py:95-105).  The cotangents returned below carry no
            # graph, i.e. they are constants of the outer differentiation -- which is what the reference computes
            # only where the element residual sits under stop_gradient.  Anything else must not pass silently.
            second_order = loss._check_second_order(params_need_grad=ctx.needs_input_grad[1])
        else:
            second_order = "constant"
        grad_u, grad_k = ctx.grads
        up = None
        for g in (g_mean, g_mean2):
            if g is not None:
                up = g if up is None else up + g
        if up is None:
            up = torch.zeros((), dtype=loss.dtype, device=grad_u.device)
        up = up.to(dtype=loss.dtype).contiguous()       # device scalar: the kernel reads it, the host never does
        nb = grad_u.shape[0]
        _lib.check(_lib.load().fol_scale_grads(_lib.stream_ptr(), loss._dt, nb, loss.total_number_of_dofs,
                                               loss.fe_mesh.GetNumberOfNodes(), _lib.ptr(ctx.scale), 1.0, _lib.ptr(up),
                                               1 if ctx.prescaled else 0,
                                               _lib.ptr(loss._dir_flag if ctx.mask_dirichlet else loss._no_flag),
                                               _lib.ptr(grad_u), _lib.ptr(grad_k) if grad_k is not None else None))
        if second_order == "hessian":
            # true potential whose tangent stiffness IS the Hessian (Neo-Hooke): hand the cotangents out through a
            # node that knows their derivatives (K v kernel + nested VJP)
            batch_params, batch_dofs = ctx.saved_tensors
            grad_k, grad_u = _SecondOrderFn.apply(loss, batch_params, batch_dofs, up,
                                                  _Payload(grad_k, grad_u, ctx.mask_dirichlet, ctx.fuse_dirichlet))
        return None, (grad_k if ctx.need[0] else None), (grad_u if ctx.need[1] else None), None, None


class _Payload:
    def __init__(self, grad_k, grad_u, mask_dirichlet, fuse_dirichlet):
        self.grad_k, self.grad_u = grad_k, grad_u
        self.mask_dirichlet, self.fuse_dirichlet = mask_dirichlet, fuse_dirichlet


class _SecondOrderFn(torch.autograd.Function):
    """The first-order cotangents of ComputeBatchLoss as differentiable functions of (params, dofs), for losses
    whose energy is a true potential with Hessian = tangent stiffness (Neo-Hooke, exponent 1):
        g_u[b] = (up/B) mask (.) F_int(u_b, K_b)          g_K[b] = (up/B) dE/dK(u_b, K_b)
    so for cotangents (w_K, w_u) on them
        d/du_b = (up/B) mask (.) [ K_T(u_b, K_b) (mask (.) w_u[b])  +  F_int(u_b; controls = w_K[b]) ]
        d/dK_b = (up/B) (mask (.) w_u[b])^T dF_int/dK          (F_int is linear in K, so d2E/dK2 = 0)
    K_T v is the matrix-free product of fol_apply_jacobian_elements without the Dirichlet row mask, the last
    line is fol_residual_adjoint_elements, F_int(u; w_K) is one more call of the batched energy kernel.
    Every kernel runs once for the whole batch (grid.y = sample)."""

    @staticmethod
    def forward(ctx, loss, batch_params, batch_dofs, up, payload):
        ctx.loss, ctx.payload = loss, payload
        ctx.save_for_backward(batch_params, batch_dofs, up)
        ctx.set_materialize_grads(False)
        gk = payload.grad_k if payload.grad_k is not None else torch.zeros_like(batch_params)
        return gk, payload.grad_u

    @staticmethod
    def backward(ctx, w_k, w_u):
        loss, pl = ctx.loss, ctx.payload
        params, dofs, up = ctx.saved_tensors
        lib = _lib.load()
        nb = dofs.shape[0]
        s = _lib.stream_ptr()
        u_full = loss.GetFullDofVector(None, dofs) if pl.fuse_dirichlet else dofs
        scale = up.to(loss.dtype) / nb
        d_dofs = torch.zeros_like(dofs)
        d_params = torch.zeros_like(params) if ctx.needs_input_grad[1] else None
        if w_u is not None:
            w_u = w_u.to(loss.dtype).contiguous().clone()
            if pl.mask_dirichlet:
                w_u[:, loss._dir_idx.to(torch.int64)] = 0
            # ONE launch per kernel for the whole batch (grid.y = sample): K_T v, its fixed-order node sums, v^T dF_int/dK
            phys = _lib.PHYSICS[loss.physics]
            params_c, u_c = params.contiguous(), u_full.contiguous()
            ye = torch.empty((nb, max(loss._ne * loss._nd, 1)), dtype=loss.dtype, device=loss.device)
            _lib.check(lib.fol_apply_jacobian_elements_batched(
                s, loss._dt, phys, loss.fe_element.code, loss.num_gp, 0, loss._ne, loss._nn, nb, _lib.ptr(loss._xyz),
                _lib.ptr(loss._conn), _lib.ptr(params_c), _lib.ptr(u_c), _lib.ptr(loss._no_flag), loss._params,
                _lib.ptr(w_u), _lib.ptr(ye)))
            _lib.check(lib.fol_residual_gather_batched(s, loss._dt, loss._nn, loss._nnode, loss.number_dofs_per_node, nb,
                                                       loss._ne, _lib.ptr(loss._adj_ptr), _lib.ptr(loss._adj),
                                                       _lib.ptr(ye), _lib.ptr(d_dofs)))
            if d_params is not None:
                dk_e = torch.empty((nb, max(loss._ne * loss._nnode, 1)), dtype=loss.dtype, device=loss.device)
                _lib.check(lib.fol_residual_adjoint_elements_batched(
                    s, loss._dt, phys, loss.fe_element.code, loss.num_gp, loss._ne, loss._nn, nb, _lib.ptr(loss._xyz),
                    _lib.ptr(loss._conn), _lib.ptr(params_c), _lib.ptr(u_c), _lib.ptr(w_u), None, loss._params,
                    _lib.ptr(dk_e)))
                _lib.check(lib.fol_residual_gather_batched(s, loss._dt, loss._nn, loss._nnode, 1, nb, loss._ne,
                                                           _lib.ptr(loss._adj_ptr), _lib.ptr(loss._adj),
                                                           _lib.ptr(dk_e), _lib.ptr(d_params)))
        if w_k is not None:
            # F_int is linear in the control field: dF_int/dK . w_K = F_int(u; controls = w_K)
            _, fint, _ = loss._energy_and_grads(w_k.to(loss.dtype).contiguous(), u_full.contiguous())
            d_dofs = d_dofs + fint
        d_dofs = d_dofs * scale
        if pl.mask_dirichlet:
            d_dofs[:, loss._dir_idx.to(torch.int64)] = 0
        if d_params is not None:
            d_params = d_params * scale
        return None, d_params, d_dofs, None, None


class FiniteElementLoss(Loss):
    """FE-based loss; subclasses fix physics, compute_dims, ordered_dofs and element_type."""

    physics = None  # key of _lib.PHYSICS

    def __init__(self, name: str, loss_settings: dict, fe_mesh):
        super().__init__(name)
        self.loss_settings = loss_settings
        self.dofs = self.loss_settings["ordered_dofs"]
        self.element_type = self.loss_settings["element_type"]
        self.fe_mesh = fe_mesh
        if "dirichlet_bc_dict" not in self.loss_settings.keys():
            fol_error("dirichlet_bc_dict should provided in the loss settings !", self.GetName())

    # ------------------------------------------------------------------ setup
    def _create_dofs_dict(self, dofs_list, dirichlet_bc_dict):
        """fe_loss.py:34-55: dof-major, then boundary name, values broadcast per node set."""
        d = len(dofs_list)
        idx, val = [], []
        for dof_index, dof in enumerate(dofs_list):
            for boundary_name, boundary_value in dirichlet_bc_dict[dof].items():
                nodes = np.asarray(self.fe_mesh.GetNodeSet(boundary_name), dtype=np.int64)
                idx.append(d * nodes + dof_index)
                val.append(np.full(nodes.shape, float(boundary_value)))
        if idx:
            self.dirichlet_indices = np.concatenate(idx).astype(np.int32)
            self.dirichlet_values = np.concatenate(val)
        else:
            self.dirichlet_indices = np.zeros(0, dtype=np.int32)
            self.dirichlet_values = np.zeros(0)
        all_indices = np.arange(d * self.fe_mesh.GetNumberOfNodes())
        self.non_dirichlet_indices = np.setdiff1d(all_indices, self.dirichlet_indices).astype(np.int32)

    def _material_params(self):
        """double[FOL_NUM_PARAMS] of include/folax_b200.h; subclasses fill it."""
        return [0.0] * _lib.NUM_PARAMS

    def Initialize(self, reinitialize=False) -> None:
        if self.initialized and not reinitialize:
            return
        _lib.require_cuda()
        lib = _lib.load()
        self.dtype = {"float64": torch.float64, "float32": torch.float32}[
            str(self.loss_settings.get("dtype", "float64")).replace("torch.", "")]
        self._dt = _lib.dtype_code(self.dtype)
        self.device = torch.device("cuda", torch.cuda.current_device())

        self.number_dofs_per_node = len(self.dofs)
        self.total_number_of_dofs = len(self.dofs) * self.fe_mesh.GetNumberOfNodes()
        self._create_dofs_dict(self.dofs, self.loss_settings["dirichlet_bc_dict"])
        self.number_of_unknown_dofs = self.non_dirichlet_indices.size

        self.fe_element = ElementInfo(self.element_type)
        if "num_gp" in self.loss_settings.keys():
            self.num_gp = self.loss_settings["num_gp"]
            if self.num_gp not in (1, 2, 3):
                raise ValueError(f" number gauss points {self.num_gp} is not supported ! ")
        else:
            self.loss_settings["num_gp"] = 1
            self.num_gp = 1
        self.fe_element.SetGaussIntegrationMethod(f"GI_GAUSS_{self.num_gp}")
        if "compute_dims" not in self.loss_settings.keys():
            raise ValueError(f"compute_dims must be provided in the loss settings of {self.GetName()}! ")
        self.dim = self.loss_settings["compute_dims"]
        self._nnode, self._edim, self._ngauss = self.fe_element.info()
        expected = lib.fol_dofs_per_node(_lib.PHYSICS[self.physics], self.fe_element.code)
        if expected != self.number_dofs_per_node:
            raise ValueError(f"{self.GetName()}: {self.number_dofs_per_node} dofs per node given, "
                             f"{self.physics} on {self.element_type} needs {expected}")
        # fe_loss.py:94-100: in parametric boundary learning the batch parameters ARE the Dirichlet values
        self._parametric = bool(self.loss_settings.get("parametric_boundary_learning"))
        self.loss_function_exponent = self.loss_settings.get("loss_function_exponent", 1.0)

        # element batching of the reference (fe_loss.py:103-114) only bounds XLA memory; kept as
        # attributes because callers may read them -- the kernels process the mesh in one launch
        ne = self.fe_mesh.GetNumberOfElements(self.element_type)
        self.adjusted_batch_size, self.num_element_batches = ne, 1

        # ---- device-resident plan
        conn = np.ascontiguousarray(self.fe_mesh.GetElementsNodes(self.element_type), dtype=np.int32)
        nn = self.fe_mesh.GetNumberOfNodes()
        if conn.ndim != 2 or conn.shape[1] != self._nnode:
            raise ValueError(f"{self.element_type} connectivity must be (ne, {self._nnode})")
        if ne and (conn.min() < 0 or conn.max() >= nn):
            raise ValueError("connectivity refers to nodes outside the mesh")
        self._ne, self._nn = ne, nn
        self._nd = self._nnode * self.number_dofs_per_node
        self._xyz = _lib.to_device(np.asarray(self.fe_mesh.GetNodesCoordinates(), dtype=np.float64), self.dtype)
        self._conn = torch.as_tensor(conn, device=self.device)
        self._dir_idx = torch.as_tensor(self.dirichlet_indices, device=self.device)
        self._dir_val = _lib.to_device(self.dirichlet_values, self.dtype)
        self._dir_flag = torch.empty(max(self.total_number_of_dofs, 1), dtype=torch.uint8, device=self.device)
        s = _lib.stream_ptr()
        _lib.check(lib.fol_dirichlet_flags(s, _lib.ptr(self._dir_idx), self._dir_idx.numel(),
                                           self.total_number_of_dofs, _lib.ptr(self._dir_flag)))
        self._adj_ptr = torch.empty(nn + 1, dtype=torch.int32, device=self.device)
        self._adj = torch.empty(max(ne * self._nnode, 1), dtype=torch.int32, device=self.device)
        work = torch.empty(max(nn, 1), dtype=torch.int32, device=self.device)
        _lib.check(lib.fol_node_adjacency(s, _lib.ptr(self._conn), ne, self._nnode, nn, _lib.ptr(self._adj_ptr),
                                          _lib.ptr(self._adj), _lib.ptr(work)))
        self._no_flag = torch.zeros_like(self._dir_flag)
        # Dirichlet value per dof, NaN where free (the batched-loss kernel overwrites while staging)
        full = np.full(max(self.total_number_of_dofs, 1), np.nan)
        full[np.asarray(self.dirichlet_indices, dtype=np.int64)] = np.asarray(self.dirichlet_values, dtype=float)
        self._dir_full = _lib.to_device(full, self.dtype)
        self._indices = None   # BCOO indices, built on the first Jacobian request
        self._geom = None      # geometry cache, built on the first batched-loss request
        self._params = _lib.params_array(self._material_params())
        self.initialized = True

    # ------------------------------------------------------------------ small API
    def GetFullDofVector(self, known_dofs, unknown_dofs):
        """fe_loss.py:91-100, 123-124: all_dofs[:, dirichlet_indices] = dirichlet_values (out of place); with
        `parametric_boundary_learning` the per-sample `known_dofs` (B, n_dirichlet) are written instead
        (ConstructFullDofVectorParametricLearning -- what the Predict paths call,
        explicit_parametric_operator_learning.py:117)."""
        u = self._as_batch(unknown_dofs, self.total_number_of_dofs).clone()
        if self._parametric:
            if known_dofs is None:
                raise ValueError(f"{self.GetName()}: parametric_boundary_learning needs known_dofs (B, n_dirichlet)")
            known = self._as_batch(known_dofs, self.dirichlet_indices.size)
            if known.shape[0] == 1 and u.shape[0] > 1:
                known = known.expand(u.shape[0], -1).contiguous()
            if known.shape[0] != u.shape[0]:
                raise ValueError(f"{self.GetName()}: {known.shape[0]} samples of known dofs for {u.shape[0]} dof vectors")
            _lib.check(_lib.load().fol_apply_dirichlet(_lib.stream_ptr(), self._dt, u.shape[0],
                                                       self.total_number_of_dofs, _lib.ptr(self._dir_idx),
                                                       self._dir_idx.numel(), _lib.ptr(known), 1, 1.0, _lib.ptr(u)))
            return u
        _lib.check(_lib.load().fol_apply_dirichlet(_lib.stream_ptr(), self._dt, u.shape[0],
                                                   self.total_number_of_dofs, _lib.ptr(self._dir_idx),
                                                   self._dir_idx.numel(), _lib.ptr(self._dir_val), 0, 1.0,
                                                   _lib.ptr(u)))
        return u

    def GetParametersVectors(self, param_vector):
        """fe_loss.py:127-128 (identity) -- losses with a mesh-resident heterogeneity field override the
        control field in parametric boundary learning (mechanical_saint_venant.py:59-66)."""
        field = self.__dict__.get("heterogeneity_field")
        if self.__dict__.get("_parametric") and field is not None:
            return field
        return param_vector

    def Finalize(self) -> None:
        pass

    def GetNumberOfUnknowns(self):
        return self.number_of_unknown_dofs

    def GetTotalNumberOfDOFs(self):
        return self.total_number_of_dofs

    def GetDOFs(self):
        return self.dofs

    def ApplyDirichletBCOnDofVector(self, full_dof_vector, load_increment: float = 1.0):
        """fe_loss.py:186-189: u[dirichlet_indices] = load_increment * dirichlet_values."""
        u = _lib.to_device(full_dof_vector, self.dtype).reshape(1, -1).clone()
        _lib.check(_lib.load().fol_apply_dirichlet(_lib.stream_ptr(), self._dt, 1, self.total_number_of_dofs,
                                                   _lib.ptr(self._dir_idx), self._dir_idx.numel(),
                                                   _lib.ptr(self._dir_val), 0, float(load_increment), _lib.ptr(u)))
        return u.reshape(-1)

    def _as_batch(self, x, width):
        t = _lib.to_device(x, self.dtype) if not (isinstance(x, torch.Tensor) and x.is_cuda
                                                   and x.dtype == self.dtype) else x
        t = torch.atleast_2d(t)
        t = t.reshape(t.shape[0], -1)
        if t.shape[1] != width:
            raise ValueError(f"{self.GetName()}: expected vectors of length {width}, got {t.shape[1]}")
        return t.contiguous()

    # ------------------------------------------------------------------ residual + Jacobian
    def _bcoo_indices(self):
        if self._indices is None:
            n = self._ne * self._nd * self._nd
            self._indices = torch.empty((n, 2), dtype=torch.int32, device=self.device)
            _lib.check(_lib.load().fol_bcoo_indices(_lib.stream_ptr(), _lib.ptr(self._conn), self._ne,
                                                    self._nnode, self.number_dofs_per_node,
                                                    _lib.ptr(self._indices)))
        return self._indices

    def _assemble(self, controls, dofs, transpose, ke_out=None, state_in=None, state_out=None):
        lib = _lib.load()
        ctrl = _lib.to_device(controls, self.dtype).reshape(-1)
        u = _lib.to_device(dofs, self.dtype).reshape(-1)
        if ctrl.numel() != self._nn or u.numel() != self.total_number_of_dofs:
            raise ValueError(f"{self.GetName()}: controls must have {self._nn} entries and dofs "
                             f"{self.total_number_of_dofs}")
        data = ke_out if ke_out is not None else torch.empty(self._ne * self._nd * self._nd, dtype=self.dtype,
                                                             device=self.device)
        re = torch.empty(max(self._ne * self._nd, 1), dtype=self.dtype, device=self.device)
        R = torch.empty(self.total_number_of_dofs, dtype=self.dtype, device=self.device)
        s = _lib.stream_ptr()
        _lib.check(lib.fol_assemble_elements(s, self._dt, _lib.PHYSICS[self.physics], self.fe_element.code,
                                             self.num_gp, int(bool(transpose)), self._ne, self._nn,
                                             _lib.ptr(self._xyz), _lib.ptr(self._conn), _lib.ptr(ctrl), _lib.ptr(u),
                                             _lib.ptr(self._dir_flag), self._params, _lib.ptr(data), _lib.ptr(re),
                                             _lib.ptr(state_in), _lib.ptr(state_out)))
        _lib.check(lib.fol_residual_gather(s, self._dt, self._nn, self._nnode, self.number_dofs_per_node,
                                           _lib.ptr(self._adj_ptr), _lib.ptr(self._adj), _lib.ptr(re), _lib.ptr(R)))
        return data, R

    def ComputeJacobianMatrixAndResidualVector(self, total_control_vars, total_primal_vars,
                                               transpose_jacobian: bool = False, new_implementation: bool = False):
        """fe_loss.py:264-318 -> (BCOO with duplicates, residual vector)."""
        data, R = self._assemble(total_control_vars, total_primal_vars, transpose_jacobian)
        jac = BCOO((data, self._bcoo_indices()), shape=(self.total_number_of_dofs, self.total_number_of_dofs))
        return jac, R

    def ApplyJacobian(self, total_control_vars, total_primal_vars, vector, transpose_jacobian: bool = False,
                      state_in=None):
        """y = J(controls, dofs) @ vector without forming J: what the consumers of the BCOO do with
        `jacobian @ v` (fe_solver.py:61) and what a Krylov solver needs every iteration.  J is the matrix
        ComputeJacobianMatrixAndResidualVector(..., transpose_jacobian) returns (Dirichlet rows included);
        the element matrices live in registers only, so the cost is the element arithmetic, not the
        nd^2 * 8 bytes per element of the assembled format."""
        lib = _lib.load()
        ctrl = _lib.to_device(total_control_vars, self.dtype).reshape(-1)
        u = _lib.to_device(total_primal_vars, self.dtype).reshape(-1)
        v = _lib.to_device(vector, self.dtype).reshape(-1).contiguous()
        if ctrl.numel() != self._nn or u.numel() != self.total_number_of_dofs or v.numel() != u.numel():
            raise ValueError(f"{self.GetName()}: controls must have {self._nn} entries, dofs and vector "
                             f"{self.total_number_of_dofs}")
        ye = torch.empty(max(self._ne * self._nd, 1), dtype=self.dtype, device=self.device)
        y = torch.empty(self.total_number_of_dofs, dtype=self.dtype, device=self.device)
        s = _lib.stream_ptr()
        _lib.check(lib.fol_apply_jacobian_elements(s, self._dt, _lib.PHYSICS[self.physics], self.fe_element.code,
                                                   self.num_gp, int(bool(transpose_jacobian)), self._ne, self._nn,
                                                   _lib.ptr(self._xyz), _lib.ptr(self._conn), _lib.ptr(ctrl),
                                                   _lib.ptr(u), _lib.ptr(self._dir_flag), self._params, _lib.ptr(v),
                                                   _lib.ptr(ye), _lib.ptr(state_in)))
        _lib.check(lib.fol_residual_gather(s, self._dt, self._nn, self._nnode, self.number_dofs_per_node,
                                           _lib.ptr(self._adj_ptr), _lib.ptr(self._adj), _lib.ptr(ye), _lib.ptr(y)))
        return y

    def _csr_plan(self):
        if self.__dict__.get("_cplan") is None:
            from .. import csr_plan
            plan = csr_plan.build(self.fe_mesh.GetElementsNodes(self.element_type), self._nn,
                                  self.number_dofs_per_node)
            self._cplan = {k: (torch.as_tensor(v, device=self.device) if isinstance(v, np.ndarray) else v)
                           for k, v in plan.items()}
        return self._cplan

    def _sell_plan(self):
        """Device-resident sliced-ELLPACK plan of the duplicate-free CSR (sell_plan.py), for the Krylov solvers."""
        if self.__dict__.get("_splan") is None:
            from .. import sell_plan
            cp = self._csr_plan()
            plan = sell_plan.build(cp["indptr"].cpu().numpy(), cp["indices"].cpu().numpy(), self.number_dofs_per_node)
            self._splan = {k: (torch.as_tensor(v, device=self.device) if isinstance(v, np.ndarray) else v)
                           for k, v in plan.items()}
        return self._splan

    def JacobianToCSR(self, jacobian):
        """Duplicate-free CSR (indptr, indices, values) of a Jacobian returned by
        ComputeJacobianMatrixAndResidualVector -- the sum the reference's solvers do on the host with
        scipy.sparse.csr_array (fe_solver.py:71-72), done on the GPU in a fixed order."""
        cp = self._csr_plan()
        vals = torch.empty(cp["nnz"], dtype=self.dtype, device=self.device)
        _lib.check(_lib.load().fol_csr_values(_lib.stream_ptr(), self._dt, cp["npairs"], self.number_dofs_per_node,
                                              self._nnode, _lib.ptr(cp["pair_ptr"]), _lib.ptr(cp["contrib"]),
                                              _lib.ptr(cp["out_base"]), _lib.ptr(cp["row_stride"]),
                                              _lib.ptr(jacobian.data), _lib.ptr(vals)))
        return cp["indptr"], cp["indices"], vals

    def ComputeElement(self, elem_xyz, elem_controls, elem_dofs):
        """Single-element evaluation (the unit-test entry point): (energy, re (nd,1), Ke (nd,nd)).
        Runs the same kernel on a one-element mesh."""
        lib = _lib.load()
        A, nd = self._nnode, self._nd
        xyz = _lib.to_device(elem_xyz, self.dtype).reshape(A, 3)
        ctrl = _lib.to_device(elem_controls, self.dtype).reshape(-1)
        if ctrl.numel() == 1:
            ctrl = ctrl.expand(A).contiguous()
        u = _lib.to_device(elem_dofs, self.dtype).reshape(nd)
        conn = torch.arange(A, dtype=torch.int32, device=self.device).reshape(1, A)
        flags = torch.zeros(nd, dtype=torch.uint8, device=self.device)
        ke = torch.empty(nd * nd, dtype=self.dtype, device=self.device)
        re = torch.empty(nd, dtype=self.dtype, device=self.device)
        _lib.check(lib.fol_assemble_elements(_lib.stream_ptr(), self._dt, _lib.PHYSICS[self.physics],
                                             self.fe_element.code, self.num_gp, 0, 1, A, _lib.ptr(xyz),
                                             _lib.ptr(conn), _lib.ptr(ctrl), _lib.ptr(u), _lib.ptr(flags),
                                             self._params, _lib.ptr(ke), _lib.ptr(re), None, None))
        energy = self._element_energy(xyz, conn, ctrl, u, re)
        return energy, re.reshape(nd, 1), ke.reshape(nd, nd)

    def _element_energy(self, xyz, conn, ctrl, u, re):
        # mechanical.py:116-117 / thermal.py:45-49: energy = u^T stop_gradient(re)
        return torch.dot(u, re)

    # ------------------------------------------------------------------ element-level helpers of the reference API
    def ComputeElementEnergy(self, elem_xyz, elem_controls, elem_dofs):
        """fe_loss.py:149-153."""
        return self.ComputeElement(elem_xyz, elem_controls, elem_dofs)[0]

    def ComputeElementsEnergies(self, total_control_vars, total_primal_vars):
        """fe_loss.py:166-173 -> (ne,) energies, the first return value of ComputeElement for every element
        (one kernel over the mesh, csrc/adjoint.cuh: element_energy)."""
        ctrl = _lib.to_device(total_control_vars, self.dtype).reshape(-1)
        u = _lib.to_device(total_primal_vars, self.dtype).reshape(-1)
        if ctrl.numel() != self._nn or u.numel() != self.total_number_of_dofs:
            raise ValueError(f"{self.GetName()}: controls must have {self._nn} entries and dofs "
                             f"{self.total_number_of_dofs}")
        en = torch.empty(max(self._ne, 1), dtype=self.dtype, device=self.device)
        _lib.check(_lib.load().fol_element_energies(
            _lib.stream_ptr(), self._dt, _lib.PHYSICS[self.physics], self.fe_element.code, self.num_gp, self._ne,
            _lib.ptr(self._xyz), _lib.ptr(self._conn), _lib.ptr(ctrl), _lib.ptr(u), _lib.ptr(self._geom_aux),
            self._params, _lib.ptr(en)))
        return en[:self._ne]

    def ComputeElementJacobianIndices(self, nodes_ids):
        """fe_loss.py:178-184 -> (nd*nd, 2) global (row, col) pairs of one element, in the order of the BCOO."""
        nodes = torch.as_tensor(np.ascontiguousarray(np.asarray(nodes_ids).reshape(1, -1), dtype=np.int32),
                                device=self.device)
        nd = nodes.shape[1] * self.number_dofs_per_node
        out = torch.empty((nd * nd, 2), dtype=torch.int32, device=self.device)
        _lib.check(_lib.load().fol_bcoo_indices(_lib.stream_ptr(), _lib.ptr(nodes), 1, nodes.shape[1],
                                                self.number_dofs_per_node, _lib.ptr(out)))
        return out

    def ApplyDirichletBCOnElementResidualAndJacobian(self, elem_res, elem_jac, elem_BC_vec, elem_mask_BC_vec):
        """fe_loss.py:191-207 on ONE element given as arrays: BC (.) re,  diag(BC) Ke + diag(diag(M Ke M)), M = diag(mask).
        (Inside the assembly kernels this is a row mask applied while storing; this method exists for callers that
        hold an element matrix.)"""
        re = _lib.to_device(elem_res, self.dtype).reshape(-1, 1)
        ke = _lib.to_device(elem_jac, self.dtype)
        bc = _lib.to_device(elem_BC_vec, self.dtype).reshape(-1)
        mask = _lib.to_device(elem_mask_BC_vec, self.dtype).reshape(-1)
        kept = mask * mask * torch.diagonal(ke)
        return bc.reshape(-1, 1) * re, bc.reshape(-1, 1) * ke + torch.diag(kept)

    def ComputeElementResidualAndJacobian(self, elem_xyz, elem_controls, elem_dofs, elem_BC, elem_mask_BC,
                                          transpose_jac: bool):
        """fe_loss.py:209-230: ComputeElement, optional transpose, then the Dirichlet treatment above."""
        _, re, ke = self.ComputeElement(elem_xyz, elem_controls, elem_dofs)
        if transpose_jac:
            ke = ke.t()
        return self.ApplyDirichletBCOnElementResidualAndJacobian(re, ke, elem_BC, elem_mask_BC)

    # ------------------------------------------------------------------ batched loss
    _geom_aux = None   # nodal field folded into the geometry cache (transient thermal: k0)

    def _geometry_cache(self):
        if self._geom is None:
            lib, phys = _lib.load(), _lib.PHYSICS[self._batch_physics()]
            width = lib.fol_geometry_width(phys, self.fe_element.code)
            self._geom = torch.empty(max(self._ne * self._ngauss * width, 1), dtype=self.dtype, device=self.device)
            _lib.check(lib.fol_geometry_cache_physics(_lib.stream_ptr(), self._dt, phys, self.fe_element.code,
                                                      self.num_gp, self._ne, _lib.ptr(self._xyz),
                                                      _lib.ptr(self._conn), _lib.ptr(self._geom_aux),
                                                      _lib.ptr(self._geom)))
        return self._geom

    _has_control_gradient = True

    def _batch_physics(self):
        """Physics tag of the batched energy kernels (the element stage uses self.physics)."""
        return self.physics

    def _energy_plan(self):
        """Device-resident integer tile plan of the fused batched-loss kernel (energy_plan.py)."""
        if self.__dict__.get("_eplan") is None:
            from .. import energy_plan
            import os
            # rows of the generic kernel's shared memory: S samples x KW element-vector rows x ecap (energy.cu)
            nd = self._nnode * self.number_dofs_per_node
            kw_rows = nd + (self._nnode if self._has_control_gradient else 0) + 1
            samples = 4 if nd <= 8 else 2
            kw = {"generic_max_elems": int(150 * 1024 / (torch.empty(0, dtype=self.dtype).element_size() * samples * kw_rows))}
            # the pipelined kernel exists where the geometry factors fit in registers (energy.cuh: geom_in_regs)
            if (self.element_type, self.num_gp) not in (("quad", 1), ("quad", 2), ("triangle", 1), ("triangle", 2),
                                                        ("tetra", 1), ("hexahedron", 1)):
                kw["max_elems"] = None
            coords = np.asarray(self.fe_mesh.GetNodesCoordinates())
            conn = self.fe_mesh.GetElementsNodes(self.element_type)
            if (self._batch_physics() == "thermal" and self.element_type == "quad" and self.num_gp == 2
                    and self.dtype == torch.float64 and energy_plan.is_affine(coords, conn)):
                # affine Quad4 meshes run the 256-thread sample-vectorised kernel (csrc/energy_qt.cuh): larger tiles,
                # fewer border elements evaluated twice (1.16x instead of 1.19x at 256x256)
                kw.update(max_elems=256, tile_nodes=225)
            if "FOL_ENERGY_MAX_ELEMS" in os.environ:      # tuning experiments only (scripts/energy_sweep.sh)
                kw.update(max_elems=int(os.environ["FOL_ENERGY_MAX_ELEMS"]),
                          tile_nodes=int(os.environ.get("FOL_ENERGY_TILE_NODES", energy_plan.TILE_NODES)))
            plan = energy_plan.build(coords, conn, **kw)
            self._eplan = {k: (torch.as_tensor(v, device=self.device) if isinstance(v, np.ndarray) else v)
                           for k, v in plan.items()}
        return self._eplan

    def _grid_plan(self):
        """Structured-grid facts for csrc/energy_grid.cu (thermal Quad4, 2 x 2 rule), or None (tile kernels).
        FOL_ENERGY_GRID=0 keeps the tile kernels (A/B measurements)."""
        if "_grid" not in self.__dict__:
            import os
            from .. import energy_plan
            g = None
            if (self._batch_physics() in ("thermal", "mechanical") and self.element_type == "quad" and self.num_gp == 2
                    and os.environ.get("FOL_ENERGY_GRID", "1") != "0"):
                g = energy_plan.grid_structure(np.asarray(self.fe_mesh.GetNodesCoordinates()),
                                               self.fe_mesh.GetElementsNodes(self.element_type))
                if g is not None:
                    g["jinv_c"] = (C.c_double * 4)(*[float(v) for v in g["jinv"]])
                    col = np.zeros(g["nx"] + 1, dtype=np.uint8)      # node columns that hold a Dirichlet node
                    nodes = np.asarray(self.dirichlet_indices, dtype=np.int64) // self.number_dofs_per_node
                    col[nodes % (g["nx"] + 1)] = 1
                    g["col_dir"] = torch.as_tensor(col, device=self.device)
            self._grid = g
        return self._grid

    def _energy_work(self, nb):
        """Scratch of the batched-loss kernel, cached per batch size (no allocation on the hot call)."""
        cache = self.__dict__.setdefault("_work_cache", {})
        if nb not in cache:
            n = _lib.load().fol_energy_work_size(self._energy_plan()["ntiles"], nb)
            cache.clear()
            cache[nb] = torch.empty(n, dtype=self.dtype, device=self.device)
        return cache[nb]

    def _energy_and_grads(self, batch_params, batch_dofs, dir_values=None, dir_flag=None, out_scale=1.0):
        """(E_b, out_scale * dE_b/du_b (un-masked assembled residual, zero where dir_flag), out_scale * dE_b/dK_b);
        the dofs are BC-applied already, or dir_values (ndof, NaN = free) is applied by the kernel."""
        lib = _lib.load()
        nb = batch_dofs.shape[0]
        grad_u = torch.empty_like(batch_dofs)
        grad_k = torch.empty_like(batch_params) if self._has_control_gradient else None
        energy = torch.empty(nb, dtype=self.dtype, device=self.device)
        grid = self._grid_plan()
        if grid is not None:              # structured Quad4 grid: the marching kernel (no connectivity, no tile plan)
            cache = self.__dict__.setdefault("_grid_work_cache", {})
            if nb not in cache:
                cache.clear()
                size = max(int(lib.fol_energy_grid_work_size(grid["nx"], grid["ny"], nb)),
                           int(lib.fol_energy_grid_mech_work_size(grid["nx"], grid["ny"], nb)))
                cache[nb] = torch.empty(size, dtype=self.dtype, device=self.device)
            if self._batch_physics() == "mechanical":          # csrc/energy_grid_mech.cu: two dofs per node, dE/dK = 0
                _lib.check(lib.fol_energy_and_grads_grid_mech(
                    _lib.stream_ptr(), self._dt, grid["nx"], grid["ny"], nb, grid["jinv_c"], grid["wdetj"],
                    _lib.ptr(_aligned16(batch_params)), _lib.ptr(_aligned16(batch_dofs)),
                    _lib.ptr(dir_values) if dir_values is not None else None,
                    _lib.ptr(dir_flag) if dir_flag is not None else None, _lib.ptr(grid["col_dir"]), float(out_scale),
                    self._params, _lib.ptr(grad_u), _lib.ptr(energy), _lib.ptr(cache[nb])))
                return energy, grad_u, grad_k
            _lib.check(lib.fol_energy_and_grads_grid(
                _lib.stream_ptr(), self._dt, grid["nx"], grid["ny"], nb, grid["jinv_c"], grid["wdetj"],
                _lib.ptr(_aligned16(batch_params)), _lib.ptr(_aligned16(batch_dofs)),
                _lib.ptr(dir_values) if dir_values is not None else None,
                _lib.ptr(dir_flag) if dir_flag is not None else None, _lib.ptr(grid["col_dir"]), float(out_scale),
                self._params,
                _lib.ptr(grad_u), _lib.ptr(grad_k), _lib.ptr(energy), _lib.ptr(cache[nb])))
            return energy, grad_u, grad_k
        geom, ep = self._geometry_cache(), self._energy_plan()
        work = self._energy_work(nb)
        _lib.check(lib.fol_energy_and_grads_flags(_lib.stream_ptr(), self._dt, _lib.PHYSICS[self._batch_physics()],
                                            self.fe_element.code, self.num_gp, self._ne, self._nn, nb,
                                            _lib.ptr(geom), _lib.ptr(self._conn), _lib.ptr(ep["adj_ptr"]),
                                            _lib.ptr(ep["adj_local"]), _lib.ptr(ep["tile_node_ptr"]),
                                            _lib.ptr(ep["tile_nodes"]), _lib.ptr(ep["tile_elem_ptr"]),
                                            _lib.ptr(ep["tile_elems"]), _lib.ptr(ep["tile_conn"]),
                                            _lib.ptr(ep["tile_lnode_ptr"]), _lib.ptr(ep["tile_lnodes"]), ep["ntiles"],
                                            ep["ecap"], ep["lcap"], ep["ncap"],
                                            _lib.ptr(batch_params), _lib.ptr(batch_dofs),
                                            _lib.ptr(dir_values) if dir_values is not None else None,
                                            _lib.ptr(dir_flag) if dir_flag is not None else None, float(out_scale),
                                            self._params,
                                            _lib.ptr(grad_u), _lib.ptr(grad_k), _lib.ptr(energy), _lib.ptr(work),
                                            _lib.MESH_AFFINE if ep.get("affine") else 0))
        return energy, grad_u, grad_k

    # how the reference's energy behaves under a SECOND differentiation:
    #   "zero"        E = u^T stop_gradient(re) (mechanical.py:116): every second derivative is zero
    #   "zero_in_u"   thermal.py:31, 45-46: T is stopped inside Se and re, so d2E/dT2 = 0, but d2E/dT dK is not
    #   "hessian"     Neo-Hooke: a true potential whose analytic tangent IS the Hessian (checked on the oracle to
    #                 1e-15): second-order terms from the matrix-free K v kernel (_SecondOrderFn)
    #   None          the other true potentials: St-Venant's reference tangent is not its Hessian (unsymmetrised
    #                 identity, saint_venant.py:24-27), the implicit scalar energies' Hessians differ from their Ke
    _second_order = None

    def _check_second_order(self, params_need_grad):
        """-> "constant" (the cotangents are constants of the outer differentiation) | "hessian" (route them
        through _SecondOrderFn); raises where neither is the reference's result."""
        if float(self.loss_function_exponent) == 1.0:       # E^p with p != 1 adds p(p-1)E^(p-2) dE dE^T
            if self._second_order == "zero":
                return "constant"
            if self._second_order == "zero_in_u" and not params_need_grad:
                return "constant"
            if self._second_order == "hessian" and not self._parametric:
                return "hessian"
        raise NotImplementedError(
            f"{self.GetName()}: second-order differentiation of ComputeBatchLoss is not available for this loss "
            "(the first-order cotangents would be treated as constants, which differs from the reference here)")

    def ComputeBatchLoss(self, batch_params, batch_dofs):
        """fe_loss.py:250-262 -> (mean_b E_b^p, (min, max, mean)); differentiable w.r.t. both inputs
        (torch.autograd.Function = the custom_vjp of the FFI call)."""
        dofs = self._as_batch(batch_dofs, self.total_number_of_dofs)
        if self._parametric:
            known = self._as_batch(batch_params, self.dirichlet_indices.size)
            full = _ApplyDirichletParametric.apply(self, known, dofs)
            ctrl = _lib.to_device(self.GetParametersVectors(batch_params), self.dtype)
            if ctrl.dim() == 1:
                ctrl = ctrl.reshape(1, -1).expand(dofs.shape[0], -1)
            params = self._as_batch(ctrl, self._nn)
            mean, mn, mx, mean2, _ = _BatchLossFn.apply(self, params, full, False, False)
        else:
            params = self._as_batch(batch_params, self._nn)
            mean, mn, mx, mean2, _ = _BatchLossFn.apply(self, params, dofs, True, True)
        return mean, (mn, mx, mean2)

    def ComputeTotalEnergy(self, total_control_vars, total_primal_vars):
        """fe_loss.py:175-176: sum of element energies for one sample (no Dirichlet overwrite)."""
        params = self._as_batch(total_control_vars, self._nn)
        dofs = self._as_batch(total_primal_vars, self.total_number_of_dofs)
        energy, _, _ = self._energy_and_grads(params, dofs)
        return energy[0]


class _ApplyDirichletParametric(torch.autograd.Function):
    """u[:, dirichlet_indices] = known (per sample): ConstructFullDofVectorParametricLearning, fe_loss.py:94-95.
    The cotangent at the Dirichlet entries flows to `known`, the rest to the free dofs."""

    @staticmethod
    def forward(ctx, loss, known, dofs):
        ctx.loss = loss
        u = dofs.clone()
        _lib.check(_lib.load().fol_apply_dirichlet(_lib.stream_ptr(), loss._dt, u.shape[0], loss.total_number_of_dofs,
                                                   _lib.ptr(loss._dir_idx), loss._dir_idx.numel(),
                                                   _lib.ptr(known.contiguous()), 1, 1.0, _lib.ptr(u)))
        return u

    @staticmethod
    def backward(ctx, g):
        idx = ctx.loss._dir_idx.to(torch.int64)
        g_known = g.index_select(1, idx)
        g_dofs = g.clone()
        g_dofs[:, idx] = 0
        return None, g_known, g_dofs


class _ApplyDirichlet(torch.autograd.Function):
    """u -> u with Dirichlet entries overwritten (fe_loss.py:91-92); the cotangent is cut there."""

    @staticmethod
    def forward(ctx, loss, dofs):
        ctx.loss = loss
        return loss.GetFullDofVector(None, dofs)

    @staticmethod
    def backward(ctx, g):
        # fol_scale_grads already zeroed the Dirichlet entries of the incoming cotangent
        return None, g
