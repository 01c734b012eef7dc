"""The reference-side binding: folax loss classes whose assembly and batched loss run in libfolax_b200 through
`jax.ffi`, with the batched loss wrapped in `jax.custom_vjp` so that `jax.grad` through the physics loss keeps
working (BASELINE.json north_star).

    import fol.loss_functions.mechanical as ref
    from folax_b200.ffi.jax_binding import register, accelerate
    register()                                                   # once per process
    MechanicalLoss3DHexaB200 = accelerate(ref.MechanicalLoss3DHexa, physics="mechanical")
    loss = MechanicalLoss3DHexaB200(name, loss_settings, fe_mesh)          # same constructor, same methods

It follows the reference's own FFI class line by line in style (`fol/loss_functions/kratos_small_displacement.py:25-27`
registers the capsules, `:96` / `:117` call `jax.ffi.ffi_call(name, ShapeDtypeStruct...)(*arrays)`), with the handlers of
`folax_b200/ffi/xla_ffi_shim.cc`.

STATUS: NEVER RUN.  JAX / jaxlib are not installable in the image this repository was built in (no wheel, no
network), so this module is checked only for syntax and for raising a clear error without JAX
(tests/test_cabi.py).  What IS tested is everything underneath: the C ABI it forwards to, and the same custom-VJP
logic in its torch form (`folax_b200/loss_functions/fe_loss.py::_BatchLossFn`).  Treat it as the starting point a
folax maintainer would debug against a real JAX install, not as a finished binding.
"""
import ctypes
import os

import numpy as np

from .. import _lib, energy_plan

_HERE = os.path.dirname(os.path.abspath(__file__))
XLA_LIB = os.path.join(_HERE, "libfolax_b200_xla.so")          # built per INTEGRATION.md (needs jaxlib's headers)
HANDLERS = ("FolAssembleElements", "FolResidualGather", "FolEnergyAndGrads", "FolGeometryCache",
            "FolApplyJacobianElements", "FolGaussInterpolate", "FolResponseElements", "FolResidualAdjointElements")


def _jax():
    try:
        import jax
        import jax.numpy as jnp
        from jax.experimental import sparse
    except ImportError as ex:
        raise _lib.FolaxError("folax_b200.ffi.jax_binding needs JAX (jax.ffi); it is not installed here. The torch "
                              "host in folax_b200.loss_functions runs the same kernels.") from ex
    return jax, jnp, sparse


def register(path=XLA_LIB):
    """kratos_small_displacement.py:25-27 for the handlers of xla_ffi_shim.cc."""
    jax, _, _ = _jax()
    if not os.path.exists(path):
        raise _lib.FolaxError(f"{path} is missing: build the XLA-FFI shim first (INTEGRATION.md, section 1)")
    lib = ctypes.CDLL(path)
    for name in HANDLERS:
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, name)), platform="CUDA")


def accelerate(loss_class, physics):
    """Subclass of a reference `FiniteElementLoss` class whose `ComputeJacobianMatrixAndResidualVector` and
    `ComputeBatchLoss` run on the B200 kernels.  `physics`: "mechanical" | "thermal" | "neohooke" | "stvenant"."""
    jax, jnp, sparse = _jax()
    phys = np.int32(_lib.PHYSICS[physics])

    class Accelerated(loss_class):
        def Initialize(self, reinitialize=False):
            super().Initialize()
            conn = np.asarray(self.fe_mesh.GetElementsNodes(self.element_type), np.int32)
            nn = self.fe_mesh.GetNumberOfNodes()
            self._b200_elem = np.int32(_lib.ELEMENTS[self.element_type])
            self._b200_conn = jnp.asarray(conn)
            flags = np.zeros(self.total_number_of_dofs, np.uint8)
            flags[np.asarray(self.dirichlet_indices)] = 1
            self._b200_flags = jnp.asarray(flags)
            order = np.argsort(conn.reshape(-1), kind="stable").astype(np.int32)       # entries e*a + local per node
            counts = np.bincount(conn.reshape(-1), minlength=nn)
            self._b200_adj = jnp.asarray(order)
            self._b200_adj_ptr = jnp.asarray(np.concatenate([[0], np.cumsum(counts)]).astype(np.int32))
            self._b200_indices = jax.vmap(self.ComputeElementJacobianIndices)(
                self.fe_mesh.GetElementsNodes(self.element_type)).reshape(-1, 2)       # fe_loss.py:313-314
            params = np.zeros(_lib.NUM_PARAMS)
            md = self.loss_settings.get("material_dict", {})
            params[0], params[1] = md.get("young_modulus", 0.0), md.get("poisson_ratio", 0.0)
            body = np.ravel(np.asarray(getattr(self, "body_force", np.zeros(0)), float))
            params[2:2 + body.size] = body
            tls = getattr(self, "thermal_loss_settings", {})
            params[5], params[6] = tls.get("beta", 0.0), tls.get("c", 1.0)
            self._b200_params = params
            plan = energy_plan.build(np.asarray(self.fe_mesh.GetNodesCoordinates()), conn)
            self._b200_plan = {k: (jnp.asarray(v) if isinstance(v, np.ndarray) else v) for k, v in plan.items()}
            self._b200_geom = None

        # ---- residual + Jacobian (fe_loss.py:264-318)
        def ComputeJacobianMatrixAndResidualVector(self, total_control_vars, total_primal_vars,
                                                   transpose_jacobian: bool = False):
            ne, a = self._b200_conn.shape
            nd = a * self.number_dofs_per_node
            dt = total_primal_vars.dtype
            ke, re = jax.ffi.ffi_call("FolAssembleElements", (jax.ShapeDtypeStruct((ne * nd * nd,), dt),
                                                              jax.ShapeDtypeStruct((ne * nd,), dt)))(
                self.fe_mesh.GetNodesCoordinates().astype(dt), self._b200_conn, total_control_vars.astype(dt),
                total_primal_vars, self._b200_flags, physics=phys, element=self._b200_elem,
                num_gp=np.int32(self.num_gp), transpose=np.int32(bool(transpose_jacobian)), params=self._b200_params)
            R = jax.ffi.ffi_call("FolResidualGather", jax.ShapeDtypeStruct((self.total_number_of_dofs,), dt))(
                self._b200_adj_ptr, self._b200_adj, re, nnode=np.int32(a),
                dofs_per_node=np.int32(self.number_dofs_per_node))
            jac = sparse.BCOO((ke, self._b200_indices), shape=(self.total_number_of_dofs, self.total_number_of_dofs))
            return jac, R

        # ---- batched loss with a custom VJP (fe_loss.py:250-262; SURVEY.md A.7 for the cotangents)
        def _b200_energy_and_grads(self, batch_params, full_dofs):
            p = self._b200_plan
            dt = full_dofs.dtype
            nb, ndof = full_dofs.shape
            nn = batch_params.shape[1]
            ne, a = self._b200_conn.shape
            if self._b200_geom is None or self._b200_geom.dtype != dt:
                width = int(_lib.load().fol_geometry_width(int(phys), int(self._b200_elem)))
                ngauss = ctypes.c_int()
                _lib.load().fol_element_info(int(self._b200_elem), int(self.num_gp), None, None, ctypes.byref(ngauss))
                self._b200_geom = jax.ffi.ffi_call("FolGeometryCache", jax.ShapeDtypeStruct((ne * ngauss.value * width,), dt))(
                    self.fe_mesh.GetNodesCoordinates().astype(dt), self._b200_conn, physics=phys,
                    element=self._b200_elem, num_gp=np.int32(self.num_gp))
            work = int(_lib.load().fol_energy_work_size(int(p["ntiles"]), int(nb)))
            out = (jax.ShapeDtypeStruct((nb, ndof), dt), jax.ShapeDtypeStruct((nb, nn), dt),
                   jax.ShapeDtypeStruct((nb,), dt), jax.ShapeDtypeStruct((work,), dt))
            grad_u, grad_k, energy, _ = jax.ffi.ffi_call("FolEnergyAndGrads", out)(
                self._b200_geom, self._b200_conn, p["adj_ptr"], p["adj_local"], p["tile_node_ptr"], p["tile_nodes"],
                p["tile_elem_ptr"], p["tile_elems"], p["tile_conn"], p["tile_lnode_ptr"], p["tile_lnodes"],
                batch_params, full_dofs, physics=phys, element=self._b200_elem, num_gp=np.int32(self.num_gp),
                ecap=np.int64(p["ecap"]), lcap=np.int64(p["lcap"]), ncap=np.int64(p["ncap"]), params=self._b200_params)
            return energy, grad_u, grad_k

        def ComputeBatchLoss(self, batch_params, batch_dofs):
            exponent = self.loss_settings.get("loss_function_exponent", 1.0)
            free = jnp.ones(self.total_number_of_dofs, batch_dofs.dtype).at[jnp.asarray(self.dirichlet_indices)].set(0.0)

            @jax.custom_vjp
            def batch_energy_loss(params, dofs):
                full = self.GetFullDofVector(params, dofs)                   # fe_loss.py:255 (Dirichlet overwrite)
                energy, _, _ = self._b200_energy_and_grads(params, full)
                return jnp.mean(energy ** exponent), energy

            def fwd(params, dofs):
                full = self.GetFullDofVector(params, dofs)
                energy, grad_u, grad_k = self._b200_energy_and_grads(params, full)
                scale = exponent * energy ** (exponent - 1.0) / energy.shape[0]
                return (jnp.mean(energy ** exponent), energy), (scale, grad_u, grad_k)

            def bwd(res, cot):
                scale, grad_u, grad_k = res
                g = cot[0]                                                   # the per-sample energies carry no cotangent
                w = (g * scale)[:, None]
                return w * grad_k, w * grad_u * free[None, :]                # zero at the Dirichlet dofs

            batch_energy_loss.defvjp(fwd, bwd)
            mean, energy = batch_energy_loss(batch_params.reshape(batch_dofs.shape[0], -1),
                                             batch_dofs.reshape(batch_dofs.shape[0], -1))
            e = jax.lax.stop_gradient(energy ** exponent)
            return mean, (jnp.min(e), jnp.max(e), jnp.mean(e))

    Accelerated.__name__ = loss_class.__name__ + "B200"
    return Accelerated
