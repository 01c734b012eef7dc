"""Oracle (test infrastructure): adjoint sensitivities of ``FiniteElementResponse`` in NumPy.

Restates fol/responses/fe_response.py:
  :59-66    the response formula, a Python expression in (control name, first dof letter)
  :91-124   ComputeResponseElementValue  = sum_g w detJ f(N.de, N_mat u_e)
  :205-213  ComputeValue
  :245-283  ComputeAdjointJacobianMatrixAndRHSVector (rhs = -scatter(d value_e / d u_e), zero at Dirichlet
            dofs; Jacobian = the loss's transposed, BC-applied Jacobian)
  :309-394  ComputeAdjointNodalShapeDerivatives   = scatter(d value_e/d x_e + lam_e^T d re/d x_e), 3 per node
  :420-524  ComputeAdjointNodalControlDerivatives = scatter(d value_e/d K_e + lam_e^T d re/d K_e)
and fol/solvers/adjoint_fe_solver.py:19-24 (the adjoint solve J^T lam = rhs).

The reference differentiates with JAX AD.  Here every derivative is a COMPLEX-STEP derivative
(Im f(x + ih)/h, h = 1e-30: exact to rounding, no subtractive cancellation) of the same NumPy element
functions that are pinned on the reference's goldens -- deliberately a different route from the closed
forms in the CUDA kernels (csrc/adjoint.cuh), so agreement is evidence for both.
Pinned on the reference's own known answers: tests/unit/test_sensitivity_analysis.py:54-80 and
tests/integration/test_mechanical_2D_sa.py:81-113 (tests/test_oracle_responses.py).
"""
import numpy as np

from . import assembly, losses
from .geometry import ELEMENTS, point_data

_H = 1e-30


def response_function(formula, control_name, dof_name):
    """fe_response.py:63-66: ``lambda <control>, <first letter of the first dof>: formula`` with ``jnp`` in
    scope.  Evaluated on arrays: control (...,), dofs (d, ...)."""
    return eval(f"lambda {control_name}, {dof_name[0]}: {formula}", {"jnp": np, "np": np})


def _element_residual(physics, element_type, num_gp, X, de, ue, params):
    if physics == "mechanical":
        return losses.mechanical_element(element_type, num_gp, X, de, ue, params["young_modulus"],
                                         params["poisson_ratio"], params.get("body_force"))[1]
    if physics == "thermal":
        return losses.thermal_element(element_type, num_gp, X, de, ue, params.get("beta", 0.0),
                                      params.get("c", 1.0))[1]
    if physics in ("neohooke", "stvenant"):
        return losses.neo_hooke_element(element_type, num_gp, X, de, ue, params["young_modulus"],
                                        params["poisson_ratio"], params.get("body_force"), law=physics)[1]
    if physics == "transient_thermal":       # (de, ue) = (current, next) temperatures; k0 per element node
        return losses.transient_thermal_element(element_type, num_gp, X, de, ue, params["k0"], params["rho"],
                                                params["cp"], params["time_step"], params.get("beta", 0.0),
                                                params.get("c", 1.0))[1]
    if physics == "allen_cahn":
        return losses.allen_cahn_element(element_type, num_gp, X, de, ue, params["dt"], params["epsilon"])[1]
    raise ValueError(physics)


def element_values(f, element_type, num_gp, d, X, de, ue):
    """fe_response.py:91-124 for all elements: X (ne,a,3), de (ne,a), ue (ne, a*d) -> (ne,)."""
    elem = ELEMENTS[element_type]
    Ns, _, detJ, w = point_data(elem, X, num_gp)
    Kg = np.einsum("ga,ea->eg", Ns, de)
    Ug = np.einsum("ga,eak->keg", Ns, ue.reshape(ue.shape[0], elem.nnode, d))
    fv = f(Kg, Ug) * np.ones_like(Kg)            # a formula may ignore one argument
    return np.einsum("g,eg,eg->e", w, detJ, fv)


def _complex_step(fun, x):
    """Gradient of the per-element scalar fun(x) (ne,) w.r.t. the trailing axis of x (ne, n)."""
    out = np.zeros(x.shape)
    for k in range(x.shape[1]):
        xp = x.astype(complex)
        xp[:, k] += 1j * _H
        out[:, k] = fun(xp).imag / _H
    return out


def element_value_grads(f, element_type, num_gp, d, X, de, ue):
    """(d value_e/d u_e (ne,nd), d value_e/d K_e (ne,a), d value_e/d x_e (ne,a*3)); fe_response.py:126-168."""
    ne, a = de.shape
    dU = _complex_step(lambda v: element_values(f, element_type, num_gp, d, X, de, v), ue)
    dK = _complex_step(lambda v: element_values(f, element_type, num_gp, d, X, v, ue), de)
    dX = _complex_step(lambda v: element_values(f, element_type, num_gp, d, v.reshape(ne, a, 3), de, ue),
                       X.reshape(ne, a * 3))
    return dU, dK, dX


def residual_adjoint_grads(physics, element_type, num_gp, X, de, ue, lam_e, params):
    """(lam_e^T d re/d K_e (ne,a), lam_e^T d re/d x_e (ne,a*3)) with re = ComputeElement(...)[1], the element
    residual BEFORE the Dirichlet mask (fe_response.py:328-330, 440-442)."""
    ne, a = de.shape

    def phi_k(v):
        return np.einsum("en,en->e", lam_e, _element_residual(physics, element_type, num_gp, X, v, ue, params))

    def phi_x(v):
        return np.einsum("en,en->e", lam_e,
                         _element_residual(physics, element_type, num_gp, v.reshape(ne, a, 3), de, ue, params))

    return _complex_step(phi_k, de), _complex_step(phi_x, X.reshape(ne, a * 3))


def _element_params(params, conn):
    """Nodal auxiliary fields (transient thermal: k0) gathered to the elements."""
    if "k0" in params and np.ndim(params["k0"]) == 1:
        params = dict(params)
        params["k0"] = np.asarray(params["k0"], float)[conn]
    return params


def _gather(physics, element_type, coords, conn, controls, dofs):
    d = assembly.dofs_per_node(physics, element_type)
    return d, coords[conn].astype(float), controls[conn].astype(float), dofs[assembly.element_dof_ids(conn, d)]


def compute_value(f, physics, element_type, num_gp, coords, conn, controls, dofs):
    d, X, de, ue = _gather(physics, element_type, coords, conn, controls, dofs)
    return element_values(f, element_type, num_gp, d, X, de, ue).sum()


def adjoint_jacobian_and_rhs(f, physics, element_type, num_gp, coords, conn, controls, dofs,
                             dirichlet_indices, params):
    """fe_response.py:245-283 -> (data, indices, rhs)."""
    d, X, de, ue = _gather(physics, element_type, coords, conn, controls, dofs)
    dU, _, _ = element_value_grads(f, element_type, num_gp, d, X, de, ue)
    rhs = np.zeros(d * coords.shape[0])
    np.add.at(rhs, assembly.element_dof_ids(conn, d).reshape(-1), dU.reshape(-1))
    rhs[dirichlet_indices] = 0.0
    rhs *= -1.0
    data, idx, _ = assembly.assemble(physics, element_type, num_gp, coords, conn, controls, dofs,
                                     dirichlet_indices, params, transpose=True)
    return data, idx, rhs


def adjoint_solve(f, physics, element_type, num_gp, coords, conn, controls, dofs, dirichlet_indices, params):
    """adjoint_fe_solver.py:19-24: the solver is handed (jac, -rhs) and solves jac x = -(-rhs)."""
    data, idx, rhs = adjoint_jacobian_and_rhs(f, physics, element_type, num_gp, coords, conn, controls, dofs,
                                              dirichlet_indices, params)
    ndof = rhs.size
    return np.linalg.solve(assembly.to_dense(data, idx, ndof), rhs)


def control_derivatives(f, physics, element_type, num_gp, coords, conn, controls, dofs, adj_dofs, params):
    """fe_response.py:486-524 (one control per node)."""
    d, X, de, ue = _gather(physics, element_type, coords, conn, controls, dofs)
    lam_e = adj_dofs[assembly.element_dof_ids(conn, d)]
    params = _element_params(params, conn)
    _, dK, _ = element_value_grads(f, element_type, num_gp, d, X, de, ue)
    rK, _ = residual_adjoint_grads(physics, element_type, num_gp, X, de, ue, lam_e, params)
    out = np.zeros(coords.shape[0])
    np.add.at(out, conn.reshape(-1), (dK + rK).reshape(-1))
    return out


def shape_derivatives(f, physics, element_type, num_gp, coords, conn, controls, dofs, adj_dofs, params):
    """fe_response.py:358-394 -> (3*nn,), node-major."""
    d, X, de, ue = _gather(physics, element_type, coords, conn, controls, dofs)
    lam_e = adj_dofs[assembly.element_dof_ids(conn, d)]
    params = _element_params(params, conn)
    _, _, dX = element_value_grads(f, element_type, num_gp, d, X, de, ue)
    _, rX = residual_adjoint_grads(physics, element_type, num_gp, X, de, ue, lam_e, params)
    out = np.zeros(3 * coords.shape[0])
    ids = (3 * conn[:, :, None] + np.arange(3)[None, None, :]).reshape(-1)
    np.add.at(out, ids, (dX + rX).reshape(-1))
    return out
