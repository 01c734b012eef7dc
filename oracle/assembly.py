"""Oracle (test infrastructure): global assembly, Dirichlet handling and the batched loss.

Restates, in NumPy float64 / integer arithmetic:
  fol/loss_functions/fe_loss.py:34-55    Dirichlet index/value vectors (dof-major, then set)
  fol/loss_functions/fe_loss.py:178-184  per-element BCOO index pairs
  fol/loss_functions/fe_loss.py:191-230  transpose-then-row-mask Dirichlet application
  fol/loss_functions/fe_loss.py:264-318  residual scatter-add + duplicate-keeping BCOO
  fol/loss_functions/fe_loss.py:250-262  batched energy loss (mean_b E_b^p, (min,max,mean))
"""
import numpy as np

from .geometry import ELEMENTS
from . import losses

DOFS_PER_NODE = {"mechanical": None, "thermal": 1}


def dirichlet_vectors(ordered_dofs, dirichlet_bc_dict, node_sets):
    """fe_loss.py:34-55.  Returns (dirichlet_indices, dirichlet_values, non_dirichlet_indices(ndof))
    as a function of ndof via closure-free tuple: caller supplies ndof to ``non_dirichlet``."""
    d = len(ordered_dofs)
    idx, val = [], []
    for k, dof in enumerate(ordered_dofs):
        for name, value in dirichlet_bc_dict.get(dof, {}).items():
            nodes = np.asarray(node_sets[name], dtype=np.int64)
            idx.append(d * nodes + k)
            val.append(np.full(nodes.shape, float(value)))
    if idx:
        return np.concatenate(idx), np.concatenate(val)
    return np.zeros(0, np.int64), np.zeros(0)


def non_dirichlet(ndof, dirichlet_indices):
    return np.setdiff1d(np.arange(ndof), dirichlet_indices)


def element_dof_ids(conn, d):
    """gdof(e, i) = d*conn[e, i//d] + i%d   (fe_loss.py:163-164, 178-181)."""
    return (d * conn[:, :, None] + np.arange(d)[None, None, :]).reshape(conn.shape[0], -1)


def bcoo_indices(conn, d):
    """fe_loss.py:178-184, 313-314: (ne*nd*nd, 2), dtype of the connectivity."""
    g = element_dof_ids(conn, d)
    nd = g.shape[1]
    rows = np.repeat(g[:, :, None], nd, axis=2)
    cols = np.repeat(g[:, None, :], nd, axis=1)
    return np.stack([rows.reshape(-1), cols.reshape(-1)], axis=1).astype(conn.dtype)


def apply_dirichlet(re, Ke, bc, transpose=False):
    """fe_loss.py:191-230: optional transpose first; rows of fixed dofs zeroed, their diagonal
    kept, columns untouched.  bc (ne, nd) is 1 on free dofs, 0 on Dirichlet dofs."""
    if transpose:
        Ke = np.swapaxes(Ke, 1, 2)
    diag = np.einsum("eii->ei", Ke)
    out = bc[:, :, None] * Ke
    idx = np.arange(Ke.shape[1])
    out[:, idx, idx] += (1.0 - bc) * diag
    return bc * re, out


def compute_elements(physics, element_type, num_gp, coords, conn, controls, dofs, params):
    """Gather (fe_loss.py:155-164) + ComputeElement.  Returns (energy, re, Ke) per element."""
    elem = ELEMENTS[element_type]
    X = coords[conn]
    de = controls[conn]
    if physics == "thermal":
        return losses.thermal_element(element_type, num_gp, X, de, dofs[conn],
                                      params.get("beta", 0.0), params.get("c", 1.0))
    if physics == "transient_thermal":      # (controls, dofs) = (current, next) temperatures
        return losses.transient_thermal_element(element_type, num_gp, X, de, dofs[conn], params["k0"][conn],
                                                params["rho"], params["cp"], params["time_step"],
                                                params.get("beta", 0.0), params.get("c", 1.0))
    if physics == "allen_cahn":             # (controls, dofs) = (current, next) phase field
        return losses.allen_cahn_element(element_type, num_gp, X, de, dofs[conn], params["dt"], params["epsilon"])
    d = elem.dim
    u = dofs[element_dof_ids(conn, d)]
    if physics == "mechanical":
        return losses.mechanical_element(element_type, num_gp, X, de, u, params["young_modulus"],
                                         params["poisson_ratio"], params.get("body_force"))
    if physics in ("neohooke", "stvenant"):
        return losses.neo_hooke_element(element_type, num_gp, X, de, u, params["young_modulus"],
                                        params["poisson_ratio"], params.get("body_force"), law=physics)
    raise ValueError(physics)


def dofs_per_node(physics, element_type):
    return 1 if physics in ("thermal", "transient_thermal", "allen_cahn") else ELEMENTS[element_type].dim


def assemble(physics, element_type, num_gp, coords, conn, controls, dofs, dirichlet_indices,
             params, transpose=False, chunk=65536):
    """ComputeJacobianMatrixAndResidualVector (fe_loss.py:264-318).
    Returns (data (ne*nd*nd,), indices (ne*nd*nd, 2), residual (ndof,))."""
    d = dofs_per_node(physics, element_type)
    ndof = d * coords.shape[0]
    bc_vec = np.ones(ndof)
    bc_vec[dirichlet_indices] = 0.0
    ne = conn.shape[0]
    nd = conn.shape[1] * d
    data = np.empty((ne, nd, nd))
    R = np.zeros(ndof)
    for s in range(0, ne, chunk):
        c = conn[s:s + chunk]
        _, re, Ke = compute_elements(physics, element_type, num_gp, coords, c, controls, dofs, params)
        g = element_dof_ids(c, d)
        re, Ke = apply_dirichlet(re, Ke, bc_vec[g], transpose)
        data[s:s + chunk] = Ke
        np.add.at(R, g.reshape(-1), re.reshape(-1))
    return data.reshape(-1), bcoo_indices(conn, d), R


def to_dense(data, indices, ndof):
    A = np.zeros((ndof, ndof))
    np.add.at(A, (indices[:, 0], indices[:, 1]), data)
    return A


def full_dof_vector(batch_dofs, dirichlet_indices, dirichlet_values):
    """fe_loss.py:91-92: overwrite Dirichlet entries of every sample."""
    out = np.array(batch_dofs, dtype=float, copy=True)
    out[:, dirichlet_indices] = dirichlet_values
    return out


def batch_loss(physics, element_type, num_gp, coords, conn, batch_controls, batch_dofs,
               dirichlet_indices, dirichlet_values, params, exponent=1.0):
    """ComputeBatchLoss (fe_loss.py:250-262): returns (mean, (min, max, mean), E_b)."""
    U = full_dof_vector(np.atleast_2d(batch_dofs), dirichlet_indices, dirichlet_values)
    K = np.atleast_2d(batch_controls)
    E = np.array([compute_elements(physics, element_type, num_gp, coords, conn, K[b], U[b], params)[0].sum()
                  for b in range(U.shape[0])]) ** exponent
    return E.mean(), (E.min(), E.max(), E.mean()), E


def batch_loss_grads(physics, element_type, num_gp, coords, conn, batch_controls, batch_dofs,
                     dirichlet_indices, dirichlet_values, params, exponent=1.0):
    """Analytic cotangents of ``batch_loss`` w.r.t. batch_dofs and batch_controls, following
    the reference's stop_gradient placement (mechanical.py:116, thermal.py:31, 45-49):
      d mean / d u_b = (p E_b^(p-1) / B) * R_unmasked(u_b)   zeroed at Dirichlet dofs,
      d mean / d K_b = 0 (mechanical) | sum_g N_a (1+beta T^c)|grad T|^2 detJ w (thermal)
                       | d(sum psi)/dK (neo-hooke)."""
    U = full_dof_vector(np.atleast_2d(batch_dofs), dirichlet_indices, dirichlet_values)
    K = np.atleast_2d(batch_controls)
    nb = U.shape[0]
    d = dofs_per_node(physics, element_type)
    g = element_dof_ids(conn, d)
    gU, gK, Eb = np.zeros_like(U), np.zeros_like(K), np.zeros(nb)
    X = coords[conn]
    for b in range(nb):
        en, re, _ = compute_elements(physics, element_type, num_gp, coords, conn, K[b], U[b], params)
        Eb[b] = en.sum()
        if physics in ("neohooke", "stvenant"):
            # energy = sum psi; its u-gradient is F_int (no body-force term)
            body = params.get("body_force")
            if body is not None:
                elem = ELEMENTS[element_type]
                from .geometry import point_data
                Ns, _, detJ, w = point_data(elem, X, num_gp)
                re = re + losses.body_force_vector(elem, Ns, detJ, w, body)
            dK = losses.neo_hooke_energy_dcontrol(element_type, num_gp, X, K[b][conn],
                                                  U[b][g], params["poisson_ratio"], law=physics)
            np.add.at(gK[b], conn.reshape(-1), dK.reshape(-1))
        elif physics in ("transient_thermal", "allen_cahn"):
            # true potentials: both gradients come from the energy itself, not from the element residual
            p_el = dict(params)
            if "k0" in p_el:
                p_el["k0"] = np.asarray(params["k0"])[conn]
            re, dK = losses.implicit_scalar_energy_grads(physics, element_type, num_gp, X, K[b][conn], U[b][conn], p_el)
            np.add.at(gK[b], conn.reshape(-1), dK.reshape(-1))
        elif physics == "thermal":
            _, dK = losses.thermal_energy_grads(element_type, num_gp, X, K[b][conn], U[b][conn],
                                                params.get("beta", 0.0), params.get("c", 1.0))
            np.add.at(gK[b], conn.reshape(-1), dK.reshape(-1))
        np.add.at(gU[b], g.reshape(-1), re.reshape(-1))
    scale = exponent * Eb ** (exponent - 1.0) / nb
    gU *= scale[:, None]
    gK *= scale[:, None]
    gU[:, dirichlet_indices] = 0.0
    return gU, gK


def assemble_j2(element_type, num_gp, coords, conn, dofs, state, dirichlet_indices, params, transpose=False):
    """ElastoplasticityLoss.ComputeJacobianMatrixAndResidualVector
    (mechanical_elastoplasticity.py:153-235) -> (new_state, data, indices, residual)."""
    from . import j2
    elem = ELEMENTS[element_type]
    d = elem.dim
    ndof = d * coords.shape[0]
    bc_vec = np.ones(ndof)
    bc_vec[dirichlet_indices] = 0.0
    g = element_dof_ids(conn, d)
    _, new_state, re, Ke = j2.j2_element(element_type, num_gp, coords[conn], dofs[g], state,
                                         params["young_modulus"], params["poisson_ratio"], params["yield_limit"],
                                         params["iso_hardening_parameter_1"], params["iso_hardening_param_2"],
                                         params.get("body_force"))
    re, Ke = apply_dirichlet(re, Ke, bc_vec[g], transpose)
    R = np.zeros(ndof)
    np.add.at(R, g.reshape(-1), re.reshape(-1))
    return new_state, Ke.reshape(-1), bcoo_indices(conn, d), R
