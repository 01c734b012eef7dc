"""Oracle (test infrastructure): SECOND, independent oracle of the batched physics loss and its gradient -- a literal
transcription of the reference's element code into torch float64 on the CPU, differentiated by torch.autograd (reverse
mode, with `.detach()` exactly where the reference writes `jax.lax.stop_gradient`).  It shares nothing with the closed-form
cotangents of oracle/assembly.py::batch_loss_grads or with the CUDA kernels.

Follows, statement by statement,
  fol/loss_functions/fe_loss.py:91-92, 166-176, 250-262   (Dirichlet overwrite, element energies, batch loss)
  fol/loss_functions/mechanical.py:37-58, 60-96, 98-117   (B, D, N matrices; ComputeElement, energy = u^T sg(Se u - Fe))
  fol/loss_functions/thermal.py:28-49                     (ComputeElement, T stop-gradiented inside kappa and re)
  fol/geometries/geometry.py:88-97                        (J = (dN^T X)^T, grad N = dN J^-1 via linalg.inv, det)
  fol/geometries/{quadrilateral_2d_4,hexahedra_3d_8,tetrahedra_3d_4}.py  (shape functions and Gauss rules: taken
                                                                          from oracle/geometry.py, which is pinned on
                                                                          tests/unit/test_geometries.py)
Parity status: the reference holds no golden for ComputeBatchLoss, for its gradient or for the static thermal element
(SURVEY.md 8c).  These rows are pinned by the agreement of the two oracles at <= 1e-12
(tests/test_oracle_batch_loss_torch.py); the kernels are then compared with oracle/assembly.py.
"""
import numpy as np
import torch

from .geometry import ELEMENTS


def _geometry(elem, X, xi):
    """geometry.py:88-97 for one element (X: (a, 3) tensor) at one point -> (N (a,), DN_DX (a, dim), detJ)."""
    N, dN = torch.as_tensor(elem.N(xi)), torch.as_tensor(elem.dN(xi))      # (a,), (a, dim)
    jac = (dN.T @ X[:, :elem.dim]).T                                        # Jacobian: jnp.dot(dN_dxi.T, points).T
    return N, dN @ torch.linalg.inv(jac), torch.linalg.det(jac)


def _b_matrix(DN_DX):
    """mechanical.py:37-58."""
    a, dim = DN_DX.shape
    if dim == 2:
        B = torch.zeros((3, 2 * a), dtype=torch.float64)
        idx = torch.arange(a)
        B[0, 2 * idx] = DN_DX[:, 0]
        B[1, 2 * idx + 1] = DN_DX[:, 1]
        B[2, 2 * idx] = DN_DX[:, 1]
        B[2, 2 * idx + 1] = DN_DX[:, 0]
        return B
    B = torch.zeros((6, 3 * a), dtype=torch.float64)
    idx = torch.arange(a) * 3
    B[0, idx] = DN_DX[:, 0]
    B[1, idx + 1] = DN_DX[:, 1]
    B[2, idx + 2] = DN_DX[:, 2]
    B[3, idx] = DN_DX[:, 1]
    B[3, idx + 1] = DN_DX[:, 0]
    B[4, idx + 1] = DN_DX[:, 2]
    B[4, idx + 2] = DN_DX[:, 1]
    B[5, idx] = DN_DX[:, 2]
    B[5, idx + 2] = DN_DX[:, 0]
    return B


def _d_matrix(E, nu, dim):
    """mechanical.py:60-82 (2-D: plane stress)."""
    if dim == 2:
        return torch.tensor([[1, nu, 0], [nu, 1, 0], [0, 0, (1 - nu) / 2]], dtype=torch.float64) * (E / (1 - nu ** 2))
    c1 = E / ((1.0 + nu) * (1.0 - 2.0 * nu))
    c2, c3, c4 = c1 * (1.0 - nu), c1 * nu, c1 * 0.5 * (1.0 - 2.0 * nu)
    D = torch.zeros((6, 6), dtype=torch.float64)
    D[:3, :3] = c3
    for i in range(3):
        D[i, i] = c2
        D[3 + i, 3 + i] = c4
    return D


def mechanical_element(elem, num_gp, X, de, uvwe, E, nu, body):
    """mechanical.py:98-117 -> energy (scalar tensor)."""
    pts, wts = elem.gauss(num_gp)
    D = _d_matrix(E, nu, elem.dim)
    nd = elem.nnode * elem.dim
    Se, Fe = torch.zeros((nd, nd), dtype=torch.float64), torch.zeros((nd, 1), dtype=torch.float64)
    body = torch.as_tensor(np.asarray(body, dtype=np.float64).reshape(-1, 1))
    for xi, w in zip(pts, wts):
        N, DN_DX, detJ = _geometry(elem, X, xi)
        N_mat = torch.zeros((elem.dim, nd), dtype=torch.float64)
        for k in range(elem.dim):
            N_mat[k, k::elem.dim] = N
        e_at_gauss = torch.dot(N, de)
        B = _b_matrix(DN_DX)
        Se = Se + w * detJ * e_at_gauss * (B.T @ (D @ B))
        Fe = Fe + w * detJ * (N_mat.T @ body)
    residual = (Se @ uvwe - Fe).detach()                                     # jax.lax.stop_gradient
    return (uvwe.T @ residual)[0, 0]


def thermal_element(elem, num_gp, X, de, te, beta, c):
    """thermal.py:28-49 -> energy (scalar tensor); the generic path passes no body force."""
    pts, wts = elem.gauss(num_gp)
    a = elem.nnode
    Se = torch.zeros((a, a), dtype=torch.float64)
    te_sg = te.detach()                                                      # :31
    for xi, w in zip(pts, wts):
        N, DN_DX, detJ = _geometry(elem, X, xi)
        conductivity = torch.dot(N, de) * (1 + beta * torch.dot(N, te_sg.reshape(-1)) ** c)
        Se = Se + conductivity * (DN_DX @ DN_DX.T) * detJ * w
    residual = Se @ te.detach()                                              # :44-46, Fe = 0
    return (te.T @ residual)[0, 0]


def batch_loss_and_grads(physics, element_type, num_gp, coords, conn, batch_controls, batch_dofs, dirichlet_indices,
                         dirichlet_values, params, exponent=1.0):
    """fe_loss.py:250-262 and its reverse-mode gradient -> (mean, E_b, d mean / d dofs, d mean / d controls)."""
    elem = ELEMENTS[element_type]
    d = 1 if physics == "thermal" else elem.dim
    X = torch.as_tensor(np.asarray(coords, dtype=np.float64))
    K = torch.tensor(np.atleast_2d(batch_controls), dtype=torch.float64, requires_grad=True)
    U = torch.tensor(np.atleast_2d(batch_dofs), dtype=torch.float64, requires_grad=True)
    didx = torch.as_tensor(np.asarray(dirichlet_indices, dtype=np.int64))
    dval = torch.as_tensor(np.asarray(dirichlet_values, dtype=np.float64))
    full = U.clone()
    full[:, didx] = dval                                                     # fe_loss.py:91-92 (.at[].set())
    energies = []
    for b in range(U.shape[0]):
        total = torch.zeros((), dtype=torch.float64)
        for nodes in np.asarray(conn):
            n = torch.as_tensor(np.asarray(nodes, dtype=np.int64))
            dofs = (d * n[:, None] + torch.arange(d)[None, :]).reshape(-1)   # fe_loss.py:163-164
            if physics == "thermal":
                total = total + thermal_element(elem, num_gp, X[n], K[b, n], full[b, dofs].reshape(-1, 1),
                                                params.get("beta", 0.0), params.get("c", 1))
            else:
                total = total + mechanical_element(elem, num_gp, X[n], K[b, n], full[b, dofs].reshape(-1, 1),
                                                   params["young_modulus"], params["poisson_ratio"],
                                                   params.get("body_force", np.zeros(elem.dim)))
        energies.append(total ** exponent)
    E = torch.stack(energies)
    mean = E.mean()
    gU, gK = torch.autograd.grad(mean, (U, K), allow_unused=True)
    zero = lambda t, like: torch.zeros_like(like) if t is None else t
    return mean.item(), E.detach().numpy(), zero(gU, U).numpy(), zero(gK, K).numpy()
