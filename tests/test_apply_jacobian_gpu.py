"""Matrix-free Jacobian products (ApplyJacobian): y = J v must equal the assembled BCOO of the oracle applied to
v, for both transpose settings (the mask is applied after the transpose, fe_loss.py:191-230) and every physics."""
import numpy as np
import pytest
import scipy.sparse as sp

import folax_b200
from folax_b200 import loss_functions as lf
from oracle import assembly
from tests import gpu_helpers as H

pytestmark = pytest.mark.gpu

CASES = [("mechanical", "hexahedron", 2), ("mechanical", "quad", 2), ("mechanical", "tetra", 1),
         ("mechanical", "triangle", 1), ("thermal", "quad", 2), ("thermal", "hexahedron", 2), ("thermal", "tetra", 1),
         ("neohooke", "tetra", 1), ("neohooke", "hexahedron", 2), ("neohooke", "quad", 2), ("stvenant", "quad", 2)]


def _dense_apply(data, idx, n, v):
    J = sp.coo_array((data, (idx[:, 0], idx[:, 1])), shape=(n, n)).tocsr()
    return J @ v


@pytest.mark.parametrize("physics,etype,num_gp", CASES)
@pytest.mark.parametrize("transpose", [False, True])
def test_apply_jacobian_matches_assembled_matrix(physics, etype, num_gp, transpose):
    mesh = H.make_mesh(etype, 3 if etype in ("hexahedron", "tetra") else 6, seed=3)
    extra = {"beta": 2.0, "c": 3} if physics == "thermal" else {}
    loss = H.make_loss(physics, etype, mesh, num_gp, "float64", extra)
    K, u = H.fields(physics, mesh, loss, seed=4)
    rng = np.random.default_rng(9)
    v = rng.standard_normal(loss.total_number_of_dofs)
    y = loss.ApplyJacobian(K, u, v, transpose_jacobian=transpose).cpu().numpy()
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes(etype)
    data, idx, _ = assembly.assemble(physics, etype, num_gp, coords, conn, K, u, loss.dirichlet_indices,
                                     H.oracle_params(loss), transpose)
    ref = _dense_apply(data, idx, loss.total_number_of_dofs, v)
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    # and against this library's own assembled Jacobian
    jac, _ = loss.ComputeJacobianMatrixAndResidualVector(K, u, transpose_jacobian=transpose)
    own = _dense_apply(jac.data.cpu().numpy(), jac.indices.cpu().numpy(), loss.total_number_of_dofs, v)
    assert np.abs(y - own).max() <= 1e-12 * np.abs(own).max()


def test_apply_jacobian_is_linear_and_deterministic():
    mesh = H.make_mesh("hexahedron", 6, seed=1)
    loss = H.make_loss("mechanical", "hexahedron", mesh, 2)
    K, u = H.fields("mechanical", mesh, loss, seed=2)
    rng = np.random.default_rng(0)
    v, w = rng.standard_normal((2, loss.total_number_of_dofs))
    a = loss.ApplyJacobian(K, u, v)
    b = loss.ApplyJacobian(K, u, w)
    c = loss.ApplyJacobian(K, u, 2.0 * v - 3.0 * w)
    assert (c - (2.0 * a - 3.0 * b)).abs().max().item() <= 1e-12 * c.abs().max().item()
    assert bool((loss.ApplyJacobian(K, u, v) == a).all())          # run-to-run bit-identical
    # Dirichlet rows: identity-like (diagonal kept, fe_loss.py:203-207)
    jac, _ = loss.ComputeJacobianMatrixAndResidualVector(K, u)
    d = loss.dirichlet_indices
    dense_diag = np.zeros(loss.total_number_of_dofs)
    idx, data = jac.indices.cpu().numpy(), jac.data.cpu().numpy()
    on = idx[:, 0] == idx[:, 1]
    np.add.at(dense_diag, idx[on, 0], data[on])
    assert np.abs(a.cpu().numpy()[d] - dense_diag[d] * v[d]).max() <= 1e-12 * np.abs(dense_diag[d] * v[d]).max()


def test_apply_jacobian_float32_and_transient():
    mesh = H.make_mesh("quad", 8, seed=2)
    loss = H.make_loss("mechanical", "quad", mesh, 2, "float32")
    K, u = H.fields("mechanical", mesh, loss, seed=3)
    v = np.random.default_rng(1).standard_normal(loss.total_number_of_dofs)
    y = loss.ApplyJacobian(K, u, v).cpu().numpy()
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("quad")
    data, idx, _ = assembly.assemble("mechanical", "quad", 2, coords, conn, K, u, loss.dirichlet_indices,
                                     H.oracle_params(loss))
    ref = _dense_apply(data, idx, loss.total_number_of_dofs, v)
    assert np.abs(y - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.parametrize("physics,etype,num_gp", [("mechanical", "hexahedron", 2), ("thermal", "quad", 2),
                                                  ("neohooke", "tetra", 1)])
@pytest.mark.parametrize("transpose", [0, 1])
def test_batched_products_equal_per_sample_products(physics, etype, num_gp, transpose):
    """fol_apply_jacobian_elements_batched + fol_residual_gather_batched (grid.y = sample) against one call per sample:
    bit for bit (same kernel, same fixed summation order)."""
    import torch
    from folax_b200 import _lib
    lib = _lib.load()
    mesh = H.make_mesh(etype, 3, seed=9)
    loss = H.make_loss(physics, etype, mesh, num_gp)
    nb = 4
    K, u = H.fields(physics, mesh, loss, seed=2, batch=nb)
    rng = np.random.default_rng(3)
    V = rng.standard_normal(u.shape)
    Kt, ut, vt = (torch.tensor(a, device="cuda") for a in (K, u, V))
    ye = torch.empty((nb, loss._ne * loss._nd), dtype=torch.float64, device="cuda")
    y = torch.empty((nb, loss.total_number_of_dofs), dtype=torch.float64, device="cuda")
    s = _lib.stream_ptr()
    _lib.check(lib.fol_apply_jacobian_elements_batched(
        s, loss._dt, _lib.PHYSICS[physics], loss.fe_element.code, num_gp, transpose, loss._ne, loss._nn, nb,
        _lib.ptr(loss._xyz), _lib.ptr(loss._conn), _lib.ptr(Kt), _lib.ptr(ut), _lib.ptr(loss._dir_flag), loss._params,
        _lib.ptr(vt), _lib.ptr(ye)))
    _lib.check(lib.fol_residual_gather_batched(s, loss._dt, loss._nn, loss._nnode, loss.number_dofs_per_node, nb, loss._ne,
                                               _lib.ptr(loss._adj_ptr), _lib.ptr(loss._adj), _lib.ptr(ye), _lib.ptr(y)))
    for b in range(nb):
        ref = loss.ApplyJacobian(K[b], u[b], V[b], transpose_jacobian=bool(transpose))
        assert torch.equal(y[b], ref), f"sample {b}"
