"""GPU checks of the element-level helpers of the reference API (fe_loss.py:149-230): ComputeElementEnergy,
ComputeElementsEnergies, ComputeElementJacobianIndices, ApplyDirichletBCOnElementResidualAndJacobian,
ComputeElementResidualAndJacobian."""
import numpy as np
import pytest
import torch

from oracle import assembly
from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("physics,etype,num_gp", [("mechanical", "hexahedron", 2), ("mechanical", "triangle", 1),
                                                  ("thermal", "quad", 2), ("thermal", "tetra", 1),
                                                  ("neohooke", "tetra", 1), ("neohooke", "quad", 2),
                                                  ("stvenant", "hexahedron", 2)])
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-12), ("float32", 1e-4)])
def test_elements_energies(physics, etype, num_gp, dtype, tol):
    mesh = gh.make_mesh(etype, 4, seed=6)
    extra = {"beta": 2.0, "c": 4.0} if physics == "thermal" else {"body_foce": [0.2, -0.4, 0.7][:3 if etype in ("hexahedron", "tetra") else 2]}
    loss = gh.make_loss(physics, etype, mesh, num_gp=num_gp, dtype=dtype, extra=extra)
    K, u = gh.fields(physics, mesh, loss, seed=2)
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes(etype)
    ref = assembly.compute_elements(physics, etype, num_gp, coords, conn, K, u, gh.oracle_params(loss))[0]
    en = loss.ComputeElementsEnergies(K, u)
    assert en.shape == (len(conn),)
    assert np.abs(en.cpu().numpy() - ref).max() <= tol * np.abs(ref).max()
    total = float(loss.ComputeTotalEnergy(K, u))
    assert abs(total - ref.sum()) <= 50 * tol * max(abs(ref.sum()), np.abs(ref).max())
    # one element through the single-element entry points
    e = 3
    g = assembly.element_dof_ids(conn, loss.number_dofs_per_node)[e]
    en_e = float(loss.ComputeElementEnergy(coords[conn[e]], K[conn[e]], u[g]))
    assert abs(en_e - ref[e]) <= 10 * tol * np.abs(ref).max()


def test_element_indices_and_dirichlet_helpers():
    mesh = gh.make_mesh("quad", 4, seed=1)
    loss = gh.make_loss("mechanical", "quad", mesh, num_gp=2)
    conn = mesh.GetElementsNodes("quad")
    coords = np.asarray(mesh.GetNodesCoordinates())
    idx = loss.ComputeElementJacobianIndices(conn[5].copy())
    assert np.array_equal(idx.cpu().numpy(), assembly.bcoo_indices(conn[5:6], 2))
    rng = np.random.default_rng(0)
    K, u = gh.fields("mechanical", mesh, loss, seed=3)
    g = assembly.element_dof_ids(conn, 2)
    e = 0                                                     # touches the `left` Dirichlet boundary
    bc = np.ones(loss.total_number_of_dofs)
    bc[loss.dirichlet_indices] = 0.0
    bc_e, mask_e = bc[g[e]], 1.0 - bc[g[e]]
    assert mask_e.any() and bc_e.any()
    _, re, ke = assembly.compute_elements("mechanical", "quad", 2, coords, conn[e:e + 1], K, u, gh.oracle_params(loss))
    for transpose in (False, True):
        re_ref, ke_ref = assembly.apply_dirichlet(re, ke, bc_e[None], transpose)
        re_g, ke_g = loss.ComputeElementResidualAndJacobian(coords[conn[e]], K[conn[e]], u[g[e]], bc_e, mask_e, transpose)
        assert np.abs(ke_g.cpu().numpy() - ke_ref[0]).max() <= 1e-12 * np.abs(ke_ref).max()
        assert np.abs(re_g.cpu().numpy().reshape(-1) - re_ref[0]).max() <= 1e-12 * np.abs(re_ref).max()
    # the standalone masking on arbitrary arrays
    A, r = rng.standard_normal((8, 8)), rng.standard_normal(8)
    r2, A2 = loss.ApplyDirichletBCOnElementResidualAndJacobian(r, A, bc_e, mask_e)
    r_ref, A_ref = assembly.apply_dirichlet(r[None], A[None], bc_e[None])
    assert np.allclose(A2.cpu().numpy(), A_ref[0], rtol=0, atol=1e-15) and np.allclose(r2.cpu().numpy().reshape(-1), r_ref[0])
