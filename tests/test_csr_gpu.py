"""Duplicate-free CSR hand-off (fol_csr_values) against SciPy's COO->CSR duplicate sum, the
conversion the reference's solvers do on the host (fe_solver.py:71-72)."""
import numpy as np
import pytest
import scipy.sparse as sp

from tests import gpu_helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("physics,etype,n", [("mechanical", "hexahedron", 5), ("thermal", "quad", 9),
                                             ("mechanical", "tetra", 3), ("neohooke", "triangle", 6)])
def test_csr_matches_scipy_duplicate_sum(physics, etype, n):
    mesh = H.make_mesh(etype, n, seed=2)
    loss = H.make_loss(physics, etype, mesh, None)
    K, u = H.fields(physics, mesh, loss, seed=5)
    jac, _ = loss.ComputeJacobianMatrixAndResidualVector(K, u)
    indptr, indices, vals = loss.JacobianToCSR(jac)
    idx = jac.indices.cpu().numpy()
    ref = sp.csr_array((jac.data.cpu().numpy(), (idx[:, 0], idx[:, 1])), shape=jac.shape)
    ref.sum_duplicates()
    ref.sort_indices()
    assert np.array_equal(indptr.cpu().numpy(), ref.indptr), "CSR row pointers must be bit-exact"
    assert np.array_equal(indices.cpu().numpy(), ref.indices), "CSR column indices must be bit-exact"
    v = vals.cpu().numpy()
    assert np.abs(v - ref.data).max() <= 1e-13 * np.abs(ref.data).max()
    # deterministic: same bits on a second call
    _, _, vals2 = loss.JacobianToCSR(jac)
    assert np.array_equal(v, vals2.cpu().numpy())
    # the CSR reproduces J @ x
    x = np.random.default_rng(0).standard_normal(jac.shape[0])
    A = sp.csr_array((v, indices.cpu().numpy(), indptr.cpu().numpy()), shape=jac.shape)
    assert np.abs(A @ x - (jac @ x).cpu().numpy()).max() <= 1e-12 * np.abs(A @ x).max()
