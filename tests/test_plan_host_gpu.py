"""The host-buffer C-ABI entry point (fol_plan_create / fol_plan_assemble_host): host arrays in, BCOO data + residual in
host arrays out -- the drop-in call a non-Python host binds.  Must equal the loss-class path bit for bit
and the oracle to 1e-12, including element counts that do not divide into the transfer chunks."""
import ctypes

import numpy as np
import pytest

import folax_b200
from folax_b200 import _lib
from oracle import assembly
from tests import gpu_helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("physics,etype,num_gp,n", [("mechanical", "hexahedron", 2, 5), ("mechanical", "hexahedron", 2, 1),
                                                    ("thermal", "quad", 2, 7), ("neohooke", "tetra", 1, 3)])
@pytest.mark.parametrize("transpose", [0, 1])
def test_host_plan_matches_loss_class_and_oracle(physics, etype, num_gp, n, transpose):
    lib = _lib.load()
    mesh = H.make_mesh(etype, n, seed=11)
    loss = H.make_loss(physics, etype, mesh, num_gp)
    K, u = H.fields(physics, mesh, loss, seed=12)
    conn = np.ascontiguousarray(mesh.GetElementsNodes(etype), dtype=np.int32)
    xyz = np.ascontiguousarray(mesh.GetNodesCoordinates(), dtype=np.float64)
    didx = np.ascontiguousarray(loss.dirichlet_indices, dtype=np.int32)
    ne, nn, nd, ndof = conn.shape[0], xyz.shape[0], loss._nd, loss.total_number_of_dofs
    plan = ctypes.c_void_p()
    _lib.check(lib.fol_plan_create(ctypes.byref(plan), _lib.F64, _lib.PHYSICS[physics], loss.fe_element.code, num_gp,
                                   ne, nn, xyz.ctypes.data, conn.ctypes.data, didx.ctypes.data, didx.size, loss._params))
    try:
        Kh, uh = np.ascontiguousarray(K, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64)
        ke, R = np.full(ne * nd * nd, np.nan), np.full(ndof, np.nan)
        for _ in range(2):   # the plan is reusable
            _lib.check(lib.fol_plan_assemble_host(plan, transpose, Kh.ctypes.data, uh.ctypes.data, ke.ctypes.data,
                                                  R.ctypes.data))
    finally:
        lib.fol_plan_destroy(plan)
    jac, Rc = loss.ComputeJacobianMatrixAndResidualVector(K, u, transpose_jacobian=bool(transpose))
    assert np.array_equal(ke, jac.data.cpu().numpy()) and np.array_equal(R, Rc.cpu().numpy())
    data, _, Rref = assembly.assemble(physics, etype, num_gp, xyz, conn, K, u, loss.dirichlet_indices,
                                      H.oracle_params(loss), bool(transpose))
    assert np.abs(ke - data).max() <= 1e-12 * np.abs(data).max()
    assert np.abs(R - Rref).max() <= 1e-11 * np.abs(Rref).max()


@pytest.mark.parametrize("physics,etype,num_gp,n", [("mechanical", "hexahedron", 2, 6), ("thermal", "quad", 2, 9),
                                                    ("neohooke", "tetra", 1, 4)])
def test_host_plan_csr_hand_off(physics, etype, num_gp, n):
    """fol_plan_set_csr / fol_plan_assemble_host_csr: the duplicate-free CSR values the reference's solvers build on the
    host with scipy.sparse.csr_array((data, (rows, cols))) (fe_solver.py:71-72) -- summed on the device, pipelined to
    the host in chunks of whole node rows.  Structure bit-exact against SciPy, values against SciPy's sum of the
    oracle's BCOO and bit-identical to the device-resident JacobianToCSR."""
    import scipy.sparse as sp
    from folax_b200 import csr_plan
    lib = _lib.load()
    mesh = H.make_mesh(etype, n, seed=5)
    loss = H.make_loss(physics, etype, mesh, num_gp)
    K, u = H.fields(physics, mesh, loss, seed=6)
    conn = np.ascontiguousarray(mesh.GetElementsNodes(etype), dtype=np.int32)
    xyz = np.ascontiguousarray(mesh.GetNodesCoordinates(), dtype=np.float64)
    didx = np.ascontiguousarray(loss.dirichlet_indices, dtype=np.int32)
    ne, nn, ndof, d = conn.shape[0], xyz.shape[0], loss.total_number_of_dofs, loss.number_dofs_per_node
    cp = csr_plan.build(conn, nn, d)
    plan = ctypes.c_void_p()
    _lib.check(lib.fol_plan_create(ctypes.byref(plan), _lib.F64, _lib.PHYSICS[physics], loss.fe_element.code, num_gp,
                                   ne, nn, xyz.ctypes.data, conn.ctypes.data, didx.ctypes.data, didx.size, loss._params))
    try:
        Kh, uh = np.ascontiguousarray(K, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64)
        vals, R = np.full(cp["nnz"], np.nan), np.full(ndof, np.nan)
        with pytest.raises(_lib.FolaxError):          # the plan must be uploaded first
            _lib.check(lib.fol_plan_assemble_host_csr(plan, 0, Kh.ctypes.data, uh.ctypes.data, vals.ctypes.data,
                                                      R.ctypes.data))
        _lib.check(lib.fol_plan_set_csr(plan, cp["npairs"], cp["nnz"], cp["pair_ptr"].ctypes.data,
                                        cp["contrib"].ctypes.data, cp["out_base"].ctypes.data,
                                        cp["row_stride"].ctypes.data))
        for _ in range(2):
            _lib.check(lib.fol_plan_assemble_host_csr(plan, 0, Kh.ctypes.data, uh.ctypes.data, vals.ctypes.data,
                                                      R.ctypes.data))
    finally:
        lib.fol_plan_destroy(plan)
    assert not np.isnan(vals).any() and not np.isnan(R).any()
    jac, Rc = loss.ComputeJacobianMatrixAndResidualVector(K, u)
    indptr, indices, dev_vals = loss.JacobianToCSR(jac)
    assert np.array_equal(vals, dev_vals.cpu().numpy()) and np.array_equal(R, Rc.cpu().numpy())
    data, idx, _ = assembly.assemble(physics, etype, num_gp, xyz, conn, K, u, loss.dirichlet_indices, H.oracle_params(loss))
    ref = sp.csr_array((data, (idx[:, 0], idx[:, 1])), shape=(ndof, ndof))
    ref.sum_duplicates()
    ref.sort_indices()
    assert np.array_equal(cp["indptr"], ref.indptr) and np.array_equal(cp["indices"], ref.indices)
    assert np.abs(vals - ref.data).max() <= 1e-12 * np.abs(ref.data).max()
