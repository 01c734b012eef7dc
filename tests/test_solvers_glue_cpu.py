"""CPU check of folax_b200.solvers / folax_b200.linalg (the API of fol/solvers on the device-resident Jacobian).

No GPU here: the C ABI is the stand-in of tests/cpu_backend.py -- the SpMV, gather and vector kernels run as their
own per-thread code compiled for the CPU (tests/host_shim/krylov_host.cu), the loss is backed by the oracle.  What
is exercised for real: the sliced-ELLPACK plan, BiCGSTAB, the Newton / load-step logic, the adjoint solve, against
SciPy and against the reference's integration golden (tests/integration/test_mechanical_2D_sa.py:81-113).
GPU runs of the same classes: tests/test_zz2_solvers_gpu.py."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

import folax_b200
from folax_b200 import linalg, sell_plan
from folax_b200.responses import FiniteElementResponse, NodalControl
from folax_b200.solvers import (AdjointFiniteElementSolver, FiniteElementLinearResidualBasedSolver,
                                FiniteElementNonLinearResidualBasedSolver)
from oracle import assembly
from tests.cpu_backend import fake_loss
from tests.test_oracle_golden import _square_mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BC = {"Ux": {"left": 0.0, "right": 0.05}, "Uy": {"left": 0.0, "right": 0.05}}
MAT = {"young_modulus": 1.0, "poisson_ratio": 0.3}


def _mech_loss(N):
    coords, conn, sets = _square_mesh(N)
    return fake_loss("mechanical", "quad", 2, coords, conn, sets, ["Ux", "Uy"], BC, MAT, [1.0, 0.3] + [0.0] * 10)


def test_sell_plan_and_spmv_on_a_fe_matrix(shim):
    """Sliced-ELLPACK product (the kernel's per-thread code) == SciPy CSR product, on a 3-D hex elasticity matrix
    whose row lengths vary (corner / edge / face / interior dofs) and whose row count is not a multiple of 32."""
    mesh = folax_b200.create_3D_box_mesh(3, 4, 2, 1.0, 1.0, 1.0)
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("hexahedron")
    rng = np.random.default_rng(0)
    nn = len(coords)
    didx, _ = assembly.dirichlet_vectors(["Ux", "Uy", "Uz"], {d: {"left": 0.0, "right": 0.1} for d in ("Ux", "Uy", "Uz")},
                                         mesh.node_sets)
    data, idx, _ = assembly.assemble("mechanical", "hexahedron", 2, coords, conn, rng.uniform(0.1, 1, nn),
                                     np.zeros(3 * nn), didx, MAT)
    A = sp.csr_array((data, (idx[:, 0], idx[:, 1])), shape=(3 * nn, 3 * nn))
    A.sum_duplicates()
    A.sort_indices()
    assert A.shape[0] % 32 != 0
    plan = sell_plan.build(A.indptr, A.indices)
    assert plan["total"] >= A.nnz and plan["total"] <= 1.5 * A.nnz          # little padding on FE matrices
    vals = np.zeros(plan["total"])
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert shim.host_gather_values(C.c_longlong(plan["total"]), p(plan["src"]), p(A.data), p(vals)) == 0
    x, y = rng.standard_normal(3 * nn), np.full(3 * nn, np.nan)
    assert shim.host_sell_spmv(C.c_longlong(plan["nrows"]), p(plan["slice_ptr"]), p(plan["cols"]), p(vals), p(x), p(y)) == 0
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-13 * np.abs(ref).max()
    diag = np.zeros(3 * nn)
    assert shim.host_gather_values(C.c_longlong(3 * nn), p(plan["diag_src"]), p(A.data), p(diag)) == 0
    assert np.array_equal(diag, A.diagonal())
    # block variant: one node index per run of 3 dofs, same values array, bit-identical product
    bplan = sell_plan.build(A.indptr, A.indices, 3)
    assert bplan["node_cols"] is not None and bplan["node_cols"].size * 3 == bplan["total"]
    yb = np.full(3 * nn, np.nan)
    assert shim.host_sell_spmv_block(3, C.c_longlong(bplan["nrows"]), p(bplan["slice_ptr"]), p(bplan["node_cols"]), p(vals),
                                     p(x), p(yb)) == 0
    assert np.array_equal(yb, y)
    # a structure that does not consist of whole runs is refused by the plan (falls back to the scalar kernel)
    B = sp.csr_array(np.triu(np.ones((6, 6))))
    assert sell_plan.build(B.indptr, B.indices, 3)["node_cols"] is None


def test_block_kernel_is_used_and_equals_the_scalar_one(cpu_backend):
    L = _mech_loss(6)                                            # 2 dofs per node
    fake = cpu_backend(L)
    K = np.random.default_rng(3).uniform(0.2, 1.0, L._nn)
    jac, _ = L.ComputeJacobianMatrixAndResidualVector(K, L.ApplyDirichletBCOnDofVector(np.zeros(L.total_number_of_dofs)))
    A = linalg.SellOperator(L, jac)
    v = torch.as_tensor(np.random.default_rng(4).standard_normal(L.total_number_of_dofs))
    y_block = A.matvec(v).clone()
    assert fake.calls.get("sell_spmv_block", 0) == 1
    A.use_block_kernel = False
    y_scalar = A.matvec(v)
    assert fake.calls["sell_spmv_block"] == 1 and torch.equal(y_block, y_scalar)


@pytest.mark.parametrize("precond", [None, "jacobi"])
def test_bicgstab_matches_direct_solve(cpu_backend, precond):
    L = _mech_loss(9)
    fake = cpu_backend(L)
    rng = np.random.default_rng(1)
    K = rng.uniform(0.2, 1.0, L._nn)
    u0 = L.ApplyDirichletBCOnDofVector(np.zeros(L.total_number_of_dofs))
    jac, R = L.ComputeJacobianMatrixAndResidualVector(K, u0)
    A = linalg.SellOperator(L, jac)
    b = -R
    x, info = linalg.bicgstab(A, b, x0=None, tol=1e-12, atol=0.0, maxiter=2000,
                              M_diagonal=A.diagonal() if precond else None)
    ref = spla.spsolve(A.to_scipy_csr().tocsc(), b.numpy())
    assert info > 0, info
    assert np.abs(x.numpy() - ref).max() <= 1e-8 * np.abs(ref).max()
    assert fake.calls["sell_spmv"] == 2 * info + 1 or fake.calls["sell_spmv"] == 2 * info   # 2 products per iteration
    # the operator itself
    v = torch.as_tensor(rng.standard_normal(L.total_number_of_dofs))
    assert np.abs(A.matvec(v).numpy() - A.to_scipy_csr() @ v.numpy()).max() <= 1e-13


def test_linear_and_adjoint_solvers_reproduce_the_reference_integration_golden(cpu_backend):
    """tests/integration/test_mechanical_2D_sa.py with its own settings (JAX-direct for both solves)."""
    with open(os.path.join(ROOT, "tests", "golden", "reference_unit_goldens.json")) as fh:
        rec = json.load(fh)["tests/integration/test_mechanical_2D_sa.py"]
    K = np.array(rec["setUp"]["assign"]["random_K"])
    L = _mech_loss(5)
    cpu_backend(L)
    resp = FiniteElementResponse("test_response", "(E**2)*U[0]", L, NodalControl("E", L.fe_mesh))
    fe_setting = {"linear_solver_settings": {"solver": "JAX-direct", "tol": 1e-6, "atol": 1e-6, "maxiter": 1000,
                                             "pre-conditioner": "ilu"},
                  "nonlinear_solver_settings": {"rel_tol": 1e-5, "abs_tol": 1e-5, "maxiter": 10, "load_incr": 5}}
    linear_fe_solver = FiniteElementLinearResidualBasedSolver("linear_fe_solver", L, fe_setting)
    adj_fe_solver = AdjointFiniteElementSolver("first_adj_fe_solver", resp, {"linear_solver_settings": {"solver": "JAX-direct"}})
    resp.Initialize()
    linear_fe_solver.Initialize()
    adj_fe_solver.Initialize()
    ndof = L.total_number_of_dofs
    FE_UV = linear_fe_solver.Solve(K, np.zeros(ndof))
    FE_adj_UV = adj_fe_solver.Solve(K, FE_UV, np.ones(ndof))
    a = rec["test_sensitivites"]["asserts"]
    np.testing.assert_allclose(resp.ComputeAdjointNodalControlDerivatives(K, FE_UV, FE_adj_UV).numpy(), a[0]["value"],
                               rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(resp.ComputeAdjointNodalShapeDerivatives(K, FE_UV, FE_adj_UV).numpy(), a[1]["value"],
                               rtol=1e-5, atol=1e-5)
    # the same two solves with the device BiCGSTAB (the reference's default linear solver)
    it = {"linear_solver_settings": {"solver": "JAX-bicgstab", "tol": 1e-12, "atol": 1e-14, "maxiter": 500}}
    s2 = FiniteElementLinearResidualBasedSolver("it", L, it)
    a2 = AdjointFiniteElementSolver("it_adj", resp, it)
    s2.Initialize()
    a2.Initialize()
    u_it = s2.Solve(K, np.zeros(ndof))
    assert s2.last_linear_solve_info > 0
    assert np.abs(u_it.numpy() - FE_UV.numpy()).max() <= 1e-9 * np.abs(FE_UV.numpy()).max()
    lam_it = a2.Solve(K, u_it, np.zeros(ndof))
    assert np.abs(lam_it.numpy() - FE_adj_UV.numpy()).max() <= 1e-8 * np.abs(FE_adj_UV.numpy()).max()


def test_newton_solver_follows_the_reference_loop(cpu_backend):
    """fe_nonlinear_residual_based_solver.py:107-170 on a Neo-Hooke quad mesh: same iterates as the loop written
    out with the oracle and SciPy, including the reference's habit of NOT applying the update of the converged
    iteration."""
    coords, conn, sets = _square_mesh(6)
    bc = {"Ux": {"left": 0.0, "right": 0.3}, "Uy": {"left": 0.0, "right": 0.05}}
    L = fake_loss("neohooke", "quad", 2, coords, conn, sets, ["Ux", "Uy"], bc, MAT, [1.0, 0.3] + [0.0] * 10)
    cpu_backend(L)
    K = np.random.default_rng(2).uniform(0.5, 1.0, L._nn)
    settings = {"linear_solver_settings": {"solver": "JAX-direct"},
                "nonlinear_solver_settings": {"rel_tol": 1e-9, "abs_tol": 1e-9, "maxiter": 8, "load_incr": 3}}
    solver = FiniteElementNonLinearResidualBasedSolver("nl", L, settings)
    solver.Initialize()
    ndof = L.total_number_of_dofs
    u = solver.Solve(K, np.zeros(ndof)).numpy()

    didx, dval = L.dirichlet_indices, L.dirichlet_values
    ref = np.zeros(ndof)
    hist = {}
    for step in range(1, 4):
        ref[didx] = step / 3 * dval
        hist[step] = []
        for i in range(1, 9):
            data, idx, R = assembly.assemble("neohooke", "quad", 2, coords, conn, K, ref, didx, MAT)
            A = sp.csr_array((data, (idx[:, 0], idx[:, 1])), shape=(ndof, ndof))
            du = spla.spsolve(A.tocsc(), -R)
            rn, dn = np.linalg.norm(R), np.linalg.norm(du)
            hist[step].append(rn)
            if rn < 1e-9 or dn < 1e-9 or i == 8:
                break
            ref = ref + du
    assert np.abs(u - ref).max() <= 1e-10 * np.abs(ref).max()
    for step in range(1, 4):
        got = solver.convergence_history[step]["res_norm"]
        assert len(got) == len(hist[step]) and len(got) >= 4
        assert np.allclose(got, hist[step], rtol=1e-6, atol=1e-13)
        assert got[-1] < 1e-7                                    # quadratic convergence reached the tolerance


# ---------------------------------------------------------------------------------- slab-partitioned solve (gloo)
def _slab_solve_worker(rank, world, port, shim_path, out):
    import torch.distributed as dist
    from folax_b200.distributed import SlabPartition
    from tests import cpu_backend as cb
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 3
        part = SlabPartition(n, n, 3 * world, 1.0, 1.0, 1.5, rank, world)
        m = part.mesh
        coords, conn = np.asarray(m.GetNodesCoordinates()), m.GetElementsNodes("hexahedron")
        dofs = ["Ux", "Uy", "Uz"]
        bc = {d: {"left": 0.0, "right": 0.1} for d in dofs}
        L = cb.fake_loss("mechanical", "hexahedron", 2, coords, conn, m.node_sets, dofs, bc, MAT, [1.0, 0.3] + [0.0] * 10)
        cb.install_globally(L, C.CDLL(shim_path))
        gids = part.global_node_ids()
        nn_glob = (n + 1) * (n + 1) * (3 * world + 1)
        Kg = np.random.default_rng(0).uniform(0.2, 1.0, nn_glob)            # the same global field on every rank
        u0 = L.ApplyDirichletBCOnDofVector(np.zeros(L.total_number_of_dofs))
        jac, R = L.ComputeJacobianMatrixAndResidualVector(Kg[gids], u0)
        part.halo_sum(R, 3)                                                  # assembled residual on the interface planes
        A = linalg.SlabOperator(L, jac, part)
        x, info = linalg.bicgstab(A, -R, x0=None, tol=1e-12, atol=0.0, maxiter=3000, M_diagonal=A.diagonal())
        out[rank] = (gids, (u0 + x).numpy(), info)
    finally:
        dist.destroy_process_group()


def test_slab_partitioned_bicgstab_matches_the_undivided_solve(shim):
    """One linear elastic solve on a hex box cut into 2 z-slabs (one process each, gloo): local SELL products +
    the interface exchange + ownership-weighted, all-reduced dot products give the solution of the undivided mesh,
    identical on both copies of the interface plane."""
    import socket
    import torch.multiprocessing as mp
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_slab_solve_worker, args=(world, port, shim._name, out), nprocs=world, join=True)
    n = 3
    mesh = folax_b200.create_3D_box_mesh(n, n, 3 * world, 1.0, 1.0, 1.5)
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("hexahedron")
    dofs = ["Ux", "Uy", "Uz"]
    didx, dval = assembly.dirichlet_vectors(dofs, {d: {"left": 0.0, "right": 0.1} for d in dofs}, mesh.node_sets)
    ndof = 3 * len(coords)
    Kg = np.random.default_rng(0).uniform(0.2, 1.0, len(coords))
    u0 = assembly.full_dof_vector(np.zeros((1, ndof)), didx, dval)[0]
    data, idx, R = assembly.assemble("mechanical", "hexahedron", 2, coords, conn, Kg, u0, didx, MAT)
    A = sp.csr_array((data, (idx[:, 0], idx[:, 1])), shape=(ndof, ndof))
    ref = (u0 + spla.spsolve(A.tocsc(), -R)).reshape(-1, 3)
    for rank in range(world):
        gids, u, info = out[rank]
        assert info > 0
        assert np.abs(u.reshape(-1, 3) - ref[gids]).max() <= 1e-8 * np.abs(ref).max()
    (g0, u_0, _), (g1, u_1, _) = out[0], out[1]
    shared = np.intersect1d(g0, g1)
    a = u_0.reshape(-1, 3)[np.searchsorted(g0, shared)]
    b = u_1.reshape(-1, 3)[np.searchsorted(g1, shared)]
    assert len(shared) == (n + 1) ** 2 and np.abs(a - b).max() <= 1e-13 * np.abs(ref).max()


def test_sell_plan_edge_cases(shim):
    """Empty matrix, empty rows, a row count that is not a multiple of the slice height, no diagonal entry."""
    empty = sp.csr_array((0, 0))
    plan = sell_plan.build(empty.indptr, empty.indices)
    assert plan["nrows"] == 0 and plan["total"] == 0 and plan["slice_ptr"].tolist() == [0]
    rng = np.random.default_rng(5)
    dense = rng.standard_normal((45, 45)) * (rng.uniform(size=(45, 45)) < 0.2)
    dense[[3, 17, 44], :] = 0.0                                   # empty rows (one of them the last)
    np.fill_diagonal(dense[:10, :10], 0.0)                        # rows without a diagonal entry
    A = sp.csr_array(dense)
    A.sort_indices()
    plan = sell_plan.build(A.indptr, A.indices, chunk_rows=32)
    assert plan["node_cols"] is None
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    vals = np.zeros(plan["total"])
    assert shim.host_gather_values(C.c_longlong(plan["total"]), p(plan["src"]), p(np.ascontiguousarray(A.data)), p(vals)) == 0
    x, y = rng.standard_normal(45), np.full(45, np.nan)
    assert shim.host_sell_spmv(C.c_longlong(45), p(plan["slice_ptr"]), p(plan["cols"]), p(vals), p(x), p(y)) == 0
    assert np.abs(y - A @ x).max() <= 1e-14 and y[3] == 0.0 and y[44] == 0.0
    diag = np.full(45, np.nan)
    assert shim.host_gather_values(C.c_longlong(45), p(plan["diag_src"]), p(np.ascontiguousarray(A.data)), p(diag)) == 0
    assert np.array_equal(diag, A.diagonal())                     # missing diagonals gather as 0


@pytest.mark.parametrize("precond", [None, "jacobi"])
@pytest.mark.parametrize("check_every", [1, 8])
def test_device_scalar_bicgstab_equals_the_host_scalar_loop(cpu_backend, precond, check_every):
    """Same iterates, same stopping iteration: the scalar recurrences run in bicg_scalar_stage (device) instead of
    Python, the vector kernels are gated by the device-side state, the host reads (state, k) once per batch."""
    L = _mech_loss(8)
    fake = cpu_backend(L)
    K = np.random.default_rng(7).uniform(0.2, 1.0, L._nn)
    u0 = L.ApplyDirichletBCOnDofVector(np.zeros(L.total_number_of_dofs))
    jac, R = L.ComputeJacobianMatrixAndResidualVector(K, u0)
    A = linalg.SellOperator(L, jac)
    diag = A.diagonal() if precond else None
    for tol, maxiter in ((1e-10, 2000), (1e-10, 7), (1e-3, 2000)):          # converged | stopped by maxiter | loose
        x_h, k_h = linalg.bicgstab(A, -R, x0=u0, tol=tol, atol=0.0, maxiter=maxiter, M_diagonal=diag)
        x_d, k_d = linalg.bicgstab_device(A, -R, x0=u0, tol=tol, atol=0.0, maxiter=maxiter, M_diagonal=diag,
                                          check_every=check_every)
        assert k_d == k_h, (tol, maxiter, k_d, k_h)
        assert torch.equal(x_d, x_h)
    assert fake.calls["bicg_scalars"] > 0 and fake.calls["vec_op_dev"] > 0
    # zero right-hand side: nothing to do, x0 comes back
    x_d, k_d = linalg.bicgstab_device(A, torch.zeros_like(R), x0=None, tol=1e-8, check_every=check_every)
    assert k_d == 0 and not x_d.any()
