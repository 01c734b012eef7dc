"""2-GPU NCCL run of the slab-partitioned linear solve (folax_b200.linalg.SlabOperator + BiCGSTAB): local SELL
products, interface exchange, ownership-weighted all-reduced dot products -- against the undivided solve of the
oracle's system (skipped on a single-GPU box; run with `gpurun --gpus 2`).  The same logic runs on 2 gloo ranks in
tests/test_solvers_glue_cpu.py."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
MAT = {"young_modulus": 1.0, "poisson_ratio": 0.3}


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from folax_b200 import linalg
        from folax_b200.distributed import SlabPartition
        from folax_b200.loss_functions import MechanicalLoss3DHexa
        n = 6
        part = SlabPartition(n, n, 4 * world, 1.0, 1.0, 2.0, rank, world)
        bc = {d: {"left": 0.0, "right": 0.1} for d in ("Ux", "Uy", "Uz")}
        loss = MechanicalLoss3DHexa("p", {"dirichlet_bc_dict": bc, "material_dict": dict(MAT)}, part.mesh)
        loss.Initialize()
        gids = part.global_node_ids()
        nn_glob = (n + 1) * (n + 1) * (4 * world + 1)
        Kg = np.random.default_rng(0).uniform(0.2, 1.0, nn_glob)
        u0 = loss.ApplyDirichletBCOnDofVector(np.zeros(loss.total_number_of_dofs))
        jac, R = loss.ComputeJacobianMatrixAndResidualVector(Kg[gids], u0)
        part.halo_sum(R, 3)
        A = linalg.SlabOperator(loss, jac, part)
        rhs = -R
        x, info = linalg.bicgstab(A, rhs, x0=None, tol=1e-11, atol=0.0, maxiter=4000, M_diagonal=A.diagonal())
        torch.cuda.synchronize()
        out[rank] = (gids, (u0 + x).cpu().numpy(), info)
    finally:
        dist.destroy_process_group()


def test_two_slabs_solve_like_the_undivided_mesh():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    import folax_b200
    from oracle import assembly
    n = 6
    mesh = folax_b200.create_3D_box_mesh(n, n, 4 * world, 1.0, 1.0, 2.0)
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
        assert np.abs(u.reshape(-1, 3) - ref[gids]).max() <= 1e-7 * np.abs(ref).max()
