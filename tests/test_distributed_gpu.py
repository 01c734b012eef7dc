"""2-GPU NCCL test of the slab-partitioned assembly with the overlapped halo-DOF exchange (skipped on a
single-GPU box; run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from folax_b200.distributed import SlabPartition, assemble_overlapped
        from folax_b200.loss_functions import MechanicalLoss3DHexa
        n = 6
        part = SlabPartition(n, n, 4 * world, 1.0, 1.0, 2.0, rank, world)
        bc = {d: {"left": 0.0, "right": 0.1} for d in ("Ux", "Uy", "Uz")}
        loss = MechanicalLoss3DHexa("p", {"dirichlet_bc_dict": bc, "material_dict": {"young_modulus": 1.0,
                                                                                    "poisson_ratio": 0.3}}, part.mesh)
        loss.Initialize()
        gids = part.global_node_ids()
        rng = np.random.default_rng(0)
        nn_glob = (n + 1) * (n + 1) * (4 * world + 1)
        Kg, ug = rng.uniform(0.1, 1, nn_glob), 0.01 * rng.standard_normal(3 * nn_glob)
        K, u = Kg[gids], ug.reshape(-1, 3)[gids].reshape(-1)
        ke = torch.empty(loss._ne * 576, dtype=torch.float64, device="cuda")
        comm = torch.cuda.Stream()
        for _ in range(3):
            data, R = assemble_overlapped(loss, part, K, u, ke, comm)
        torch.cuda.synchronize()
        R_nccl = R.clone()
        # same exchange over NVLink peer memory: first the FUSED path (the element-stage launch visits the interface
        # layers first and pushes the planes itself, csrc/assemble_hex_common.cuh), then the layered path
        # (csrc/halo.cu: plane gather fused with the push, separate add) on a fresh halo object
        import folax_b200.distributed as D
        for fused in (True, False):
            D.FUSED_HALO = fused
            assert part.enable_peer_halo(loss)
            ke.fill_(float("nan"))
            for _ in range(5):                     # several steps: both buffer parities, counters re-armed / counting
                data, R = assemble_overlapped(loss, part, K, u, ke, comm)
            torch.cuda.synchronize()
            assert bool((R == R_nccl).all()), f"peer-memory halo sum (fused={fused}) differs from the NCCL send/recv path"
            dist.barrier()
            part.close_peer_halo()
            dist.barrier()
        D.FUSED_HALO = True
        data, R = data.clone(), R.clone()           # the J2 section below reuses the `ke` buffer
        # J2 elastoplasticity with Gauss-point history through the fused path against the NCCL path
        from folax_b200.loss_functions import ElastoplasticityLoss3DHexa
        mat = {"young_modulus": 3.0, "poisson_ratio": 0.3, "iso_hardening_parameter_1": 0.4,
               "iso_hardening_param_2": 10.0, "yield_limit": 0.2}
        lj = ElastoplasticityLoss3DHexa("ep", {"dirichlet_bc_dict": bc, "material_dict": mat}, part.mesh)
        lj.Initialize()
        st0 = torch.zeros(lj.GetStateShape(), dtype=torch.float64, device="cuda")
        st_a, st_b = torch.empty_like(st0), torch.empty_like(st0)
        ke2 = torch.empty_like(ke)
        uj = 0.5 * u
        _, Rj_nccl = assemble_overlapped(lj, part, K, uj, ke2, comm, state_in=st0, state_out=st_a)
        torch.cuda.synchronize()
        Rj_nccl = Rj_nccl.clone()
        assert part.enable_peer_halo(lj)
        for _ in range(3):
            _, Rj = assemble_overlapped(lj, part, K, uj, ke, comm, state_in=st0, state_out=st_b)
        torch.cuda.synchronize()
        assert bool((Rj == Rj_nccl).all()) and torch.equal(st_a, st_b) and torch.equal(ke, ke2), "fused J2 slab step"
        assert float((st_b[..., -1] > 0).double().mean()) > 0.05
        dist.barrier()
        part.close_peer_halo()
        out[rank] = (gids, R.cpu().numpy(), data.cpu().numpy(), part.element_offset)
    finally:
        dist.destroy_process_group()


def test_overlapped_halo_exchange_matches_single_domain():
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
    rng = np.random.default_rng(0)
    Kg, ug = rng.uniform(0.1, 1, len(coords)), 0.01 * rng.standard_normal(3 * len(coords))
    didx, _ = assembly.dirichlet_vectors(["Ux", "Uy", "Uz"], {d: {"left": 0.0, "right": 0.1} for d in
                                                               ("Ux", "Uy", "Uz")}, mesh.node_sets)
    data, _, R = assembly.assemble("mechanical", "hexahedron", 2, coords, conn, Kg, ug, didx,
                                   {"young_modulus": 1.0, "poisson_ratio": 0.3})
    data = data.reshape(len(conn), -1)
    for rank in range(world):
        gids, Rl, dl, eoff = out[rank]
        assert np.abs(Rl.reshape(-1, 3) - R.reshape(-1, 3)[gids]).max() <= 4e-12 * np.abs(R).max()
        dl = dl.reshape(-1, 576)
        assert np.abs(dl - data[eoff:eoff + len(dl)]).max() <= 1e-12 * np.abs(data).max()
