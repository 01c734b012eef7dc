"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: slab partition + halo-DOF sum against a
single-domain oracle assembly, gradient all-reduce, sample sharding."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from folax_b200.distributed import (EarlyReduceLinear, GradientReducer, SlabPartition, allreduce_gradients,
                                    allreduce_loss_statistics, shard_batch)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _halo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import assembly
        n, mat = 3, {"young_modulus": 1.0, "poisson_ratio": 0.3}
        part = SlabPartition(n, n, 2 * world, 1.0, 1.0, 2.0, rank, world)
        m = part.mesh
        coords, conn = np.asarray(m.GetNodesCoordinates()), m.GetElementsNodes("hexahedron")
        gids = part.global_node_ids()
        rng = np.random.default_rng(0)                      # same global fields on every rank
        nn_glob = (n + 1) * (n + 1) * (2 * world + 1)
        Kg, ug = rng.uniform(0.1, 1, nn_glob), 0.01 * rng.standard_normal(3 * nn_glob)
        K = Kg[gids]
        u = ug.reshape(-1, 3)[gids].reshape(-1)
        didx, _ = assembly.dirichlet_vectors(["Ux", "Uy", "Uz"], {d: {"left": 0.0, "right": 0.1} for d in
                                                                   ("Ux", "Uy", "Uz")}, m.node_sets)
        data, idx, R = assembly.assemble("mechanical", "hexahedron", 2, coords, conn, K, u, didx, mat)
        Rt = torch.tensor(R)
        part.halo_sum(Rt, 3)
        out[rank] = (gids, Rt.numpy(), data, part.element_offset)
    finally:
        dist.destroy_process_group()


def test_slab_partition_halo_sum_matches_single_domain():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_halo_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    # single-domain reference on the whole box
    import folax_b200
    from oracle import assembly
    n, mat = 3, {"young_modulus": 1.0, "poisson_ratio": 0.3}
    mesh = folax_b200.create_3D_box_mesh(n, n, 2 * world, 1.0, 1.0, 2.0)
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("hexahedron")
    rng = np.random.default_rng(0)
    Kg, ug = rng.uniform(0.1, 1, len(coords)), 0.01 * rng.standard_normal(3 * len(coords))
    didx, _ = assembly.dirichlet_vectors(["Ux", "Uy", "Uz"], {d: {"left": 0.0, "right": 0.1} for d in
                                                               ("Ux", "Uy", "Uz")}, mesh.node_sets)
    data, idx, R = assembly.assemble("mechanical", "hexahedron", 2, coords, conn, Kg, ug, didx, mat)
    data = data.reshape(len(conn), -1)
    for rank in range(world):
        gids, Rl, dl, eoff = out[rank]
        np.testing.assert_allclose(Rl.reshape(-1, 3), R.reshape(-1, 3)[gids], atol=1e-13)
        dl = dl.reshape(-1, data.shape[1])
        np.testing.assert_allclose(dl, data[eoff:eoff + len(dl)], atol=1e-14)   # Jacobian blocks need no exchange


def _grad_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 3)).double()
        x = torch.arange(24, dtype=torch.float64).reshape(6, 4) / 10
        sl = shard_batch(6, rank, world)
        loss = net(x[sl]).pow(2).sum() / 6
        loss.backward()
        ref = [p.grad.clone() for p in net.parameters()]
        allreduce_gradients(list(net.parameters()), bucket_bytes=64)        # every gradient is "large": reduced in place
        out[rank] = [p.grad.numpy().copy() for p in net.parameters()]
        for p, g in zip(net.parameters(), ref):
            p.grad = g.clone()
        allreduce_gradients(list(net.parameters()), bucket_bytes=1 << 20)   # all in one flattened bucket
        for p, g in zip(net.parameters(), out[rank]):
            np.testing.assert_allclose(p.grad.numpy(), g, atol=1e-15)
    finally:
        dist.destroy_process_group()


def _early_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        first = torch.nn.Linear(4, 8)                   # same construction order (= same initial weights) as the check
        last = EarlyReduceLinear(8, 3)
        net = torch.nn.Sequential(first, torch.nn.Tanh(), last).double()
        reducer = GradientReducer(list(net.parameters()), exclude=list(last.parameters()))
        last.reducer = reducer
        x = torch.arange(24, dtype=torch.float64).reshape(6, 4) / 10
        sl = shard_batch(6, rank, world)
        (net(x[sl]).pow(2).sum() / 6).backward()       # hooks + the early reduce of the last layer: all grads summed
        reducer.wait()
        out[rank] = [p.grad.numpy().copy() for p in net.parameters()]
        reducer.close()
    finally:
        dist.destroy_process_group()


def test_early_reduce_linear_sums_like_the_plain_all_reduce():
    """GradientReducer (hooks) + EarlyReduceLinear (weight / bias gradients reduced from inside backward, before the
    input gradient): every rank ends with the gradient of the undivided batch, the excluded layer reduced exactly once."""
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_early_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 3)).double()
    x = torch.arange(24, dtype=torch.float64).reshape(6, 4) / 10
    (net(x).pow(2).sum() / 6).backward()
    for rank in range(world):
        for g, p in zip(out[rank], net.parameters()):
            np.testing.assert_allclose(g, p.grad.numpy(), rtol=1e-13, atol=1e-15)
    # single process: a plain Linear
    lone = EarlyReduceLinear(8, 3).double()
    y = lone(torch.ones(2, 8, dtype=torch.float64))
    assert y.shape == (2, 3)


def test_data_parallel_gradient_allreduce():
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_grad_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 3)).double()
    x = torch.arange(24, dtype=torch.float64).reshape(6, 4) / 10
    (net(x).pow(2).sum() / 6).backward()
    for rank in range(world):
        for g, p in zip(out[rank], net.parameters()):
            np.testing.assert_allclose(g, p.grad.numpy(), atol=1e-14)


def _stats_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        E = torch.tensor([3.0, -1.0, 4.0, 1.5, 9.0, 2.5], dtype=torch.float64)[shard_batch(6, rank, world)]
        mean, (mn, mx, mean2) = allreduce_loss_statistics(E.mean(), (E.min(), E.max(), E.mean()))
        out[rank] = (float(mean), float(mn), float(mx), float(mean2))
    finally:
        dist.destroy_process_group()


def test_loss_statistics_over_ranks():
    """(mean, (min, max, mean)) of fe_loss.py:262 over a batch sharded on 2 ranks."""
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_stats_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    E = np.array([3.0, -1.0, 4.0, 1.5, 9.0, 2.5])
    for rank in range(world):
        assert np.allclose(out[rank], (E.mean(), E.min(), E.max(), E.mean()), rtol=0, atol=1e-15)
    m, (mn, mx, m2) = allreduce_loss_statistics(torch.tensor(2.0), (torch.tensor(1.0), torch.tensor(3.0), torch.tensor(2.0)))
    assert (float(m), float(mn), float(mx), float(m2)) == (2.0, 1.0, 3.0, 2.0)        # single process: pass-through


def test_shard_batch():
    assert shard_batch(8, 1, 4) == slice(2, 4)
    with pytest.raises(ValueError):
        shard_batch(10, 0, 4)
