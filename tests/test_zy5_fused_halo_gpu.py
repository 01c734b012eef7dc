"""Fused slab step (fol_assemble_elements_halo + fol_residual_gather_halo, csrc/assemble_hex_common.cuh) on ONE GPU with
no neighbour connected: the interface-first tile order, the in-kernel plane gather and the work counters must leave
exactly what the plain two-launch path leaves -- Jacobian data, residual and (J2) Gauss-point history bit for bit --
over several steps (the counters are re-armed by the completing kernel).  The NVLink peer stores themselves need two
GPUs: tests/test_distributed_gpu.py and the halo_check of bench.py at N > 1."""
import ctypes as C

import numpy as np
import pytest
import torch

import folax_b200
from folax_b200 import _lib
from folax_b200.distributed import SlabPartition, assemble_overlapped
from folax_b200.loss_functions import ElastoplasticityLoss3DHexa, MechanicalLoss3DHexa

pytestmark = pytest.mark.gpu

BC = {d: {"left": 0.0, "right": 0.1} for d in ("Ux", "Uy", "Uz")}
J2MAT = {"young_modulus": 3.0, "poisson_ratio": 0.3, "iso_hardening_parameter_1": 0.4,
         "iso_hardening_param_2": 10.0, "yield_limit": 0.2}


def _lonely_halo(part, loss):
    """A halo object without connected neighbours: every push is a plain gather, every add a no-op."""
    h = C.c_void_p()
    _lib.check(_lib.load().fol_halo_create(C.byref(h), loss._dt, part.plane_nodes * 3))
    part._halo, part._halo_step = h, 0
    return h


@pytest.mark.parametrize("shape", [(6, 5, 7), (9, 7, 2), (5, 5, 1), (16, 16, 12)])
def test_fused_mechanical_step_equals_plain_step(shape):
    nx, ny, nz = shape        # (6,5,7): layers of 30 elements (tiles straddle layers); nz = 2, 1: everything is interface
    part = SlabPartition(nx, ny, 2 * nz, 1.0, 1.0, 2.0, 0, 2)
    folax_b200.perturb_interior_nodes(part.mesh, 0.15, 2)
    loss = MechanicalLoss3DHexa("m", {"dirichlet_bc_dict": BC, "material_dict": {"young_modulus": 1.0, "poisson_ratio": 0.3},
                                      "body_foce": [0.1, -0.2, 0.3]}, part.mesh)
    loss.Initialize()
    rng = np.random.default_rng(1)
    K = torch.tensor(rng.uniform(0.1, 1.0, loss._nn), device="cuda")
    ke_ref = torch.empty(loss._ne * 576, dtype=torch.float64, device="cuda")
    ke = torch.full_like(ke_ref, float("nan"))
    _lonely_halo(part, loss)
    try:
        for step in range(3):
            u = torch.tensor(0.01 * rng.standard_normal(loss.total_number_of_dofs), device="cuda")
            _, R_ref = loss._assemble(K, u, False, ke_out=ke_ref)
            ke.fill_(float("nan"))
            _, R = assemble_overlapped(loss, part, K, u, ke, None)
            torch.cuda.synchronize()
            assert torch.equal(ke, ke_ref), f"step {step}: Jacobian data differ"
            assert torch.equal(R, R_ref), f"step {step}: residual differs"
    finally:
        part.close_peer_halo()


def test_fused_j2_step_equals_plain_step():
    part = SlabPartition(6, 6, 10, 1.0, 1.0, 2.0, 0, 2)
    loss = ElastoplasticityLoss3DHexa("ep", {"dirichlet_bc_dict": BC, "material_dict": dict(J2MAT)}, part.mesh)
    loss.Initialize()
    rng = np.random.default_rng(3)
    K = torch.ones(loss._nn, dtype=torch.float64, device="cuda")
    state = torch.zeros(loss.GetStateShape(), dtype=torch.float64, device="cuda")
    ke = torch.empty(loss._ne * 576, dtype=torch.float64, device="cuda")
    u = torch.zeros(loss.total_number_of_dofs, dtype=torch.float64, device="cuda")
    _lonely_halo(part, loss)
    try:
        for step in range(3):
            u = u + torch.tensor(0.006 * rng.standard_normal(u.shape[0]), device="cuda")
            st_ref, jac_ref, R_ref = loss.ComputeJacobianMatrixAndResidualVector(K, u, state)
            st = torch.full_like(state, float("nan"))
            ke.fill_(float("nan"))
            _, R = assemble_overlapped(loss, part, K, u, ke, None, state_in=state, state_out=st)
            torch.cuda.synchronize()
            assert torch.equal(ke, jac_ref.data) and torch.equal(R, R_ref) and torch.equal(st, st_ref), f"step {step}"
            plastic = float((st_ref[..., -1] > state[..., -1]).double().mean())
            state = st_ref
        assert 0.05 < plastic <= 1.0
    finally:
        part.close_peer_halo()
