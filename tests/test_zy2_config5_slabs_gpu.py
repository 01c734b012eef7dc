"""configs[4] of BASELINE.json on one GPU: J2 elastoplasticity on a hex box, element-partitioned into z-slabs with
per-Gauss-point history.  The two slabs of a 2-rank partition are assembled one after the other on cuda:0 through
`assemble_overlapped` (the layered launch order of the multi-GPU path: interface layers first, interior after);
the neighbour exchange is replaced by adding the two interface planes here, so what is checked is that the
element-range launches read / write the right slices of the Jacobian data AND of the Gauss-point state, against
the oracle on the undivided mesh.  (The exchange itself: tests/test_distributed_gpu.py, 2 GPUs.)"""
import numpy as np
import pytest
import torch

import folax_b200
from folax_b200.distributed import SlabPartition, assemble_overlapped
from folax_b200.loss_functions import ElastoplasticityLoss3DHexa
from oracle import assembly

pytestmark = pytest.mark.gpu

MAT = {"young_modulus": 3.0, "poisson_ratio": 0.3, "iso_hardening_parameter_1": 0.4,
       "iso_hardening_param_2": 10.0, "yield_limit": 0.2}
BC = {d: {"left": 0.0, "right": 0.1} for d in ("Ux", "Uy", "Uz")}


def test_two_slabs_with_history_match_the_undivided_mesh():
    n, nz, world = 4, 8, 2
    mesh = folax_b200.create_3D_box_mesh(n, n, nz, 1.0, 1.0, 2.0)
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("hexahedron")
    nn, ne = len(coords), len(conn)
    didx, _ = assembly.dirichlet_vectors(["Ux", "Uy", "Uz"], BC, mesh.node_sets)
    rng = np.random.default_rng(3)
    state = np.zeros((ne, 8, 7))
    u = np.zeros(3 * nn)
    comm = torch.cuda.Stream()
    plastic = 0
    for step in range(2):                            # second step starts from a non-zero history
        u = u + 0.005 * rng.standard_normal(u.shape)     # about half of the points yield
        ref_state, data, _, R = assembly.assemble_j2("hexahedron", 2, coords, conn, u, state, didx, MAT)
        data = data.reshape(ne, -1)
        R_parts = []
        for rank in range(world):
            part = SlabPartition(n, n, nz, 1.0, 1.0, 2.0, rank, world)
            part.halo_sum = lambda residual, d, group=None: residual          # exchange done by hand below
            loss = ElastoplasticityLoss3DHexa("ep", {"dirichlet_bc_dict": BC, "material_dict": dict(MAT)}, part.mesh)
            loss.Initialize()
            gids = part.global_node_ids()
            e0, nel = part.element_offset, loss._ne
            assert part.nz_local >= 3                 # layered path: two interface launches + the interior one
            ke = torch.empty(nel * 576, dtype=torch.float64, device="cuda")
            st_out = torch.full((nel, 8, 7), float("nan"), dtype=torch.float64, device="cuda")
            _, Rl = assemble_overlapped(loss, part, np.ones(len(gids)), u.reshape(-1, 3)[gids].reshape(-1), ke, comm,
                                        state_in=state[e0:e0 + nel], state_out=st_out)
            torch.cuda.synchronize()
            assert np.abs(ke.cpu().numpy().reshape(nel, -1) - data[e0:e0 + nel]).max() <= 1e-11 * np.abs(data).max()
            got = st_out.cpu().numpy()
            assert not np.isnan(got).any(), "an element range did not write its state slice"
            assert np.abs(got - ref_state[e0:e0 + nel]).max() <= 1e-11 * max(np.abs(ref_state).max(), 1e-300)
            R_parts.append((gids, Rl.cpu().numpy().reshape(-1, 3)))
        Rsum = np.zeros((nn, 3))
        for gids, Rl in R_parts:
            Rsum[gids] += Rl                          # interface plane: the two partial sums meet
        assert np.abs(Rsum.reshape(-1) - R).max() <= 1e-11 * np.abs(R).max()
        plastic += int((ref_state[..., -1] > state[..., -1]).sum())
        state = ref_state
    assert plastic > 0.05 * 2 * ne * 8, "the test must exercise the plastic branch"


def test_state_arguments_are_validated():
    part = SlabPartition(2, 2, 3, 1.0, 1.0, 1.0, 0, 1)
    from folax_b200.loss_functions import MechanicalLoss3DHexa
    loss = MechanicalLoss3DHexa("m", {"dirichlet_bc_dict": BC, "material_dict": {"young_modulus": 1.0,
                                                                                "poisson_ratio": 0.3}}, part.mesh)
    loss.Initialize()
    ke = torch.empty(loss._ne * 576, dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError):
        assemble_overlapped(loss, part, np.ones(loss._nn), np.zeros(loss.total_number_of_dofs), ke,
                            torch.cuda.Stream(), state_in=np.zeros((loss._ne, 8, 7)),
                            state_out=torch.zeros((loss._ne, 8, 7), dtype=torch.float64, device="cuda"))
