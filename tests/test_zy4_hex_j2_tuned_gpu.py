"""Tuned Hex8 J2 element-stage kernel (csrc/assemble_hex_j2.cu, BASELINE.json configs[4]): against the oracle on a small
mesh, against the generic kernel on a mesh with a ragged last tile / Dirichlet rows / body force / non-zero history, and
on the 64^3 box through a size-independent property (the two kernels agree on every element; second load step)."""
import numpy as np
import pytest
import torch

import folax_b200
from folax_b200 import _lib
from folax_b200 import loss_functions as lf
from oracle import assembly

pytestmark = pytest.mark.gpu

MAT = {"young_modulus": 3.0, "poisson_ratio": 0.3, "iso_hardening_parameter_1": 0.4,
       "iso_hardening_param_2": 10.0, "yield_limit": 0.2}
BC = {d: {"left": 0.0, "right": 0.1} for d in ("Ux", "Uy", "Uz")}


def _loss(mesh, body=None):
    settings = {"dirichlet_bc_dict": BC, "material_dict": dict(MAT)}
    if body is not None:
        settings["body_foce"] = body
    loss = lf.ElastoplasticityLoss3DHexa("ep", settings, mesh)
    loss.Initialize()
    return loss


def _with_tuned(flag, fn):
    lib = _lib.load()
    prev = lib.fol_set_tuned_kernels(flag)
    try:
        return fn()
    finally:
        lib.fol_set_tuned_kernels(prev)


def test_tuned_kernel_matches_oracle_over_two_load_steps():
    mesh = folax_b200.perturb_interior_nodes(folax_b200.create_3D_box_mesh(3, 2, 3, 1.0, 0.8, 1.1), 0.2, 1)
    loss = _loss(mesh, body=[0.3, -0.2, 0.5])
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("hexahedron")
    rng = np.random.default_rng(2)
    state = np.zeros(loss.GetStateShape())
    u = np.zeros(loss.total_number_of_dofs)
    K = np.ones(len(coords))
    plastic = 0
    for step in range(2):
        u = u + 0.03 * rng.standard_normal(u.shape)
        new_state, jac, R = _with_tuned(1, lambda: loss.ComputeJacobianMatrixAndResidualVector(K, u, state))
        ref_state, data, idx, Rref = assembly.assemble_j2("hexahedron", 2, coords, conn, u, state, loss.dirichlet_indices,
                                                          {**MAT, "body_force": np.array([0.3, -0.2, 0.5])})
        assert np.array_equal(jac.indices.cpu().numpy(), idx)
        assert np.abs(jac.data.cpu().numpy() - data).max() <= 1e-11 * np.abs(data).max()
        assert np.abs(R.cpu().numpy() - Rref).max() <= 1e-11 * np.abs(Rref).max()
        assert np.abs(new_state.cpu().numpy() - ref_state).max() <= 1e-11 * np.abs(ref_state).max()
        plastic += int((ref_state[..., -1] > state[..., -1]).sum())
        state = ref_state
    assert plastic > 0.2 * 2 * state.shape[0] * 8


@pytest.mark.parametrize("n,amp", [((5, 3, 7), 0.02), ((5, 3, 7), 0.4), ((64, 64, 64), 0.05)])
def test_tuned_kernel_matches_generic_kernel(n, amp):
    """(5,3,7): 105 elements = 26 tiles + 1 ragged element; amplitudes for mostly-elastic and all-plastic states.
    64^3: the size-independent property at a size the oracle's Python loops cannot reach."""
    nx, ny, nz = n
    mesh = folax_b200.perturb_interior_nodes(folax_b200.create_3D_box_mesh(nx, ny, nz, 1.0, 1.0, 1.0), 0.15, 3)
    loss = _loss(mesh, body=[0.1, 0.2, -0.3])
    g = torch.Generator(device="cuda").manual_seed(5)
    h = 1.0 / max(n)
    K = torch.ones(loss._nn, dtype=torch.float64, device="cuda")
    state = torch.zeros(loss.GetStateShape(), dtype=torch.float64, device="cuda")
    u = torch.zeros(loss.total_number_of_dofs, dtype=torch.float64, device="cuda")
    for step in range(2):
        u = u + amp * h * torch.randn(u.shape, generator=g, device="cuda", dtype=torch.float64)
        st_t, jac_t, R_t = _with_tuned(1, lambda: loss.ComputeJacobianMatrixAndResidualVector(K, u, state))
        st_g, jac_g, R_g = _with_tuned(0, lambda: loss.ComputeJacobianMatrixAndResidualVector(K, u, state))
        scale = jac_g.data.abs().max()
        assert (jac_t.data - jac_g.data).abs().max() <= 1e-12 * scale
        assert (R_t - R_g).abs().max() <= 1e-12 * R_g.abs().max()
        assert (st_t - st_g).abs().max() <= 1e-13 * max(float(st_g.abs().max()), 1e-300)
        # the element-local test of the elastic branch leaves the history bit-identical
        elastic = st_g[..., -1] == state[..., -1]
        assert torch.equal(st_t[elastic], state[elastic])
        frac = float((~elastic).double().mean())
        state = st_g
    assert (frac > 0.9) if amp >= 0.4 else (0.02 < frac < 0.98), frac
