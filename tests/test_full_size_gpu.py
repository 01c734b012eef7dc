"""BASELINE.json full-size configuration (128^3 Hex8 elements, 6.44 M dofs, float64) checked through
size-independent properties -- the oracle cannot assemble 2 M elements in seconds, so parity at this
size rests on: (i) a random sample of elements compared against the oracle, (ii) linearity of the
residual, (iii) rigid-body null space, (iv) symmetry of the element blocks, (v) run-to-run
bit-identity, (vi) tuned kernel == generic kernel to rounding."""
import numpy as np
import pytest
import torch

import folax_b200
from folax_b200 import _lib
from folax_b200.loss_functions import MechanicalLoss3DHexa
from oracle import assembly

pytestmark = pytest.mark.gpu

N = 128
MAT = {"young_modulus": 1.0, "poisson_ratio": 0.3}


@pytest.fixture(scope="module")
def big():
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs a large-memory GPU")
    mesh = folax_b200.create_3D_box_mesh(N, N, N, 1.0, 1.0, 1.0)
    bc = {d: {"left": 0.0, "right": 0.1} for d in ("Ux", "Uy", "Uz")}
    loss = MechanicalLoss3DHexa("c2", {"dirichlet_bc_dict": bc, "num_gp": 2, "material_dict": dict(MAT)}, mesh)
    loss.Initialize()
    free = MechanicalLoss3DHexa("c2_free", {"dirichlet_bc_dict": {"Ux": {}, "Uy": {}, "Uz": {}}, "num_gp": 2,
                                            "material_dict": dict(MAT)}, mesh)
    free.Initialize()
    g = torch.Generator(device="cuda").manual_seed(0)
    K = torch.rand(loss._nn, generator=g, device="cuda", dtype=torch.float64) * 0.9 + 0.1
    u = torch.randn(loss.total_number_of_dofs, generator=g, device="cuda", dtype=torch.float64) * 0.01
    return mesh, loss, free, K, u


def test_sampled_elements_match_oracle(big):
    mesh, loss, _, K, u = big
    jac, R = loss.ComputeJacobianMatrixAndResidualVector(K, u)
    assert jac.data.numel() == 2097152 * 576 and jac.shape == (6440067, 6440067)
    rng = np.random.default_rng(1)
    conn = mesh.GetElementsNodes("hexahedron")
    left_layer = np.arange(0, len(conn), N)[:200]                      # elements touching the Dirichlet face x = 0
    sample = np.unique(np.concatenate([rng.integers(0, len(conn), 2000), left_layer, [0, len(conn) - 1]]))
    coords = np.asarray(mesh.GetNodesCoordinates())
    Kh, uh = K.cpu().numpy(), u.cpu().numpy()
    ref, _, _ = assembly.assemble("mechanical", "hexahedron", 2, coords, conn[sample], Kh, uh,
                                  loss.dirichlet_indices, MAT)
    got = jac.data.view(-1, 576)[torch.as_tensor(sample, device="cuda")].cpu().numpy().reshape(-1)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    # BCOO indices of the sampled elements, bit-exact
    idx = jac.indices.view(-1, 576, 2)[torch.as_tensor(sample, device="cuda")].cpu().numpy().reshape(-1, 2)
    assert np.array_equal(idx, assembly.bcoo_indices(conn[sample], 3))
    # run-to-run bit-identity at full size
    jac2, R2 = loss.ComputeJacobianMatrixAndResidualVector(K, u)
    assert torch.equal(jac.data, jac2.data) and torch.equal(R, R2)
    del jac2
    # tuned kernel vs generic kernel
    lib = _lib.load()
    prev = lib.fol_set_tuned_kernels(0)
    try:
        jg, Rg = loss.ComputeJacobianMatrixAndResidualVector(K, u)
    finally:
        lib.fol_set_tuned_kernels(prev)
    scale = jac.data.abs().max().item()
    assert (jac.data - jg.data).abs().max().item() <= 1e-12 * scale
    assert (R - Rg).abs().max().item() <= 4e-12 * R.abs().max().item()


def test_linearity_rigid_modes_and_symmetry(big):
    mesh, _, free, K, u = big
    g = torch.Generator(device="cuda").manual_seed(5)
    u2 = torch.randn(u.shape, generator=g, device="cuda", dtype=torch.float64) * 0.01
    J, R1 = free.ComputeJacobianMatrixAndResidualVector(K, u)
    # symmetry of the element matrices (to rounding: the Gauss-point scale rides on one DMMA operand)
    blocks = J.data.view(-1, 24, 24)
    assert (blocks[::97] - blocks[::97].transpose(1, 2)).abs().max().item() <= 1e-15 * blocks.abs().max().item()
    # rows of an unconstrained element matrix sum to zero per displacement component (rigid translation)
    tr = torch.zeros(24, 3, dtype=torch.float64, device="cuda")
    for c in range(3):
        tr[c::3, c] = 1.0
    assert (blocks[::97] @ tr).abs().max().item() <= 1e-13 * blocks.abs().max().item()
    del J, blocks
    _, R2 = free.ComputeJacobianMatrixAndResidualVector(K, u2)
    _, R12 = free.ComputeJacobianMatrixAndResidualVector(K, 2.0 * u - 3.0 * u2)
    scale = R1.abs().max().item()
    assert (2.0 * R1 - 3.0 * R2 - R12).abs().max().item() <= 1e-11 * scale
    t = torch.tensor([0.3, -0.1, 0.2], dtype=torch.float64, device="cuda").repeat(free._nn)
    _, Rt = free.ComputeJacobianMatrixAndResidualVector(K, t)
    assert Rt.abs().max().item() <= 1e-12
    # total force balance: the assembled internal forces of a free body sum to zero
    assert abs(R1.view(-1, 3).sum(0)).max().item() <= 1e-9 * scale


def test_fol_config3_full_size_batched_loss():
    """configs[2] at full size: 1024 conductivity fields on the 256x256 thermal quad mesh.  The pipelined kernel runs
    with several sample chunks per tile here (unlike the small parity cases): sampled rows against the oracle, the whole
    batch against the generic kernel, and run-to-run bit-identity."""
    import os
    import torch
    from folax_b200.loss_functions import ThermalLoss2DQuad
    mesh = folax_b200.create_2D_square_mesh(1.0, 257)
    loss = ThermalLoss2DQuad("fol", {"dirichlet_bc_dict": {"T": {"left": 1.0, "right": 0.1}}, "beta": 2.0, "c": 4}, mesh)
    loss.Initialize()
    nn, B = mesh.GetNumberOfNodes(), 1024
    g = torch.Generator(device="cuda").manual_seed(3)
    K = torch.rand((B, nn), generator=g, device="cuda", dtype=torch.float64) * 0.9 + 0.1
    u = torch.rand((B, nn), generator=g, device="cuda", dtype=torch.float64)

    def run():
        kk, uu = K.clone().requires_grad_(True), u.clone().requires_grad_(True)
        mean, (mn, mx, _) = loss.ComputeBatchLoss(kk, uu)
        mean.backward()
        return mean.detach(), mn.detach(), mx.detach(), uu.grad, kk.grad

    a = run()
    b = run()
    assert all(bool((x == y).all()) for x, y in zip(a, b))                     # deterministic
    # generic kernel (one thread per owned node, <= 192 per tile) on its own 160-node tile plan; the node sums have the
    # same fixed order whatever the tiling, the per-tile energy shares are added in a different order
    os.environ.update(FOL_ENERGY_V1="1", FOL_ENERGY_MAX_ELEMS="192", FOL_ENERGY_TILE_NODES="160")
    try:
        loss_g = ThermalLoss2DQuad("fol_g", {"dirichlet_bc_dict": {"T": {"left": 1.0, "right": 0.1}}, "beta": 2.0, "c": 4}, mesh)
        loss_g.Initialize()
        loss, loss_keep = loss_g, loss
        c = run()
        loss = loss_keep
    finally:
        for k in ("FOL_ENERGY_V1", "FOL_ENERGY_MAX_ELEMS", "FOL_ENERGY_TILE_NODES"):
            del os.environ[k]
    for x, y in zip(a, c):
        assert (x - y).abs().max().item() <= 1e-13 * y.abs().max().item()
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("quad")
    rows = [0, 511, 1023]
    Ks, us = K[rows].cpu().numpy(), u[rows].cpu().numpy()
    args = ("thermal", "quad", 2, coords, conn, Ks, us, loss.dirichlet_indices, loss.dirichlet_values,
            {"beta": 2.0, "c": 4})
    gU, gK = assembly.batch_loss_grads(*args)                                  # scaled by 1/3 (its own batch)
    _, _, Eb = assembly.batch_loss(*args)
    gu, gk = a[3][rows].cpu().numpy() * B / 3.0, a[4][rows].cpu().numpy() * B / 3.0
    assert np.abs(gu - gU).max() <= 1e-12 * np.abs(gU).max()
    assert np.abs(gk - gK).max() <= 1e-12 * np.abs(gK).max()
    energies = loss._energy_and_grads(K, loss.GetFullDofVector(None, u))[0][rows].cpu().numpy()
    assert np.abs(energies - Eb).max() <= 1e-12 * np.abs(Eb).max()


def test_j2_config5_full_size_per_gpu():
    """configs[4] at the per-GPU size of the 8-GPU partition (128^3 Hex8 with (ne, 8, 7) history, two load steps):
    sampled elements against the oracle's literal 7-unknown replay (Ke, residual contributions, new history), the elastic
    points' history bit-identical, yield consistency of every plastic point (sigma_eq = y(xi_new) within the Newton
    tolerance), run-to-run bit-identity, tuned == generic kernel on the sample."""
    from folax_b200.loss_functions import ElastoplasticityLoss3DHexa
    from oracle import j2
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs a large-memory GPU")
    mat = {"young_modulus": 3.0, "poisson_ratio": 0.3, "iso_hardening_parameter_1": 0.4, "iso_hardening_param_2": 10.0,
           "yield_limit": 0.2}
    mesh = folax_b200.perturb_interior_nodes(folax_b200.create_3D_box_mesh(N, N, N, 1.0, 1.0, 1.0), 0.1, 2)
    bc = {d: {"left": 0.0, "right": 0.1} for d in ("Ux", "Uy", "Uz")}
    loss = ElastoplasticityLoss3DHexa("c5", {"dirichlet_bc_dict": bc, "num_gp": 2, "material_dict": mat}, mesh)
    loss.Initialize()
    g = torch.Generator(device="cuda").manual_seed(7)
    K = torch.ones(loss._nn, dtype=torch.float64, device="cuda")
    u1 = (0.02 / N) * torch.randn(loss.total_number_of_dofs, generator=g, device="cuda", dtype=torch.float64)
    st0 = torch.zeros(loss.GetStateShape(), dtype=torch.float64, device="cuda")
    st1, _, _ = loss.ComputeJacobianMatrixAndResidualVector(K, u1, st0)
    u2 = 2.0 * u1
    st2, jac, R = loss.ComputeJacobianMatrixAndResidualVector(K, u2, st1)
    st2b, jac_b, Rb = loss.ComputeJacobianMatrixAndResidualVector(K, u2, st1)
    assert torch.equal(jac.data, jac_b.data) and torch.equal(R, Rb) and torch.equal(st2, st2b)     # deterministic
    del jac_b
    plastic = st2[..., -1] > st1[..., -1]
    frac = float(plastic.double().mean())
    assert 0.2 < frac < 1.0, frac
    assert torch.equal(st2[~plastic], st1[~plastic])                      # elastic points: history untouched, bit for bit
    # sampled elements against the oracle
    rng = np.random.default_rng(3)
    conn = mesh.GetElementsNodes("hexahedron")
    sample = np.unique(np.concatenate([rng.integers(0, len(conn), 60), [0, len(conn) - 1], np.arange(0, len(conn), N)[:6]]))
    coords = np.asarray(mesh.GetNodesCoordinates())
    dofs = assembly.element_dof_ids(conn[sample], 3)
    uh = u2.cpu().numpy()
    sidx = torch.as_tensor(sample, device="cuda")
    _, st_ref, re_ref, Ke_ref = j2.j2_element("hexahedron", 2, coords[conn[sample]], uh[dofs], st1[sidx].cpu().numpy(),
                                              mat["young_modulus"], mat["poisson_ratio"], mat["yield_limit"],
                                              mat["iso_hardening_parameter_1"], mat["iso_hardening_param_2"])
    bc_vec = np.ones(loss.total_number_of_dofs)
    bc_vec[loss.dirichlet_indices] = 0.0
    _, Ke_ref = assembly.apply_dirichlet(re_ref, Ke_ref, bc_vec[dofs], False)
    got = jac.data.view(-1, 576)[sidx].cpu().numpy().reshape(Ke_ref.shape)
    assert np.abs(got - Ke_ref).max() <= 1e-11 * np.abs(Ke_ref).max()
    assert np.abs(st2[sidx].cpu().numpy() - st_ref).max() <= 1e-11 * np.abs(st_ref).max()
    # tuned == generic kernel on a slice of elements (the generic kernel at 128^3 takes a while: 4096 elements)
    lib = _lib.load()
    sl = slice(N * N * 5, N * N * 5 + 4096)
    sub = folax_b200.Mesh("", ".")
    sub.node_ids, sub.nodes_coordinates = np.arange(len(coords)), coords
    sub.elements_nodes = {"hexahedron": conn[sl]}
    sub.node_sets = mesh.node_sets
    lsub = ElastoplasticityLoss3DHexa("c5s", {"dirichlet_bc_dict": bc, "num_gp": 2, "material_dict": mat}, sub)
    lsub.Initialize()
    prev = lib.fol_set_tuned_kernels(0)
    try:
        st_g, jac_g, _ = lsub.ComputeJacobianMatrixAndResidualVector(K, u2, st1[sl])
    finally:
        lib.fol_set_tuned_kernels(prev)
    blk = jac.data.view(-1, 576)[sl]
    assert (blk.reshape(-1) - jac_g.data).abs().max().item() <= 1e-12 * jac_g.data.abs().max().item()
    assert (st2[sl] - st_g).abs().max().item() <= 1e-13 * max(st_g.abs().max().item(), 1e-300)
    # yield consistency of the kernel's own output on the sampled elements: sigma_eq(eps - eps_p_new) = y(xi_new) within
    # the Newton stop tolerance (1e-6 on the residual norm) at every plastic point
    del jac
    from oracle.geometry import ELEMENTS, point_data
    from oracle.losses import b_matrix
    _, gradN, _, _ = point_data(ELEMENTS["hexahedron"], coords[conn[sample]], 2)
    eps = np.einsum("egvn,en->egv", b_matrix(gradN), uh[dofs])                      # (ns, 8, 6), engineering shears
    st_new = st2[sidx].cpu().numpy()
    ee = eps - st_new[..., :6]                                                       # tensor components, shears unhalved
    G = mat["young_modulus"] / (2.0 * (1.0 + mat["poisson_ratio"]))
    dev = ee.copy()
    dev[..., :3] -= ee[..., :3].mean(axis=-1, keepdims=True)
    s_dev = 2.0 * G * dev
    sig_eq = np.sqrt(1.5 * ((s_dev[..., :3] ** 2).sum(-1) + 2.0 * (s_dev[..., 3:] ** 2).sum(-1)))
    y_new = mat["yield_limit"] + mat["iso_hardening_parameter_1"] * (1.0 - np.exp(-mat["iso_hardening_param_2"] * st_new[..., 6]))
    pl = st_new[..., 6] > st1[sidx].cpu().numpy()[..., 6]
    assert pl.sum() > 50
    assert np.abs(sig_eq[pl] - y_new[pl]).max() <= 2e-6
    assert (sig_eq[~pl] <= y_new[~pl] + 1e-12).all()                                 # elastic points are inside the surface
