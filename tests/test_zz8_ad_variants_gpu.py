"""GPU parity of the *_AD.py loss variants (mechanical_neohooke_AD.py, mechanical_saint_venant_AD.py): the reference's
own 19-digit goldens through ComputeElement, mesh assembly against the oracle (complex-step Jacobian of the AD
residual), the reference's AD == analytic assertion at F = I, and the batched-loss behaviour."""
import json
import os

import numpy as np
import pytest
import torch

import folax_b200
from folax_b200.loss_functions import mechanical_neohooke_AD as nh_ad
from folax_b200.loss_functions import mechanical_saint_venant as sv
from folax_b200.loss_functions import mechanical_saint_venant_AD as sv_ad
from oracle import assembly, losses
from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _one_element_mesh(etype, coords):
    m = folax_b200.Mesh("", ".")
    m.node_ids = np.arange(len(coords))
    m.nodes_coordinates = np.asarray(coords, float)
    m.elements_nodes = {etype: m.node_ids.reshape(1, -1)}
    return m


@pytest.mark.parametrize("test,cls,etype,ckey", [
    ("test_tetra", nh_ad.NeoHookeMechanicalLoss3DTetra, "tetra", "tet_points_coordinates"),
    ("test_hexa", nh_ad.NeoHookeMechanicalLoss3DHexa, "hexahedron", "hex_points_coordinates"),
    ("test_quad", nh_ad.NeoHookeMechanicalLoss2DQuad, "quad", "quad_points_coordinates")])
def test_reference_ad_goldens(test, cls, etype, ckey):
    """tests/unit/test_neo_hooke_mechanical_loss_AD.py:21-190, same calls; float64 here, so far tighter than the
    reference's own tolerance."""
    with open(os.path.join(ROOT, "tests", "golden", "reference_unit_goldens.json")) as fh:
        rec = json.load(fh)["tests/unit/test_neo_hooke_mechanical_loss_AD.py"][test]
    X = np.array(rec["assign"][ckey], float)
    a = len(X)
    dofs = ["Ux", "Uy", "Uz"][:3 if etype != "quad" else 2]
    d = len(dofs)
    loss = cls("mechanical_loss", {"dirichlet_bc_dict": {k: {} for k in dofs},
                                   "material_dict": {"young_modulus": 1, "poisson_ratio": 0.3},
                                   "body_foce": np.array([[1], [2], [3]][:d])}, _one_element_mesh(etype, X))
    loss.Initialize()
    en, re, ke = loss.ComputeElement(X, np.ones(a), np.ones((a * d, 1)))
    K_ref, r_ref = np.array(rec["asserts"][0]["value"]), np.array(rec["asserts"][1]["value"])
    assert np.abs(ke.cpu().numpy() - K_ref).max() <= 1e-12 * np.abs(K_ref).max()
    assert np.abs(re.cpu().numpy().reshape(-1) - r_ref).max() <= 1e-12 * max(np.abs(r_ref).max(), 1.0)
    en_ref = losses.ad_variant_element(etype, loss.num_gp, X[None], np.ones((1, a)), np.ones((1, a * d)), 0.3,
                                       np.array([1.0, 2.0, 3.0][:d]), law="neohooke_ad")[0][0]
    # u = ones => F = I and the true energy is 0: en_ref is rounding noise (1e-19..1e-18), so the tolerance is
    # scaled by the size of the terms that cancel (|u|^T |Ke| |u|, |u|^T |re|), never by the noise-level reference
    scale = max(abs(en_ref), np.abs(K_ref).sum(), np.abs(r_ref).sum(), 1.0)
    assert abs(float(en) - en_ref) <= 1e-12 * scale


CLASSES = {("neohooke_ad", "hexahedron"): nh_ad.NeoHookeMechanicalLoss3DHexa, ("neohooke_ad", "tetra"): nh_ad.NeoHookeMechanicalLoss3DTetra,
           ("neohooke_ad", "quad"): nh_ad.NeoHookeMechanicalLoss2DQuad, ("neohooke_ad", "triangle"): nh_ad.NeoHookeMechanicalLoss2DTri,
           ("stvenant_ad", "hexahedron"): sv_ad.SaintVenantMechanicalLoss3DHexa, ("stvenant_ad", "tetra"): sv_ad.SaintVenantMechanicalLoss3DTetra,
           ("stvenant_ad", "quad"): sv_ad.SaintVenantMechanicalLoss2DQuad, ("stvenant_ad", "triangle"): sv_ad.SaintVenantMechanicalLoss2DTri}


@pytest.mark.parametrize("law", ["neohooke_ad", "stvenant_ad"])
@pytest.mark.parametrize("etype", ["hexahedron", "tetra", "quad", "triangle"])
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-11), ("float32", 1e-5)])
def test_mesh_assembly_against_oracle(law, etype, dtype, tol):
    mesh = gh.make_mesh(etype, 3, perturb=0.2, seed=4)
    dofs = gh.dofs_of("mechanical", etype)
    d = len(dofs)
    body = [0.2, -0.4, 0.7][:d]
    loss = CLASSES[(law, etype)]("ad", {"dirichlet_bc_dict": {k: {"left": 0.0, "right": 0.1} for k in dofs},
                                        "material_dict": {"young_modulus": 1.0, "poisson_ratio": 0.3},
                                        "body_foce": body, "dtype": dtype}, mesh)
    loss.Initialize()
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes(etype)
    nn = len(coords)
    rng = np.random.default_rng(8)
    K, u = rng.uniform(0.2, 1.0, nn), 0.03 * rng.standard_normal(nn * d)
    if dtype == "float32":
        # north_star's float32 bar (1e-5) is about the ARITHMETIC: the oracle gets the inputs the kernel gets, i.e. the
        # float32 roundings of coordinates, controls and dofs (measured 3-4e-7; against un-rounded inputs the input
        # rounding alone is 1e-4 for these finite-strain laws, which is what the earlier 2e-4 tolerance absorbed)
        coords, K, u = (x.astype(np.float32).astype(np.float64) for x in (coords, K, u))
    g = assembly.element_dof_ids(conn, d)
    bc = np.ones(nn * d)
    bc[loss.dirichlet_indices] = 0.0
    en_ref, re_ref, Ke_ref = losses.ad_variant_element(etype, loss.num_gp, coords[conn], K[conn], u[g], 0.3,
                                                       np.array(body), law=law)
    for transpose in (False, True):
        jac, R = loss.ComputeJacobianMatrixAndResidualVector(K, u, transpose_jacobian=transpose)
        re_m, Ke_m = assembly.apply_dirichlet(re_ref, Ke_ref, bc[g], transpose)
        R_ref = np.zeros(nn * d)
        np.add.at(R_ref, g.reshape(-1), re_m.reshape(-1))
        assert np.array_equal(jac.indices.cpu().numpy(), assembly.bcoo_indices(conn, d))
        assert np.abs(jac.data.cpu().numpy() - Ke_m.reshape(-1)).max() <= tol * np.abs(Ke_m).max()
        assert np.abs(R.cpu().numpy() - R_ref).max() <= 10 * tol * np.abs(R_ref).max()
    assert abs(float(loss.ComputeTotalEnergy(K, u)) - en_ref.sum()) <= 10 * tol * abs(en_ref.sum())


def test_saint_venant_ad_vs_analytic():
    """test_saint_venant_mechanical_loss.py:21-42 (u = ones => F = I: equal), and the batched loss, which
    differentiates the ENERGY (the same function in both classes): identical values and gradients."""
    X = np.array([[0.1, 0.1, 0.1], [0.28739360416666665, 0.27808503701741405, 0.05672979583333333],
                  [0.0, 1.0, 0.0], [0.0, 1.0, 0.1]])
    settings = {"dirichlet_bc_dict": {"Ux": {}, "Uy": {}, "Uz": {}},
                "material_dict": {"young_modulus": 1, "poisson_ratio": 0.3}, "body_foce": np.array([[1], [2], [3]])}
    a = sv.SaintVenantMechanicalLoss3DTetra("a", dict(settings), _one_element_mesh("tetra", X))
    b = sv_ad.SaintVenantMechanicalLoss3DTetra("b", dict(settings), _one_element_mesh("tetra", X))
    a.Initialize()
    b.Initialize()
    en, re, ke = a.ComputeElement(X, np.ones(4), np.ones((12, 1)))
    en_ad, re_ad, ke_ad = b.ComputeElement(X, np.ones(4), np.ones((12, 1)))
    np.testing.assert_allclose(ke.cpu().numpy(), ke_ad.cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(re.cpu().numpy(), re_ad.cpu().numpy(), rtol=1e-5, atol=1e-6)

    mesh = gh.make_mesh("quad", 4, seed=3)
    bcd = {k: {"left": 0.0, "right": 0.1} for k in ("Ux", "Uy")}
    la = sv.SaintVenantMechanicalLoss2DQuad("a", {"dirichlet_bc_dict": bcd, "material_dict": dict(gh.MATERIAL)}, mesh)
    lb = sv_ad.SaintVenantMechanicalLoss2DQuad("b", {"dirichlet_bc_dict": bcd, "material_dict": dict(gh.MATERIAL)}, mesh)
    la.Initialize()
    lb.Initialize()
    K, u = gh.fields("stvenant", mesh, la, seed=1, batch=3)
    out = []
    for L in (la, lb):
        Kt = torch.tensor(K, device="cuda", requires_grad=True)
        ut = torch.tensor(u, device="cuda", requires_grad=True)
        mean, _ = L.ComputeBatchLoss(Kt, ut)
        mean.backward()
        out.append((float(mean), Kt.grad.clone(), ut.grad.clone()))
    assert out[0][0] == out[1][0] and torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])


def test_neo_hooke_ad_batch_loss_is_refused():
    mesh = gh.make_mesh("quad", 3)
    loss = nh_ad.NeoHookeMechanicalLoss2DQuad("n", {"dirichlet_bc_dict": {k: {"left": 0.0} for k in ("Ux", "Uy")},
                                                    "material_dict": dict(gh.MATERIAL)}, mesh)
    loss.Initialize()
    with pytest.raises(NotImplementedError):
        loss.ComputeBatchLoss(np.ones((1, mesh.GetNumberOfNodes())), np.zeros((1, loss.total_number_of_dofs)))
    from folax_b200 import _lib
    with pytest.raises(_lib.FolaxError):                        # no matrix-free product for the AD variants
        loss.ApplyJacobian(np.ones(mesh.GetNumberOfNodes()), np.zeros(loss.total_number_of_dofs),
                           np.ones(loss.total_number_of_dofs))
