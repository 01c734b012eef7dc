"""The reference's own single-element known-answer tests, run through the CUDA kernels
(ComputeElement -> fol_assemble_elements on a one-element mesh)."""
import numpy as np
import pytest

import folax_b200
from folax_b200 import loss_functions as lf

pytestmark = pytest.mark.gpu

MECH = "tests/unit/test_mechanical_loss.py"
NH = "tests/unit/test_neo_hooke_mechanical_loss.py"
SA = "tests/unit/test_sensitivity_analysis.py"

CASES = [("test_tetra", "tetra", "tet_points_coordinates", [1, 2, 3]),
         ("test_hexa", "hexahedron", "hex_points_coordinates", [1, 2, 3]),
         ("test_quad", "quad", "quad_points_coordinates", [1, 2])]
MECH_CLS = {"tetra": lf.MechanicalLoss3DTetra, "hexahedron": lf.MechanicalLoss3DHexa, "quad": lf.MechanicalLoss2DQuad}
NH_CLS = {"tetra": lf.NeoHookeMechanicalLoss3DTetra, "hexahedron": lf.NeoHookeMechanicalLoss3DHexa,
          "quad": lf.NeoHookeMechanicalLoss2DQuad}


def _one_element_mesh(etype, coords):
    m = folax_b200.Mesh("", ".")
    m.node_ids = np.arange(len(coords))
    m.nodes_coordinates = np.asarray(coords, float)
    m.elements_nodes = {etype: m.node_ids.reshape(1, -1)}
    return m


def _settings(body):
    dofs = ["Ux", "Uy", "Uz"][: len(body)]
    return {"dirichlet_bc_dict": {d: {} for d in dofs}, "material_dict": {"young_modulus": 1, "poisson_ratio": 0.3},
            "body_foce": np.array(body).reshape(-1, 1)}


@pytest.mark.parametrize("test,etype,coords,body", CASES)
def test_mechanical_reference_goldens(goldens, test, etype, coords, body):
    """test_mechanical_loss.py:16-123 with the reference's own tolerances."""
    rec = goldens[MECH][test]
    X = rec["assign"][coords]
    loss = MECH_CLS[etype]("mechanical_loss", _settings(body), _one_element_mesh(etype, X))
    loss.Initialize()
    a = len(X)
    nd = a * len(body)
    en, re, ke = loss.ComputeElement(np.array(X), np.ones(a), np.ones((nd, 1)))
    k, r = rec["asserts"]
    np.testing.assert_allclose(ke.cpu().numpy(), np.array(k["value"]), rtol=k["rtol"], atol=k["atol"])
    np.testing.assert_allclose(re.cpu().numpy().flatten(), np.array(r["value"]), rtol=r["rtol"], atol=r["atol"])
    assert np.isclose(en.item(), re.cpu().numpy().sum())


@pytest.mark.parametrize("test,etype,coords,body", CASES)
def test_neo_hooke_reference_goldens_f64(goldens, test, etype, coords, body):
    """test_neo_hooke_mechanical_loss.py:16-190: 19-digit goldens, checked to 1e-12."""
    rec = goldens[NH][test]
    X = rec["assign"][coords]
    loss = NH_CLS[etype]("nh_loss", _settings(body), _one_element_mesh(etype, X))
    loss.Initialize()
    a = len(X)
    nd = a * len(body)
    en, re, ke = loss.ComputeElement(np.array(X), np.ones(a), np.ones((nd, 1)))
    K_ref, r_ref = np.array(rec["asserts"][0]["value"]), np.array(rec["asserts"][1]["value"])
    assert np.abs(ke.cpu().numpy() - K_ref).max() <= 1e-12 * np.abs(K_ref).max()
    assert np.abs(re.cpu().numpy().flatten() - r_ref).max() <= 1e-12 * np.abs(r_ref).max()
    assert abs(en.item()) <= 1e-14            # F = I: zero strain energy


def test_global_assembly_golden(goldens):
    """test_sensitivity_analysis.py:41-62: transposed, BC-applied global Jacobian row 8."""
    rec = goldens[SA]["test_quad"]
    mesh = folax_b200.create_2D_square_mesh(L=1, N=3)
    bc = {"Ux": {"left": 0.0, "right": 0.05}, "Uy": {"left": 0.0, "right": 0.05}}
    loss = lf.MechanicalLoss2DQuad("m2d", {"dirichlet_bc_dict": bc, "num_gp": 2,
                                            "material_dict": {"young_modulus": 1, "poisson_ratio": 0.3}}, mesh)
    loss.Initialize()
    jac, _ = loss.ComputeJacobianMatrixAndResidualVector(np.array(rec["assign"]["random_K"]),
                                                         np.array(rec["assign"]["random_FE_UV"]),
                                                         transpose_jacobian=True)
    g = rec["asserts"][0]
    np.testing.assert_allclose(jac.todense().cpu().numpy()[8, :], np.array(g["value"]), rtol=g["rtol"], atol=g["atol"])
