"""GPU parity of the J2 elastoplastic element / assembly (return mapping + forward-mode tangent)."""
import numpy as np
import pytest
import torch

import folax_b200
from folax_b200 import loss_functions as lf
from oracle import assembly, j2
from tests import gpu_helpers as H

pytestmark = pytest.mark.gpu

EP = "tests/unit/test_elastoplasticity.py"
MAT = {"young_modulus": 3.0, "poisson_ratio": 0.3, "iso_hardening_parameter_1": 0.4,
       "iso_hardening_param_2": 10.0, "yield_limit": 0.2}


def _one_element_mesh(etype, coords):
    m = folax_b200.Mesh("", ".")
    m.node_ids = np.arange(len(coords))
    m.nodes_coordinates = np.asarray(coords, float)
    m.elements_nodes = {etype: m.node_ids.reshape(1, -1)}
    return m


@pytest.mark.parametrize("test,cls,etype,coords,body,ng,ns", [
    ("test_tetra", lf.ElastoplasticityLoss3DTetra, "tetra", "tet_points_coordinates", [1, 2, 3], 1, 7),
    ("test_quad", lf.ElastoplasticityLoss2DQuad, "quad", "quad_points_coordinates", [1, 2], 4, 4)])
def test_reference_goldens(goldens, test, cls, etype, coords, body, ng, ns):
    """test_elastoplasticity.py:19-160 with its own tolerances."""
    rec = goldens[EP][test]
    X = rec["assign"][coords]
    dofs = ["Ux", "Uy", "Uz"][: len(body)]
    loss = cls("ep", {"dirichlet_bc_dict": {d: {} for d in dofs}, "material_dict": dict(MAT),
                      "body_foce": np.array(body).reshape(-1, 1)}, _one_element_mesh(etype, X))
    loss.Initialize()
    nd = len(X) * len(body)
    en, st, re, ke = loss.ComputeElement(np.array(X), np.array([1.0]), np.ones((nd, 1)), np.zeros((ng, ns)))
    k, r = rec["asserts"]
    np.testing.assert_allclose(ke.cpu().numpy(), np.array(k["value"]), rtol=k["rtol"] or 1e-5, atol=k["atol"] or 1e-6)
    np.testing.assert_allclose(re.cpu().numpy().flatten(), np.array(r["value"]), rtol=r["rtol"] or 1e-5,
                               atol=r["atol"] or 1e-6)
    assert st.shape == (ng, ns) and not st.any()


@pytest.mark.parametrize("cls,etype,n", [(lf.ElastoplasticityLoss3DTetra, "tetra", 2),
                                          (lf.ElastoplasticityLoss2DQuad, "quad", 4),
                                          (lf.ElastoplasticityLoss3DHexa, "hexahedron", 2)])
def test_plastic_assembly_matches_oracle(cls, etype, n):
    """Two load steps: mixed elastic / plastic points with non-zero history on the second."""
    mesh = H.make_mesh(etype, n, seed=4)
    dofs = H.dofs_of("mechanical", etype)
    loss = cls("ep", {"dirichlet_bc_dict": {d: {"left": 0.0, "right": 0.1} for d in dofs},
                      "material_dict": dict(MAT)}, mesh)
    loss.Initialize()
    rng = np.random.default_rng(9)
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes(etype)
    state = np.zeros(loss.GetStateShape())
    K = np.ones(mesh.GetNumberOfNodes())
    u = np.zeros(loss.total_number_of_dofs)
    plastic_seen = 0
    for step in range(2):
        u = u + 0.12 * rng.standard_normal(u.shape) * (1.0 / n)
        new_state, jac, R = loss.ComputeJacobianMatrixAndResidualVector(K, u, state, transpose_jacobian=bool(step))
        ref_state, data, idx, Rref = assembly.assemble_j2(etype, loss.num_gp, coords, conn, u, state,
                                                          loss.dirichlet_indices, MAT, transpose=bool(step))
        assert np.array_equal(jac.indices.cpu().numpy(), idx)
        sc = np.abs(data).max()
        assert np.abs(jac.data.cpu().numpy() - data).max() <= 1e-11 * sc
        assert np.abs(R.cpu().numpy() - Rref).max() <= 1e-11 * max(np.abs(Rref).max(), 1e-300)
        assert np.abs(new_state.cpu().numpy() - ref_state).max() <= 1e-11 * max(np.abs(ref_state).max(), 1e-300)
        plastic_seen += int((ref_state[..., -1] > state[..., -1]).sum())
        state = ref_state
    frac = plastic_seen / (2 * state.shape[0] * state.shape[1])
    assert 0.1 < frac <= 1.0, f"test must exercise the plastic branch (fraction {frac})"
