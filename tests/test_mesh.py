"""Host-side mesh container and generators (fol/mesh_input_output/mesh.py, usefull_functions.py:196-258)."""
import os

import numpy as np
import pytest

import folax_b200
from oracle import geometry

MESHES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "meshes")


def test_square_mesh_numbering():
    m = folax_b200.create_2D_square_mesh(L=1, N=3)
    assert m.GetNumberOfNodes() == 9 and m.GetNumberOfElements("quad") == 4
    np.testing.assert_array_equal(m.GetElementsNodes("quad"), [[0, 1, 4, 3], [1, 2, 5, 4], [3, 4, 7, 6], [4, 5, 8, 7]])
    np.testing.assert_array_equal(m.GetNodeSet("left"), [0, 3, 6])
    np.testing.assert_array_equal(m.GetNodeSet("right"), [2, 5, 8])
    np.testing.assert_allclose(m.GetNodesCoordinates()[4], [0.5, 0.5, 0.0])


@pytest.mark.parametrize("maker,etype", [(folax_b200.create_3D_box_mesh, "hexahedron"),
                                         (folax_b200.create_3D_tetra_box_mesh, "tetra")])
def test_box_meshes_positive_and_fill_volume(maker, etype):
    m = maker(3, 2, 4, 1.5, 1.0, 2.0)
    elem = geometry.ELEMENTS[etype]
    X = m.GetNodesCoordinates()[m.GetElementsNodes(etype)]
    order = 2 if etype == "hexahedron" else 1
    _, _, detJ, w = geometry.point_data(elem, X, order)
    assert (detJ > 0).all()
    np.testing.assert_allclose((detJ * w).sum(), 3.0, rtol=1e-12)
    assert m.GetNumberOfElements(etype) == 24 * (6 if etype == "tetra" else 1)
    assert len(m.GetNodeSet("left")) == 3 * 5 and len(m.GetNodeSet("right")) == 15


def test_mdpa_reader_counts():
    """test_mesh_io.py:26-29: 85 nodes, 249 tets, node-set sizes 66 / 13."""
    m = folax_b200.Mesh("mdpa_io", file_name="coarse_sphere.mdpa", case_dir=MESHES)
    m.Initialize()
    assert m.GetNumberOfNodes() == 85
    assert m.GetNumberOfElements("tetra") == 249
    assert len(m.GetNodeSet("Skin_Part")) == 66
    assert len(m.GetNodeSet("Partial_Skin_Part")) == 13
    X = m.GetNodesCoordinates()[m.GetElementsNodes("tetra")]
    _, _, detJ, _ = geometry.point_data(geometry.ELEMENTS["tetra"], X, 1)
    assert (detJ > 0).all()


def test_orientation_fix():
    m = folax_b200.create_3D_tetra_box_mesh(1, 1, 1, 1, 1, 1)
    conn = m.elements_nodes["tetra"].copy()
    conn[2, [0, 1]] = conn[2, [1, 0]]
    m.elements_nodes["tetra"] = conn.copy()
    m.CheckAndOrientElements()
    assert (m.elements_nodes["tetra"][2] != conn[2]).any()
    X = m.GetNodesCoordinates()[m.GetElementsNodes("tetra")]
    assert (geometry.point_data(geometry.ELEMENTS["tetra"], X, 1)[2] > 0).all()


def test_regression_loss_matches_its_definition():
    """fol/loss_functions/regression_loss.py:69-96: mean / min / max of the squared error over (batch, -1)."""
    import torch
    from folax_b200.loss_functions import RegressionLoss
    mesh = folax_b200.create_2D_square_mesh(1.0, 3)
    loss = RegressionLoss("reg", {"nodal_unknows": ["Ux", "Uy"]}, mesh)
    loss.Initialize()
    assert loss.non_dirichlet_indices.size == 18 and loss.GetFullDofVector(None, 5) == 5
    rng = np.random.default_rng(0)
    gt, pred = rng.standard_normal((4, 9, 2)), rng.standard_normal((4, 18))
    p = torch.tensor(pred, requires_grad=True)
    mean, (mn, mx, mean2) = loss.ComputeBatchLoss(gt, p)
    err = (gt.reshape(4, -1) - pred) ** 2
    assert np.allclose([float(mean), float(mn), float(mx), float(mean2)], [err.mean(), err.min(), err.max(), err.mean()])
    mean.backward()
    assert np.allclose(p.grad.numpy(), -2.0 * (gt.reshape(4, -1) - pred) / err.size)
