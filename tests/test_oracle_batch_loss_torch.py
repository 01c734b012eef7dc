"""ComputeBatchLoss, its gradient and the static thermal element have no golden in the reference.  They
are pinned here by two independently written oracles agreeing at <= 1e-12:
  oracle/losses_torch.py   literal transcription of mechanical.py / thermal.py / fe_loss.py in torch float64, gradient
                           by torch.autograd with .detach() where the reference has stop_gradient;
  oracle/assembly.py       NumPy restatement with closed-form cotangents (what the GPU parity tests compare against).
Runs without a GPU."""
import numpy as np
import pytest

import folax_b200
from oracle import assembly, losses_torch
from tests import gpu_helpers as H


def _mesh(etype):
    if etype == "quad":
        m = folax_b200.create_2D_square_mesh(1.0, 4)
    elif etype == "hexahedron":
        m = folax_b200.create_3D_box_mesh(2, 2, 2, 1.0, 1.1, 0.9)
    else:
        m = folax_b200.create_3D_tetra_box_mesh(2, 1, 1, 1.0, 1.0, 1.0)
    return folax_b200.perturb_interior_nodes(m, 0.2, 7)


@pytest.mark.parametrize("physics,etype,num_gp,params,exponent", [
    ("thermal", "quad", 2, {"beta": 2.0, "c": 4}, 1.0),
    ("thermal", "quad", 2, {"beta": 0.0, "c": 1}, 2.0),
    ("thermal", "hexahedron", 2, {"beta": 1.5, "c": 2}, 1.0),
    ("thermal", "tetra", 1, {"beta": 0.5, "c": 3}, 1.0),
    ("mechanical", "quad", 2, {"young_modulus": 1.0, "poisson_ratio": 0.3, "body_force": np.array([0.2, -0.1])}, 1.0),
    ("mechanical", "hexahedron", 2, {"young_modulus": 2.0, "poisson_ratio": 0.25}, 1.0),
    ("mechanical", "tetra", 1, {"young_modulus": 1.0, "poisson_ratio": 0.3, "body_force": np.array([0.1, 0.2, 0.3])}, 2.0)])
def test_two_oracles_agree(physics, etype, num_gp, params, exponent):
    mesh = _mesh(etype)
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes(etype)
    dofs = ["T"] if physics == "thermal" else H.dofs_of(physics, etype)
    bc = {d: ({"left": 1.0, "right": 0.1} if physics == "thermal" else {"left": 0.0, "right": 0.1}) for d in dofs}
    didx, dval = assembly.dirichlet_vectors(dofs, bc, mesh.node_sets)
    rng = np.random.default_rng(5)
    B, nn = 3, len(coords)
    K = rng.uniform(0.1, 1.0, (B, nn))
    U = rng.uniform(0.1, 1.0, (B, nn * len(dofs))) if physics == "thermal" else 0.05 * rng.standard_normal((B, nn * len(dofs)))
    mean_t, E_t, gU_t, gK_t = losses_torch.batch_loss_and_grads(physics, etype, num_gp, coords, conn, K, U, didx, dval,
                                                               params, exponent)
    mean_n, _, E_n = assembly.batch_loss(physics, etype, num_gp, coords, conn, K, U, didx, dval, params, exponent)
    gU_n, gK_n = assembly.batch_loss_grads(physics, etype, num_gp, coords, conn, K, U, didx, dval, params, exponent)
    assert abs(mean_t - mean_n) <= 1e-12 * abs(mean_n)
    assert np.abs(E_t - E_n).max() <= 1e-12 * np.abs(E_n).max()
    assert np.abs(gU_t - gU_n).max() <= 1e-12 * np.abs(gU_n).max()
    assert not gU_t[:, didx].any()                                  # the cotangent is cut at the overwritten entries
    if physics == "thermal":
        assert np.abs(gK_t - gK_n).max() <= 1e-12 * np.abs(gK_n).max()
    else:
        assert not gK_t.any() and not gK_n.any()                    # mechanical.py:116: the residual is stop-gradiented
