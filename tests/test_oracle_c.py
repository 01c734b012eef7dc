"""The C/OpenMP restatement (CPU baseline) against the NumPy oracle and the reference hex golden."""
import numpy as np

import folax_b200
from oracle import assembly, c_oracle


def test_c_oracle_matches_numpy_oracle():
    mesh = folax_b200.perturb_interior_nodes(folax_b200.create_3D_box_mesh(4, 3, 5, 1.0, 1.0, 1.0), 0.2, 1)
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("hexahedron")
    rng = np.random.default_rng(0)
    K, u = rng.uniform(0.1, 1, len(coords)), 0.01 * rng.standard_normal(3 * len(coords))
    didx, _ = assembly.dirichlet_vectors(["Ux", "Uy", "Uz"], {d: {"left": 0.0, "right": 0.1} for d in
                                                               ("Ux", "Uy", "Uz")}, mesh.node_sets)
    for transpose in (False, True):
        data, _, R = assembly.assemble("mechanical", "hexahedron", 2, coords, conn, K, u, didx,
                                       {"young_modulus": 1.3, "poisson_ratio": 0.3, "body_force": [0.1, 0.2, 0.3]},
                                       transpose)
        dc, Rc = c_oracle.hex_mech_assemble(coords, conn, K, u, didx, 1.3, 0.3, [0.1, 0.2, 0.3], transpose)
        assert np.abs(dc - data).max() <= 1e-13 * np.abs(data).max()
        assert np.abs(Rc - R).max() <= 1e-12 * np.abs(R).max()


def test_c_oracle_reference_golden(goldens):
    rec = goldens["tests/unit/test_neo_hooke_mechanical_loss.py"]["test_hexa"]
    X = np.array(rec["assign"]["hex_points_coordinates"], float)
    data, R = c_oracle.hex_mech_assemble(X, np.arange(8)[None], np.ones(8), np.ones(24), [], 1.0, 0.3, [1, 2, 3])
    K_ref = np.array(rec["asserts"][0]["value"])
    assert np.abs(data.reshape(24, 24) - K_ref).max() <= 1e-12 * np.abs(K_ref).max()
    assert np.abs(R - np.array(rec["asserts"][1]["value"])).max() <= 1e-12


def test_c_thermal_batch_loss_matches_numpy_oracle():
    """oracle/c/quad_thermal_loss.c (CPU baseline of the FOL loss+grad metric) against the NumPy oracle."""
    mesh = folax_b200.perturb_interior_nodes(folax_b200.create_2D_square_mesh(1.0, 9), 0.2, 3)
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("quad")
    rng = np.random.default_rng(2)
    nb, nn = 5, len(coords)
    K, u = rng.uniform(0.1, 1, (nb, nn)), rng.uniform(0.1, 1, (nb, nn))
    didx, dval = assembly.dirichlet_vectors(["T"], {"T": {"left": 1.0, "right": 0.1}}, mesh.node_sets)
    for beta, c in ((0.0, 1.0), (2.0, 4.0)):
        par = {"beta": beta, "c": c}
        _, _, Eb = assembly.batch_loss("thermal", "quad", 2, coords, conn, K, u, didx, dval, par)
        gU, gK = assembly.batch_loss_grads("thermal", "quad", 2, coords, conn, K, u, didx, dval, par)
        U = assembly.full_dof_vector(u, didx, dval)
        E_c, gU_c, gK_c = c_oracle.quad_thermal_batch_loss_grads(coords, conn, K, U, beta, c)
        assert np.abs(E_c - Eb).max() <= 1e-13 * np.abs(Eb).max()
        gU_c = gU_c / nb
        gU_c[:, didx] = 0.0                       # the mean's 1/B and the Dirichlet cut (fe_loss.py:262, 91-92)
        assert np.abs(gU_c - gU).max() <= 1e-13 * np.abs(gU).max()
        assert np.abs(gK_c / nb - gK).max() <= 1e-13 * np.abs(gK).max()
