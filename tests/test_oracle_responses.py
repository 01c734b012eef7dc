"""The response / adjoint-sensitivity oracle (oracle/responses.py) against the reference's own known answers:
tests/unit/test_sensitivity_analysis.py:54-80 (explicit K, u, adjoint vector on a 3x3-node quad mesh) and
tests/integration/test_mechanical_2D_sa.py:81-113 (FE solve + adjoint solve + derivatives, 5x5 nodes)."""
import json
import os

import numpy as np
import pytest

from oracle import assembly, responses
from tests.test_oracle_golden import _square_mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BC = {"Ux": {"left": 0.0, "right": 0.05}, "Uy": {"left": 0.0, "right": 0.05}}
MAT = {"young_modulus": 1.0, "poisson_ratio": 0.3}


@pytest.fixture(scope="module")
def goldens():
    with open(os.path.join(ROOT, "tests", "golden", "reference_unit_goldens.json")) as fh:
        return json.load(fh)


def test_unit_sensitivity_goldens(goldens):
    rec = goldens["tests/unit/test_sensitivity_analysis.py"]["test_quad"]
    coords, conn, sets = _square_mesh(3)
    u = np.array(rec["assign"]["random_FE_UV"])
    K = np.array(rec["assign"]["random_K"])
    lam = np.array(rec["assign"]["random_adj_FE_UV"])
    didx, _ = assembly.dirichlet_vectors(["Ux", "Uy"], BC, sets)
    f = responses.response_function("(E**2)*U[0]", "E", "Ux")
    data, idx, rhs = responses.adjoint_jacobian_and_rhs(f, "mechanical", "quad", 2, coords, conn, K, u, didx, MAT)
    a = rec["asserts"]
    np.testing.assert_allclose(assembly.to_dense(data, idx, 18)[8], a[0]["value"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(rhs, a[1]["value"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(responses.control_derivatives(f, "mechanical", "quad", 2, coords, conn, K, u, lam, MAT),
                               a[2]["value"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(responses.shape_derivatives(f, "mechanical", "quad", 2, coords, conn, K, u, lam, MAT),
                               a[3]["value"], rtol=1e-5, atol=1e-5)


def test_integration_sensitivity_goldens(goldens):
    rec = goldens["tests/integration/test_mechanical_2D_sa.py"]
    K = np.array(rec["setUp"]["assign"]["random_K"])
    coords, conn, sets = _square_mesh(5)
    didx, dval = assembly.dirichlet_vectors(["Ux", "Uy"], BC, sets)
    ndof = 2 * coords.shape[0]
    u0 = assembly.full_dof_vector(np.zeros((1, ndof)), didx, dval)[0]
    data, idx, R = assembly.assemble("mechanical", "quad", 2, coords, conn, K, u0, didx, MAT)
    u = u0 + np.linalg.solve(assembly.to_dense(data, idx, ndof), -R)
    f = responses.response_function("(E**2)*U[0]", "E", "Ux")
    lam = responses.adjoint_solve(f, "mechanical", "quad", 2, coords, conn, K, u, didx, MAT)
    a = rec["test_sensitivites"]["asserts"]
    cd = responses.control_derivatives(f, "mechanical", "quad", 2, coords, conn, K, u, lam, MAT)
    sd = responses.shape_derivatives(f, "mechanical", "quad", 2, coords, conn, K, u, lam, MAT)
    np.testing.assert_allclose(cd, a[0]["value"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(sd, a[1]["value"], rtol=1e-5, atol=1e-5)
    # tighter than the reference's own tolerance: 7 printed digits of a float32 run
    assert np.abs(cd - np.array(a[0]["value"])).max() <= 2e-9
    assert np.abs(sd - np.array(a[1]["value"])).max() <= 2e-9


def test_complex_step_against_central_differences():
    """The differentiation route of the oracle itself: complex step vs central differences of the element value."""
    rng = np.random.default_rng(1)
    coords, conn, _ = _square_mesh(4)
    coords = coords + np.concatenate([0.03 * rng.standard_normal((16, 2)), np.zeros((16, 1))], axis=1)
    K, u, lam = rng.uniform(0.2, 1, 16), rng.standard_normal(32), rng.standard_normal(32)
    g = assembly.element_dof_ids(conn, 2)
    X, de, ue, le = coords[conn], K[conn], u[g], lam[g]
    rK, rX = responses.residual_adjoint_grads("mechanical", "quad", 2, X, de, ue, le, MAT)
    h = 1e-6

    def phi(Xp, dep):
        return np.einsum("en,en->e", le, responses._element_residual("mechanical", "quad", 2, Xp, dep, ue, MAT))

    for k in range(4):
        dp, dm = de.copy(), de.copy()
        dp[:, k] += h
        dm[:, k] -= h
        assert np.abs((phi(X, dp) - phi(X, dm)) / (2 * h) - rK[:, k]).max() <= 1e-8
        for c in range(2):
            Xp, Xm = X.copy(), X.copy()
            Xp[:, k, c] += h
            Xm[:, k, c] -= h
            assert np.abs((phi(Xp, de) - phi(Xm, de)) / (2 * h) - rX[:, 3 * k + c]).max() <= 1e-7
    assert np.all(rX[:, 2::3] == 0.0)
