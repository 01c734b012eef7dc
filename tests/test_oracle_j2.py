"""J2 plasticity oracle: elastic branch against the reference goldens, plastic branch against finite
differences and yield consistency (no reference golden exists for it)."""
import numpy as np
import pytest

from oracle import j2

EP = "tests/unit/test_elastoplasticity.py"
MAT = dict(E=3.0, nu=0.3, y0=0.2, h1=0.4, h2=10.0)


@pytest.mark.parametrize("test,etype,num_gp,coords,body,ns", [
    ("test_tetra", "tetra", 1, "tet_points_coordinates", [1, 2, 3], 7),
    ("test_quad", "quad", 2, "quad_points_coordinates", [1, 2], 4)])
def test_elastic_branch_goldens(goldens, test, etype, num_gp, coords, body, ns):
    """test_elastoplasticity.py:70-90, 144-160 (zero state, rigid translation; float32 goldens)."""
    rec = goldens[EP][test]
    X = np.array(rec["assign"][coords], float)[None]
    a = X.shape[1]
    d = len(body)
    ng = 1 if etype == "tetra" else 4
    en, st, re, Ke = j2.j2_element(etype, num_gp, X, np.ones((1, a * d)), np.zeros((1, ng, ns)), body=body, **MAT)
    k, r = rec["asserts"]
    np.testing.assert_allclose(Ke[0], np.array(k["value"]), rtol=k["rtol"] or 1e-5, atol=k["atol"] or 1e-6)
    np.testing.assert_allclose(re[0], np.array(r["value"]), rtol=r["rtol"] or 1e-5, atol=r["atol"] or 1e-6)
    assert not st.any()


def _stress_only(eps, state, dim):
    return j2.j2_point(eps, state, dim=dim, tol=1e-13, **MAT)[0]


@pytest.mark.parametrize("dim", [3, 2])
def test_plastic_branch_consistency(dim):
    rng = np.random.default_rng(3)
    V = 6 if dim == 3 else 3
    hits = 0
    for trial in range(12):
        eps = rng.standard_normal(V) * 0.15
        state = np.zeros(V + 1)
        if trial % 2:
            state[:V] = rng.standard_normal(V) * 0.01
            if dim == 3:
                state[:3] -= state[:3].mean()          # plastic strain is deviatoric
            state[-1] = 0.02
        sig, tan, st = j2.j2_point(eps, state, dim=dim, **MAT)
        if st[-1] == state[-1]:
            continue                                    # elastic point
        hits += 1
        assert st[-1] > state[-1]
        # tangent vs central finite differences of the (tightly converged) stress update
        h = 1e-6
        fd = np.zeros((V, V))
        for k in range(V):
            ep, em = eps.copy(), eps.copy()
            ep[k] += h
            em[k] -= h
            fd[:, k] = (_stress_only(ep, state, dim) - _stress_only(em, state, dim)) / (2 * h)
        assert np.abs(fd - tan).max() <= 2e-4 * np.abs(tan).max()
        # yield consistency of the returned stress (3-D check: sigma_eq = y(xi_new) within the Newton tol)
        if dim == 3:
            S = np.array([[sig[0], sig[3], sig[5]], [sig[3], sig[1], sig[4]], [sig[5], sig[4], sig[2]]])
            s = S - np.trace(S) / 3 * np.eye(3)
            q = np.sqrt(1.5) * np.linalg.norm(s)
            assert abs(q - (MAT["y0"] + MAT["h1"] * (1 - np.exp(-MAT["h2"] * st[-1])))) <= 1e-5
    assert hits >= 4


def test_elastic_tangent_has_double_shear_stiffness():
    """engineering shears enter the tensor unhalved, so the shear stiffness is 2G."""
    sig, tan, _ = j2.j2_point(np.array([0, 0, 0, 1e-4, 0, 0.0]), np.zeros(7), dim=3, **MAT)
    G = MAT["E"] / (2 * (1 + MAT["nu"]))
    assert np.isclose(tan[3, 3], 2 * G) and np.isclose(sig[3], 2 * G * 1e-4)
