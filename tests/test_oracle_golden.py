"""Pin the NumPy oracle against the known-answer arrays of the reference's own unit tests
(tests/golden/reference_unit_goldens.json, extracted from /root/reference/tests/unit by
tests/golden/make_golden.py).  Tolerances are the reference tests' own rtol/atol for float32
goldens, and 1e-12 for the 19-digit (float64-grade) Neo-Hooke goldens."""
import numpy as np
import pytest

from oracle import assembly, geometry, losses

GEO = "tests/unit/test_geometries.py"
MECH = "tests/unit/test_mechanical_loss.py"
NH = "tests/unit/test_neo_hooke_mechanical_loss.py"
SA = "tests/unit/test_sensitivity_analysis.py"


def _check(actual, rec, rtol=None, atol=None):
    np.testing.assert_allclose(np.asarray(actual), np.asarray(rec["value"], float),
                               rtol=rec["rtol"] if rtol is None else rtol,
                               atol=rec["atol"] if atol is None else atol)


@pytest.mark.parametrize("test,etype,coords", [
    ("test_tri2D3", "triangle", "tri_points_coordinates"),
    ("test_quad2D4", "quad", "quad_points_coordinates"),
    ("test_tet3D4", "tetra", "tet_points_coordinates"),
    ("test_hex3D8", "hexahedron", "hex_points_coordinates"),
])
def test_geometry_goldens(goldens, test, etype, coords):
    """test_geometries.py:45-263 -- Gauss rules, N, dN/dxi, J, grad N for orders 1-3."""
    elem = geometry.ELEMENTS[etype]
    X = np.array(goldens[GEO]["setUp"]["assign"][coords], float)
    order, n = 1, 0
    for rec in goldens[GEO][test]["asserts"]:
        expr = rec["expr"]
        val = np.asarray(rec["value"], float)
        if expr == "points":
            order = next(o for o in (1, 2, 3) if len(elem.gauss(o)[0]) == len(val)
                         and np.allclose(elem.gauss(o)[0], val, atol=1e-5))
            _check(elem.gauss(order)[0], rec)
        elif expr == "weights":
            _check(elem.gauss(order)[1], rec)
        else:
            pts = elem.gauss(order)[0]
            if expr.startswith("shape_function_values"):
                act = np.stack([elem.N(p) for p in pts])
            elif expr.startswith("shape_function_grads"):
                act = np.stack([elem.dN(p) for p in pts])
            elif expr.startswith("jacobian"):
                act = np.stack([geometry.jacobian(elem, X, p) for p in pts])
            elif expr == "shape_function_g_grads":
                act = geometry.point_data(elem, X, order)[1]
            else:
                raise AssertionError(expr)
            if expr.endswith("[0]"):
                act = act[0]
            _check(act, rec)
        n += 1
    assert n >= 9


CASES = [("test_tetra", "tetra", 1, "tet_points_coordinates", [1, 2, 3]),
         ("test_hexa", "hexahedron", 2, "hex_points_coordinates", [1, 2, 3]),
         ("test_quad", "quad", 2, "quad_points_coordinates", [1, 2])]


@pytest.mark.parametrize("test,etype,num_gp,coords,body", CASES)
def test_mechanical_goldens(goldens, test, etype, num_gp, coords, body):
    """test_mechanical_loss.py:37-63, 90-94, 119-123 (float32 goldens, E=1, nu=0.3, u=1)."""
    rec = goldens[MECH][test]
    X = np.array(rec["assign"][coords], float)[None]
    a = X.shape[1]
    nd = a * geometry.ELEMENTS[etype].dim
    en, re, Ke = losses.mechanical_element(etype, num_gp, X, np.ones((1, a)), np.ones((1, nd)),
                                           1.0, 0.3, body)
    _check(Ke[0], rec["asserts"][0])
    _check(re[0], rec["asserts"][1])
    assert np.isclose(en[0], np.ones(nd) @ re[0])


@pytest.mark.parametrize("test,etype,num_gp,coords,body", CASES)
def test_neo_hooke_goldens_f64(goldens, test, etype, num_gp, coords, body):
    """test_neo_hooke_mechanical_loss.py:38-66, 93-145, 170-190: 19-digit goldens."""
    rec = goldens[NH][test]
    X = np.array(rec["assign"][coords], float)[None]
    a = X.shape[1]
    nd = a * geometry.ELEMENTS[etype].dim
    _, re, Ke = losses.neo_hooke_element(etype, num_gp, X, np.ones((1, a)), np.ones((1, nd)),
                                         1.0, 0.3, body)
    K_ref = np.asarray(rec["asserts"][0]["value"], float)
    r_ref = np.asarray(rec["asserts"][1]["value"], float)
    assert np.abs(Ke[0] - K_ref).max() <= 1e-12 * np.abs(K_ref).max()
    assert np.abs(re[0] - r_ref).max() <= 1e-12 * np.abs(r_ref).max()


@pytest.mark.parametrize("test,etype,num_gp,coords,body", CASES[:2])
def test_linear_equals_neo_hooke_at_identity_f64(goldens, test, etype, num_gp, coords, body):
    """At u = const (F = I) the Neo-Hooke tangent equals the linear-elastic stiffness in 3-D, so
    the 19-digit goldens pin MechanicalLoss to float64 accuracy as well."""
    rec = goldens[NH][test]
    X = np.array(rec["assign"][coords], float)[None]
    a = X.shape[1]
    _, re, Ke = losses.mechanical_element(etype, num_gp, X, np.ones((1, a)), np.ones((1, 3 * a)),
                                          1.0, 0.3, body)
    K_ref = np.asarray(rec["asserts"][0]["value"], float)
    assert np.abs(Ke[0] - K_ref).max() <= 1e-12 * np.abs(K_ref).max()


def _square_mesh(N, L=1.0):
    # numbering of fol/tools/usefull_functions.py:213-258 (restated for the fixture)
    x = np.linspace(0, L, N)
    Xg, Yg = np.meshgrid(x, x)
    coords = np.stack([Xg.ravel(), Yg.ravel(), np.zeros(N * N)], axis=1)
    Ne = N - 1
    conn = np.array([[i * N + j, i * N + j + 1, (i + 1) * N + j + 1, (i + 1) * N + j]
                     for i in range(Ne) for j in range(Ne)], dtype=np.int32)
    sets = {"left": np.arange(0, N * N, N), "right": np.arange(N - 1, N * N, N)}
    return coords, conn, sets


def test_global_assembly_dirichlet_transpose_golden(goldens):
    """test_sensitivity_analysis.py:41-62: 3x3-node quad mesh, explicit K and u; row 8 of the
    transposed, BC-applied dense Jacobian (float32 golden, rtol 1e-5 / atol 1e-5)."""
    rec = goldens[SA]["test_quad"]
    coords, conn, sets = _square_mesh(3)
    u = np.array(rec["assign"]["random_FE_UV"], float)
    K = np.array(rec["assign"]["random_K"], float)
    bc = {"Ux": {"left": 0.0, "right": 0.05}, "Uy": {"left": 0.0, "right": 0.05}}
    didx, dval = assembly.dirichlet_vectors(["Ux", "Uy"], bc, sets)
    data, idx, R = assembly.assemble("mechanical", "quad", 2, coords, conn, K, u, didx,
                                     {"young_modulus": 1.0, "poisson_ratio": 0.3}, transpose=True)
    dense = assembly.to_dense(data, idx, 18)
    _check(dense[8], rec["asserts"][0])
    assert idx.dtype == conn.dtype and idx.shape == (4 * 64, 2)
    # Dirichlet rows: zero off-diagonal, kept diagonal
    for r in didx:
        off = dense[r].copy()
        off[r] = 0.0
        assert np.all(off == 0.0) and dense[r, r] != 0.0


TT = "tests/unit/test_nonlinear_transient_thermal.py"
AC = "tests/unit/test_allencahn_loss.py"


@pytest.mark.parametrize("test,etype,num_gp,coords", [
    ("test_tetra", "tetra", 1, "tet_points_coordinates"), ("test_hexa", "hexahedron", 2, "hex_points_coordinates"),
    ("test_tri", "triangle", 1, "tri_points_coordinates"), ("test_quad", "quad", 2, "quad_points_coordinates")])
def test_transient_thermal_goldens(goldens, test, etype, num_gp, coords):
    """test_nonlinear_transient_thermal.py:26-165: rho=cp=1, beta=1.5, c=1, dt=0.005, Tc=1, Tn=0, k0=1."""
    rec = goldens[TT][test]
    X = np.array(rec["assign"][coords], float)[None]
    a = X.shape[1]
    en, re, Ke = losses.transient_thermal_element(etype, num_gp, X, np.ones((1, a)), np.zeros((1, a)),
                                                  np.ones((1, a)), 1.0, 1.0, 0.005, 1.5, 1.0)
    _check(Ke[0], rec["asserts"][0])
    _check(re[0], rec["asserts"][1])


@pytest.mark.parametrize("test,etype,num_gp,coords", [
    ("test_hexa", "hexahedron", 2, "hex_points_coordinates"), ("test_tri", "triangle", 1, "tri_points_coordinates"),
    ("test_quad", "quad", 2, "quad_points_coordinates")])
def test_allen_cahn_goldens(goldens, test, etype, num_gp, coords):
    """test_allencahn_loss.py:19-105: dt=0.001, epsilon=0.2, phi_c=1, phi_n=0."""
    rec = goldens[AC][test]
    X = np.array(rec["assign"][coords], float)[None]
    a = X.shape[1]
    en, re, Ke = losses.allen_cahn_element(etype, num_gp, X, np.ones((1, a)), np.zeros((1, a)), 0.001, 0.2)
    _check(Ke[0].reshape(-1), rec["asserts"][0])
    _check(re[0], rec["asserts"][1])


@pytest.mark.parametrize("physics,etype,num_gp", [("transient_thermal", "quad", 2), ("transient_thermal", "tetra", 1),
                                                  ("allen_cahn", "quad", 2), ("allen_cahn", "hexahedron", 2)])
def test_implicit_scalar_batch_gradients_vs_finite_differences(physics, etype, num_gp):
    """Pins the analytic cotangents of the implicit-Euler scalar losses (true potentials: jax.grad of the energy of
    transient_thermal.py:42-73 / phase_field.py:38-70) on central differences of batch_loss."""
    from oracle import assembly
    from tests import gpu_helpers as H
    mesh = H.make_mesh(etype, 2 if etype in ("hexahedron", "tetra") else 3, seed=3)
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes(etype)
    nn = len(coords)
    rng = np.random.default_rng(5)
    cur, nxt = rng.uniform(0.2, 1.0, (2, nn)), rng.uniform(0.2, 1.0, (2, nn))
    if physics == "transient_thermal":
        params = {"rho": 1.3, "cp": 0.7, "beta": 1.5, "c": 3, "k0": rng.uniform(0.5, 1.5, nn), "time_step": 0.05}
    else:
        params = {"dt": 0.02, "epsilon": 0.3}
    didx = np.asarray(mesh.GetNodeSet("left"), dtype=np.int64)
    dval = np.full(didx.size, 0.7)
    args = (physics, etype, num_gp, coords, conn)
    gU, gK = assembly.batch_loss_grads(*args, cur, nxt, didx, dval, params, exponent=2.0)
    free = np.setdiff1d(np.arange(nn), didx)
    h = 1e-6
    for _ in range(5):
        b, i, j = rng.integers(2), rng.choice(free), rng.integers(nn)
        for arr, idx, grad in ((nxt, i, gU), (cur, j, gK)):
            ap, am = arr.copy(), arr.copy()
            ap[b, idx] += h
            am[b, idx] -= h
            pair = (lambda a: (cur, a)) if arr is nxt else (lambda a: (a, nxt))
            fp = assembly.batch_loss(*args, *pair(ap), didx, dval, params, exponent=2.0)[0]
            fm = assembly.batch_loss(*args, *pair(am), didx, dval, params, exponent=2.0)[0]
            assert abs((fp - fm) / (2 * h) - grad[b, idx]) <= 2e-6 * max(1.0, np.abs(grad).max())
    assert not gU[:, didx].any()


def _tf32(x):
    """Round float32 values to TF32 (10 explicit mantissa bits, round to nearest)."""
    b = np.asarray(x, np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def test_kratos_ffi_class_jacobian_golden(goldens):
    """tests/unit/test_kratos_ffi_mechanical_loss.py:72: the dense Jacobian of KratosSmallDisplacement3DTetra on one
    tet (E = 1, nu = 0.3).  Its values carry only 11 significant bits: they are the TF32 rounding of the float32
    stiffness (JAX's default matmul precision on the GPU that produced them, in the 0/1 mask products of
    fe_loss.py:203-207).  The constant-strain Tet4 of the oracle, rounded the same way, reproduces them exactly --
    which pins the Kratos element (SmallDisplacementElement3D4N + LinearElastic3DLaw) to MechanicalLoss3DTetra with a
    unit control field.  (The energy / residual goldens of that test need JAX's PRNG and cannot be used here.)"""
    rec = goldens["tests/unit/test_kratos_ffi_mechanical_loss.py"]["test_tetra"]
    X = np.array(rec["assign"]["tet_points_coordinates"], float)[None]
    jac = np.array([a for a in rec["asserts"] if "jac" in a["expr"]][0]["value"])
    _, _, Ke = losses.mechanical_element("tetra", 1, X, np.ones((1, 4)), np.zeros((1, 12)), 1.0, 0.3, None)
    assert np.abs(_tf32(Ke[0].reshape(-1)) - jac).max() <= 1e-12
    assert np.abs(Ke[0].reshape(-1) - jac).max() <= 2.0 ** -11 * np.abs(jac).max()
