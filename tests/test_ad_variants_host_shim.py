"""CPU check of the *_AD.py loss variants (folax_b200/csrc/assemble_ad_threads.cuh: residual in a generic scalar type,
stiffness by dual-number sweeps, as the reference's `jax.jacfwd(residual)`): the kernel's thread body, compiled for
the CPU (tests/host_shim/assemble_ad_host.cu), against the reference's own 19-digit goldens
(tests/unit/test_neo_hooke_mechanical_loss_AD.py:38-190) and against the oracle (oracle/losses.py: ad_variant_element,
complex-step Jacobian) on meshes, including the transpose switch and the Dirichlet row mask."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle import assembly, losses
from tests.gpu_helpers import make_mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ELEM = {"hexahedron": 0, "quad": 1, "tetra": 2, "triangle": 3}
LAW = {"neohooke_ad": 7, "stvenant_ad": 8}


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _run(shim, law, etype, num_gp, coords, conn, K, u, flags, nu, body, transpose):
    ne, a = conn.shape
    d = 3 if etype in ("hexahedron", "tetra") else 2
    nd = a * d
    params = np.zeros(12)
    params[0], params[1] = 1.0, nu
    params[2:2 + d] = body
    ke, re, en = np.full(ne * nd * nd, np.nan), np.full(ne * nd, np.nan), np.full(ne, np.nan)
    assert shim.host_assemble_ad(LAW[law], ELEM[etype], num_gp, int(transpose), C.c_longlong(ne), _p(coords), _p(conn),
                                 _p(K), _p(u), _p(flags), _p(params), _p(ke), _p(re), _p(en)) == 0
    return ke.reshape(ne, nd, nd), re.reshape(ne, nd), en


@pytest.mark.parametrize("test,etype,num_gp,ckey", [("test_tetra", "tetra", 1, "tet_points_coordinates"),
                                                    ("test_hexa", "hexahedron", 2, "hex_points_coordinates"),
                                                    ("test_quad", "quad", 2, "quad_points_coordinates")])
def test_reference_ad_goldens(shim, test, etype, num_gp, ckey):
    with open(os.path.join(ROOT, "tests", "golden", "reference_unit_goldens.json")) as fh:
        rec = json.load(fh)["tests/unit/test_neo_hooke_mechanical_loss_AD.py"][test]
    X = np.ascontiguousarray(rec["assign"][ckey], dtype=np.float64)
    a = X.shape[0]
    d = 3 if etype != "quad" else 2
    conn = np.arange(a, dtype=np.int32)[None]
    ke, re, en = _run(shim, "neohooke_ad", etype, num_gp, X, conn, np.ones(a), np.ones(a * d),
                      np.zeros(a * d, np.uint8), 0.3, [1.0, 2.0, 3.0][:d], False)
    K_ref, r_ref = np.array(rec["asserts"][0]["value"]), np.array(rec["asserts"][1]["value"])
    assert np.abs(ke[0] - K_ref).max() <= 1e-12 * np.abs(K_ref).max()
    assert np.abs(re[0] - r_ref).max() <= 1e-12 * max(np.abs(r_ref).max(), 1.0)


@pytest.mark.parametrize("law", list(LAW))
@pytest.mark.parametrize("etype,num_gp", [("hexahedron", 2), ("tetra", 1), ("tetra", 2), ("quad", 2), ("quad", 3),
                                          ("triangle", 1), ("hexahedron", 1)])
@pytest.mark.parametrize("transpose", [False, True])
def test_mesh_assembly_against_oracle(shim, law, etype, num_gp, transpose):
    mesh = make_mesh(etype, 2, perturb=0.2, seed=4)
    coords = np.ascontiguousarray(mesh.GetNodesCoordinates(), dtype=np.float64)
    conn = np.ascontiguousarray(mesh.GetElementsNodes(etype), dtype=np.int32)
    d = 3 if etype in ("hexahedron", "tetra") else 2
    nn = len(coords)
    rng = np.random.default_rng(8)
    K, u = rng.uniform(0.2, 1.0, nn), 0.03 * rng.standard_normal(nn * d)
    dofs = ["Ux", "Uy", "Uz"][:d]
    didx, _ = assembly.dirichlet_vectors(dofs, {k: {"left": 0.0, "right": 0.1} for k in dofs}, mesh.node_sets)
    flags = np.zeros(nn * d, np.uint8)
    flags[didx] = 1
    body = np.array([0.2, -0.4, 0.7][:d])
    ke, re, en = _run(shim, law, etype, num_gp, coords, conn, K, u, flags, 0.3, body, transpose)
    g = assembly.element_dof_ids(conn, d)
    en_ref, re_ref, Ke_ref = losses.ad_variant_element(etype, num_gp, coords[conn], K[conn], u[g], 0.3, body, law=law)
    bc = 1.0 - flags[g].astype(float)
    re_m, Ke_m = assembly.apply_dirichlet(re_ref, Ke_ref, bc, transpose)
    assert np.abs(ke - Ke_m).max() <= 1e-12 * np.abs(Ke_m).max()
    assert np.abs(re - re_m).max() <= 1e-12 * np.abs(re_m).max()
    assert np.abs(en - en_ref).max() <= 1e-12 * max(np.abs(en_ref).max(), 1e-300)
    # exact zeros where the mask says so
    rows = flags[g].astype(bool)
    off = ke.copy()
    idx = np.arange(ke.shape[1])
    off[:, idx, idx] = 0.0
    assert np.all(off[rows] == 0.0)


def test_saint_venant_ad_equals_the_analytic_class_at_identity(shim):
    """What the reference's own test asserts (test_saint_venant_mechanical_loss.py:41-42, u = ones => F = I):
    AD and analytic St-Venant agree there (and only there: away from F = I the AD stress carries doubled shear)."""
    X = np.array([[0.1, 0.1, 0.1], [0.28739360416666665, 0.27808503701741405, 0.05672979583333333],
                  [0.0, 1.0, 0.0], [0.0, 1.0, 0.1]])
    conn = np.arange(4, dtype=np.int32)[None]
    ke, re, _ = _run(shim, "stvenant_ad", "tetra", 1, X, conn, np.ones(4), np.ones(12), np.zeros(12, np.uint8), 0.3,
                     [1.0, 2.0, 3.0], False)
    _, re_a, Ke_a = losses.neo_hooke_element("tetra", 1, X[None], np.ones((1, 4)), np.ones((1, 12)), 1.0, 0.3,
                                             np.array([1.0, 2.0, 3.0]), law="stvenant")
    np.testing.assert_allclose(ke[0], Ke_a[0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(re[0], re_a[0], rtol=1e-5, atol=1e-6)
