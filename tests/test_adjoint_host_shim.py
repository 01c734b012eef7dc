"""CPU check of the adjoint-sensitivity element routines (folax_b200/csrc/adjoint*.cuh).

The per-thread bodies of the kernels are __host__ __device__ functions; tests/host_shim/adjoint_host.cu loops
them over the elements on the CPU (test infrastructure, never part of libfolax_b200).  Here they are compared,
in float64, with the complex-step oracle (oracle/responses.py) -- closed forms against an independent
differentiation route -- on every element type and integration order.  The GPU run of the same functions is
tests/test_zz1_responses_gpu.py."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import assembly, responses
from tests.gpu_helpers import make_mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ELEM = {"hexahedron": 0, "quad": 1, "tetra": 2, "triangle": 3}
NGP = {("hexahedron", 1): 1, ("hexahedron", 2): 8, ("hexahedron", 3): 27, ("quad", 1): 1, ("quad", 2): 4,
       ("quad", 3): 9, ("tetra", 1): 1, ("tetra", 2): 4, ("tetra", 3): 8, ("triangle", 1): 1,
       ("triangle", 2): 3, ("triangle", 3): 4}




def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _case(etype, physics, seed=0):
    mesh = make_mesh(etype, 3, perturb=0.25, seed=seed)
    coords = np.ascontiguousarray(mesh.GetNodesCoordinates(), dtype=np.float64)
    conn = np.ascontiguousarray(mesh.GetElementsNodes(etype), dtype=np.int32)
    d = assembly.dofs_per_node(physics, etype)
    rng = np.random.default_rng(seed + 7)
    nn = coords.shape[0]
    K = rng.uniform(0.1, 1.0, nn)
    u = rng.uniform(0.1, 1.0, nn * d)
    lam = rng.standard_normal(nn * d)
    return coords, conn, d, K, u, lam


PHYS = {"mechanical": 0, "thermal": 1, "neohooke": 2, "stvenant": 4, "transient_thermal": 5, "allen_cahn": 6}


def _params(physics, dim, conn=None, nn=0):
    """(double[12] of include/folax_b200.h, the oracle's parameter dict, nodal aux field or None)."""
    arr = np.zeros(12)
    if physics in ("mechanical", "neohooke", "stvenant"):
        arr[0], arr[1] = 1.3, 0.3
        arr[2:2 + dim] = [0.2, -0.4, 0.7][:dim]
        return arr, {"young_modulus": 1.3, "poisson_ratio": 0.3, "body_force": arr[2:2 + dim].copy()}, None
    if physics == "thermal":
        arr[5], arr[6] = 2.0, 4.0
        return arr, {"beta": 2.0, "c": 4.0}, None
    if physics == "transient_thermal":
        arr[5], arr[6], arr[8], arr[9], arr[10] = 1.5, 2.0, 1.2, 0.8, 0.05
        k0 = np.random.default_rng(42).uniform(0.5, 1.5, nn)
        return arr, {"beta": 1.5, "c": 2.0, "rho": 1.2, "cp": 0.8, "time_step": 0.05, "k0": k0[conn]}, k0
    arr[10], arr[11] = 0.01, 0.3
    return arr, {"dt": 0.01, "epsilon": 0.3}, None


@pytest.mark.parametrize("physics", list(PHYS))
@pytest.mark.parametrize("etype", list(ELEM))
@pytest.mark.parametrize("num_gp", [1, 2, 3])
def test_residual_adjoint_sensitivities(shim, physics, etype, num_gp):
    """fe_response.py:312-331, 424-442: lam_e^T d re/d K_e and lam_e^T d re/d x_e -- the closed forms
    (mechanical, thermal) and the forward-mode sweeps (the other losses) of csrc/adjoint.cuh against complex-step
    differentiation of the oracle's ComputeElement."""
    coords, conn, d, K, u, lam = _case(etype, physics, seed=3)
    if physics in ("neohooke", "stvenant"):
        u = 0.05 * (u - 0.5)                                            # keep det F > 0
    ne, a = conn.shape
    dim = 3 if etype in ("hexahedron", "tetra") else 2
    arr, par, aux = _params(physics, dim, conn, coords.shape[0])
    dk, dx = np.full((ne, a), 0.5), np.full((ne, a * 3), -0.25)        # accumulate on top of these
    assert shim.host_residual_adjoint_elements(PHYS[physics], ELEM[etype], num_gp, 1, C.c_longlong(ne), _p(coords),
                                               _p(conn), _p(K), _p(u), _p(lam), _p(aux), _p(arr), _p(dk), _p(dx)) == 0
    g = assembly.element_dof_ids(conn, d)
    rK, rX = responses.residual_adjoint_grads(physics, etype, num_gp, coords[conn], K[conn], u[g], lam[g], par)
    assert np.abs(dk - 0.5 - rK).max() <= 1e-11 * np.abs(rK).max()
    assert np.abs(dx + 0.25 - rX).max() <= 1e-11 * np.abs(rX).max()
    if dim == 2:
        assert np.all(dx[:, 2::3] == -0.25)                             # unused coordinate: exact zeros added
    # overwrite mode, one output only
    dk2 = np.full((ne, a), 9.0)
    assert shim.host_residual_adjoint_elements(PHYS[physics], ELEM[etype], num_gp, 0, C.c_longlong(ne), _p(coords),
                                               _p(conn), _p(K), _p(u), _p(lam), _p(aux), _p(arr), _p(dk2), None) == 0
    assert np.abs(dk2 - rK).max() <= 1e-11 * np.abs(rK).max()


@pytest.mark.parametrize("physics", list(PHYS))
@pytest.mark.parametrize("etype,num_gp", [("hexahedron", 2), ("quad", 3), ("tetra", 2), ("triangle", 1)])
def test_product_routes_equal_whole_element_forward_mode(shim, physics, etype, num_gp):
    """The routes the kernels use (closed forms; point-law derivatives + closed-form geometry) against the
    whole-element dual-number sweeps of csrc/adjoint.cuh on the same inputs."""
    coords, conn, d, K, u, lam = _case(etype, physics, seed=5)
    if physics in ("neohooke", "stvenant"):
        u = 0.05 * (u - 0.5)
    ne, a = conn.shape
    arr, _, aux = _params(physics, 3 if etype in ("hexahedron", "tetra") else 2, conn, coords.shape[0])
    dk, dx, dk2, dx2 = np.zeros((ne, a)), np.zeros((ne, a * 3)), np.zeros((ne, a)), np.zeros((ne, a * 3))
    assert shim.host_residual_adjoint_elements(PHYS[physics], ELEM[etype], num_gp, 0, C.c_longlong(ne), _p(coords),
                                               _p(conn), _p(K), _p(u), _p(lam), _p(aux), _p(arr), _p(dk), _p(dx)) == 0
    assert shim.host_residual_adjoint_dual_reference(PHYS[physics], ELEM[etype], num_gp, C.c_longlong(ne), _p(coords),
                                                     _p(conn), _p(K), _p(u), _p(lam), _p(aux), _p(arr), _p(dk2),
                                                     _p(dx2)) == 0
    assert np.abs(dk - dk2).max() <= 1e-12 * np.abs(dk2).max() and np.abs(dx - dx2).max() <= 1e-12 * np.abs(dx2).max()


@pytest.mark.parametrize("etype", list(ELEM))
@pytest.mark.parametrize("num_gp", [1, 2, 3])
def test_gauss_interpolation_and_response(shim, etype, num_gp):
    """fe_response.py:91-168 with the formula 'jnp.sin(E)*U[0]**2 + E*U[-1]' (partials taken analytically
    here, by torch autograd in the product)."""
    coords, conn, d, K, u, _ = _case(etype, "mechanical")
    ne, a = conn.shape
    g = NGP[(etype, num_gp)]
    kg, ug = np.zeros((ne, g)), np.zeros((d, ne, g))
    assert shim.host_gauss_interpolate(ELEM[etype], num_gp, d, C.c_longlong(ne), _p(conn), _p(K), _p(u), _p(kg),
                                       _p(ug)) == 0
    from oracle.geometry import ELEMENTS
    Ns = np.stack([ELEMENTS[etype].N(p) for p in ELEMENTS[etype].gauss(num_gp)[0]])
    np.testing.assert_allclose(kg, np.einsum("ga,ea->eg", Ns, K[conn]), rtol=1e-14, atol=1e-15)
    ue = u[assembly.element_dof_ids(conn, d)].reshape(ne, a, d)
    np.testing.assert_allclose(ug, np.einsum("ga,eak->keg", Ns, ue), rtol=1e-14, atol=1e-15)

    f = responses.response_function("jnp.sin(E)*U[0]**2 + E*U[-1]", "E", "Ux")
    fv = np.ascontiguousarray(f(kg, ug))
    fk = np.ascontiguousarray(np.cos(kg) * ug[0] ** 2 + ug[-1])
    fu = np.zeros_like(ug)
    fu[0] += 2.0 * np.sin(kg) * ug[0]
    fu[-1] += kg
    val, du, dk, dx = np.zeros(ne), np.zeros((ne, a * d)), np.zeros((ne, a)), np.zeros((ne, a * 3))
    assert shim.host_response_elements(ELEM[etype], num_gp, d, C.c_longlong(ne), _p(coords), _p(conn), _p(fv), _p(fk),
                                       _p(fu), _p(val), _p(du), _p(dk), _p(dx)) == 0
    X, de, uel = coords[conn], K[conn], u[assembly.element_dof_ids(conn, d)]
    ref_val = responses.element_values(f, etype, num_gp, d, X, de, uel)
    rU, rK, rX = responses.element_value_grads(f, etype, num_gp, d, X, de, uel)
    for got, ref in ((val, ref_val), (du, rU), (dk, rK), (dx, rX)):
        assert np.abs(got - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1e-300)




@pytest.mark.parametrize("physics", list(PHYS))
@pytest.mark.parametrize("etype,num_gp", [("hexahedron", 2), ("hexahedron", 3), ("quad", 2), ("tetra", 1), ("tetra", 2),
                                          ("triangle", 1), ("triangle", 3)])
def test_element_energies(shim, physics, etype, num_gp):
    """ComputeElementsEnergies (fe_loss.py:149-176): the first return value of ComputeElement per element."""
    coords, conn, d, K, u, _ = _case(etype, physics, seed=9)
    if physics in ("neohooke", "stvenant"):
        u = 0.05 * (u - 0.5)
    ne = conn.shape[0]
    dim = 3 if etype in ("hexahedron", "tetra") else 2
    arr, par, aux = _params(physics, dim, conn, coords.shape[0])
    if "k0" in par:
        par = dict(par, k0=aux)                                          # compute_elements gathers the nodal field itself
    en = np.full(ne, np.nan)
    assert shim.host_element_energies(PHYS[physics], ELEM[etype], num_gp, C.c_longlong(ne), _p(coords), _p(conn), _p(K),
                                      _p(u), _p(aux), _p(arr), _p(en)) == 0
    ref = assembly.compute_elements(physics, etype, num_gp, coords, conn, K, u, par)[0]
    assert np.abs(en - ref).max() <= 1e-12 * np.abs(ref).max()
