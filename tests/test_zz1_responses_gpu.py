"""GPU parity of the adjoint sensitivities: folax_b200.responses.FiniteElementResponse through
the C ABI (fol_gauss_interpolate, fol_response_elements, fol_residual_adjoint_elements, fol_sum) against the
reference's known answers and the complex-step oracle.  (Named to run after the rest of the GPU suite.)"""
import json
import os

import numpy as np
import pytest

import folax_b200
from folax_b200.loss_functions import MechanicalLoss2DQuad
from folax_b200.responses import FiniteElementResponse, NodalControl
from oracle import assembly, responses
from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BC = {"Ux": {"left": 0.0, "right": 0.05}, "Uy": {"left": 0.0, "right": 0.05}}


@pytest.fixture(scope="module")
def goldens():
    with open(os.path.join(ROOT, "tests", "golden", "reference_unit_goldens.json")) as fh:
        return json.load(fh)


def _quad_response(N):
    mesh = folax_b200.create_2D_square_mesh(1.0, N)
    loss = MechanicalLoss2DQuad("mechanical_loss_2d", {"dirichlet_bc_dict": BC, "num_gp": 2,
                                                       "material_dict": dict(gh.MATERIAL)}, mesh)
    resp = FiniteElementResponse("test_response", "(E**2)*U[0]", loss, NodalControl("E", mesh))
    resp.Initialize()
    return mesh, loss, resp


def test_reference_unit_goldens(goldens):
    """tests/unit/test_sensitivity_analysis.py:54-80, same calls, same tolerances."""
    rec = goldens["tests/unit/test_sensitivity_analysis.py"]["test_quad"]
    _, _, resp = _quad_response(3)
    u = np.array(rec["assign"]["random_FE_UV"])
    K = np.array(rec["assign"]["random_K"])
    lam = np.array(rec["assign"]["random_adj_FE_UV"])
    jac, rhs = resp.ComputeAdjointJacobianMatrixAndRHSVector(K, u)
    a = rec["asserts"]
    np.testing.assert_allclose(jac.todense()[8, :].cpu().numpy(), a[0]["value"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(rhs.cpu().numpy(), a[1]["value"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(resp.ComputeAdjointNodalControlDerivatives(K, u, lam).cpu().numpy(), a[2]["value"],
                               rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(resp.ComputeAdjointNodalShapeDerivatives(K, u, lam).cpu().numpy(), a[3]["value"],
                               rtol=1e-5, atol=1e-5)


def test_reference_integration_golden(goldens):
    """tests/integration/test_mechanical_2D_sa.py:81-113: FE solve -> adjoint solve -> derivatives (the two
    linear solves are the caller's, done densely on the host here)."""
    rec = goldens["tests/integration/test_mechanical_2D_sa.py"]
    K = np.array(rec["setUp"]["assign"]["random_K"])
    mesh, loss, resp = _quad_response(5)
    u0 = loss.ApplyDirichletBCOnDofVector(np.zeros(loss.GetTotalNumberOfDOFs()))
    jac, R = loss.ComputeJacobianMatrixAndResidualVector(K, u0)
    u = u0.cpu().numpy() + np.linalg.solve(jac.todense().cpu().numpy(), -R.cpu().numpy())
    adj_jac, adj_rhs = resp.ComputeAdjointJacobianMatrixAndRHSVector(K, u)
    lam = np.linalg.solve(adj_jac.todense().cpu().numpy(), adj_rhs.cpu().numpy())
    a = rec["test_sensitivites"]["asserts"]
    cd = resp.ComputeAdjointNodalControlDerivatives(K, u, lam).cpu().numpy()
    sd = resp.ComputeAdjointNodalShapeDerivatives(K, u, lam).cpu().numpy()
    np.testing.assert_allclose(cd, a[0]["value"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(sd, a[1]["value"], rtol=1e-5, atol=1e-5)
    assert np.abs(cd - np.array(a[0]["value"])).max() <= 2e-9 and np.abs(sd - np.array(a[1]["value"])).max() <= 2e-9


CASES = [("mechanical", "quad", 2), ("mechanical", "hexahedron", 2), ("mechanical", "tetra", 1),
         ("mechanical", "triangle", 2), ("mechanical", "hexahedron", 3), ("thermal", "quad", 2),
         ("thermal", "hexahedron", 2), ("thermal", "tetra", 2), ("thermal", "triangle", 1)]


@pytest.mark.parametrize("physics,etype,num_gp", CASES)
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-11), ("float32", 2e-4)])
def test_against_oracle(physics, etype, num_gp, dtype, tol):
    mesh = gh.make_mesh(etype, 4, perturb=0.2, seed=2)
    extra = {"beta": 2.0, "c": 4.0} if physics == "thermal" else {"body_foce": [0.2, -0.4, 0.7][:3 if etype in ("hexahedron", "tetra") else 2]}
    loss = gh.make_loss(physics, etype, mesh, num_gp=num_gp, dtype=dtype, extra=extra)
    name = "T" if physics == "thermal" else "U"
    formula = f"jnp.sin(K)*{name}[0]**2 + K*{name}[-1]"
    resp = FiniteElementResponse("r", formula, loss, NodalControl("K", mesh))
    resp.Initialize()
    rng = np.random.default_rng(11)
    nn, ndof = mesh.GetNumberOfNodes(), loss.total_number_of_dofs
    K, u, lam = rng.uniform(0.1, 1.0, nn), rng.uniform(0.1, 1.0, ndof), rng.standard_normal(ndof)
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes(etype)
    par = gh.oracle_params(loss)
    f = responses.response_function(formula, "K", loss.dofs[0])
    ref_val = responses.compute_value(f, physics, etype, num_gp, coords, conn, K, u)
    assert abs(float(resp.ComputeValue(K, u)) - ref_val) <= tol * abs(ref_val)
    _, _, ref_rhs = responses.adjoint_jacobian_and_rhs(f, physics, etype, num_gp, coords, conn, K, u,
                                                       loss.dirichlet_indices, par)
    jac, rhs = resp.ComputeAdjointJacobianMatrixAndRHSVector(K, u)
    assert np.abs(rhs.cpu().numpy() - ref_rhs).max() <= tol * np.abs(ref_rhs).max()
    assert np.all(rhs.cpu().numpy()[loss.dirichlet_indices] == 0.0)
    ref_cd = responses.control_derivatives(f, physics, etype, num_gp, coords, conn, K, u, lam, par)
    ref_sd = responses.shape_derivatives(f, physics, etype, num_gp, coords, conn, K, u, lam, par)
    cd = resp.ComputeAdjointNodalControlDerivatives(K, u, lam).cpu().numpy()
    sd = resp.ComputeAdjointNodalShapeDerivatives(K, u, lam).cpu().numpy()
    assert np.abs(cd - ref_cd).max() <= tol * np.abs(ref_cd).max()
    assert np.abs(sd - ref_sd).max() <= tol * np.abs(ref_sd).max()
    # run-to-run bit-identical (fixed-order node sums)
    assert np.array_equal(sd, resp.ComputeAdjointNodalShapeDerivatives(K, u, lam).cpu().numpy())


EXTRA = [("neohooke", "quad", 2), ("neohooke", "tetra", 1), ("neohooke", "hexahedron", 2), ("stvenant", "triangle", 1),
         ("stvenant", "hexahedron", 2), ("transient_thermal", "quad", 2), ("transient_thermal", "tetra", 1),
         ("allen_cahn", "quad", 2), ("allen_cahn", "hexahedron", 2)]


@pytest.mark.parametrize("physics,etype,num_gp", EXTRA)
def test_forward_mode_physics_against_oracle(physics, etype, num_gp):
    """Residual sensitivities of the finite-strain and implicit-Euler scalar losses (forward-mode sweeps in
    csrc/adjoint.cuh) + the response part, float64, against the complex-step oracle."""
    from folax_b200 import loss_functions as lf
    mesh = gh.make_mesh(etype, 3, perturb=0.2, seed=5)
    rng = np.random.default_rng(13)
    nn = mesh.GetNumberOfNodes()
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes(etype)
    if physics in ("neohooke", "stvenant"):
        loss = gh.make_loss(physics, etype, mesh, num_gp=num_gp, extra={"body_foce": [0.2, -0.4, 0.7][:3 if etype in ("hexahedron", "tetra") else 2]})
        par = gh.oracle_params(loss)
        K, u = gh.fields(physics, mesh, loss, seed=3)
        name = "U"
    elif physics == "transient_thermal":
        cls = {"quad": lf.TransientThermalLoss2DQuad, "tetra": lf.TransientThermalLoss3DTetra}[etype]
        k0 = rng.uniform(0.5, 1.5, nn)
        loss = cls("tt", {"dirichlet_bc_dict": {"T": {"left": 1.0, "right": 0.1}}, "c": 3, "num_gp": num_gp,
                          "material_dict": {"rho": 1.3, "cp": 0.7, "beta": 1.5, "k0": k0},
                          "time_integration_dict": {"time_step": 0.01}}, mesh)
        loss.Initialize()
        par = {"rho": 1.3, "cp": 0.7, "beta": 1.5, "c": 3, "k0": k0, "time_step": 0.01}
        K, u = rng.uniform(0.2, 1.0, nn), rng.uniform(0.2, 1.0, nn)      # (current, next) temperatures
        name = "T"
    else:
        cls = {"quad": lf.AllenCahnLoss2DQuad, "hexahedron": lf.AllenCahnLoss3DHexa}[etype]
        loss = cls("ac", {"dirichlet_bc_dict": {"Phi": {"left": 1.0}}, "num_gp": num_gp,
                          "material_dict": {"rho": 1.0, "cp": 1.0, "dt": 0.002, "epsilon": 0.3}}, mesh)
        loss.Initialize()
        par = {"dt": 0.002, "epsilon": 0.3}
        K, u = rng.uniform(-1.0, 1.0, nn), rng.uniform(-1.0, 1.0, nn)    # (current, next) phase field
        name = "P"
    assert loss.num_gp == num_gp
    lam = rng.standard_normal(loss.total_number_of_dofs)
    formula = f"jnp.cos(K)*{name}[0]**2 + K*{name}[-1]"
    resp = FiniteElementResponse("r", formula, loss, NodalControl("K", mesh))
    resp.Initialize()
    f = responses.response_function(formula, "K", loss.dofs[0])
    ref_cd = responses.control_derivatives(f, physics, etype, num_gp, coords, conn, K, u, lam, par)
    ref_sd = responses.shape_derivatives(f, physics, etype, num_gp, coords, conn, K, u, lam, par)
    cd = resp.ComputeAdjointNodalControlDerivatives(K, u, lam).cpu().numpy()
    sd = resp.ComputeAdjointNodalShapeDerivatives(K, u, lam).cpu().numpy()
    assert np.abs(cd - ref_cd).max() <= 1e-11 * np.abs(ref_cd).max()
    assert np.abs(sd - ref_sd).max() <= 1e-11 * np.abs(ref_sd).max()
    _, _, ref_rhs = responses.adjoint_jacobian_and_rhs(f, physics, etype, num_gp, coords, conn, K, u,
                                                       loss.dirichlet_indices, par)
    _, rhs = resp.ComputeAdjointJacobianMatrixAndRHSVector(K, u)
    assert np.abs(rhs.cpu().numpy() - ref_rhs).max() <= 1e-11 * np.abs(ref_rhs).max()


def test_history_dependent_loss_raises():
    from folax_b200 import _lib
    from folax_b200.loss_functions import ElastoplasticityLoss2DQuad
    mesh = gh.make_mesh("quad", 3)
    loss = ElastoplasticityLoss2DQuad("ep", {"dirichlet_bc_dict": BC, "material_dict": {
        "young_modulus": 3.0, "poisson_ratio": 0.3, "iso_hardening_parameter_1": 0.4, "iso_hardening_param_2": 10.0,
        "yield_limit": 0.2}}, mesh)
    resp = FiniteElementResponse("r", "K*U[0]", loss, NodalControl("K", mesh))
    resp.Initialize()
    nn, ndof = mesh.GetNumberOfNodes(), loss.total_number_of_dofs
    with pytest.raises(_lib.FolaxError):
        resp.ComputeAdjointNodalControlDerivatives(np.ones(nn), np.zeros(ndof), np.ones(ndof))
