"""GPU parity of ComputeBatchLoss and its VJP (fe_loss.py:250-262 + the JAX-AD gradient restated
analytically in oracle/assembly.py, itself checked by finite differences below)."""
import numpy as np
import pytest
import torch

from oracle import assembly
from tests import gpu_helpers as H

pytestmark = pytest.mark.gpu

CASES = [("thermal", "quad", 2, {"beta": 0.0, "c": 1}), ("thermal", "quad", 2, {"beta": 2.0, "c": 4}),
         ("thermal", "hexahedron", 2, {}), ("thermal", "tetra", 1, {}), ("thermal", "triangle", 1, {}),
         ("mechanical", "quad", 2, {}), ("mechanical", "hexahedron", 2, {"body_foce": [0.1, 0.2, -0.3]}),
         ("mechanical", "tetra", 1, {}), ("mechanical", "triangle", 1, {}),
         ("neohooke", "tetra", 1, {}), ("neohooke", "hexahedron", 2, {}), ("neohooke", "quad", 2, {}),
         ("stvenant", "tetra", 1, {}), ("stvenant", "quad", 2, {}), ("stvenant", "hexahedron", 2, {})]


@pytest.mark.parametrize("physics,etype,num_gp,extra", CASES)
@pytest.mark.parametrize("exponent", [1.0, 2.0])
def test_batch_loss_and_vjp(physics, etype, num_gp, extra, exponent):
    mesh = H.make_mesh(etype, 3 if etype in ("hexahedron", "tetra") else 6, seed=5)
    loss = H.make_loss(physics, etype, mesh, num_gp, "float64", {**extra, "loss_function_exponent": exponent})
    B = 5
    K, u = H.fields(physics, mesh, loss, seed=7, batch=B)
    Kt = torch.tensor(K, device="cuda", requires_grad=True)
    ut = torch.tensor(u, device="cuda", requires_grad=True)
    mean, (mn, mx, mean2) = loss.ComputeBatchLoss(Kt, ut)
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes(etype)
    args = (physics, etype, num_gp, coords, conn, K, u, loss.dirichlet_indices, loss.dirichlet_values,
            H.oracle_params(loss))
    ref_mean, (rmin, rmax, _), Eb = assembly.batch_loss(*args, exponent=exponent)
    tol = 1e-12 * max(abs(Eb).max(), 1e-300)
    assert abs(mean.item() - ref_mean) <= tol and abs(mean2.item() - ref_mean) <= tol
    assert abs(mn.item() - rmin) <= tol and abs(mx.item() - rmax) <= tol
    (2.5 * mean).backward()
    gU, gK = assembly.batch_loss_grads(*args, exponent=exponent)
    gu = ut.grad.cpu().numpy() / 2.5
    assert np.abs(gu - gU).max() <= 1e-12 * max(np.abs(gU).max(), 1e-300)
    assert not gu[:, loss.dirichlet_indices].any()
    if physics == "mechanical":
        assert Kt.grad is None or not Kt.grad.any()
    else:
        gk = Kt.grad.cpu().numpy() / 2.5
        assert np.abs(gk - gK).max() <= 1e-12 * max(np.abs(gK).max(), 1e-300)


def test_oracle_gradient_vs_finite_differences():
    """Pins the analytic VJP of the oracle: neo-hooke energy is a true potential, so central
    differences of batch_loss must reproduce batch_loss_grads."""
    mesh = H.make_mesh("tetra", 2, seed=1)
    loss = H.make_loss("neohooke", "tetra", mesh, 1)
    K, u = H.fields("neohooke", mesh, loss, seed=2, batch=2)
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("tetra")
    args = dict(dirichlet_indices=loss.dirichlet_indices, dirichlet_values=loss.dirichlet_values,
                params=H.oracle_params(loss), exponent=2.0)
    gU, gK = assembly.batch_loss_grads("neohooke", "tetra", 1, coords, conn, K, u, **args)
    rng = np.random.default_rng(0)
    for _ in range(6):
        b, i = rng.integers(2), rng.integers(u.shape[1])
        h = 1e-6
        up, um = u.copy(), u.copy()
        up[b, i] += h
        um[b, i] -= h
        fd = (assembly.batch_loss("neohooke", "tetra", 1, coords, conn, K, up, **args)[0]
              - assembly.batch_loss("neohooke", "tetra", 1, coords, conn, K, um, **args)[0]) / (2 * h)
        assert abs(fd - gU[b, i]) <= 1e-6 * max(1.0, abs(gU).max())
        j = rng.integers(K.shape[1])
        Kp, Km = K.copy(), K.copy()
        Kp[b, j] += h
        Km[b, j] -= h
        fd = (assembly.batch_loss("neohooke", "tetra", 1, coords, conn, Kp, u, **args)[0]
              - assembly.batch_loss("neohooke", "tetra", 1, coords, conn, Km, u, **args)[0]) / (2 * h)
        assert abs(fd - gK[b, j]) <= 1e-6 * max(1.0, abs(gK).max())


def test_total_energy_and_float32():
    mesh = H.make_mesh("quad", 8, seed=2)
    loss = H.make_loss("thermal", "quad", mesh, 2, "float32", {"beta": 2.0, "c": 4})
    K, u = H.fields("thermal", mesh, loss, seed=3, batch=3)
    K32, u32 = K.astype(np.float32), u.astype(np.float32)
    mean, _ = loss.ComputeBatchLoss(K32, u32)
    coords = np.asarray(mesh.GetNodesCoordinates()).astype(np.float32).astype(np.float64)
    ref, _, Eb = assembly.batch_loss("thermal", "quad", 2, coords, mesh.GetElementsNodes("quad"),
                                     K32.astype(np.float64), u32.astype(np.float64), loss.dirichlet_indices,
                                     loss.dirichlet_values.astype(np.float32).astype(np.float64),
                                     H.oracle_params(loss))
    assert abs(mean.item() - ref) <= 1e-5 * abs(Eb).max()
    e0 = loss.ComputeTotalEnergy(K32[0], u32[0])
    from oracle.assembly import compute_elements
    ref0 = compute_elements("thermal", "quad", 2, coords, mesh.GetElementsNodes("quad"), K32[0].astype(np.float64),
                            u32[0].astype(np.float64), H.oracle_params(loss))[0].sum()
    assert abs(e0.item() - ref0) <= 1e-5 * abs(ref0)


def test_parametric_boundary_learning():
    """mechanical_saint_venant.py:59-66 + fe_loss.py:94-100: the batch parameters are the Dirichlet values of
    every sample, the control field is the mesh's heterogeneity field; cotangents at the Dirichlet dofs flow
    to the parameters."""
    from folax_b200.loss_functions import SaintVenantMechanicalLoss3DTetra
    mesh = H.make_mesh("tetra", 2, seed=3)
    rng = np.random.default_rng(4)
    mesh["K"] = rng.uniform(0.5, 1.0, mesh.GetNumberOfNodes())
    bc = {d: {"left": 0.0, "right": 0.1} for d in ("Ux", "Uy", "Uz")}
    loss = SaintVenantMechanicalLoss3DTetra("pbl", {"dirichlet_bc_dict": bc, "parametric_boundary_learning": True,
                                                    "material_dict": {"young_modulus": 1.0, "poisson_ratio": 0.3}}, mesh)
    loss.Initialize()
    B, nd = 3, loss.dirichlet_indices.size
    known = 0.02 * rng.standard_normal((B, nd))
    u = 0.01 * rng.standard_normal((B, loss.total_number_of_dofs))
    kt = torch.tensor(known, device="cuda", requires_grad=True)
    ut = torch.tensor(u, device="cuda", requires_grad=True)
    mean, _ = loss.ComputeBatchLoss(kt, ut)
    mean.backward()
    # oracle: same energy with the Dirichlet entries replaced per sample, control = heterogeneity field
    full = u.copy()
    full[:, loss.dirichlet_indices] = known
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("tetra")
    Kf = np.tile(mesh["K"], (B, 1))
    none = np.zeros(0, dtype=np.int64)
    ref, _, Eb = assembly.batch_loss("stvenant", "tetra", 1, coords, conn, Kf, full, none, np.zeros(0),
                                     H.oracle_params(loss))
    gU, _ = assembly.batch_loss_grads("stvenant", "tetra", 1, coords, conn, Kf, full, none, np.zeros(0),
                                      H.oracle_params(loss))
    assert abs(mean.item() - ref) <= 1e-12 * abs(Eb).max()
    scale = np.abs(gU).max()
    assert np.abs(kt.grad.cpu().numpy() - gU[:, loss.dirichlet_indices]).max() <= 1e-12 * scale
    want = gU.copy()
    want[:, loss.dirichlet_indices] = 0.0
    assert np.abs(ut.grad.cpu().numpy() - want).max() <= 1e-12 * scale
    # fe_loss.py:94-95, 123-124: GetFullDofVector writes the PER-SAMPLE known dofs in this mode (what the Predict
    # paths call, explicit_parametric_operator_learning.py:117), not the settings' boundary values
    got = loss.GetFullDofVector(known, u).cpu().numpy()
    assert np.array_equal(got, full)
    with pytest.raises(ValueError):
        loss.GetFullDofVector(None, u)
