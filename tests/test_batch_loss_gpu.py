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
         ("neohooke", "tetra", 1, {}), ("neohooke", "hexahedron", 2, {}), ("neohooke", "quad", 2, {})]


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
