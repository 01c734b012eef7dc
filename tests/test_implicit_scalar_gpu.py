"""Transient thermal and Allen-Cahn losses: the reference's single-element goldens through the kernels and
mesh-level parity against the oracle."""
import numpy as np
import pytest

import folax_b200
from folax_b200 import loss_functions as lf
from oracle import assembly
from tests import gpu_helpers as H

pytestmark = pytest.mark.gpu

TT = "tests/unit/test_nonlinear_transient_thermal.py"
AC = "tests/unit/test_allencahn_loss.py"


def _one(etype, coords):
    m = folax_b200.Mesh("", ".")
    m.node_ids = np.arange(len(coords))
    m.nodes_coordinates = np.asarray(coords, float)
    m.elements_nodes = {etype: m.node_ids.reshape(1, -1)}
    return m


@pytest.mark.parametrize("test,cls,etype,coords", [
    ("test_tetra", lf.TransientThermalLoss3DTetra, "tetra", "tet_points_coordinates"),
    ("test_hexa", lf.TransientThermalLoss3DHexa, "hexahedron", "hex_points_coordinates"),
    ("test_tri", lf.TransientThermalLoss2DTri, "triangle", "tri_points_coordinates"),
    ("test_quad", lf.TransientThermalLoss2DQuad, "quad", "quad_points_coordinates")])
def test_transient_thermal_reference_goldens(goldens, test, cls, etype, coords):
    """test_nonlinear_transient_thermal.py:26-165 with its own tolerances."""
    rec = goldens[TT][test]
    X = rec["assign"][coords]
    loss = cls("tt", {"dirichlet_bc_dict": {"T": {}}, "material_dict": {"rho": 1.0, "cp": 1.0, "beta": 1.5, "c": 1.0},
                      "time_integration_dict": {"method": "implicit-euler", "time_step": 0.005}}, _one(etype, X))
    loss.Initialize()
    a = len(X)
    en, re, ke = loss.ComputeElement(np.array(X), np.ones(a), np.zeros((a, 1)), np.ones((a, 1)))
    k, r = rec["asserts"]
    np.testing.assert_allclose(ke.cpu().numpy(), np.array(k["value"]), rtol=k["rtol"], atol=k["atol"])
    np.testing.assert_allclose(re.cpu().numpy().flatten(), np.array(r["value"]), rtol=r["rtol"], atol=r["atol"])


@pytest.mark.parametrize("test,cls,etype,coords", [
    ("test_hexa", lf.AllenCahnLoss3DHexa, "hexahedron", "hex_points_coordinates"),
    ("test_tri", lf.AllenCahnLoss2DTri, "triangle", "tri_points_coordinates"),
    ("test_quad", lf.AllenCahnLoss2DQuad, "quad", "quad_points_coordinates")])
def test_allen_cahn_reference_goldens(goldens, test, cls, etype, coords):
    """test_allencahn_loss.py:19-105 with its own tolerances."""
    rec = goldens[AC][test]
    X = rec["assign"][coords]
    loss = cls("ac", {"dirichlet_bc_dict": {"Phi": {}}, "material_dict": {"rho": 1.0, "cp": 1.0, "dt": 0.001,
                                                                          "epsilon": 0.2}}, _one(etype, X))
    loss.Initialize()
    a = len(X)
    en, re, ke = loss.ComputeElement(np.array(X), np.ones(a), np.zeros((a, 1)))
    k, r = rec["asserts"]
    np.testing.assert_allclose(ke.cpu().numpy().flatten(), np.array(k["value"]), rtol=k["rtol"], atol=k["atol"])
    np.testing.assert_allclose(re.cpu().numpy().flatten(), np.array(r["value"]), rtol=r["rtol"], atol=r["atol"])


@pytest.mark.parametrize("kind,etype", [("tt", "quad"), ("tt", "hexahedron"), ("tt", "tetra"), ("ac", "quad"),
                                        ("ac", "triangle"), ("ac", "hexahedron")])
def test_mesh_assembly_matches_oracle(kind, etype):
    mesh = H.make_mesh(etype, 3 if etype in ("hexahedron", "tetra") else 6, seed=8)
    rng = np.random.default_rng(2)
    nn = mesh.GetNumberOfNodes()
    cur, nxt = rng.uniform(0.2, 1.0, nn), rng.uniform(0.2, 1.0, nn)
    if kind == "tt":
        cls = {"quad": lf.TransientThermalLoss2DQuad, "hexahedron": lf.TransientThermalLoss3DHexa,
               "tetra": lf.TransientThermalLoss3DTetra}[etype]
        k0 = rng.uniform(0.5, 1.5, nn)
        loss = cls("tt", {"dirichlet_bc_dict": {"T": {"left": 1.0, "right": 0.1}}, "c": 3,
                          "material_dict": {"rho": 1.3, "cp": 0.7, "beta": 1.5, "k0": k0},
                          "time_integration_dict": {"time_step": 0.01}}, mesh)
        params = {"rho": 1.3, "cp": 0.7, "beta": 1.5, "c": 3, "k0": k0, "time_step": 0.01}
        phys = "transient_thermal"
    else:
        cls = {"quad": lf.AllenCahnLoss2DQuad, "triangle": lf.AllenCahnLoss2DTri,
               "hexahedron": lf.AllenCahnLoss3DHexa}[etype]
        loss = cls("ac", {"dirichlet_bc_dict": {"Phi": {"left": 1.0}}, "material_dict": {"rho": 1.0, "cp": 1.0,
                                                                                         "dt": 0.002, "epsilon": 0.3}}, mesh)
        params = {"dt": 0.002, "epsilon": 0.3}
        phys = "allen_cahn"
    loss.Initialize()
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes(etype)
    for transpose in (False, True):
        jac, R = loss.ComputeJacobianMatrixAndResidualVector(cur, nxt, transpose_jacobian=transpose)
        data, idx, Rref = assembly.assemble(phys, etype, loss.num_gp, coords, conn, cur, nxt, loss.dirichlet_indices,
                                            params, transpose)
        assert np.array_equal(jac.indices.cpu().numpy(), idx)
        assert np.abs(jac.data.cpu().numpy() - data).max() <= 1e-12 * np.abs(data).max()
        assert np.abs(R.cpu().numpy() - Rref).max() <= 4e-12 * np.abs(Rref).max()
    en_ref = assembly.compute_elements(phys, etype, loss.num_gp, coords, conn, cur, nxt, params)[0].sum()
    assert abs(loss.ComputeTotalEnergy(cur, nxt).item() - en_ref) <= 1e-12 * abs(en_ref)


@pytest.mark.parametrize("kind,etype", [("tt", "quad"), ("tt", "tetra"), ("tt", "hexahedron"), ("ac", "quad"),
                                        ("ac", "triangle"), ("ac", "hexahedron")])
@pytest.mark.parametrize("exponent", [1.0, 2.0])
def test_batch_loss_and_vjp_match_oracle(kind, etype, exponent):
    """ComputeBatchLoss of the implicit-Euler losses: (params, dofs) = (current, next) fields; the energies are true
    potentials, so both cotangents come from the energy (oracle pinned on finite differences in the CPU suite)."""
    import torch
    mesh = H.make_mesh(etype, 3 if etype in ("hexahedron", "tetra") else 6, seed=8)
    rng = np.random.default_rng(3)
    nn = mesh.GetNumberOfNodes()
    B = 5
    cur, nxt = rng.uniform(0.2, 1.0, (B, nn)), rng.uniform(0.2, 1.0, (B, nn))
    if kind == "tt":
        cls = {"quad": lf.TransientThermalLoss2DQuad, "hexahedron": lf.TransientThermalLoss3DHexa,
               "tetra": lf.TransientThermalLoss3DTetra}[etype]
        k0 = rng.uniform(0.5, 1.5, nn)
        loss = cls("tt", {"dirichlet_bc_dict": {"T": {"left": 1.0, "right": 0.1}}, "c": 3,
                          "loss_function_exponent": exponent,
                          "material_dict": {"rho": 1.3, "cp": 0.7, "beta": 1.5, "k0": k0},
                          "time_integration_dict": {"time_step": 0.01}}, mesh)
        params = {"rho": 1.3, "cp": 0.7, "beta": 1.5, "c": 3, "k0": k0, "time_step": 0.01}
        phys = "transient_thermal"
    else:
        cls = {"quad": lf.AllenCahnLoss2DQuad, "triangle": lf.AllenCahnLoss2DTri,
               "hexahedron": lf.AllenCahnLoss3DHexa}[etype]
        loss = cls("ac", {"dirichlet_bc_dict": {"Phi": {"left": 1.0}}, "loss_function_exponent": exponent,
                          "material_dict": {"rho": 1.0, "cp": 1.0, "dt": 0.002, "epsilon": 0.3}}, mesh)
        params = {"dt": 0.002, "epsilon": 0.3}
        phys = "allen_cahn"
    loss.Initialize()
    ct = torch.tensor(cur, device="cuda", requires_grad=True)
    nt = torch.tensor(nxt, device="cuda", requires_grad=True)
    mean, (mn, mx, mean2) = loss.ComputeBatchLoss(ct, nt)
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes(etype)
    args = (phys, etype, loss.num_gp, coords, conn, cur, nxt, loss.dirichlet_indices, loss.dirichlet_values, params)
    ref_mean, (rmin, rmax, _), Eb = assembly.batch_loss(*args, exponent=exponent)
    tol = 1e-12 * abs(Eb).max()
    assert abs(mean.item() - ref_mean) <= tol and abs(mn.item() - rmin) <= tol and abs(mx.item() - rmax) <= tol
    (1.5 * mean).backward()
    gN, gC = assembly.batch_loss_grads(*args, exponent=exponent)
    gn, gc = nt.grad.cpu().numpy() / 1.5, ct.grad.cpu().numpy() / 1.5
    assert np.abs(gn - gN).max() <= 1e-12 * np.abs(gN).max()
    assert np.abs(gc - gC).max() <= 1e-12 * np.abs(gC).max()
    assert not gn[:, loss.dirichlet_indices].any()
