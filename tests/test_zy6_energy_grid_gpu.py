"""Structured-grid thermal loss + VJP kernel (csrc/energy_grid.cu, fol_energy_and_grads_grid) against the oracle
(oracle/assembly.py::batch_loss / batch_loss_grads: fe_loss.py:250-262, thermal.py:28-49 and the analytic cotangents)
and against the tile kernels on the same inputs.  Grids chosen to hit every ownership rule of the kernel: widths below,
at and above a warp / a panel (1, 31, 32, 33, 64, 100, 256, 257, 300 element columns), heights around the chunk size,
sheared parallelograms (full J^-1) and axis-aligned ones (diagonal J^-1), Dirichlet rows fused into the staging or
pre-applied, with and without the control gradient, both precisions, exponent 1 and 2."""
import os

import numpy as np
import pytest
import torch

import folax_b200
from folax_b200 import energy_plan
from folax_b200.loss_functions import MechanicalLoss2DQuad, ThermalLoss2DQuad
from folax_b200.mesh import Mesh, _finish
from oracle import assembly

pytestmark = pytest.mark.gpu


def grid_mesh(nx, ny, hx=0.1, hy=0.07, shear=0.0):
    """nx x ny Quad4 grid, row-major nodes, element [n, n+1, n+nx+2, n+nx+1]; `shear` tilts the rows (parallelograms)."""
    c, r = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1))
    coords = np.stack([(c * hx + shear * r * hy).ravel(), (r * hy).ravel(), np.zeros((nx + 1) * (ny + 1))], axis=1)
    e_r, e_c = np.divmod(np.arange(nx * ny), nx)
    n0 = e_r * (nx + 1) + e_c
    conn = np.stack([n0, n0 + 1, n0 + nx + 2, n0 + nx + 1], axis=1)
    ids = np.arange((nx + 1) * (ny + 1))
    sets = {"left": ids[ids % (nx + 1) == 0], "right": ids[ids % (nx + 1) == nx]}
    return _finish(Mesh("grid_io", "grid."), coords, conn, "quad", sets)


def make(mesh, dtype="float64", beta=2.0, c=4, exponent=1.0, bc=None):
    loss = ThermalLoss2DQuad("t", {"dirichlet_bc_dict": bc or {"T": {"left": 1.0, "right": 0.1}}, "beta": beta, "c": c,
                                   "dtype": dtype, "loss_function_exponent": exponent}, mesh)
    loss.Initialize()
    return loss


def oracle(loss, mesh, K, u, exponent=1.0):
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("quad")
    args = ("thermal", "quad", 2, coords, conn, K, u, loss.dirichlet_indices, loss.dirichlet_values,
            {"beta": loss.thermal_loss_settings["beta"], "c": loss.thermal_loss_settings["c"]})
    mean, _, Eb = assembly.batch_loss(*args, exponent=exponent)
    gU, gK = assembly.batch_loss_grads(*args, exponent=exponent)
    return mean, Eb, gU, gK


def run(loss, K, u):
    Kt = torch.tensor(K, device="cuda", dtype=loss.dtype, requires_grad=True)
    ut = torch.tensor(u, device="cuda", dtype=loss.dtype, requires_grad=True)
    mean, _ = loss.ComputeBatchLoss(Kt, ut)
    mean.backward()
    return mean.item(), ut.grad.cpu().numpy(), Kt.grad.cpu().numpy()


SHAPES = [(1, 1), (5, 3), (31, 9), (32, 8), (33, 17), (64, 40), (100, 7), (256, 12), (257, 5), (300, 33), (16, 130)]


@pytest.mark.parametrize("nx,ny", SHAPES)
@pytest.mark.parametrize("shear", [0.0, 0.3])
def test_grid_kernel_matches_the_oracle(nx, ny, shear):
    mesh = grid_mesh(nx, ny, shear=shear)
    loss = make(mesh)
    assert loss._grid_plan() is not None and (loss._grid_plan()["nx"], loss._grid_plan()["ny"]) == (nx, ny)
    rng = np.random.default_rng(nx * 1000 + ny)
    B = 3
    K, u = rng.uniform(0.1, 1.0, (B, mesh.GetNumberOfNodes())), rng.uniform(0.1, 1.0, (B, mesh.GetNumberOfNodes()))
    mean, gu, gk = run(loss, K, u)
    ref_mean, Eb, gU, gK = oracle(loss, mesh, K, u)
    assert abs(mean - ref_mean) <= 1e-12 * np.abs(Eb).max()
    assert np.abs(gu - gU).max() <= 1e-12 * np.abs(gU).max()
    assert np.abs(gk - gK).max() <= 1e-12 * np.abs(gK).max()
    assert not gu[:, loss.dirichlet_indices].any()
    # deterministic
    mean2, gu2, gk2 = run(loss, K, u)
    assert mean2 == mean and np.array_equal(gu2, gu) and np.array_equal(gk2, gk)


@pytest.mark.parametrize("rows", [8, 11, 64])
def test_every_chunk_height_gives_the_same_sums(rows, monkeypatch):
    """The chunk height only moves the ownership of node rows between CTAs: bit-identical gradients; the energy
    shares are summed per warp and chunk, so the energies agree to rounding."""
    mesh = grid_mesh(70, 45)
    rng = np.random.default_rng(1)
    K, u = rng.uniform(0.1, 1.0, (2, mesh.GetNumberOfNodes())), rng.uniform(0.1, 1.0, (2, mesh.GetNumberOfNodes()))
    base = run(make(mesh), K, u)
    # FOL_ENERGY_GRID_ROWS is read once per process by the launcher: compare through the C ABI in a subprocess
    import subprocess, sys, json
    code = (
        "import numpy as np, torch, json, sys\n"
        "sys.path.insert(0, %r)\n"
        "from tests.test_zy6_energy_grid_gpu import grid_mesh, make, run\n"
        "mesh = grid_mesh(70, 45)\n"
        "rng = np.random.default_rng(1)\n"
        "K, u = rng.uniform(0.1, 1.0, (2, mesh.GetNumberOfNodes())), rng.uniform(0.1, 1.0, (2, mesh.GetNumberOfNodes()))\n"
        "m, gu, gk = run(make(mesh), K, u)\n"
        "print(json.dumps([m, float(np.abs(gu).sum()), float(np.abs(gk).sum()), gu.tobytes().hex()[:64]]))\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FOL_ENERGY_GRID_ROWS=str(rows))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    m, su, sk, _ = json.loads(out.stdout.strip().splitlines()[-1])
    assert abs(m - base[0]) <= 1e-13 * abs(base[0])
    assert su == float(np.abs(base[1]).sum()) and sk == float(np.abs(base[2]).sum())


def test_grid_kernel_equals_the_tile_kernels(monkeypatch):
    """Same mesh through both routes (FOL_ENERGY_GRID=0 keeps the tile kernels): equal to rounding."""
    mesh = folax_b200.create_2D_square_mesh(1.0, 48)
    rng = np.random.default_rng(3)
    K, u = rng.uniform(0.1, 1.0, (4, 48 * 48)), rng.uniform(0.1, 1.0, (4, 48 * 48))
    grid = run(make(mesh), K, u)
    monkeypatch.setenv("FOL_ENERGY_GRID", "0")
    loss = make(mesh)
    assert loss._grid_plan() is None
    tile = run(loss, K, u)
    assert abs(grid[0] - tile[0]) <= 1e-13 * abs(tile[0])
    assert np.abs(grid[1] - tile[1]).max() <= 1e-13 * np.abs(tile[1]).max()
    assert np.abs(grid[2] - tile[2]).max() <= 1e-13 * np.abs(tile[2]).max()


@pytest.mark.parametrize("beta,c", [(0.0, 1), (1.5, 1), (0.7, 2), (2.0, 3), (2.0, 4), (0.5, 2.5)])
@pytest.mark.parametrize("exponent", [1.0, 2.0])
def test_conductivity_laws_and_exponent(beta, c, exponent):
    mesh = grid_mesh(40, 21)
    loss = make(mesh, beta=beta, c=c, exponent=exponent)
    rng = np.random.default_rng(5)
    K, u = rng.uniform(0.1, 1.0, (3, mesh.GetNumberOfNodes())), rng.uniform(0.1, 1.0, (3, mesh.GetNumberOfNodes()))
    mean, gu, gk = run(loss, K, u)
    ref_mean, Eb, gU, gK = oracle(loss, mesh, K, u, exponent)
    assert abs(mean - ref_mean) <= 1e-12 * (np.abs(Eb) ** exponent).max()
    assert np.abs(gu - gU).max() <= 1e-12 * np.abs(gU).max()
    assert np.abs(gk - gK).max() <= 1e-12 * np.abs(gK).max()


def test_grid_kernel_float32():
    mesh = grid_mesh(96, 50)
    loss = make(mesh, dtype="float32")
    rng = np.random.default_rng(6)
    K = rng.uniform(0.1, 1.0, (4, mesh.GetNumberOfNodes())).astype(np.float32)
    u = rng.uniform(0.1, 1.0, (4, mesh.GetNumberOfNodes())).astype(np.float32)
    mean, gu, gk = run(loss, K, u)
    ref_mean, Eb, gU, gK = oracle(loss, mesh, K.astype(np.float64), u.astype(np.float64))
    assert abs(mean - ref_mean) <= 1e-5 * np.abs(Eb).max()
    assert np.abs(gu - gU).max() <= 1e-5 * np.abs(gU).max()
    assert np.abs(gk - gK).max() <= 1e-5 * np.abs(gK).max()


@pytest.mark.parametrize("B", [1, 2, 3, 7])
@pytest.mark.parametrize("shear", [0.0, 0.25])
def test_grid_kernel_float32_sample_pairs(B, shear):
    """float32 runs two samples per lane on the packed FP32 instructions: even, odd and single-sample batches (an odd
    batch evaluates its last sample twice and stores it once), each sample against the float64 oracle."""
    mesh = grid_mesh(70, 23, shear=shear)
    loss = make(mesh, dtype="float32")
    rng = np.random.default_rng(60 + B)
    K = rng.uniform(0.1, 1.0, (B, mesh.GetNumberOfNodes())).astype(np.float32)
    u = rng.uniform(0.1, 1.0, (B, mesh.GetNumberOfNodes())).astype(np.float32)
    mean, gu, gk = run(loss, K, u)
    ref_mean, Eb, gU, gK = oracle(loss, mesh, K.astype(np.float64), u.astype(np.float64))
    assert abs(mean - ref_mean) <= 1e-5 * np.abs(Eb).max()
    for b in range(B):
        assert np.abs(gu[b] - gU[b]).max() <= 1e-5 * np.abs(gU).max(), b
        assert np.abs(gk[b] - gK[b]).max() <= 1e-5 * np.abs(gK).max(), b
    # the pairing does not change a sample's result: sample 0 alone gives the same bits
    if B > 1:
        _, gu1, gk1 = run(make(mesh, dtype="float32"), K[:1], u[:1])
        scale = np.float32(1.0 / B)                         # the kernel writes fl(scale * R) with scale = 1 / B
        assert np.array_equal((scale * gu1[0]).astype(np.float32), gu[0])
        assert np.array_equal((scale * gk1[0]).astype(np.float32), gk[0])


def test_dirichlet_on_interior_and_top_rows_float32():
    nx, ny = 70, 20
    mesh = grid_mesh(nx, ny)
    ids = np.arange((nx + 1) * (ny + 1))
    mesh.node_sets["left"] = ids[(ids % (nx + 1) == 32) | (ids // (nx + 1) == ny) | (ids == 5)].astype(np.int32)
    mesh.node_sets["right"] = ids[(ids % (nx + 1) == 64) & (ids // (nx + 1) < ny)].astype(np.int32)
    loss = make(mesh, dtype="float32")
    rng = np.random.default_rng(9)
    K = rng.uniform(0.1, 1.0, (3, len(ids))).astype(np.float32)
    u = rng.uniform(0.1, 1.0, (3, len(ids))).astype(np.float32)
    mean, gu, gk = run(loss, K, u)
    ref_mean, Eb, gU, gK = oracle(loss, mesh, K.astype(np.float64), u.astype(np.float64))
    assert abs(mean - ref_mean) <= 1e-5 * np.abs(Eb).max()
    assert np.abs(gu - gU).max() <= 1e-5 * np.abs(gU).max()
    assert np.abs(gk - gK).max() <= 1e-5 * np.abs(gK).max()
    assert not gu[:, loss.dirichlet_indices].any()


def test_dirichlet_on_interior_and_top_rows():
    """Dirichlet nodes anywhere (not only on the left / right columns): overwrite while staging, cut at the store,
    also on the columns two warps share and on the top row."""
    nx, ny = 70, 20
    mesh = grid_mesh(nx, ny)
    ids = np.arange((nx + 1) * (ny + 1))
    mesh.node_sets["left"] = ids[(ids % (nx + 1) == 32) | (ids // (nx + 1) == ny) | (ids == 5)].astype(np.int32)
    mesh.node_sets["right"] = ids[(ids % (nx + 1) == 64) & (ids // (nx + 1) < ny)].astype(np.int32)
    loss = make(mesh)
    rng = np.random.default_rng(8)
    K, u = rng.uniform(0.1, 1.0, (2, len(ids))), rng.uniform(0.1, 1.0, (2, len(ids)))
    mean, gu, gk = run(loss, K, u)
    ref_mean, Eb, gU, gK = oracle(loss, mesh, K, u)
    assert abs(mean - ref_mean) <= 1e-12 * np.abs(Eb).max()
    assert np.abs(gu - gU).max() <= 1e-12 * np.abs(gU).max()
    assert np.abs(gk - gK).max() <= 1e-12 * np.abs(gK).max()
    assert not gu[:, loss.dirichlet_indices].any()


# ---------------------------------------------------------------------------------------------- elasticity (two dofs / node)
MAT = {"young_modulus": 1.3, "poisson_ratio": 0.3}


def make_mech(mesh, dtype="float64", exponent=1.0, body=None):
    settings = {"dirichlet_bc_dict": {"Ux": {"left": 0.0, "right": 0.05}, "Uy": {"left": 0.0, "right": -0.02}},
                "material_dict": dict(MAT), "dtype": dtype, "loss_function_exponent": exponent}
    if body is not None:
        settings["body_foce"] = body
    loss = MechanicalLoss2DQuad("m", settings, mesh)
    loss.Initialize()
    return loss


def oracle_mech(loss, mesh, K, u, exponent=1.0):
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("quad")
    params = {"young_modulus": MAT["young_modulus"], "poisson_ratio": MAT["poisson_ratio"]}
    if "body_foce" in loss.loss_settings:
        params["body_force"] = np.asarray(loss.loss_settings["body_foce"], float).reshape(-1)
    args = ("mechanical", "quad", 2, coords, conn, K, u, loss.dirichlet_indices, loss.dirichlet_values, params)
    mean, _, Eb = assembly.batch_loss(*args, exponent=exponent)
    gU, _ = assembly.batch_loss_grads(*args, exponent=exponent)
    return mean, Eb, gU


def run_mech(loss, K, u):
    Kt = torch.tensor(K, device="cuda", dtype=loss.dtype, requires_grad=True)
    ut = torch.tensor(u, device="cuda", dtype=loss.dtype, requires_grad=True)
    mean, _ = loss.ComputeBatchLoss(Kt, ut)
    mean.backward()
    assert Kt.grad is None or not Kt.grad.any()            # dE/dK = 0: Se sits under stop_gradient (mechanical.py:116)
    return mean.item(), ut.grad.cpu().numpy()


@pytest.mark.parametrize("nx,ny", SHAPES)
@pytest.mark.parametrize("shear,body", [(0.0, None), (0.3, [0.3, -0.2])])
def test_mech_grid_kernel_matches_the_oracle(nx, ny, shear, body):
    mesh = grid_mesh(nx, ny, shear=shear)
    loss = make_mech(mesh, body=body)
    assert loss._grid_plan() is not None
    rng = np.random.default_rng(nx * 1000 + ny + 7)
    B, nn = 3, mesh.GetNumberOfNodes()
    K, u = rng.uniform(0.1, 1.0, (B, nn)), 0.01 * rng.standard_normal((B, 2 * nn))
    mean, gu = run_mech(loss, K, u)
    ref_mean, Eb, gU = oracle_mech(loss, mesh, K, u)
    assert abs(mean - ref_mean) <= 1e-12 * np.abs(Eb).max()
    assert np.abs(gu - gU).max() <= 1e-12 * np.abs(gU).max()
    assert not gu[:, loss.dirichlet_indices].any()
    mean2, gu2 = run_mech(loss, K, u)
    assert mean2 == mean and np.array_equal(gu2, gu)       # deterministic


def test_mech_grid_kernel_equals_the_tile_kernels(monkeypatch):
    mesh = folax_b200.create_2D_square_mesh(1.0, 48)
    rng = np.random.default_rng(4)
    K, u = rng.uniform(0.1, 1.0, (4, 48 * 48)), 0.01 * rng.standard_normal((4, 2 * 48 * 48))
    grid = run_mech(make_mech(mesh, body=[0.1, 0.2]), K, u)
    monkeypatch.setenv("FOL_ENERGY_GRID", "0")
    loss = make_mech(mesh, body=[0.1, 0.2])
    assert loss._grid_plan() is None
    tile = run_mech(loss, K, u)
    assert abs(grid[0] - tile[0]) <= 1e-13 * abs(tile[0])
    assert np.abs(grid[1] - tile[1]).max() <= 1e-13 * np.abs(tile[1]).max()


@pytest.mark.parametrize("B", [1, 2, 5])
@pytest.mark.parametrize("exponent", [1.0, 2.0])
def test_mech_grid_kernel_float32_and_exponent(B, exponent):
    mesh = grid_mesh(70, 23, shear=0.2)
    loss = make_mech(mesh, dtype="float32", exponent=exponent, body=[0.2, 0.1])
    rng = np.random.default_rng(70 + B)
    nn = mesh.GetNumberOfNodes()
    K = rng.uniform(0.1, 1.0, (B, nn)).astype(np.float32)
    u = (0.01 * rng.standard_normal((B, 2 * nn))).astype(np.float32)
    mean, gu = run_mech(loss, K, u)
    ref_mean, Eb, gU = oracle_mech(loss, mesh, K.astype(np.float64), u.astype(np.float64), exponent)
    assert abs(mean - ref_mean) <= 2e-5 * (np.abs(Eb) ** exponent).max()
    assert np.abs(gu - gU).max() <= 2e-5 * np.abs(gU).max()


def test_mech_dirichlet_dofs_anywhere():
    """Dirichlet dofs on single components of interior nodes, on the columns two warps share and on the top row."""
    nx, ny = 70, 20
    mesh = grid_mesh(nx, ny)
    ids = np.arange((nx + 1) * (ny + 1))
    mesh.node_sets["left"] = ids[(ids % (nx + 1) == 32) | (ids // (nx + 1) == ny) | (ids == 5)].astype(np.int32)
    mesh.node_sets["right"] = ids[(ids % (nx + 1) == 64) & (ids // (nx + 1) < ny)].astype(np.int32)
    loss = MechanicalLoss2DQuad("m", {"dirichlet_bc_dict": {"Ux": {"left": 0.01}, "Uy": {"right": -0.02}},
                                      "material_dict": dict(MAT)}, mesh)
    loss.Initialize()
    rng = np.random.default_rng(8)
    K, u = rng.uniform(0.1, 1.0, (2, len(ids))), 0.01 * rng.standard_normal((2, 2 * len(ids)))
    mean, gu = run_mech(loss, K, u)
    ref_mean, Eb, gU = oracle_mech(loss, mesh, K, u)
    assert abs(mean - ref_mean) <= 1e-12 * np.abs(Eb).max()
    assert np.abs(gu - gU).max() <= 1e-12 * np.abs(gU).max()
    assert not gu[:, loss.dirichlet_indices].any()
