"""GPU parity of the drop-in for the reference's own FFI loss class, KratosSmallDisplacement3DTetra
(fol/loss_functions/kratos_small_displacement.py): same constructor and calls as
tests/unit/test_kratos_ffi_mechanical_loss.py:29-72."""
import json
import os

import numpy as np
import pytest

import folax_b200
from folax_b200.loss_functions import KratosSmallDisplacement3DTetra
from oracle import assembly
from tests.test_oracle_golden import _tf32

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_test_tetra():
    with open(os.path.join(ROOT, "tests", "golden", "reference_unit_goldens.json")) as fh:
        rec = json.load(fh)["tests/unit/test_kratos_ffi_mechanical_loss.py"]["test_tetra"]
    X = np.array(rec["assign"]["tet_points_coordinates"], float)
    fe_mesh = folax_b200.Mesh("", ".")
    fe_mesh.node_ids = np.arange(len(X))
    fe_mesh.nodes_coordinates = X
    fe_mesh.elements_nodes = {"tetra": fe_mesh.node_ids.reshape(1, -1)}
    loss = KratosSmallDisplacement3DTetra("mechanical_loss_3d", loss_settings={
        "dirichlet_bc_dict": {"Ux": {}, "Uy": {}, "Uz": {}}, "material_dict": {"young_modulus": 1, "poisson_ratio": 0.3},
        "body_foce": np.array([[0], [0], [0]])}, fe_mesh=fe_mesh)
    loss.Initialize()
    batch = np.random.default_rng(42).uniform(size=(5, 12))      # the reference draws these with JAX's PRNG
    jac, res = loss.ComputeJacobianMatrixAndResidualVector(batch[0], batch[0])
    dense = jac.todense().cpu().numpy()
    golden = np.array([a for a in rec["asserts"] if "jac" in a["expr"]][0]["value"])
    # the golden is the TF32 rounding of the float32 stiffness (tests/test_oracle_golden.py)
    assert np.abs(_tf32(dense.reshape(-1)) - golden).max() <= 1e-12
    np.testing.assert_allclose(dense.reshape(-1), golden, rtol=2.0 ** -10, atol=1e-5)
    # residual and energy of this linear element: R = K u, E = u . R; the control argument is ignored
    assert np.abs(res.cpu().numpy() - dense @ batch[0]).max() <= 1e-14
    for b in range(5):
        en = float(loss.ComputeTotalEnergy(batch[b], batch[b]))
        assert abs(en - batch[b] @ dense @ batch[b]) <= 1e-13
        assert en == float(loss.ComputeTotalEnergy(np.zeros(4), batch[b]))
    with pytest.raises(SystemExit):
        loss.ComputeElement(X, np.ones(4), np.ones(12))


def test_mesh_equals_mechanical_tetra_with_unit_control():
    mesh = folax_b200.create_3D_tetra_box_mesh(3, 3, 3, 1.0, 1.0, 1.0)
    folax_b200.perturb_interior_nodes(mesh, 0.2, 2)
    bc = {d: {"left": 0.0, "right": 0.1} for d in ("Ux", "Uy", "Uz")}
    mat = {"young_modulus": 2.5, "poisson_ratio": 0.25}
    loss = KratosSmallDisplacement3DTetra("k", {"dirichlet_bc_dict": bc, "material_dict": dict(mat),
                                                "body_foce": [1.0, 2.0, 3.0]}, mesh)     # ignored, as by the FFI call
    loss.Initialize()
    coords, conn = np.asarray(mesh.GetNodesCoordinates()), mesh.GetElementsNodes("tetra")
    rng = np.random.default_rng(1)
    u = 0.01 * rng.standard_normal(loss.total_number_of_dofs)
    for transpose in (False, True):
        jac, R = loss.ComputeJacobianMatrixAndResidualVector(rng.uniform(0.1, 1, len(coords)), u, transpose)
        data, idx, Rref = assembly.assemble("mechanical", "tetra", 1, coords, conn, np.ones(len(coords)), u,
                                            loss.dirichlet_indices, mat, transpose)
        assert np.array_equal(jac.indices.cpu().numpy(), idx)
        assert np.abs(jac.data.cpu().numpy() - data).max() <= 1e-12 * np.abs(data).max()
        assert np.abs(R.cpu().numpy() - Rref).max() <= 1e-12 * np.abs(Rref).max()
