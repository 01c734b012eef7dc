"""Host-side integer plan of the CSR hand-off against SciPy's structure (CPU)."""
import numpy as np
import scipy.sparse as sp

import folax_b200
from folax_b200 import csr_plan, energy_plan
from oracle import assembly


def test_csr_plan_structure_matches_scipy():
    mesh = folax_b200.create_3D_tetra_box_mesh(2, 3, 2, 1, 1, 1)
    conn, nn = mesh.GetElementsNodes("tetra"), mesh.GetNumberOfNodes()
    p = csr_plan.build(conn, nn, 3)
    idx = assembly.bcoo_indices(conn, 3)
    ref = sp.csr_array((np.ones(len(idx)), (idx[:, 0], idx[:, 1])), shape=(3 * nn, 3 * nn))
    ref.sum_duplicates()
    ref.sort_indices()
    assert np.array_equal(p["indptr"], ref.indptr) and np.array_equal(p["indices"], ref.indices)
    # every BCOO entry is referenced exactly once, in ascending order within a pair
    assert np.array_equal(np.sort(p["contrib"]), np.arange(len(conn) * 16))
    for a, b in zip(p["pair_ptr"][:-1], p["pair_ptr"][1:]):
        assert (np.diff(p["contrib"][a:b]) > 0).all()


def test_energy_tile_plan_is_consistent():
    mesh = folax_b200.perturb_interior_nodes(folax_b200.create_2D_square_mesh(1.0, 40), 0.2)
    conn = mesh.GetElementsNodes("quad")
    p = energy_plan.build(mesh.GetNodesCoordinates(), conn)
    nn = mesh.GetNumberOfNodes()
    assert sorted(p["tile_nodes"]) == list(range(nn))                       # a permutation of the nodes
    assert p["tile_node_ptr"][-1] == nn and np.diff(p["tile_node_ptr"]).max() <= energy_plan.TILE_NODES
    A = conn.shape[1]
    tile_of = np.empty(nn, int)
    for t in range(p["ntiles"]):
        tile_of[p["tile_nodes"][p["tile_node_ptr"][t]:p["tile_node_ptr"][t + 1]]] = t
    for n in range(0, nn, 37):
        t = tile_of[n]
        elems = p["tile_elems"][p["tile_elem_ptr"][t]:p["tile_elem_ptr"][t + 1]]
        ent = p["adj_local"][p["adj_ptr"][n]:p["adj_ptr"][n + 1]]
        got = sorted((int(elems[e // A]), int(e % A)) for e in ent)
        want = sorted((int(e), int(a)) for e, a in zip(*np.nonzero(conn == n)))
        assert got == want
    assert p["ecap"] == np.diff(p["tile_elem_ptr"]).max()


def test_library_plans_equal_the_numpy_restatements():
    """csr_plan.build / sell_plan.build run in the library (csrc/plan_host.cu, host threads); the NumPy versions are the
    readable restatement: every array of both plans must agree entry for entry, dtypes included."""
    from folax_b200 import sell_plan
    cases = [("hexahedron", folax_b200.create_3D_box_mesh(9, 7, 5, 1, 1, 1), 3),
             ("tetra", folax_b200.create_3D_tetra_box_mesh(5, 4, 6, 1, 1, 1), 3),
             ("quad", folax_b200.create_2D_square_mesh(1.0, 23), 2),
             ("quad", folax_b200.create_2D_square_mesh(1.0, 23), 1),
             ("hexahedron", folax_b200.create_3D_box_mesh(30, 29, 31, 1, 1, 1), 3),   # > 2^14 nodes: the threaded path
             ("hexahedron", folax_b200.create_3D_box_mesh(1, 1, 1, 1, 1, 1), 3)]
    for et, mesh, d in cases:
        conn, nn = mesh.GetElementsNodes(et), mesh.GetNumberOfNodes()
        a, b = csr_plan.build(conn, nn, d), csr_plan.build_numpy(conn, nn, d)
        assert set(a) == set(b)
        for k, v in b.items():
            if isinstance(v, np.ndarray):
                assert a[k].dtype == v.dtype and np.array_equal(a[k], v), (et, d, k)
            else:
                assert a[k] == v, (et, d, k)
        for dd in sorted({1, d}):
            sa = sell_plan.build(a["indptr"], a["indices"], dd)
            sb = sell_plan.build_numpy(b["indptr"], b["indices"], dd)
            assert set(sa) == set(sb)
            for k, v in sb.items():
                if isinstance(v, np.ndarray):
                    assert sa[k].dtype == v.dtype and np.array_equal(sa[k], v), (et, d, dd, k)
                else:
                    assert sa[k] == v or (sa[k] is None and v is None), (et, d, dd, k)
    # a scalar matrix whose rows are not made of node runs: no block columns from either builder
    import scipy.sparse as sp
    A = sp.random(70, 70, density=0.1, random_state=0, format="csr") + sp.identity(70, format="csr")
    A.sort_indices()
    for dd in (1, 2, 3):
        sa, sb = sell_plan.build(A.indptr, A.indices, dd), sell_plan.build_numpy(A.indptr, A.indices, dd)
        assert (sa["node_cols"] is None) == (sb["node_cols"] is None)
        for k in ("slice_ptr", "cols", "src", "diag_src"):
            assert np.array_equal(sa[k], sb[k]), (dd, k)


def test_csr_plan_rejects_bad_connectivity():
    import pytest
    from folax_b200 import _lib
    with pytest.raises(_lib.FolaxError):
        csr_plan.build(np.array([[0, 1, 2, 7]]), 4, 3)     # node 7 of a 4-node mesh
