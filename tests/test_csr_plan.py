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
