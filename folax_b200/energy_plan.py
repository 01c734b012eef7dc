"""Integer tile plan of the fused batched-loss kernel (host side, built once per mesh).

Nodes are split into compact tiles of at most TILE_NODES nodes by recursive coordinate bisection
(balanced: every tile holds about nn / ntiles nodes; the cut axis is the longest extent of the
current box), so every tile is a compact patch; per tile the plan lists the elements touching it
and re-expresses the node->element adjacency in tile-local element indices.  Pure integer work,
deterministic; the adjacency order equals fol_node_adjacency's (ascending e*A + a).
`max_elems` bounds the longest per-tile element list (the kernel's one-thread-per-element fast path
needs it <= its block size): the node target is lowered until the bound holds.
"""
import numpy as np

TILE_NODES = 160
MAX_ELEMS = 192     # block size of the pipelined kernel (csrc/energy2_launch.cuh: ENERGY2_BLOCK)


def _morton_keys(coords, bits=16):
    X = np.asarray(coords, dtype=np.float64)
    lo, hi = X.min(0), X.max(0)
    span = np.where(hi > lo, hi - lo, 1.0)
    q = np.minimum(((X - lo) / span * (2 ** bits - 1)).astype(np.uint64), 2 ** bits - 1)
    key = np.zeros(len(X), dtype=np.uint64)
    for b in range(bits):
        for d in range(3):
            key |= ((q[:, d] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + d)
    return key


def _rcb_tiles(coords, tile_nodes):
    """tile id per node: recursive coordinate bisection into ceil(nn / tile_nodes) balanced leaves."""
    X = np.asarray(coords, dtype=np.float64)
    nn = len(X)
    tile_of_node = np.zeros(nn, dtype=np.int64)
    ntiles = max(1, -(-nn // tile_nodes))
    stack = [(np.arange(nn), ntiles, 0)]
    while stack:
        idx, k, first = stack.pop()
        if k == 1:
            tile_of_node[idx] = first
            continue
        k1 = k // 2
        n1 = int(round(len(idx) * k1 / k))
        n1 = min(k1 * tile_nodes, max(len(idx) - (k - k1) * tile_nodes, n1))
        P = X[idx]
        axis = int(np.argmax(P.max(0) - P.min(0))) if len(idx) else 0
        order = np.argsort(P[:, axis], kind="stable")                 # ties: ascending node id
        stack.append((idx[order[:n1]], k1, first))
        stack.append((idx[order[n1:]], k - k1, first + k1))
    return tile_of_node, ntiles


def _recompute(plan, ne):
    return float(plan["tile_elem_ptr"][-1]) / max(ne, 1)


def build(coords, conn, tile_nodes=TILE_NODES, max_elems=MAX_ELEMS, method="rcb", generic_max_elems=None):
    """Returns dict of int32 arrays: adj_ptr, adj_local, tile_node_ptr, tile_nodes, tile_elem_ptr,
    tile_elems, tile_conn, tile_lnode_ptr, tile_lnodes and ints ntiles, ecap, lcap, ncap (longest
    per-tile element / local-node / owned-node list).
    max_elems: bound wanted by the pipelined kernel (one thread per tile element; None = not applicable).
    The bounded plan is kept unless its smaller tiles recompute clearly more border elements than the
    unbounded one (3-D / high-valence meshes); the generic kernel then loops over the element list, which
    only has to fit its shared-memory rows (generic_max_elems)."""
    ne = len(conn)
    affine = is_affine(coords, conn)

    def tag(plan):
        plan["affine"] = affine
        return plan
    return tag(_choose(coords, conn, tile_nodes, max_elems, method, generic_max_elems, ne))


def is_affine(coords, conn):
    """True when every Quad4 is a parallelogram (x0 - x1 + x2 - x3 = 0 to rounding): the isoparametric map is affine
    and the Jacobian the same at every Gauss point (FOL_MESH_AFFINE of include/folax_b200.h).  Other element types:
    False (Tri3 / Tet4 are always affine but their kernels already keep one gradient set per element)."""
    conn = np.asarray(conn)
    if conn.ndim != 2 or conn.shape[1] != 4 or len(conn) == 0 or np.asarray(coords).shape[1] < 2:
        return False
    X = np.asarray(coords, dtype=np.float64)[:, :2][conn]
    if np.asarray(coords).shape[1] > 2 and np.ptp(np.asarray(coords)[:, 2]) != 0.0:
        return False                                     # a 4-node element of a 3-D mesh is a tetrahedron, not a quad
    skew = np.abs(X[:, 0] - X[:, 1] + X[:, 2] - X[:, 3]).max(axis=1)
    size = np.abs(X[:, 2] - X[:, 0]).max(axis=1)
    return bool((skew <= 1e-13 * size).all())


def grid_structure(coords, conn, rtol=1e-12):
    """Facts behind fol_energy_and_grads_grid (csrc/energy_grid.cu), or None: the Quad4 mesh is an nx x ny grid with
    row-major node numbers (node(c, r) = r (nx + 1) + c), elements [n, n + 1, n + nx + 2, n + nx + 1] in row-major
    order (what usefull_functions.py:213-258 builds) and ONE element shape, a parallelogram, to `rtol` of its size.
    Returns {"nx", "ny", "jinv" (row-major d xi_j / d x_k), "wdetj" (2 x 2 rule: weight 1 x det J)}."""
    conn = np.asarray(conn)
    X = np.asarray(coords, dtype=np.float64)
    if conn.ndim != 2 or conn.shape[1] != 4 or len(conn) == 0 or X.shape[1] < 2:
        return None
    if X.shape[1] > 2 and np.ptp(X[:, 2]) != 0.0:
        return None
    ne, nn = len(conn), len(X)
    nx = int(conn[0, 3] - conn[0, 0]) - 1                    # node stride between rows = nx + 1
    if nx < 1 or ne % nx != 0:
        return None
    ny = ne // nx
    if (nx + 1) * (ny + 1) != nn:
        return None
    r, c = np.divmod(np.arange(ne, dtype=np.int64), nx)
    n0 = r * (nx + 1) + c
    want = np.stack([n0, n0 + 1, n0 + nx + 2, n0 + nx + 1], axis=1)
    if not np.array_equal(conn.astype(np.int64), want):
        return None
    P = X[:, :2][conn]                                        # (ne, 4, 2)
    d_xi = ((P[:, 1] - P[:, 0]) + (P[:, 2] - P[:, 3])) / 4.0  # d x / d xi
    d_eta = ((P[:, 3] - P[:, 0]) + (P[:, 2] - P[:, 1])) / 4.0
    skew = np.abs(P[:, 0] - P[:, 1] + P[:, 2] - P[:, 3]).max()
    size = max(np.abs(d_xi[0]).max(), np.abs(d_eta[0]).max())
    if size == 0.0 or skew > rtol * size:
        return None
    if np.abs(d_xi - d_xi[0]).max() > rtol * size or np.abs(d_eta - d_eta[0]).max() > rtol * size:
        return None
    # one shape for all elements: the mean edge vectors (exact zeros stay exact zeros: axis-aligned grids keep a
    # diagonal J^-1, which the kernel exploits)
    a, b = d_xi.mean(axis=0), d_eta.mean(axis=0)
    a = np.where(np.abs(d_xi).max(axis=0) == 0.0, 0.0, a)
    b = np.where(np.abs(d_eta).max(axis=0) == 0.0, 0.0, b)
    A = np.array([[a[0], b[0]], [a[1], b[1]]])                # A[k][j] = d x_k / d xi_j
    det = A[0, 0] * A[1, 1] - A[0, 1] * A[1, 0]
    if not det > 0.0:
        return None
    jinv = np.array([A[1, 1], -A[0, 1], -A[1, 0], A[0, 0]]) / det   # row-major d xi_j / d x_k
    return {"nx": nx, "ny": ny, "jinv": jinv, "wdetj": float(det)}


def _choose(coords, conn, tile_nodes, max_elems, method, generic_max_elems, ne):
    target, generic = tile_nodes, None
    for _ in range(12):
        generic = _build(coords, conn, target, method)
        if generic_max_elems is None or generic["ecap"] <= generic_max_elems or target <= 4:
            break
        target = max(4, min(target - 1, int(target * generic_max_elems / generic["ecap"])))
    if max_elems is None or generic["ecap"] <= max_elems:
        return generic
    target = tile_nodes
    for _ in range(10):
        target = max(4, min(target - 1, int(target * max_elems / generic["ecap"] if target == tile_nodes
                                            else target * max_elems / plan["ecap"])))
        plan = _build(coords, conn, target, method)
        if plan["ecap"] <= max_elems:
            return plan if _recompute(plan, ne) <= 1.15 * _recompute(generic, ne) else generic
        if target <= 4:
            break
    return generic


def _build(coords, conn, tile_nodes, method):
    conn = np.asarray(conn, dtype=np.int64)
    ne, A = conn.shape
    nn = len(coords)
    if method == "rcb":
        tile_of_node, ntiles = _rcb_tiles(coords, tile_nodes)
    else:
        morton = np.argsort(_morton_keys(coords), kind="stable")     # compact tiles: cut the Z-curve
        tile_of_node = np.empty(nn, dtype=np.int64)
        tile_of_node[morton] = np.arange(nn) // tile_nodes
        ntiles = int((nn + tile_nodes - 1) // tile_nodes)
    # inside a tile, nodes are kept in ascending global id: neighbouring threads then touch
    # neighbouring addresses (coalesced gradient writes, conflict-free shared-memory gathers)
    order = np.lexsort((np.arange(nn), tile_of_node))
    rank = np.empty(nn, dtype=np.int64)
    rank[order] = np.arange(nn)
    tile_node_ptr = np.concatenate([[0], np.cumsum(np.bincount(tile_of_node, minlength=ntiles))])
    # adjacency sorted by node, then by e*A + a (stable sort keeps ascending entry ids)
    flat_nodes = conn.reshape(-1)
    entries = np.argsort(flat_nodes, kind="stable")                   # values: e*A + a
    counts = np.bincount(flat_nodes, minlength=nn)
    adj_ptr = np.concatenate([[0], np.cumsum(counts)])
    node_of_entry = flat_nodes[entries]
    elem_of_entry = entries // A
    a_of_entry = entries - elem_of_entry * A
    key = tile_of_node[node_of_entry] * ne + elem_of_entry          # (tile, element) pairs
    ukeys = np.unique(key)                                            # sorted
    tile_elems = ukeys % ne
    tile_of_ukey = ukeys // ne
    tile_elem_ptr = np.searchsorted(tile_of_ukey, np.arange(ntiles + 1))
    local = np.searchsorted(ukeys, key) - tile_elem_ptr[tile_of_node[node_of_entry]]
    adj_local = local * A + a_of_entry
    ecap = int(np.diff(tile_elem_ptr).max()) if ntiles else 1
    # local node lists: the tile's own nodes first (tile order), then the halo nodes its elements touch
    tile_of_telem = np.repeat(np.arange(ntiles), np.diff(tile_elem_ptr))
    tnodes = conn[tile_elems]                                         # (len(tile_elems), A) global ids
    owned = tile_of_node[tnodes] == tile_of_telem[:, None]
    halo_key = np.unique((tile_of_telem[:, None] * nn + tnodes)[~owned])          # (tile, node) pairs, sorted
    halo_tile, halo_node = halo_key // nn, halo_key % nn
    n_owned = np.diff(tile_node_ptr)
    n_halo = np.bincount(halo_tile, minlength=ntiles)
    tile_lnode_ptr = np.concatenate([[0], np.cumsum(n_owned + n_halo)])
    tile_lnodes = np.empty(int(tile_lnode_ptr[-1]), dtype=np.int64)
    halo_ptr = np.concatenate([[0], np.cumsum(n_halo)])
    own_pos = tile_lnode_ptr[:-1][tile_of_node[order]] + (rank[order] - tile_node_ptr[:-1][tile_of_node[order]])
    tile_lnodes[own_pos] = order
    halo_pos = tile_lnode_ptr[:-1][halo_tile] + n_owned[halo_tile] + (np.arange(len(halo_key)) - halo_ptr[halo_tile])
    tile_lnodes[halo_pos] = halo_node
    # element connectivity in tile-local numbering
    local_owned = rank[tnodes] - tile_node_ptr[:-1][tile_of_telem][:, None]
    local_halo = n_owned[tile_of_telem][:, None] + (np.searchsorted(halo_key, tile_of_telem[:, None] * nn + tnodes)
                                                    - halo_ptr[tile_of_telem][:, None])
    tile_conn = np.where(owned, local_owned, local_halo)
    lcap = int(np.diff(tile_lnode_ptr).max()) if ntiles else 1
    i32 = lambda x: np.ascontiguousarray(x, dtype=np.int32)
    return {"adj_ptr": i32(adj_ptr), "adj_local": i32(adj_local), "tile_node_ptr": i32(tile_node_ptr),
            "tile_nodes": i32(order), "tile_elem_ptr": i32(tile_elem_ptr), "tile_elems": i32(tile_elems),
            "tile_conn": i32(tile_conn), "tile_lnode_ptr": i32(tile_lnode_ptr), "tile_lnodes": i32(tile_lnodes),
            "ntiles": ntiles, "ecap": max(ecap, 1), "lcap": max(lcap, 1),
            "ncap": max(int(np.diff(tile_node_ptr).max()) if ntiles else 1, 1)}
