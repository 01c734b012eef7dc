"""Mesh container + structured generators: the input provider of the hot path.

Mirrors the parts of the reference mesh API that the loss classes touch
(fol/mesh_input_output/mesh.py:144-172: GetNumberOfNodes, GetNodesCoordinates,
GetElementsNodes, GetNodeSet, ...) and the numbering of its generators
(fol/tools/usefull_functions.py:196-258).  Host-side, NumPy arrays; the loss classes upload them
once in Initialize().  Only the ASCII Kratos ``.mdpa`` format is read (meshio / gmsh are not
available in this image; other formats raise).
"""
import os

import numpy as np

_MDPA_TYPES = {"Triangle2D3": "triangle", "Triangle3D3": "triangle", "Quadrilateral2D4": "quad",
               "Quadrilateral3D4": "quad", "Tetrahedra3D4": "tetra", "Hexahedra3D8": "hexahedron"}


class Mesh:
    def __init__(self, io_name: str, file_name: str, case_dir: str = ".", scale_factor: float = 1):
        self.__name = io_name
        self.file_name = file_name
        parts = file_name.split(".")
        self.mesh_format = parts[1] if len(parts) > 1 else ""
        self.case_dir = case_dir
        self.scale_factor = scale_factor
        self.node_ids = np.zeros(0, dtype=np.int32)
        self.nodes_coordinates = np.zeros((0, 3))
        self.elements_nodes = {}
        self.node_sets = {}
        self.element_sets = {}
        self.point_data = {}
        self.is_initialized = False

    def GetName(self) -> str:
        return self.__name

    def Initialize(self) -> None:
        if self.is_initialized:
            return
        if self.mesh_format != "mdpa":
            raise NotImplementedError(f"mesh format '{self.mesh_format}' needs meshio, which is not available; "
                                      "use .mdpa or the structured generators")
        self._read_mdpa(os.path.join(self.case_dir, self.file_name))
        self.CheckAndOrientElements()
        self.is_initialized = True

    def _read_mdpa(self, path):
        """Same block structure the reference parses (mesh.py:97-110, 188-237): 1-based ids."""
        with open(path) as fh:
            lines = [ln.strip() for ln in fh]
        i = 0
        while i < len(lines):
            ln = lines[i]
            if ln.startswith("Begin Nodes"):
                rows = []
                i += 1
                while not lines[i].startswith("End Nodes"):
                    if lines[i]:
                        rows.append([float(v) for v in lines[i].split()])
                    i += 1
                data = np.array(rows)
                self.nodes_coordinates = data[:, 1:4] * self.scale_factor
                self.node_ids = np.arange(len(data), dtype=np.int32)
            elif ln.startswith("Begin Elements "):
                etype = next((v for k, v in _MDPA_TYPES.items() if k in ln[15:]), None)
                rows = []
                i += 1
                while not lines[i].startswith("End Elements"):
                    if lines[i]:
                        vals = [int(v) for v in lines[i].split()]
                        rows.append(np.array(vals[2:]) - 1)
                    i += 1
                arr = np.array(rows, dtype=np.int32)
                if etype in self.elements_nodes:
                    arr = np.vstack((self.elements_nodes[etype], arr))
                self.elements_nodes[etype] = arr
            elif ln.startswith("Begin SubModelPart "):
                name = ln[19:]
                if lines[i + 1].startswith("Begin SubModelPartNodes"):
                    ids = []
                    i += 2
                    while not lines[i].startswith("End SubModelPartNodes"):
                        if lines[i]:
                            ids.append(int(lines[i]) - 1)
                        i += 1
                    self.node_sets[name] = np.array(ids, dtype=np.int32)
            i += 1

    def CheckAndOrientElements(self):
        """Swap the first two nodes of elements whose Jacobian determinant at the first Gauss point
        is negative (mesh.py:123-142).  Host-side one-off; tetra / triangle use the constant
        Jacobian, hex / quad the centroid."""
        X = np.asarray(self.nodes_coordinates, dtype=np.float64)
        for etype, conn in list(self.elements_nodes.items()):
            conn = np.array(conn)
            P = X[conn]
            if etype == "tetra":
                J = np.stack([P[:, 1] - P[:, 0], P[:, 2] - P[:, 0], P[:, 3] - P[:, 0]], axis=-1)
            elif etype == "triangle":
                J = np.stack([P[:, 1, :2] - P[:, 0, :2], P[:, 2, :2] - P[:, 0, :2]], axis=-1)
            elif etype == "quad":
                dN = 0.25 * np.array([[-1, -1], [1, -1], [1, 1], [-1, 1.0]])
                J = np.einsum("eai,aj->eij", P[:, :, :2], dN)
            elif etype == "hexahedron":
                sx = np.array([-1, 1, 1, -1, -1, 1, 1, -1.0])
                sy = np.array([-1, -1, 1, 1, -1, -1, 1, 1.0])
                sz = np.array([-1, -1, -1, -1, 1, 1, 1, 1.0])
                J = np.einsum("eai,aj->eij", P, 0.125 * np.stack([sx, sy, sz], axis=1))
            else:
                continue
            neg = np.linalg.det(J) < 0
            if neg.any():
                conn[neg, 0], conn[neg, 1] = conn[neg, 1].copy(), conn[neg, 0].copy()
                self.elements_nodes[etype] = conn

    # getters of mesh.py:144-172
    def GetNodesIds(self):
        return self.node_ids

    def GetNumberOfNodes(self) -> int:
        return len(self.node_ids)

    def GetNodesCoordinates(self):
        return self.nodes_coordinates

    def GetNodesX(self):
        return self.nodes_coordinates[:, 0]

    def GetNodesY(self):
        return self.nodes_coordinates[:, 1]

    def GetNodesZ(self):
        return self.nodes_coordinates[:, 2]

    def GetElementsIds(self, element_type):
        return np.arange(len(self.elements_nodes[element_type]))

    def GetNumberOfElements(self, element_type) -> int:
        return len(self.elements_nodes[element_type])

    def GetElementsNodes(self, element_type):
        return self.elements_nodes[element_type]

    def GetNodeSet(self, set_name):
        return self.node_sets[set_name]

    def HasPointData(self, data_name):
        return data_name in self.point_data

    def __getitem__(self, key):
        return self.point_data[key]

    def __setitem__(self, key, value):
        self.point_data[key] = np.array(value)

    def Finalize(self, export_dir: str = ".", export_format: str = "vtk") -> None:
        pass


def _finish(mesh, coords, conn, etype, sets):
    mesh.node_ids = np.arange(len(coords), dtype=np.int32)
    mesh.nodes_coordinates = np.ascontiguousarray(coords, dtype=np.float64)
    mesh.elements_nodes = {etype: np.ascontiguousarray(conn, dtype=np.int32)}
    mesh.node_sets = {k: np.asarray(v, dtype=np.int32) for k, v in sets.items()}
    mesh.is_initialized = True
    return mesh


def create_2D_square_mesh(L, N):
    """Structured quad mesh with the reference numbering (usefull_functions.py:213-258):
    nodes row-major x-fastest, element [n, n+1, n+nx+1, n+nx], left/right node sets."""
    x = np.linspace(0, L, N)
    Xg, Yg = np.meshgrid(x, x)
    coords = np.stack([Xg.ravel(), Yg.ravel(), np.zeros(N * N)], axis=1)
    i, j = np.meshgrid(np.arange(N - 1), np.arange(N - 1), indexing="ij")
    n0 = (i * N + j).ravel()
    conn = np.stack([n0, n0 + 1, n0 + N + 1, n0 + N], axis=1)
    sets = {"left": np.arange(0, N * N, N), "right": np.arange(N - 1, N * N, N)}
    return _finish(Mesh("square_io", "square."), coords, conn, "quad", sets)


def _box_nodes(Nx, Ny, Nz, Lx, Ly, Lz):
    x, y, z = np.linspace(0, Lx, Nx + 1), np.linspace(0, Ly, Ny + 1), np.linspace(0, Lz, Nz + 1)
    Zg, Yg, Xg = np.meshgrid(z, y, x, indexing="ij")
    coords = np.stack([Xg.ravel(), Yg.ravel(), Zg.ravel()], axis=1)
    k, j, i = np.meshgrid(np.arange(Nz), np.arange(Ny), np.arange(Nx), indexing="ij")
    n0 = (i + (Nx + 1) * (j + (Ny + 1) * k)).ravel()
    return coords, n0, (Nx + 1), (Nx + 1) * (Ny + 1)


def _box_sets(coords, Lx):
    ids = np.arange(len(coords))
    return {"left": ids[np.isclose(coords[:, 0], 0.0, atol=1e-5)],
            "right": ids[np.isclose(coords[:, 0], Lx, atol=1e-5)]}


def create_3D_box_mesh(Nx, Ny, Nz, Lx, Ly, Lz, case_dir=None):
    """Structured Hex8 box (the reference meshes the same box with gmsh, usefull_functions.py:
    146-211, which is unavailable here).  Nodes x-fastest; element node order is the one
    Hexahedra3D8 expects (hexahedra_3d_8.py:79-88: bottom face counter-clockwise, then top);
    `left` / `right` = nodes with x = 0 / x = Lx (usefull_functions.py:196-205)."""
    coords, n0, sx, sy = _box_nodes(Nx, Ny, Nz, Lx, Ly, Lz)
    conn = np.stack([n0, n0 + 1, n0 + 1 + sx, n0 + sx,
                     n0 + sy, n0 + 1 + sy, n0 + 1 + sx + sy, n0 + sx + sy], axis=1)
    return _finish(Mesh("box_io", "box."), coords, conn, "hexahedron", _box_sets(coords, Lx))


def create_3D_tetra_box_mesh(Nx, Ny, Nz, Lx, Ly, Lz):
    """Tet4 box: 6-tetrahedra (Kuhn) split of every cell, all positively oriented."""
    coords, n0, sx, sy = _box_nodes(Nx, Ny, Nz, Lx, Ly, Lz)
    off = {0: 1, 1: sx, 2: sy}
    tets = []
    for perm, even in (((0, 1, 2), True), ((1, 2, 0), True), ((2, 0, 1), True),
                       ((0, 2, 1), False), ((2, 1, 0), False), ((1, 0, 2), False)):
        v1 = n0 + off[perm[0]]
        v2 = v1 + off[perm[1]]
        v3 = v2 + off[perm[2]]
        tets.append(np.stack([n0, v1, v2, v3] if even else [n0, v2, v1, v3], axis=1))
    conn = np.stack(tets, axis=1).reshape(-1, 4)
    return _finish(Mesh("tet_box_io", "box."), coords, conn, "tetra", _box_sets(coords, Lx))


def perturb_interior_nodes(mesh, amplitude, seed=0):
    """Smooth-free random jitter of interior nodes (fraction `amplitude` of the smallest spacing),
    so structured meshes do not consist of identical elements."""
    X = np.array(mesh.nodes_coordinates, dtype=np.float64)
    lo, hi = X.min(0), X.max(0)
    interior = np.ones(len(X), bool)
    dims = [d for d in range(3) if hi[d] > lo[d]]
    for d in dims:
        interior &= ~np.isclose(X[:, d], lo[d]) & ~np.isclose(X[:, d], hi[d])
    h = min(np.diff(np.unique(np.round(X[:, d], 12))).min() for d in dims)
    rng = np.random.default_rng(seed)
    for d in dims:
        X[interior, d] += amplitude * h * rng.uniform(-1, 1, interior.sum())
    mesh.nodes_coordinates = X
    return mesh
