"""Oracle (test infrastructure): element library restated in NumPy float64.

Follows the reference element classes:
  fol/geometries/geometry.py:88-97          (J = (dN^T X)^T, gradN = dN . J^-1)
  fol/geometries/hexahedra_3d_8.py:17-114   (Hex8 Gauss rules 1-3, N, dN/dxi)
  fol/geometries/quadrilateral_2d_4.py:17-70
  fol/geometries/tetrahedra_3d_4.py:17-60
  fol/geometries/triangle_2d_3.py:17-58
"""
import numpy as np

_S3 = 1.0 / np.sqrt(3.0)
_S35 = np.sqrt(3.0 / 5.0)


def _hex_N(p):
    x, y, z = p
    sx = np.array([-1, 1, 1, -1, -1, 1, 1, -1.0])
    sy = np.array([-1, -1, 1, 1, -1, -1, 1, 1.0])
    sz = np.array([-1, -1, -1, -1, 1, 1, 1, 1.0])
    return 0.125 * (1 + sx * x) * (1 + sy * y) * (1 + sz * z)


def _hex_dN(p):
    x, y, z = p
    sx = np.array([-1, 1, 1, -1, -1, 1, 1, -1.0])
    sy = np.array([-1, -1, 1, 1, -1, -1, 1, 1.0])
    sz = np.array([-1, -1, -1, -1, 1, 1, 1, 1.0])
    return np.stack([0.125 * sx * (1 + sy * y) * (1 + sz * z),
                     0.125 * sy * (1 + sx * x) * (1 + sz * z),
                     0.125 * sz * (1 + sx * x) * (1 + sy * y)], axis=1)


def _quad_N(p):
    x, y = p[0], p[1]
    sx = np.array([-1, 1, 1, -1.0])
    sy = np.array([-1, -1, 1, 1.0])
    return 0.25 * (1 + sx * x) * (1 + sy * y)


def _quad_dN(p):
    x, y = p[0], p[1]
    sx = np.array([-1, 1, 1, -1.0])
    sy = np.array([-1, -1, 1, 1.0])
    return np.stack([0.25 * sx * (1 + sy * y), 0.25 * sy * (1 + sx * x)], axis=1)


def _tet_N(p):
    return np.array([1.0 - (p[0] + p[1] + p[2]), p[0], p[1], p[2]])


def _tet_dN(p):
    return np.array([[-1.0, -1.0, -1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])


def _tri_N(p):
    return np.array([1.0 - p[0] - p[1], p[0], p[1]])


def _tri_dN(p):
    return np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])


def _tensor_rule_3d(pts1, w1):
    # x fastest, then y, then z  (hexahedra_3d_8.py:37-76)
    pts, w = [], []
    for k, zk in enumerate(pts1):
        for j, yj in enumerate(pts1):
            for i, xi in enumerate(pts1):
                pts.append([xi, yj, zk])
                w.append(w1[i] * w1[j] * w1[k])
    return np.array(pts), np.array(w)


def _tensor_rule_2d(pts1, w1):
    pts, w = [], []
    for j, yj in enumerate(pts1):
        for i, xi in enumerate(pts1):
            pts.append([xi, yj, 0.0])
            w.append(w1[i] * w1[j])
    return np.array(pts), np.array(w)


def _hex_gauss(order):
    if order == 1:
        return np.array([[0.0, 0.0, 0.0]]), np.array([8.0])
    if order == 2:  # ordered like the nodes (hexahedra_3d_8.py:23-33)
        s = _S3
        pts = np.array([[-s, -s, -s], [s, -s, -s], [s, s, -s], [-s, s, -s],
                        [-s, -s, s], [s, -s, s], [s, s, s], [-s, s, s]])
        return pts, np.ones(8)
    if order == 3:
        return _tensor_rule_3d([-_S35, 0.0, _S35], [5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0])
    raise ValueError(order)


def _quad_gauss(order):
    if order == 1:
        return np.array([[0.0, 0.0, 0.0]]), np.array([4.0])
    if order == 2:
        s = _S3
        return np.array([[-s, -s, 0], [s, -s, 0], [s, s, 0], [-s, s, 0.0]]), np.ones(4)
    if order == 3:
        return _tensor_rule_2d([-_S35, 0.0, _S35], [5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0])
    raise ValueError(order)


def _tet_gauss(order):
    if order == 1:
        return np.array([[0.25, 0.25, 0.25]]), np.array([1.0 / 6.0])
    if order == 2:  # 8-digit literals, tetrahedra_3d_4.py:24-27
        a, b = 0.58541020, 0.13819660
        return np.array([[a, b, b], [b, a, b], [b, b, a], [b, b, b]]), np.full(4, 1.0 / 24.0)
    if order == 3:  # tetrahedra_3d_4.py:31-44
        a1, b1 = 0.015835909865720057993, 0.32805469671142664734
        a2, b2 = 0.67914317820120795168, 0.10695227393293068277
        pts = np.array([[a1, b1, b1], [b1, a1, b1], [b1, b1, a1], [b1, b1, b1],
                        [a2, b2, b2], [b2, a2, b2], [b2, b2, a2], [b2, b2, b2]])
        w = np.array([0.02308799441864369039] * 4 + [0.01857867224802297628] * 4)
        return pts, w
    raise ValueError(order)


def _tri_gauss(order):
    if order == 1:
        return np.array([[1 / 3.0, 1 / 3.0, 0.0]]), np.array([0.5])
    if order == 2:
        return (np.array([[1 / 6.0, 1 / 6.0, 0], [2 / 3.0, 1 / 6.0, 0], [1 / 6.0, 2 / 3.0, 0.0]]),
                np.full(3, 1 / 6.0))
    if order == 3:
        return (np.array([[0.2, 0.2, 0], [0.6, 0.2, 0], [0.2, 0.6, 0], [1 / 3.0, 1 / 3.0, 0.0]]),
                np.array([25 / 96.0, 25 / 96.0, 25 / 96.0, -27 / 96.0]))
    raise ValueError(order)


class Element:
    def __init__(self, name, nnode, dim, N, dN, gauss):
        self.name, self.nnode, self.dim = name, nnode, dim
        self.N, self.dN, self.gauss = N, dN, gauss


ELEMENTS = {
    "hexahedron": Element("hexahedron", 8, 3, _hex_N, _hex_dN, _hex_gauss),
    "quad": Element("quad", 4, 2, _quad_N, _quad_dN, _quad_gauss),
    "tetra": Element("tetra", 4, 3, _tet_N, _tet_dN, _tet_gauss),
    "triangle": Element("triangle", 3, 2, _tri_N, _tri_dN, _tri_gauss),
}


def jacobian(elem, X, point):
    """geometry.py:94-97 -- J = (dN^T X)^T ; X is (..., a, 3); 2-D elements use X[..., :2]."""
    dN = elem.dN(point)
    Xd = X[..., : elem.dim]
    return np.swapaxes(np.einsum("aj,...ai->...ji", dN, Xd), -1, -2)


def point_data(elem, X, order, transposed_inverse=False):
    """Per Gauss point: N (g,a), gradN (..., g, a, dim), detJ (..., g), weights (g).
    transposed_inverse=True reproduces `B_mat = invJ @ dN^T` of transient_thermal.py:57-58 and
    phase_field.py:47-48 (the inverse Jacobian enters un-transposed there)."""
    pts, w = elem.gauss(order)
    Ns = np.stack([elem.N(p) for p in pts])
    grads, dets = [], []
    for p in pts:
        J = jacobian(elem, X, p)                      # (..., dim, dim)
        inv = np.linalg.inv(J)
        if transposed_inverse:
            inv = np.swapaxes(inv, -1, -2)
        grads.append(np.einsum("aj,...jk->...ak", elem.dN(p), inv))
        dets.append(np.linalg.det(J))
    return Ns, np.stack(grads, axis=-3), np.stack(dets, axis=-1), w
