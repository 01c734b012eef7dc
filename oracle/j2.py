"""Oracle (test infrastructure): J2 elastoplasticity with isotropic hardening, restated in NumPy.

Follows
  fol/loss_functions/mechanical_elastoplasticity.py:32-94   (element: strain -> evaluate -> B^T sigma,
                                                             tangent = jacfwd of the element residual)
  fol/constitutive_material_models/plasticity.py:34-72, 122-325  (return mapping, plastic corrector)
  fol/constitutive_material_models/utils.py:57-100, 140-174, 216-250  (array<->tensor maps, Newton)

The reference gets the tangent by forward-mode AD *through* its Newton while-loop
(``jax.jacfwd(compute_residual_flat)``).  Because B is constant in u, that equals
sum_g w detJ B^T (d sigma / d eps) B with d sigma / d eps the forward-mode derivative of the
algorithm -- reproduced here with a small forward-mode dual-number class whose tangents are
carried through the same iteration (same start x0 = 0, same stopping test on primal values,
same 1e-12 regulariser).  Pure-Python loops per Gauss point: small cases only.

Parity status: the elastic branch is pinned by test_elastoplasticity.py:70-160; the plastic branch
has no golden in the reference.  It is pinned by the agreement (<= 1e-12) of this restatement with
oracle/j2_torch.py -- a literal torch transcription of the reference differentiated by torch.func --
and cross-checked by finite differences and yield consistency (tests/test_oracle_j2.py,
tests/test_j2_host_shim.py).
"""
import numpy as np

from .geometry import ELEMENTS, point_data
from .losses import b_matrix, body_force_vector


class Dual:
    """value + tangent vector (forward mode)."""
    __slots__ = ("v", "d")

    def __init__(self, v, d):
        self.v, self.d = float(v), np.asarray(d, dtype=float)

    @staticmethod
    def lift(x, n):
        return x if isinstance(x, Dual) else Dual(x, np.zeros(n))

    def _n(self):
        return self.d.shape[0]

    def __add__(self, o):
        o = Dual.lift(o, self._n())
        return Dual(self.v + o.v, self.d + o.d)

    __radd__ = __add__

    def __neg__(self):
        return Dual(-self.v, -self.d)

    def __sub__(self, o):
        o = Dual.lift(o, self._n())
        return Dual(self.v - o.v, self.d - o.d)

    def __rsub__(self, o):
        return Dual.lift(o, self._n()) - self

    def __mul__(self, o):
        o = Dual.lift(o, self._n())
        return Dual(self.v * o.v, self.v * o.d + o.v * self.d)

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = Dual.lift(o, self._n())
        return Dual(self.v / o.v, (self.d * o.v - self.v * o.d) / (o.v * o.v))

    def __rtruediv__(self, o):
        return Dual.lift(o, self._n()) / self


def dsqrt(x):
    r = np.sqrt(x.v)
    return Dual(r, x.d / (2.0 * r) if r != 0.0 else np.zeros_like(x.d))


def dexp(x):
    e = np.exp(x.v)
    return Dual(e, e * x.d)


def _array_to_tensor(a):
    """utils.py:79-100 (ArrayToTensor): [xx,yy,zz,xy,yz,xz] -> symmetric 3x3 (no shear halving)."""
    return [[a[0], a[3], a[5]], [a[3], a[1], a[4]], [a[5], a[4], a[2]]]


def _tensor_to_array(t):
    """utils.py:57-77 (TensorToArray)."""
    return [t[0][0], t[1][1], t[2][2], t[0][1], t[1][2], t[0][2]]


def _c_elastic(eps, lam, G):
    """plasticity.py:63-70: sigma = lam tr(eps) I + 2 G eps."""
    tr = eps[0][0] + eps[1][1] + eps[2][2]
    return [[lam * tr * (1.0 if i == j else 0.0) + 2.0 * G * eps[i][j] for j in range(3)] for i in range(3)]


def _dev_and_eq(sig):
    """utils.py:140-174: deviator and sqrt(1.5) * Frobenius norm of it."""
    tr3 = (sig[0][0] + sig[1][1] + sig[2][2]) / 3.0
    s = [[sig[i][j] - (tr3 if i == j else 0.0) for j in range(3)] for i in range(3)]
    ss = 0.0
    for i in range(3):
        for j in range(3):
            ss = ss + s[i][j] * s[i][j]
    return s, np.sqrt(1.5) * dsqrt(ss)


def _solve(Jm, rhs):
    """Gaussian elimination with partial pivoting on primal values, in dual arithmetic."""
    n = len(rhs)
    A = [[Jm[i][j] for j in range(n)] + [rhs[i]] for i in range(n)]
    for c in range(n):
        p = max(range(c, n), key=lambda r: abs(A[r][c].v))
        A[c], A[p] = A[p], A[c]
        for r in range(c + 1, n):
            f = A[r][c] / A[c][c]
            for k in range(c, n + 1):
                A[r][k] = A[r][k] - f * A[c][k]
    x = [None] * n
    for i in reversed(range(n)):
        acc = A[i][n]
        for k in range(i + 1, n):
            acc = acc - A[i][k] * x[k]
        x[i] = acc / A[i][i]
    return x


def j2_point(eps_voigt, state, E, nu, y0, h1, h2, dim, tol=1e-6, max_iter=50):
    """One Gauss point.  eps_voigt: strain in the order of the linear B rows ([xx,yy,zz,xy,yz,xz] or
    [xx,yy,xy], engineering shears).  Returns (sigma (V,), tangent d sigma/d eps (V,V), new_state)."""
    V = len(eps_voigt)
    lam = E * nu / ((1 + nu) * (1 - 2 * nu))
    G = E / (2 * (1 + nu))
    e = [Dual(eps_voigt[i], np.eye(V)[i]) for i in range(V)]
    zero = Dual(0.0, np.zeros(V))
    if dim == 3:
        # mechanical_elastoplasticity.py:50-55: shear entries are the engineering shears, unhalved
        eps = [[e[0], e[3], e[5]], [e[3], e[1], e[4]], [e[5], e[4], e[2]]]
        ps = _array_to_tensor([Dual(state[i], np.zeros(V)) for i in range(6)])
    else:
        # :45-49 and plasticity.py:152-158: plane strain embedding, eps_p_zz = -(eps_p_xx + eps_p_yy)
        eps = [[e[0], e[2], zero], [e[2], e[1], zero], [zero, zero, zero]]
        p = [Dual(state[i], np.zeros(V)) for i in range(3)]
        ps = [[p[0], p[2], zero], [p[2], p[1], zero], [zero, zero, -(p[0] + p[1])]]
    xi = float(state[-1])

    def hardening(x):
        return y0 + h1 * (1.0 - dexp(-h2 * x))

    ee = [[eps[i][j] - ps[i][j] for j in range(3)] for i in range(3)]
    sig_tr = _c_elastic(ee, lam, G)
    s_tr, q_tr = _dev_and_eq(sig_tr)
    f_trial = q_tr.v - (y0 + h1 * (1.0 - np.exp(-h2 * xi)))
    if f_trial < 0.0:                                   # plasticity.py:241-245 elastic return
        sig, ps_new, xi_new = sig_tr, ps, xi
    else:                                               # plastic corrector, plasticity.py:247-325
        x = [Dual(0.0, np.zeros(V)) for _ in range(7)]

        def residual(x):
            dp = _array_to_tensor(x[:6])
            eps_p_new = [[ps[i][j] + dp[i][j] for j in range(3)] for i in range(3)]
            sg = _c_elastic([[eps[i][j] - eps_p_new[i][j] for j in range(3)] for i in range(3)], lam, G)
            s, q = _dev_and_eq(sg)
            n = [[s[i][j] / (q + 1e-12) for j in range(3)] for i in range(3)]
            nv = _tensor_to_array(n)
            r = [x[k] - x[6] * nv[k] for k in range(6)]
            r.append(q - hardening(xi + x[6]))
            return r, s, q

        def jacobian(x, s, q):
            """d r / d x, analytic (equals jacfwd(residual) up to rounding)."""
            n = x[0]._n()
            qe = q + 1e-12
            Jm = [[Dual(0.0, np.zeros(n)) for _ in range(7)] for _ in range(7)]
            for k in range(6):
                unit = [Dual(1.0 if m == k else 0.0, np.zeros(n)) for m in range(6)]
                Tk = _array_to_tensor(unit)
                tr3 = (Tk[0][0] + Tk[1][1] + Tk[2][2]) / 3.0
                ds = [[-2.0 * G * (Tk[i][j] - (tr3 if i == j else 0.0)) for j in range(3)] for i in range(3)]
                sds = 0.0
                for i in range(3):
                    for j in range(3):
                        sds = sds + s[i][j] * ds[i][j]
                dq = 1.5 * sds / q
                dn = [[ds[i][j] / qe - s[i][j] * dq / (qe * qe) for j in range(3)] for i in range(3)]
                dnv = _tensor_to_array(dn)
                for m in range(6):
                    Jm[m][k] = (1.0 if m == k else 0.0) - x[6] * dnv[m]
                Jm[6][k] = dq
            nv = _tensor_to_array([[s[i][j] / qe for j in range(3)] for i in range(3)])
            for m in range(6):
                Jm[m][6] = -nv[m]
            Jm[6][6] = -(h1 * h2) * dexp(-h2 * (xi + x[6]))
            return Jm

        k = 0
        while True:                                     # utils.py:230-248
            r, s, q = residual(x)
            if not (np.sqrt(sum(ri.v ** 2 for ri in r)) > tol and k < max_iter):
                break
            dx = _solve(jacobian(x, s, q), [-ri for ri in r])
            x = [x[i] + dx[i] for i in range(7)]
            k += 1
        dp = _array_to_tensor(x[:6])
        ps_new = [[ps[i][j] + dp[i][j] for j in range(3)] for i in range(3)]
        xi_new = xi + x[6]
        sig = _c_elastic([[eps[i][j] - ps_new[i][j] for j in range(3)] for i in range(3)], lam, G)
    val = lambda z: z.v if isinstance(z, Dual) else float(z)
    tan = lambda z: z.d if isinstance(z, Dual) else np.zeros(V)
    if dim == 3:
        out = _tensor_to_array(sig)
        ps_arr = _tensor_to_array(ps_new)
    else:
        out = [sig[0][0], sig[1][1], sig[0][1]]         # TensorToArray of the 2x2 block
        ps_arr = [ps_new[0][0], ps_new[1][1], ps_new[0][1]]
    sigma = np.array([val(o) for o in out])
    tangent = np.stack([tan(o) for o in out])
    new_state = np.array([val(p) for p in ps_arr] + [val(xi_new)])
    return sigma, tangent, new_state


def j2_element(element_type, num_gp, X, u, state, E, nu, y0, h1, h2, body=None):
    """mechanical_elastoplasticity.py:32-94 for a batch of elements.
    X (ne,a,3), u (ne,nd), state (ne,g,7|4) -> energy (ne), new_state, re (ne,nd), Ke (ne,nd,nd)."""
    elem = ELEMENTS[element_type]
    Ns, gradN, detJ, w = point_data(elem, X, num_gp)
    B = b_matrix(gradN)                                   # (ne, g, V, nd)
    ne, ng = detJ.shape
    nd = u.shape[1]
    re, Ke = np.zeros((ne, nd)), np.zeros((ne, nd, nd))
    new_state = np.zeros_like(state, dtype=float)
    for e in range(ne):
        for g in range(ng):
            eps = B[e, g] @ u[e]
            sig, tan, st = j2_point(eps, state[e, g], E, nu, y0, h1, h2, elem.dim)
            wd = w[g] * detJ[e, g]
            re[e] += wd * (B[e, g].T @ sig)
            Ke[e] += wd * (B[e, g].T @ tan @ B[e, g])
            new_state[e, g] = st
    body = np.zeros(elem.dim) if body is None else body
    re -= body_force_vector(elem, Ns, detJ, w, body)
    return np.einsum("en,en->e", u, re), new_state, re, Ke
