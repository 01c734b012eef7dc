"""Oracle (test infrastructure): ctypes loader of the plain-C / OpenMP restatement in oracle/c
(Hex8 elasticity assembly).  Used as the CPU baseline of bench.py and cross-checked against the NumPy
oracle in tests/test_oracle_c.py.  Never imported by the product."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "liboracle_hex.so")


def build():
    subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "c")], check=True)
    return LIB


def load():
    srcs = [os.path.join(_HERE, "c", f) for f in ("hex_mech.c", "quad_thermal_loss.c", "Makefile")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(f) for f in srcs):
        build()
    lib = ctypes.CDLL(LIB)
    lib.oracle_hex_mech_assemble.restype = None
    lib.oracle_hex_mech_assemble.argtypes = [ctypes.c_int64, ctypes.c_int64] + [ctypes.c_void_p] * 5 + \
        [ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.oracle_quad_thermal_loss_grads.restype = None
    lib.oracle_quad_thermal_loss_grads.argtypes = [ctypes.c_int64] * 3 + [ctypes.c_void_p] * 4 + \
        [ctypes.c_double, ctypes.c_double] + [ctypes.c_void_p] * 3
    lib.oracle_set_threads.argtypes = [ctypes.c_int]
    lib.oracle_max_threads.restype = ctypes.c_int
    return lib


def set_threads(n):
    """Thread count of the OpenMP loops, whatever OMP_NUM_THREADS the launcher exported; returns the count in effect."""
    lib = load()
    lib.oracle_set_threads(int(n))
    return int(lib.oracle_max_threads())


def hex_mech_assemble(coords, conn, ctrl, u, dirichlet_indices, E, nu, body=None, transpose=False, out=None):
    lib = load()
    coords = np.ascontiguousarray(coords, np.float64)
    conn = np.ascontiguousarray(conn, np.int32)
    ctrl = np.ascontiguousarray(ctrl, np.float64)
    u = np.ascontiguousarray(u, np.float64)
    nn, ne = len(coords), len(conn)
    flags = np.zeros(3 * nn, np.uint8)
    flags[np.asarray(dirichlet_indices, np.int64)] = 1
    body = np.zeros(3) if body is None else np.ascontiguousarray(body, np.float64)
    data = np.empty(ne * 576) if out is None else out
    R = np.empty(3 * nn)
    lib.oracle_hex_mech_assemble(ne, nn, coords.ctypes.data, conn.ctypes.data, ctrl.ctypes.data, u.ctypes.data,
                                 flags.ctypes.data, float(E), float(nu), body.ctypes.data, int(transpose),
                                 data.ctypes.data, R.ctypes.data)
    return data, R


def quad_thermal_batch_loss_grads(coords, conn, batch_controls, batch_dofs_full, beta=0.0, c=1.0):
    """Batched thermal Quad4 (2x2 rule) energies and their gradients: (E_b (nb), dE_b/dT (nb, nn), dE_b/dK (nb, nn)).
    batch_dofs_full must carry the Dirichlet values already (fe_loss.py:255)."""
    lib = load()
    coords = np.ascontiguousarray(coords, np.float64)
    conn = np.ascontiguousarray(conn, np.int32)
    K = np.ascontiguousarray(np.atleast_2d(batch_controls), np.float64)
    U = np.ascontiguousarray(np.atleast_2d(batch_dofs_full), np.float64)
    nb, nn = U.shape
    energy, gU, gK = np.empty(nb), np.empty((nb, nn)), np.empty((nb, nn))
    lib.oracle_quad_thermal_loss_grads(nb, len(conn), nn, coords.ctypes.data, conn.ctypes.data, K.ctypes.data,
                                       U.ctypes.data, float(beta), float(c), energy.ctypes.data, gU.ctypes.data,
                                       gK.ctypes.data)
    return energy, gU, gK
