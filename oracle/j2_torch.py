"""Oracle (test infrastructure): SECOND, independent oracle of the J2 point update -- a literal transcription of the
reference into torch float64 on the CPU, differentiated by torch.func (an AD engine that shares nothing with the
hand-written dual numbers of oracle/j2.py or of the CUDA kernels).

Follows, statement by statement,
  fol/constitutive_material_models/plasticity.py:136-200   (evaluate: array -> tensor, 3-D path)
  fol/constitutive_material_models/plasticity.py:202-245   (_return_mapping: trial state, cond on f_trial < 0)
  fol/constitutive_material_models/plasticity.py:247-325   (_plastic_corrector: residual in 7 unknowns, x0 = 0)
  fol/constitutive_material_models/utils.py:57-100, 140-174 (TensorToArray / ArrayToTensor, deviator, von Mises)
  fol/constitutive_material_models/utils.py:216-250        (NewtonSolver.solve: while ||r|| > tol and k < max_iter:
                                                            x += solve(jacfwd(residual)(x), -r))
  fol/loss_functions/mechanical_elastoplasticity.py:45-55, 92 (strain matrix from B u, tangent = jacfwd)
The while-loop and the cond are Python control flow on primal values -- what jax.lax.while_loop / cond evaluate --
and the tangent d sigma / d eps is torch.func.jvp THROUGH that loop, one strain direction at a time, with the
Newton matrix from a nested torch.func.jacfwd and the step from torch.linalg.solve, as in the reference.

Parity status: the reference holds no golden for the plastic branch (its tests use zero state and rigid motion), so
the plastic branch is pinned by the agreement of three independently written routes: this file (literal + torch AD),
oracle/j2.py (literal 7-unknown replay with hand-written dual numbers and LU) and the kernels' reduced two-unknown
form (csrc/j2_point.cuh), at <= 1e-12 (tests/test_oracle_j2.py, tests/test_j2_host_shim.py).
"""
import numpy as np
import torch
from torch.func import jacfwd, jvp


def _array_to_tensor(v):
    """utils.py:79-100."""
    return torch.stack([torch.stack([v[0], v[3], v[5]]), torch.stack([v[3], v[1], v[4]]),
                        torch.stack([v[5], v[4], v[2]])])


def _tensor_to_array(t):
    """utils.py:57-77."""
    return torch.stack([t[0, 0], t[1, 1], t[2, 2], t[0, 1], t[1, 2], t[0, 2]])


def evaluate(strain, state, E, nu, y0, h1, h2, tol=1e-6, max_iter=50):
    """plasticity.py:136-325 for a 3x3 strain tensor and a state [eps_p (6), xi] -> (stress array (6), new state (7))."""
    lam = E * nu / ((1 + nu) * (1 - 2 * nu))
    G = E / (2 * (1 + nu))
    I3 = torch.eye(3, dtype=torch.float64)

    def c_elastic(e):                                         # plasticity.py:63-70
        return lam * torch.trace(e) * I3 + 2.0 * G * e

    def deviatoric(t):                                        # utils.py:140-151
        return t - I3 * (torch.trace(t) / 3)

    def von_mises(sig):                                       # utils.py:153-174
        s = deviatoric(sig)
        return np.sqrt(1.5) * torch.sqrt(torch.tensordot(s, s, dims=2))

    def hardening(xi):                                        # plasticity.py:132-134
        return y0 + h1 * (1.0 - torch.exp(-h2 * xi))

    ps, xi = _array_to_tensor(state[:-1]), state[-1]
    stress_trial = c_elastic(strain - ps)
    f_trial = von_mises(stress_trial) - hardening(xi)
    if f_trial.item() < 0.0:                                  # plasticity.py:241-245
        return _tensor_to_array(stress_trial), state

    def residual(dx):                                         # plasticity.py:266-301
        deps_p, dlambda = _array_to_tensor(dx[:-1]), dx[-1]
        sigma = c_elastic(strain - (ps + deps_p))
        s, sigma_eq = deviatoric(sigma), von_mises(sigma)
        n_voigt = _tensor_to_array(s / (sigma_eq + 1e-12))
        r_flow = dx[:-1] - dlambda * n_voigt
        r_yield = sigma_eq - hardening(xi + dlambda)
        return torch.cat([r_flow, r_yield.reshape(1)])

    x, k = torch.zeros(7, dtype=torch.float64), 0             # plasticity.py:304
    while True:                                               # utils.py:230-248
        r = residual(x)
        if not (torch.linalg.norm(r).item() > tol and k < max_iter):
            break
        x = x + torch.linalg.solve(jacfwd(residual)(x), -r)
        k += 1
    ps_new = ps + _array_to_tensor(x[:6])
    stress = c_elastic(strain - ps_new)
    return _tensor_to_array(stress), torch.cat([_tensor_to_array(ps_new), (xi + x[6]).reshape(1)])


def j2_point(eps_voigt, state, E, nu, y0, h1, h2):
    """3-D Gauss point: eps in the order of the linear B rows [xx,yy,zz,xy,yz,xz] (engineering shears entered
    unhalved, mechanical_elastoplasticity.py:50-55) -> (sigma (6), d sigma / d eps (6,6), new_state (7))."""
    eps_voigt = torch.as_tensor(np.asarray(eps_voigt, dtype=np.float64))
    state = torch.as_tensor(np.asarray(state, dtype=np.float64))
    mat = (float(E), float(nu), float(y0), float(h1), float(h2))

    def stress_of(ev):
        return evaluate(_array_to_tensor(ev), state, *mat)[0]

    sigma, new_state = evaluate(_array_to_tensor(eps_voigt), state, *mat)
    eye = torch.eye(6, dtype=torch.float64)
    cols = [jvp(stress_of, (eps_voigt,), (eye[k],))[1] for k in range(6)]
    return sigma.numpy(), torch.stack(cols, dim=1).numpy(), new_state.numpy()
