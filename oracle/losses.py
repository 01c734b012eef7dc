"""Oracle (test infrastructure): ``ComputeElement`` of the FiniteElementLoss family in NumPy f64.

Every function is vectorised over a leading element axis and returns ``(energy (ne,),
re (ne, nd), Ke (ne, nd, nd))`` exactly like the reference's ``ComputeElement`` returns
``(energy, residual, stiffness)`` for one element.

  mechanical   fol/loss_functions/mechanical.py:37-117
  thermal      fol/loss_functions/thermal.py:28-49
  neo-hooke    fol/loss_functions/mechanical_neohooke.py:49-91, 107-275
               fol/constitutive_material_models/neo_hooke.py:14-109, utils.py:14-32, 103-130, 191-196
"""
import numpy as np

from .geometry import ELEMENTS, point_data


# ----------------------------------------------------------------------------- mechanical
def d_matrix(dim, E, nu):
    """mechanical.py:60-82 (2-D is plane stress)."""
    if dim == 2:
        return np.array([[1, nu, 0], [nu, 1, 0], [0, 0, (1 - nu) / 2]]) * (E / (1 - nu ** 2))
    c1 = E / ((1.0 + nu) * (1.0 - 2.0 * nu))
    c2, c3, c4 = c1 * (1.0 - nu), c1 * nu, c1 * 0.5 * (1.0 - 2.0 * nu)
    D = np.zeros((6, 6))
    D[:3, :3] = c3
    D[0, 0] = D[1, 1] = D[2, 2] = c2
    D[3, 3] = D[4, 4] = D[5, 5] = c4
    return D


def b_matrix(gradN):
    """mechanical.py:37-58.  gradN (..., a, dim) -> B (..., nstrain, a*dim)."""
    a, dim = gradN.shape[-2:]
    lead = gradN.shape[:-2]
    if dim == 2:
        B = np.zeros(lead + (3, 2 * a), dtype=gradN.dtype)
        B[..., 0, 0::2] = gradN[..., 0]
        B[..., 1, 1::2] = gradN[..., 1]
        B[..., 2, 0::2] = gradN[..., 1]
        B[..., 2, 1::2] = gradN[..., 0]
        return B
    B = np.zeros(lead + (6, 3 * a), dtype=gradN.dtype)      # rows [xx, yy, zz, xy, yz, xz]
    B[..., 0, 0::3] = gradN[..., 0]
    B[..., 1, 1::3] = gradN[..., 1]
    B[..., 2, 2::3] = gradN[..., 2]
    B[..., 3, 0::3] = gradN[..., 1]
    B[..., 3, 1::3] = gradN[..., 0]
    B[..., 4, 1::3] = gradN[..., 2]
    B[..., 4, 2::3] = gradN[..., 1]
    B[..., 5, 0::3] = gradN[..., 2]
    B[..., 5, 2::3] = gradN[..., 0]
    return B


def n_matrix(N, dim):
    """mechanical.py:84-96.  N (g, a) -> (g, dim, a*dim)."""
    g, a = N.shape
    M = np.zeros((g, dim, dim * a))
    for k in range(dim):
        M[:, k, k::dim] = N
    return M


def body_force_vector(elem, Ns, detJ, w, body):
    """Fe = sum_g w detJ N_mat^T b  -> (ne, nd)."""
    Nm = n_matrix(Ns, elem.dim)                                   # (g, dim, nd)
    return np.einsum("g,eg,gkn,k->en", w, detJ, Nm, np.asarray(body, float).reshape(-1))


def mechanical_element(element_type, num_gp, X, de, u, E, nu, body=None):
    """mechanical.py:98-117.  X (ne,a,3), de (ne,a), u (ne,nd)."""
    elem = ELEMENTS[element_type]
    Ns, gradN, detJ, w = point_data(elem, X, num_gp)
    D = d_matrix(elem.dim, E, nu)
    B = b_matrix(gradN)                                          # (ne, g, s, nd)
    e_gp = np.einsum("ga,ea->eg", Ns, de)
    Se = np.einsum("g,eg,eg,egsn,st,egtm->enm", w, detJ, e_gp, B, D, B, optimize=True)
    body = np.zeros(elem.dim) if body is None else body
    Fe = body_force_vector(elem, Ns, detJ, w, body)
    re = np.einsum("enm,em->en", Se, u) - Fe
    return np.einsum("en,en->e", u, re), re, Se


# ----------------------------------------------------------------------------- thermal
def thermal_element(element_type, num_gp, X, de, T, beta=0.0, c=1.0, body_force=0.0):
    """thermal.py:28-49.  T (ne, a)."""
    elem = ELEMENTS[element_type]
    Ns, gradN, detJ, w = point_data(elem, X, num_gp)
    T_gp = np.einsum("ga,ea->eg", Ns, T)
    kappa = np.einsum("ga,ea->eg", Ns, de) * (1.0 + beta * T_gp ** c)
    Se = np.einsum("eg,egai,egbi,eg,g->eab", kappa, gradN, gradN, detJ, w, optimize=True)
    Fe = np.einsum("g,eg,ga->ea", w, detJ, Ns) * body_force
    re = np.einsum("eab,eb->ea", Se, T) - Fe
    return np.einsum("ea,ea->e", T, re), re, Se


def thermal_energy_grads(element_type, num_gp, X, de, T, beta=0.0, c=1.0):
    """Analytic cotangents of the thermal element energy under the reference's stop_gradient
    placement (thermal.py:31, 45-49): dE/dT = re (element residual), dE/dK_a = sum_g N_a
    (1 + beta T_g^c) |grad T_g|^2 detJ w.   Returns (dE/dT (ne,a), dE/dK (ne,a))."""
    elem = ELEMENTS[element_type]
    Ns, gradN, detJ, w = point_data(elem, X, num_gp)
    _, re, _ = thermal_element(element_type, num_gp, X, de, T, beta, c)
    T_gp = np.einsum("ga,ea->eg", Ns, T)
    gT = np.einsum("egai,ea->egi", gradN, T)
    dK = np.einsum("ga,eg,eg,eg,g->ea", Ns, 1.0 + beta * T_gp ** c,
                   np.einsum("egi,egi->eg", gT, gT), detJ, w)
    return re, dK


# ----------------------------------------------------------------------------- neo-hooke
_VOIGT3 = [(0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1)]      # utils.py:14-32, 118-128
_VOIGT2 = [(0, 0), (1, 1), (0, 1)]                               # utils.py:17-22, 107-116


def _diad_special(A, B):
    """utils.py:191-196."""
    return 0.5 * (np.einsum("...ik,...jl->...ijkl", A, B) + np.einsum("...il,...jk->...ijkl", A, B))


def neo_hooke_point(F, k, mu):
    """neo_hooke.py:14-58 (2-D) and :64-109 (3-D).  F (..., d, d); k, mu (...).
    Returns psi (...), S voigt (..., nv), C voigt (..., nv, nv)."""
    d = F.shape[-1]
    C = np.einsum("...ki,...kj->...ij", F, F)
    invC = np.linalg.inv(C)
    J = np.linalg.det(F)
    p = 0.5 * k * (J - 1.0 / J)
    dp_dJ = 0.5 * k * (1.0 + J ** (-2.0))
    trC = np.trace(C, axis1=-2, axis2=-1)
    Jm = J ** (-2.0 / d)
    psi = (k / 4.0) * (J ** 2 - 2.0 * np.log(J) - 1.0) + 0.5 * mu * (Jm * trC - d)
    eye = np.eye(d)
    S_vol = (J * p)[..., None, None] * invC
    # S_iso = J^(-2/d) * P : (mu I),  P = I4 - (1/d) invC (x) C
    S_iso = (Jm * mu)[..., None, None] * (eye - (1.0 / d) * trC[..., None, None] * invC)
    S = S_vol + S_iso
    ii = np.einsum("...ij,...kl->...ijkl", invC, invC)
    P_bar = _diad_special(invC, invC) - (1.0 / d) * ii
    C_vol = (J * p + dp_dJ * J ** 2)[..., None, None, None, None] * ii \
        - (2.0 * J * p)[..., None, None, None, None] * _diad_special(invC, invC)
    C_iso = ((2.0 / d) * Jm * mu * trC)[..., None, None, None, None] * P_bar \
        - (2.0 / d) * (np.einsum("...ij,...kl->...ijkl", invC, S_iso)
                       + np.einsum("...ij,...kl->...ijkl", S_iso, invC))
    C4 = C_vol + C_iso
    vo = _VOIGT3 if d == 3 else _VOIGT2
    Sv = np.stack([S[..., i, j] for (i, j) in vo], axis=-1)
    if d == 3:
        Cv = np.stack([np.stack([C4[..., i, j, kk, l] for (kk, l) in vo], axis=-1)
                       for (i, j) in vo], axis=-2)
    else:
        # utils.py:107-116: lower triangle is a copy of the upper one
        c00, c01, c02 = C4[..., 0, 0, 0, 0], C4[..., 0, 0, 1, 1], C4[..., 0, 0, 0, 1]
        c11, c12, c22 = C4[..., 1, 1, 1, 1], C4[..., 1, 1, 0, 1], C4[..., 0, 1, 0, 1]
        Cv = np.stack([np.stack([c00, c01, c02], -1), np.stack([c01, c11, c12], -1),
                       np.stack([c02, c12, c22], -1)], axis=-2)
    return psi, Sv, Cv


def st_venant_point(F, lam, mu):
    """saint_venant.py:11-33.  psi = lam/2 tr(E)^2 + mu tr(E E), S = lam tr(E) I + 2 mu E; the Voigt tangent
    is taken from lam d_ij d_kl + 2 mu d_ik d_jl (UNsymmetrised identity), i.e. shear entries 2 mu."""
    d = F.shape[-1]
    E = 0.5 * (np.einsum("...ki,...kj->...ij", F, F) - np.eye(d))
    tr = np.trace(E, axis1=-2, axis2=-1)
    psi = 0.5 * lam * tr ** 2 + mu * np.einsum("...ij,...ji->...", E, E)
    S = lam[..., None, None] * tr[..., None, None] * np.eye(d) + 2.0 * mu[..., None, None] * E
    vo = _VOIGT3 if d == 3 else _VOIGT2
    Sv = np.stack([S[..., i, j] for (i, j) in vo], axis=-1)
    eye = np.eye(d)
    C4 = lam[..., None, None, None, None] * np.einsum("ij,kl->ijkl", eye, eye) \
        + 2.0 * mu[..., None, None, None, None] * np.einsum("ik,jl->ijkl", eye, eye)
    if d == 3:
        Cv = np.stack([np.stack([C4[..., i, j, kk, l] for (kk, l) in vo], axis=-1) for (i, j) in vo], axis=-2)
    else:
        c00, c01, c02 = C4[..., 0, 0, 0, 0], C4[..., 0, 0, 1, 1], C4[..., 0, 0, 0, 1]
        c11, c12, c22 = C4[..., 1, 1, 1, 1], C4[..., 1, 1, 0, 1], C4[..., 0, 1, 0, 1]
        Cv = np.stack([np.stack([c00, c01, c02], -1), np.stack([c01, c11, c12], -1),
                       np.stack([c02, c12, c22], -1)], axis=-2)
    return psi, Sv, Cv


def neo_hooke_b_matrix(gradN, F):
    """mechanical_neohooke.py:49-91.  gradN (..., a, d), F (..., d, d) -> B (..., nv, a*d)."""
    a, d = gradN.shape[-2:]
    lead = gradN.shape[:-2]
    g = gradN
    if d == 2:
        B = np.zeros(lead + (3, 2 * a), dtype=np.result_type(gradN, F))
        for c in range(2):
            B[..., 0, c::2] = F[..., c, 0, None] * g[..., 0]
            B[..., 1, c::2] = F[..., c, 1, None] * g[..., 1]
            B[..., 2, c::2] = F[..., c, 1, None] * g[..., 0] + F[..., c, 0, None] * g[..., 1]
        return B
    B = np.zeros(lead + (6, 3 * a), dtype=np.result_type(gradN, F))
    for c in range(3):
        B[..., 0, c::3] = F[..., c, 0, None] * g[..., 0]
        B[..., 1, c::3] = F[..., c, 1, None] * g[..., 1]
        B[..., 2, c::3] = F[..., c, 2, None] * g[..., 2]
        B[..., 3, c::3] = F[..., c, 1, None] * g[..., 2] + F[..., c, 2, None] * g[..., 1]
        B[..., 4, c::3] = F[..., c, 0, None] * g[..., 2] + F[..., c, 2, None] * g[..., 0]
        B[..., 5, c::3] = F[..., c, 0, None] * g[..., 1] + F[..., c, 1, None] * g[..., 0]
    return B


def neo_hooke_element(element_type, num_gp, X, de, u, E, nu, body=None, law="neohooke"):
    """mechanical_neohooke.py:243-275.  ``E`` scales the control field: E_gp = N . de (the
    reference ignores ``young_modulus`` in ComputeElement, :253-255, so pass de = E*control
    -- here ``E`` multiplies nothing and is kept only for a uniform signature)."""
    elem = ELEMENTS[element_type]
    d = elem.dim
    Ns, gradN, detJ, w = point_data(elem, X, num_gp)
    ne = X.shape[0]
    U = u.reshape(ne, elem.nnode, d)
    H = np.einsum("egai,eaj->egji", gradN, U)                    # H[j,i] = du_j/dX_i
    F = H + np.eye(d)
    e_gp = np.einsum("ga,ea->eg", Ns, de)
    k_gp = e_gp / (3.0 * (1.0 - 2.0 * nu))
    mu_gp = e_gp / (2.0 * (1.0 + nu))
    if law == "stvenant":     # mechanical_saint_venant.py:282-289
        psi, Sv, Cv = st_venant_point(F, e_gp * nu / ((1.0 + nu) * (1.0 - 2.0 * nu)), mu_gp)
    else:
        psi, Sv, Cv = neo_hooke_point(F, k_gp, mu_gp)
    B = neo_hooke_b_matrix(gradN, F)
    wd = w[None, :] * detJ
    Kmat = np.einsum("eg,egsn,egst,egtm->enm", wd, B, Cv, B, optimize=True)
    vo = _VOIGT3 if d == 3 else _VOIGT2
    S_mat = np.zeros(Sv.shape[:-1] + (d, d), dtype=Sv.dtype)
    for v, (i, j) in enumerate(vo):
        S_mat[..., i, j] = Sv[..., v]
        S_mat[..., j, i] = Sv[..., v]
    geo = np.einsum("eg,egai,egij,egbj->eab", wd, gradN, S_mat, gradN, optimize=True)
    Kgeo = np.einsum("eab,ij->eaibj", geo, np.eye(d)).reshape(ne, elem.nnode * d, elem.nnode * d)
    Fint = np.einsum("eg,egsn,egs->en", wd, B, Sv)
    body = np.zeros(d) if body is None else body
    Fe = body_force_vector(elem, Ns, detJ, w, body)
    energy = np.einsum("eg,eg->e", wd, psi)
    return energy, Fint - Fe, Kmat + Kgeo


def neo_hooke_energy_dcontrol(element_type, num_gp, X, de, u, nu, law="neohooke"):
    """d(energy)/d(de): psi is linear in (k, mu) and both are linear in E_gp = N.de."""
    elem = ELEMENTS[element_type]
    d = elem.dim
    Ns, gradN, detJ, w = point_data(elem, X, num_gp)
    ne = X.shape[0]
    U = u.reshape(ne, elem.nnode, d)
    F = np.einsum("egai,eaj->egji", gradN, U) + np.eye(d)
    one = np.ones(F.shape[:-2])
    if law == "stvenant":
        psi1, _, _ = st_venant_point(F, one * nu / ((1.0 + nu) * (1.0 - 2.0 * nu)), one / (2.0 * (1.0 + nu)))
    else:
        psi1, _, _ = neo_hooke_point(F, one / (3.0 * (1.0 - 2.0 * nu)), one / (2.0 * (1.0 + nu)))
    return np.einsum("g,eg,eg,ga->ea", w, detJ, psi1, Ns)


# ----------------------------------------------------------------------------- implicit-Euler scalar losses
def transient_thermal_element(element_type, num_gp, X, Tc, Tn, k0, rho, cp, dt, beta, c):
    """transient_thermal.py:42-73.  Tc / Tn: current / next nodal temperatures (ne, a); k0 nodal heterogeneity.
    Returns (energy, re, Ke) with Ke = Me + dt * Se_dR (the linearisation of the conductivity included)."""
    elem = ELEMENTS[element_type]
    Ns, gradN, detJ, w = point_data(elem, X, num_gp, transposed_inverse=True)   # B_mat = invJ @ dN^T (:57-58)
    wd = w[None, :] * detJ
    Tn_g = np.einsum("ga,ea->eg", Ns, Tn)
    Tc_g = np.einsum("ga,ea->eg", Ns, Tc)
    k_g = np.einsum("ga,ea->eg", Ns, k0)
    K_g = k_g * (1.0 + beta * Tn_g ** c)
    with np.errstate(divide="ignore", invalid="ignore"):
        dk = k_g * beta * c * Tn_g ** (c - 1.0)
    dk = np.where(np.isfinite(dk), dk, 0.0)
    BtB = np.einsum("egai,egbi->egab", gradN, gradN)
    Se = np.einsum("eg,eg,egab->eab", K_g, wd, BtB)
    Me = rho * cp * np.einsum("ga,gb,eg->eab", Ns, Ns, wd)
    gT = np.einsum("egai,ea->egi", gradN, Tn)
    gTB = np.einsum("egi,egbi->egb", gT, gradN)                   # (B Tn)^T B
    Se_dR = np.einsum("eg,eg,ga,egb->eab", dk, wd, Ns, gTB) + Se
    Te = rho * cp * 0.5 / dt * np.einsum("eg,eg->e", wd, (Tn_g - Tc_g) ** 2)
    energy = 0.5 * np.einsum("ea,eab,eb->e", Tn, Se, Tn) + Te
    re = np.einsum("eab,eb->ea", Me + dt * Se, Tn) - np.einsum("eab,eb->ea", Me, Tc)
    return energy, re, Me + dt * Se_dR


def allen_cahn_element(element_type, num_gp, X, pc, pn, dt, eps):
    """phase_field.py:38-70."""
    elem = ELEMENTS[element_type]
    Ns, gradN, detJ, w = point_data(elem, X, num_gp, transposed_inverse=True)   # B_mat = invJ @ dN^T (:47-48)
    wd = w[None, :] * detJ
    pn_g = np.einsum("ga,ea->eg", Ns, pn)
    pc_g = np.einsum("ga,ea->eg", Ns, pc)
    Se = np.einsum("eg,egai,egbi->eab", wd, gradN, gradN)
    NN = np.einsum("ga,gb->gab", Ns, Ns)
    Me = np.einsum("gab,eg->eab", NN, wd)
    Fe = np.einsum("eg,eg->e", wd, 0.25 * (pn_g ** 2 - 1.0) ** 2)
    Fe_res = np.einsum("ga,eg,eg->ea", Ns, (pn_g ** 2 - 1.0) * pn_g, wd)
    Te = 0.5 / dt * np.einsum("eg,eg->e", wd, (pn_g - pc_g) ** 2)
    dFe = np.einsum("gab,eg,eg->eab", NN, 3.0 * pn_g ** 2 - 1.0, wd)
    energy = 0.5 * np.einsum("ea,eab,eb->e", pn, Se, pn) + Fe / eps ** 2 + Te
    re = np.einsum("eab,eb->ea", Me + dt * Se, pn) - (np.einsum("eab,eb->ea", Me, pc) - dt / eps ** 2 * Fe_res)
    return energy, re, Me + dt * Se - dt / eps ** 2 * dFe


def implicit_scalar_energy_grads(physics, element_type, num_gp, X, fc, fn, params):
    """Gradients of the element energy of the implicit-Euler scalar losses (the first return value of
    transient_thermal.py:42-73 / phase_field.py:38-70, fully differentiated as jax.grad does -- the energies
    carry no stop_gradient) w.r.t. the next field fn and the current field fc, both (ne, a)."""
    elem = ELEMENTS[element_type]
    Ns, gradN, detJ, w = point_data(elem, X, num_gp, transposed_inverse=True)
    wd = w[None, :] * detJ
    fn_g = np.einsum("ga,ea->eg", Ns, fn)
    fc_g = np.einsum("ga,ea->eg", Ns, fc)
    gf = np.einsum("egai,ea->egi", gradN, fn)                     # grad of the next field
    flux = np.einsum("egai,egi->ega", gradN, gf)                  # grad N_a . grad f
    g2 = np.einsum("egi,egi->eg", gf, gf)
    if physics == "transient_thermal":
        rho, cp, dt = params["rho"], params["cp"], params["time_step"]
        beta, c = params.get("beta", 0.0), params.get("c", 1.0)
        k_g = np.einsum("ga,ea->eg", Ns, params["k0"])
        K_g = k_g * (1.0 + beta * fn_g ** c)
        with np.errstate(divide="ignore", invalid="ignore"):
            dk = k_g * beta * c * fn_g ** (c - 1.0)
        dk = np.where(np.isfinite(dk), dk, 0.0)
        rate = rho * cp / dt * wd * (fn_g - fc_g)
        dn = np.einsum("eg,ega->ea", K_g * wd, flux) + np.einsum("eg,ga->ea", 0.5 * dk * wd * g2 + rate, Ns)
    else:
        dt, eps = params["dt"], params["epsilon"]
        rate = wd / dt * (fn_g - fc_g)
        dn = np.einsum("eg,ega->ea", wd, flux) + np.einsum("eg,ga->ea", wd / eps ** 2 * (fn_g ** 2 - 1.0) * fn_g + rate, Ns)
    dc = -np.einsum("eg,ga->ea", rate, Ns)
    return dn, dc


# ----------------------------------------------------------------------------- the *_AD.py variants
def _ad_point(F, k, mu, lam, law):
    """The AD material models differentiate an energy written in the VOIGT vector of C (or E), which
    `VoigtToTensor` (utils.py:34-57) enters twice per shear component: their Voigt shear stresses are therefore
    2x the tensor components (neo_hooke.py:112-157, 159-209; saint_venant.py:36-64).  Returns
    (psi, S_voigt) with that doubling; F (..., d, d)."""
    d = F.shape[-1]
    C = np.einsum("...ki,...kj->...ij", F, F)
    eye = np.eye(d)
    if law == "neohooke_ad":
        invC = np.linalg.inv(C)
        J = np.sqrt(np.linalg.det(C))
        lnJ = np.log(J)
        trC = np.trace(C, axis1=-2, axis2=-1)
        if d == 3:      # psi = mu/2 (J^-2/3 tr C - 3) - mu ln J + lam/2 ln^2 J      (neo_hooke.py:128-134)
            Jm = J ** (-2.0 / 3.0)
            psi = 0.5 * mu * (Jm * trC - 3.0) - mu * lnJ + 0.5 * lam * lnJ ** 2
            S = (mu * Jm)[..., None, None] * (eye - (trC / 3.0)[..., None, None] * invC) \
                + (lam * lnJ - mu)[..., None, None] * invC
        else:           # psi = mu/2 (tr C - 2) - mu ln J + lam/2 ln^2 J            (neo_hooke.py:178-181)
            psi = 0.5 * mu * (trC - 2.0) - mu * lnJ + 0.5 * lam * lnJ ** 2
            S = mu[..., None, None] * (eye - invC) + (lam * lnJ)[..., None, None] * invC
    else:               # stvenant_ad: psi = lam/2 tr(E)^2 + mu tr(E E)                (saint_venant.py:51-54)
        E = 0.5 * (C - eye)
        tr = np.trace(E, axis1=-2, axis2=-1)
        psi = 0.5 * lam * tr ** 2 + mu * np.einsum("...ij,...ji->...", E, E)
        S = (lam * tr)[..., None, None] * eye + 2.0 * mu[..., None, None] * E
    vo = _VOIGT3 if d == 3 else _VOIGT2
    Sv = np.stack([S[..., i, j] * (1.0 if i == j else 2.0) for (i, j) in vo], axis=-1)
    return psi, Sv


def ad_variant_residual(element_type, num_gp, X, de, u, nu, body, law):
    """(energy, Fint - Fe) of mechanical_neohooke_AD.py:254-293 / mechanical_saint_venant_AD.py:271-310."""
    elem = ELEMENTS[element_type]
    d = elem.dim
    Ns, gradN, detJ, w = point_data(elem, X, num_gp)
    ne = X.shape[0]
    U = u.reshape(ne, elem.nnode, d)
    F = np.einsum("egai,eaj->egji", gradN, U) + np.eye(d)
    e_gp = np.einsum("ga,ea->eg", Ns, de)
    k_gp = e_gp / (3.0 * (1.0 - 2.0 * nu))
    mu_gp = e_gp / (2.0 * (1.0 + nu))
    lam_gp = e_gp * nu / ((1.0 + nu) * (1.0 - 2.0 * nu))
    psi, Sv = _ad_point(F, k_gp, mu_gp, lam_gp, law)
    B = neo_hooke_b_matrix(gradN, F)
    wd = w[None, :] * detJ
    Fint = np.einsum("eg,egsn,egs->en", wd, B, Sv)
    body = np.zeros(d) if body is None else body
    Fe = body_force_vector(elem, Ns, detJ, w, body)
    return np.einsum("eg,eg->e", wd, psi), Fint - Fe


def ad_variant_element(element_type, num_gp, X, de, u, nu, body=None, law="neohooke_ad"):
    """ComputeElement of the AD variants: the stiffness is `jax.jacfwd(residual)` there
    (mechanical_neohooke_AD.py:275, 283); here a complex-step Jacobian of the same residual."""
    energy, re = ad_variant_residual(element_type, num_gp, X, de, u, nu, body, law)
    nd = u.shape[1]
    Ke = np.zeros((u.shape[0], nd, nd))
    h = 1e-30
    for j in range(nd):
        up = u.astype(complex)
        up[:, j] += 1j * h
        Ke[:, :, j] = ad_variant_residual(element_type, num_gp, X, de, up, nu, body, law)[1].imag / h
    return energy.real if np.iscomplexobj(energy) else energy, re, Ke
