// J2 (von Mises) plasticity with saturating isotropic hardening: stress update by return mapping and
// its tangent by forward-mode differentiation THROUGH the Newton iteration.
//
//   fol/constitutive_material_models/plasticity.py:122-325 (evaluate, _return_mapping, _plastic_corrector)
//   fol/constitutive_material_models/utils.py:57-100 (array<->tensor), :140-174, :216-250 (NewtonSolver)
//   fol/loss_functions/mechanical_elastoplasticity.py:45-55, 92 (strain tensor, jacfwd tangent)
//
// The reference solves 7 unknowns x = [d eps_p (6), d lambda] by Newton from x0 = 0 (stop on ||r||_2 <= 1e-6 or
// 50 steps, 1e-12 regulariser in n = s / (sigma_eq + 1e-12)) and differentiates the while-loop with jax.jacfwd.
// Because B is constant, the element tangent is sum_g w detJ B^T (d sigma / d eps) B with d sigma / d eps the
// forward-mode derivative of that algorithm.
//
// What is computed here is THAT iteration, written in the two coordinates it actually moves in.  With
// e = eps - eps_p_old (the trial elastic strain, tensor components [xx,yy,zz,xy,yz,xz], engineering shears entered
// unhalved as the reference does) and d = dev(e):
//   * the flow residual is r_flow = x - dl n(x) with n parallel to dev(e - x); from x0 = 0 every Newton iterate of
//     the 7x7 system stays on the line x = c d (the residual and the Newton matrix map span{d} x R into itself, and
//     Newton's method commutes with the restriction), so the 7-unknown iteration IS the 2-unknown iteration on
//     (c, dl):   rho(c, dl) = c - dl t / (q + 1e-12),   phi(c, dl) = q - y(xi + dl),
//                t = 2G (1 - c),  q = sqrt(3/2) |t| m,   m = ||d||_F (shears counted twice, utils.py:153-157),
//     with the reference's stop test ||r||_2^2 = rho^2 sum_k d_k^2 + phi^2 (plain 2-norm of the 6 + 1 entries);
//   * the iterates depend on the strain only through d (linearly) and m, so d x_k / d eps =
//     c_k Dev + (d c_k / d m) d (x) grad m, grad m = (d_normal, 2 d_shear) / m: ONE scalar tangent direction carried
//     through the loop (scalar dual numbers below) gives the derivative of every iterate, i.e. exactly what
//     jax.jacfwd propagates through the while-loop -- not the converged (implicit-function) tangent, which differs at
//     the 1e-6 level of the stop test.
// Result: d sigma / d eps = C_el - a Dev - b d (x) (w . d) with a = 2G c, b = 2G c'/m, w = (1,1,1,2,2,2): three
// scalars and d per point instead of 36 numbers and a 7 x 7 dual-number iterate (which took 255 registers and 670 M
// local-memory loads per launch at 128^3 in round 1).  Agreement with the literal 7-unknown dual-number replay
// (oracle/j2.py) and with torch.func forward-mode AD through a literal transcription of the reference
// (oracle/j2_torch.py): 1e-14 relative on sigma, tangent and state (tests/test_oracle_j2.py, tests/test_j2_host_shim.py).
#pragma once
#include <math.h>

#ifndef FOL_HD
#define FOL_HD __host__ __device__
#endif

namespace fol {

// value + one tangent (d / d m)
template <class T>
struct D1 {
  T v, d;
};
template <class T> FOL_HD inline D1<T> d1(T v, T d = (T)0) { return D1<T>{v, d}; }
template <class T> FOL_HD inline D1<T> operator+(D1<T> a, D1<T> b) { return {a.v + b.v, a.d + b.d}; }
template <class T> FOL_HD inline D1<T> operator-(D1<T> a, D1<T> b) { return {a.v - b.v, a.d - b.d}; }
template <class T> FOL_HD inline D1<T> operator-(D1<T> a) { return {-a.v, -a.d}; }
template <class T> FOL_HD inline D1<T> operator*(D1<T> a, D1<T> b) { return {a.v * b.v, a.v * b.d + a.d * b.v}; }
template <class T> FOL_HD inline D1<T> operator*(T a, D1<T> b) { return {a * b.v, a * b.d}; }
template <class T> FOL_HD inline D1<T> operator/(D1<T> a, D1<T> b) {
  const T ib = (T)1 / b.v, r = a.v * ib;
  return {r, (a.d - r * b.d) * ib};
}
template <class T> FOL_HD inline D1<T> d1exp(D1<T> a) {
  const T e = (T)exp((double)a.v);
  return {e, e * a.d};
}

// result of one point update, 6-component tensor representation [xx,yy,zz,xy,yz,xz] (utils.py:57-100)
template <class T>
struct J2Point {
  T sig[6];   // stress
  T d[6];     // dev of the trial elastic strain
  T a, b;     // d sigma / d eps = C_el - a Dev - b d (x) (w . d);  0, 0 on the elastic branch
  T c, dl;    // plastic strain increment = c d, increment of the cumulative plastic strain
};

// e_tot: total strain tensor components, ep: old plastic strain, xi: old cumulative plastic strain
template <class T>
FOL_HD inline void j2_radial(const T (&e_tot)[6], const T (&ep)[6], T xi, T lam, T G, T y0, T h1, T h2,
                             J2Point<T>& out) {
  const T tol = (T)1e-6;
  const int max_iter = 50;
  T e[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) e[k] = e_tot[k] - ep[k];
  const T tr = e[0] + e[1] + e[2];
  const T tr3 = tr * ((T)1 / (T)3);
  T mF2 = (T)0, m22 = (T)0;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    out.d[k] = (k < 3) ? e[k] - tr3 : e[k];
    const T dd = out.d[k] * out.d[k];
    mF2 += (k < 3) ? dd : (T)2 * dd;
    m22 += dd;
  }
  const T kap = (T)1.2247448713915890491;   // sqrt(3/2)
  const T mF = (T)sqrt((double)mF2);
  const T q_tr = kap * ((T)2 * G) * mF;     // von Mises stress of the trial state
  const T f_trial = q_tr - (y0 + h1 * ((T)1 - (T)exp((double)(-h2 * xi))));
  D1<T> c = d1((T)0), dl = d1((T)0);
  if (!(f_trial < (T)0)) {                  // plasticity.py:241-245: plastic corrector unless f_trial < 0
    const D1<T> m = d1(mF, (T)1);
    const T t_c = (T)(-2) * G;
    for (int it = 0;; ++it) {
      const D1<T> t = ((T)2 * G) * (d1((T)1) - c);
      const T sg = (t.v >= (T)0) ? (T)1 : (T)-1;              // q = kap m |t| (sqrt of s:s in the reference)
      const D1<T> q = (kap * sg) * (m * t);
      const D1<T> iq = d1((T)1) / (q + d1((T)1e-12));
      const D1<T> hx = d1exp((-h2) * (d1(xi) + dl));
      const D1<T> rho = c - dl * t * iq;
      const D1<T> phi = q - d1(y0 + h1) + h1 * hx;
      const T nrm = (T)sqrt((double)(rho.v * rho.v * m22 + phi.v * phi.v));
      if (!(nrm > tol && it < max_iter)) break;                // utils.py:222-226
      // Newton matrix of (rho, phi) in (c, dl) -- the restriction of jacfwd(residual) to the line x = c d
      const D1<T> q_c = (kap * sg * t_c) * m;
      const D1<T> iq_c = -(iq * iq * q_c);
      const D1<T> a11 = d1((T)1) - dl * (t_c * iq + t * iq_c);
      const D1<T> a12 = -(t * iq);
      const D1<T> a21 = q_c;
      const D1<T> a22 = (-(h1 * h2)) * hx;
      const D1<T> det = a11 * a22 - a12 * a21;
      c = c + (a12 * phi - rho * a22) / det;
      dl = dl + (a21 * rho - a11 * phi) / det;
    }
  }
  out.c = c.v;
  out.dl = dl.v;
  out.a = (T)2 * G * c.v;
  out.b = (mF > (T)0) ? (T)2 * G * c.d / mF : (T)0;
  const T ltr = lam * tr;
#pragma unroll
  for (int k = 0; k < 6; ++k) out.sig[k] = (k < 3 ? ltr : (T)0) + (T)2 * G * e[k] - out.a * out.d[k];
}

// Entry C_kj of the tangent in tensor components k, j of [xx,yy,zz,xy,yz,xz]
template <class T>
FOL_HD inline T j2_tangent_entry(const J2Point<T>& p, T lam, T G, int k, int j) {
  T v = (T)0;
  if (k < 3 && j < 3) v = lam + p.a * ((T)1 / (T)3);
  if (k == j) v += (T)2 * G - p.a;
  return v - p.b * p.d[k] * ((j < 3) ? p.d[j] : (T)2 * p.d[j]);
}

// eps: total strain in the Voigt order of the linear B matrix (engineering shears), state: history
// [eps_p (V), xi]; writes sigma (V), tangent d sigma / d eps (V*V, row-major), new state.
// 2-D is plane strain: eps_zz = 0, eps_p_zz = -(eps_p_xx + eps_p_yy) (plasticity.py:152-158), outputs the
// [xx, yy, xy] entries.
template <class T, int D>
FOL_HD inline void j2_point(const T* eps, const T* state, T E, T nu, T y0, T h1, T h2, T* sigma, T* tangent,
                            T* state_new) {
  constexpr int V = (D == 3) ? 6 : 3;
  const T lam = E * nu / (((T)1 + nu) * ((T)1 - (T)2 * nu));
  const T G = E / ((T)2 * ((T)1 + nu));
  T et[6], ep[6];
  if constexpr (D == 3) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      et[k] = eps[k];
      ep[k] = state[k];
    }
  } else {
    et[0] = eps[0]; et[1] = eps[1]; et[2] = (T)0; et[3] = eps[2]; et[4] = (T)0; et[5] = (T)0;
    ep[0] = state[0]; ep[1] = state[1]; ep[2] = -(state[0] + state[1]); ep[3] = state[2]; ep[4] = (T)0; ep[5] = (T)0;
  }
  J2Point<T> p;
  j2_radial<T>(et, ep, state[V], lam, G, y0, h1, h2, p);
  constexpr int map3[6] = {0, 1, 2, 3, 4, 5};
  constexpr int map2[3] = {0, 1, 3};
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int ck = (D == 3) ? map3[k] : map2[k];
    sigma[k] = p.sig[ck];
    state_new[k] = ep[ck] + p.c * p.d[ck];
#pragma unroll
    for (int j = 0; j < V; ++j) tangent[k * V + j] = j2_tangent_entry<T>(p, lam, G, ck, (D == 3) ? map3[j] : map2[j]);
  }
  state_new[V] = state[V] + p.dl;
}

}  // namespace fol
