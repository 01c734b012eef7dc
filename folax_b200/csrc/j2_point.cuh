// J2 (von Mises) plasticity with saturating isotropic hardening: stress update by return mapping and
// its tangent by forward-mode differentiation THROUGH the Newton iteration.
//
//   fol/constitutive_material_models/plasticity.py:122-325 (evaluate, _return_mapping, _plastic_corrector)
//   fol/constitutive_material_models/utils.py:57-100 (array<->tensor), :140-174, :216-250 (NewtonSolver)
//   fol/loss_functions/mechanical_elastoplasticity.py:45-55, 92 (strain tensor, jacfwd tangent)
//
// The reference differentiates its while-loop with jax.jacfwd; because B is constant, that is
// sum_g w detJ B^T (d sigma/d eps) B with d sigma/d eps the forward-mode derivative of the algorithm.
// Here the iteration (x0 = 0, stop on ||r|| <= 1e-6 or 50 steps tested on primal values, 1e-12 regulariser in
// n = s/(sigma_eq + 1e-12)) is replayed with its derivative: the iterate x is a dual number (value + V strain
// tangents); each step sets up the solve with the analytic 7x7 Newton Jacobian ONCE in real arithmetic (closed form:
// a structured block, a rank-one update and a scalar Schur complement) and reuses it for the tangent of the step,
// J dx' = -(r' + J' dx) -- algebraically what dual-number elimination does, at a fraction of the arithmetic and
// without a 7x8 matrix of dual numbers in local memory.
#pragma once
#include <math.h>

namespace fol {

template <class T, int N>
struct Dual {
  T v;
  T d[N];
};

template <class T, int N>
__device__ __forceinline__ Dual<T, N> dconst(T v) {
  Dual<T, N> r;
  r.v = v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = (T)0;
  return r;
}
template <class T, int N>
__device__ __forceinline__ Dual<T, N> operator+(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r;
  r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <class T, int N>
__device__ __forceinline__ Dual<T, N> operator-(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r;
  r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <class T, int N>
__device__ __forceinline__ Dual<T, N> operator-(const Dual<T, N>& a) {
  Dual<T, N> r;
  r.v = -a.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
  return r;
}
template <class T, int N>
__device__ __forceinline__ Dual<T, N> operator*(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r;
  r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.v * b.d[i] + b.v * a.d[i];
  return r;
}
template <class T, int N>
__device__ __forceinline__ Dual<T, N> operator*(T a, const Dual<T, N>& b) {
  Dual<T, N> r;
  r.v = a * b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a * b.d[i];
  return r;
}
template <class T, int N>
__device__ __forceinline__ Dual<T, N> operator/(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r;
  const T ib = (T)1 / b.v;
  r.v = a.v * ib;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
  return r;
}
template <class T, int N>
__device__ __forceinline__ Dual<T, N> dsqrt(const Dual<T, N>& a) {
  Dual<T, N> r;
  r.v = (T)sqrt((double)a.v);
  const T h = r.v != (T)0 ? (T)0.5 / r.v : (T)0;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * h;
  return r;
}
template <class T, int N>
__device__ __forceinline__ Dual<T, N> dexp(const Dual<T, N>& a) {
  Dual<T, N> r;
  r.v = (T)exp((double)a.v);
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = r.v * a.d[i];
  return r;
}

// symmetric 3x3 tensor stored as [xx, yy, zz, xy, yz, xz] -- the array order of utils.py:57-100
template <class S>
struct Sym3 {
  S c[6];
};

// deviator and sigma_eq = sqrt(3/2) ||s||_F of sigma = lam tr(e) I + 2G e  (plasticity.py:63-70, utils.py:140-174)
template <class T, int N>
__device__ __forceinline__ void stress_dev_eq(const Sym3<Dual<T, N>>& e, T lam, T G, Sym3<Dual<T, N>>& sig,
                                              Sym3<Dual<T, N>>& s, Dual<T, N>& q) {
  using D = Dual<T, N>;
  const D tr = e.c[0] + e.c[1] + e.c[2];
#pragma unroll
  for (int k = 0; k < 6; ++k) sig.c[k] = (k < 3) ? (lam * tr + ((T)2 * G) * e.c[k]) : (((T)2 * G) * e.c[k]);
  const D m = ((T)1 / (T)3) * (sig.c[0] + sig.c[1] + sig.c[2]);
#pragma unroll
  for (int k = 0; k < 6; ++k) s.c[k] = (k < 3) ? (sig.c[k] - m) : sig.c[k];
  D ss = s.c[0] * s.c[0] + s.c[1] * s.c[1] + s.c[2] * s.c[2];
  ss = ss + (T)2 * (s.c[3] * s.c[3] + s.c[4] * s.c[4] + s.c[5] * s.c[5]);
  q = (T)sqrt(1.5) * dsqrt(ss);
}

// eps: total strain in the Voigt order of the linear B matrix (engineering shears), state: history
// [eps_p (V), xi]; writes sigma (V), tangent d sigma / d eps (V*V, row-major), new state.
template <class T, int D>
__device__ void j2_point(const T* eps, const T* state, T E, T nu, T y0, T h1, T h2, T* sigma, T* tangent,
                         T* state_new) {
  constexpr int V = (D == 3) ? 6 : 3;
  using Du = Dual<T, V>;
  const T lam = E * nu / (((T)1 + nu) * ((T)1 - (T)2 * nu));
  const T G = E / ((T)2 * ((T)1 + nu));
  const T tol = (T)1e-6;
  const int max_iter = 50;

  // total strain tensor: engineering shears enter unhalved (mechanical_elastoplasticity.py:45-55);
  // 2-D is plane strain (plasticity.py:152-158)
  Sym3<Du> et, ep;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    et.c[k] = dconst<T, V>((T)0);
    ep.c[k] = dconst<T, V>((T)0);
  }
  if constexpr (D == 3) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      et.c[k].v = eps[k];
      et.c[k].d[k] = (T)1;
      ep.c[k].v = state[k];
    }
  } else {
    et.c[0].v = eps[0]; et.c[0].d[0] = (T)1;
    et.c[1].v = eps[1]; et.c[1].d[1] = (T)1;
    et.c[3].v = eps[2]; et.c[3].d[2] = (T)1;
    ep.c[0].v = state[0];
    ep.c[1].v = state[1];
    ep.c[3].v = state[2];
    ep.c[2].v = -(state[0] + state[1]);
  }
  const T xi = state[V];

  Sym3<Du> ee, sig, s;
  Du q;
#pragma unroll
  for (int k = 0; k < 6; ++k) ee.c[k] = et.c[k] - ep.c[k];
  stress_dev_eq<T, V>(ee, lam, G, sig, s, q);
  const T f_trial = q.v - (y0 + h1 * ((T)1 - (T)exp((double)(-h2 * xi))));

  Sym3<Du> ep_new = ep;
  Du xi_new = dconst<T, V>(xi);
  if (!(f_trial < (T)0)) {
    // plastic corrector: unknowns x = [d eps_p (6), d lambda], x0 = 0 (plasticity.py:262-301).
    // Forward mode through x <- x + dx, J dx = -r:  the tangent of dx obeys J dx' = -(r' + J' dx), so one REAL
    // 7x7 factorisation per iteration serves the primal step and the V tangent right-hand sides; r' and J' dx
    // are the dual parts of r and of the product J(x, eps) w evaluated in dual arithmetic with w = dx held fixed.
    Du x[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) x[k] = dconst<T, V>((T)0);
    // The tangents are carried one strain direction at a time with SCALAR dual numbers (D1): the primal quantities
    // of the step are evaluated once, then each direction re-evaluates the (cheap) residual and J w with its own
    // seed.  Same numbers as V-wide duals, a fraction of the live registers.
    using D1 = Dual<T, 1>;
    constexpr int comp3[6] = {0, 1, 2, 3, 4, 5}, comp2[3] = {0, 1, 3};   // tensor component seeded by direction t
    for (int it = 0;; ++it) {
      // residual at x (primal)
      T rv[7], sv[6], q2v, iq, hxv;
      {
        Sym3<D1> e2, sg2, s2;
        D1 q2;
#pragma unroll
        for (int k = 0; k < 6; ++k) e2.c[k] = dconst<T, 1>(et.c[k].v - ep.c[k].v - x[k].v);
        stress_dev_eq<T, 1>(e2, lam, G, sg2, s2, q2);
        q2v = q2.v;
        iq = (T)1 / (q2v + (T)1e-12);
        hxv = (T)exp((double)(-h2 * (xi + x[6].v)));
        T nrm = (T)0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          sv[k] = s2.c[k].v;
          rv[k] = x[k].v - x[6].v * (sv[k] * iq);
          nrm += rv[k] * rv[k];
        }
        rv[6] = q2v - ((y0 + h1) - h1 * hxv);
        nrm += rv[6] * rv[6];
        if (!((T)sqrt((double)nrm) > tol && it < max_iter)) break;   // utils.py:222-226
      }

      // Newton matrix J = d r / d x in closed form.  With alpha = 2G dl / (q + 1e-12), Pd = the deviatoric projector
      // on the normal components (identity on the shears) and dq_k = d q / d x_k:
      //   J = [ I + alpha Pd + dl s (dq iq^2)^T   | -s iq     ]      M = I + alpha Pd inverts in closed form
      //       [ dq^T                               | -h1 h2 hx ]      (normal block (I + alpha/3 11^T)/(1+alpha)),
      // the rest is a rank-one update (Sherman-Morrison) and a scalar Schur complement: ~60 flops per right-hand
      // side, no 7x7 factorisation and no matrix held in registers (same solution as the LU to ~1e-14).
      const T dl = x[6].v;
      const T alpha = (T)2 * G * dl * iq, ia = (T)1 / ((T)1 + alpha), a3 = ia * alpha * ((T)1 / (T)3);
      T dq[6], vv[6], Mu[6], Ac[6];
      const T trs3 = (sv[0] + sv[1] + sv[2]) * ((T)1 / (T)3), rq = (T)1 / q2v;
#pragma unroll
      for (int m = 0; m < 6; ++m) {
        dq[m] = (m < 3) ? (T)(-3) * G * (sv[m] - trs3) * rq : (T)(-6) * G * sv[m] * rq;
        vv[m] = dq[m] * iq * iq;
        Mu[m] = dl * sv[m];
      }
      auto minv = [&](T (&b)[6]) {
        const T add = a3 * (b[0] + b[1] + b[2]);
#pragma unroll
        for (int m = 0; m < 6; ++m) b[m] = b[m] * ia + (m < 3 ? add : (T)0);
      };
      minv(Mu);
      T den = (T)1;
#pragma unroll
      for (int m = 0; m < 6; ++m) den += vv[m] * Mu[m];
      const T iden = (T)1 / den;
      auto ainv = [&](T (&b)[6]) {
        minv(b);
        T f = (T)0;
#pragma unroll
        for (int m = 0; m < 6; ++m) f += vv[m] * b[m];
        f *= iden;
#pragma unroll
        for (int m = 0; m < 6; ++m) b[m] -= Mu[m] * f;
      };
#pragma unroll
      for (int m = 0; m < 6; ++m) Ac[m] = -(sv[m] * iq);
      ainv(Ac);
      T schur = (-(h1 * h2)) * hxv;
#pragma unroll
      for (int m = 0; m < 6; ++m) schur -= dq[m] * Ac[m];
      const T ischur = (T)1 / schur;
      auto solve7 = [&](T (&b)[7]) {
        T y[6];
#pragma unroll
        for (int m = 0; m < 6; ++m) y[m] = b[m];
        ainv(y);
        T z = b[6];
#pragma unroll
        for (int m = 0; m < 6; ++m) z -= dq[m] * y[m];
        z *= ischur;
#pragma unroll
        for (int m = 0; m < 6; ++m) b[m] = y[m] - Ac[m] * z;
        b[6] = z;
      };
      T w[7];
#pragma unroll
      for (int m = 0; m < 7; ++m) w[m] = -rv[m];
      solve7(w);
      const T wm = (w[0] + w[1] + w[2]) * ((T)1 / (T)3);
      T W[6];   // -2G dev(w): d s / d x applied to the step
#pragma unroll
      for (int m = 0; m < 6; ++m) W[m] = (T)(-2) * G * (m < 3 ? (w[m] - wm) : w[m]);

      // tangent of the step, direction by direction: J dx' = -(r' + J' dx)
#pragma unroll
      for (int t = 0; t < V; ++t) {
        const int ct = (D == 3) ? comp3[t] : comp2[t];
        Sym3<D1> e2, sg2, s2;
        D1 q2, xt[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          xt[k].v = x[k].v;
          xt[k].d[0] = x[k].d[t];
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          e2.c[k].v = et.c[k].v - ep.c[k].v - x[k].v;
          e2.c[k].d[0] = (k == ct ? (T)1 : (T)0) - x[k].d[t];
        }
        stress_dev_eq<T, 1>(e2, lam, G, sg2, s2, q2);
        const D1 iqe = dconst<T, 1>((T)1) / (q2 + dconst<T, 1>((T)1e-12));
        const D1 hx = dexp((-h2) * (dconst<T, 1>(xi) + xt[6]));
        // r' : dual part of the residual;  (J w)' : dual part of J(x, eps) w with w fixed
        D1 sW = dconst<T, 1>((T)0);
#pragma unroll
        for (int m = 0; m < 6; ++m) sW = sW + ((m < 3 ? (T)1 : (T)2) * W[m]) * s2.c[m];
        const D1 dqW = ((T)1.5 * sW) / q2;
        const D1 dqqW = dqW * iqe * iqe;
        T b[7];
#pragma unroll
        for (int m = 0; m < 6; ++m) {
          const D1 n = s2.c[m] * iqe;
          const D1 r = xt[m] - xt[6] * n;
          const D1 g = dconst<T, 1>(w[m]) - xt[6] * (W[m] * iqe - s2.c[m] * dqqW) - w[6] * n;
          b[m] = -(r.d[0] + g.d[0]);
        }
        {
          const D1 r6 = q2 - (dconst<T, 1>(y0 + h1) - h1 * hx);
          const D1 g6 = dqW + ((-(h1 * h2)) * w[6]) * hx;
          b[6] = -(r6.d[0] + g6.d[0]);
        }
        solve7(b);
#pragma unroll
        for (int m = 0; m < 7; ++m) x[m].d[t] += b[m];
      }
#pragma unroll
      for (int m = 0; m < 7; ++m) x[m].v += w[m];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) ep_new.c[k] = ep.c[k] + x[k];
    xi_new = dconst<T, V>(xi) + x[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) ee.c[k] = et.c[k] - ep_new.c[k];
    stress_dev_eq<T, V>(ee, lam, G, sig, s, q);
  }

  // outputs in the TensorToArray order ([xx,yy,zz,xy,yz,xz] | [xx,yy,xy])
  constexpr int map3[6] = {0, 1, 2, 3, 4, 5};
  constexpr int map2[3] = {0, 1, 3};
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int c = (D == 3) ? map3[k] : map2[k];
    sigma[k] = sig.c[c].v;
#pragma unroll
    for (int m = 0; m < V; ++m) tangent[k * V + m] = sig.c[c].d[m];
    state_new[k] = ep_new.c[c].v;
  }
  state_new[V] = xi_new.v;
}

}  // namespace fol
