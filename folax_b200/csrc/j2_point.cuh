// J2 (von Mises) plasticity with saturating isotropic hardening: stress update by return mapping and
// its tangent by forward-mode differentiation THROUGH the Newton iteration.
//
//   fol/constitutive_material_models/plasticity.py:122-325 (evaluate, _return_mapping, _plastic_corrector)
//   fol/constitutive_material_models/utils.py:57-100 (array<->tensor), :140-174, :216-250 (NewtonSolver)
//   fol/loss_functions/mechanical_elastoplasticity.py:45-55, 92 (strain tensor, jacfwd tangent)
//
// The reference differentiates its while-loop with jax.jacfwd; because B is constant, that is
// sum_g w detJ B^T (d sigma/d eps) B with d sigma/d eps the forward-mode derivative of the algorithm.
// Here every quantity of the iteration is a dual number (value + V strain tangents), so the same
// iteration (x0 = 0, stop on ||r|| <= 1e-6 or 50 steps tested on primal values, 1e-12 regulariser in
// n = s/(sigma_eq + 1e-12)) is replayed with its derivative.  The 7x7 Newton Jacobian is written out
// analytically (in dual arithmetic, so its own strain-derivative is carried too).
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
    // plastic corrector: unknowns x = [d eps_p (6), d lambda], x0 = 0 (plasticity.py:262-301)
    Du x[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) x[k] = dconst<T, V>((T)0);
    for (int it = 0;; ++it) {
      // residual at x
      Sym3<Du> e2, sg2, s2;
      Du q2;
#pragma unroll
      for (int k = 0; k < 6; ++k) e2.c[k] = et.c[k] - ep.c[k] - x[k];
      stress_dev_eq<T, V>(e2, lam, G, sg2, s2, q2);
      const Du qe = q2 + dconst<T, V>((T)1e-12);
      const Du iqe = dconst<T, V>((T)1) / qe;
      Du r[7];
      T nrm = (T)0;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        r[k] = x[k] - x[6] * (s2.c[k] * iqe);
        nrm += r[k].v * r[k].v;
      }
      const Du hx = dexp((-h2) * (dconst<T, V>(xi) + x[6]));
      r[6] = q2 - (dconst<T, V>(y0 + h1) - h1 * hx);
      nrm += r[6].v * r[6].v;
      if (!((T)sqrt((double)nrm) > tol && it < max_iter)) break;   // utils.py:222-226

      // analytic Jacobian d r / d x (columns k < 6: d/d(d eps_p)_k, column 6: d/d(d lambda))
      Du Jm[7][8];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        // d s / d x_k = -2G dev(T_k), T_k the unit array-tensor (off-diagonals symmetric)
        T ds[6];
#pragma unroll
        for (int m = 0; m < 6; ++m) ds[m] = (T)0;
        if (k < 3) {
#pragma unroll
          for (int m = 0; m < 3; ++m) ds[m] = (T)(-2) * G * ((m == k ? (T)1 : (T)0) - (T)1 / (T)3);
        } else {
          ds[k] = (T)(-2) * G;
        }
        // s : ds  (off-diagonal entries count twice in the Frobenius product)
        Du sds = dconst<T, V>((T)0);
#pragma unroll
        for (int m = 0; m < 6; ++m)
          if (ds[m] != (T)0) sds = sds + ((m < 3 ? (T)1 : (T)2) * ds[m]) * s2.c[m];
        const Du dq = ((T)1.5 * sds) / q2;
        const Du dqq = dq * iqe * iqe;
#pragma unroll
        for (int m = 0; m < 6; ++m) {
          const Du dn = ds[m] * iqe - s2.c[m] * dqq;
          Jm[m][k] = dconst<T, V>(m == k ? (T)1 : (T)0) - x[6] * dn;
        }
        Jm[6][k] = dq;
      }
#pragma unroll
      for (int m = 0; m < 6; ++m) Jm[m][6] = -(s2.c[m] * iqe);
      Jm[6][6] = (-(h1 * h2)) * hx;
#pragma unroll
      for (int m = 0; m < 7; ++m) Jm[m][7] = -r[m];

      // solve J dx = -r: Gaussian elimination with partial pivoting on primal values
      for (int c = 0; c < 7; ++c) {
        int p = c;
        T best = fabs((double)Jm[c][c].v);
        for (int rr = c + 1; rr < 7; ++rr) {
          const T a = fabs((double)Jm[rr][c].v);
          if (a > best) { best = a; p = rr; }
        }
        if (p != c) {
          for (int k = c; k < 8; ++k) {
            const Du tmp = Jm[c][k];
            Jm[c][k] = Jm[p][k];
            Jm[p][k] = tmp;
          }
        }
        const Du ipiv = dconst<T, V>((T)1) / Jm[c][c];
        for (int rr = c + 1; rr < 7; ++rr) {
          const Du f = Jm[rr][c] * ipiv;
          for (int k = c; k < 8; ++k) Jm[rr][k] = Jm[rr][k] - f * Jm[c][k];
        }
      }
      Du dx[7];
      for (int i = 6; i >= 0; --i) {
        Du acc = Jm[i][7];
        for (int k = i + 1; k < 7; ++k) acc = acc - Jm[i][k] * dx[k];
        dx[i] = acc / Jm[i][i];
      }
#pragma unroll
      for (int k = 0; k < 7; ++k) x[k] = x[k] + dx[k];
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
