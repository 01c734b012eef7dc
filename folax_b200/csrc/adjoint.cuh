// Adjoint sensitivities of a finite-element response: the element routines behind
// folax_b200/responses/fe_response.py, i.e. the arithmetic of fol/responses/fe_response.py
//   :91-124   ComputeResponseElementValue        sum_g w detJ f(N.de, N_mat u_e)
//   :126-168  its gradients w.r.t. u_e, de, x_e  (JAX AD in the reference)
//   :312-331  ComputeLossElementShapeGrad         lam_e^T d re/d x_e   (jacrev of ComputeElement(...)[1])
//   :424-442  ComputeLossElementControlGrad       lam_e^T d re/d de
// The reference differentiates with JAX AD; here the derivatives are closed forms.  With W = grad(dx)
// (W_kj = sum_b dx_bk dN_b/dx_j) a node perturbation changes the geometry by
//   d(detJ) = detJ tr(W),      d(grad v) = -(grad v) W     (v any nodal field; N itself does not change)
// so for  phi = lam_e^T re  of
//   mechanical.py:98-117   phi = sum_g w detJ [(N.de) sigma(u):grad(lam) - b.(N lam)]
//       d phi/d de_a  = sum_g w detJ N_a q,                       q = sigma(u):grad(lam)
//       d phi/d x_bk  = sum_g w detJ [(N.de)(q dN_b/dx_k - (M grad N_b)_k) - b.(N lam) dN_b/dx_k],
//                       M = grad(lam)^T sigma(u) + grad(u)^T sigma(lam)
//   thermal.py:28-49       phi = sum_g w detJ kappa_g grad(lam).grad(T),  kappa_g = (N.de)(1 + beta (N.T)^c)
//       d phi/d de_a  = sum_g w detJ N_a (1 + beta T_g^c) grad(lam).grad(T)
//       d phi/d x_bk  = sum_g w detJ kappa_g [gl.gt dN_b/dx_k - gl_k (gt.grad N_b) - gt_k (gl.grad N_b)]
// Every routine is a sequential per-element function, __host__ __device__: the kernels of adjoint.cu run one
// element per thread, and tests/host_shim compiles the same functions for the CPU to check them against the
// complex-step oracle where no GPU is available.  Not a hot path (one call per design iteration, next to a
// linear solve); written for clarity.
#pragma once
#include <math.h>

#include "common.cuh"
#include "elements.cuh"

namespace fol {

// enum values of assemble.cuh, repeated to keep this header free of the assembly kernels
enum : int { ADJ_MECH = 0, ADJ_THERMAL = 1 };

template <int ELEM, int ORDER, class T>
struct ElemPoint {
  static constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM);
  T N[A];
  T gN[A][D];
  T wd;   // w * detJ
};

template <int ELEM, int ORDER, class T>
__host__ __device__ inline void eval_point(const T* X, int g, ElemPoint<ELEM, ORDER, T>& p) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM);
  double xi[3], w;
  gauss_point<ELEM, ORDER>(g, xi, w);
  T dN[A][D];
  shape_functions<ELEM, T>(xi, p.N, dN);
  const T det = global_gradients<ELEM, T, false>(X, dN, p.gN);
  p.wd = (T)w * det;
}

// Kg[g] = N_g . de ;  Ug[k * ustride + g] = sum_a N_g[a] ue[a*DPN + k]      (fe_response.py:112-115)
template <class T, int ELEM, int ORDER>
__host__ __device__ inline void gauss_interpolate_element(int dpn, const T* de, const T* ue, T* Kg, T* Ug,
                                                          long long ustride) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM), NGP = elem_ngauss(ELEM, ORDER);
  for (int g = 0; g < NGP; ++g) {
    double xi[3], w;
    gauss_point<ELEM, ORDER>(g, xi, w);
    T N[A], dN[A][D];
    shape_functions<ELEM, T>(xi, N, dN);
    T kg = (T)0;
    for (int a = 0; a < A; ++a) kg += N[a] * de[a];
    Kg[g] = kg;
    for (int k = 0; k < dpn; ++k) {
      T acc = (T)0;
      for (int a = 0; a < A; ++a) acc += N[a] * ue[a * dpn + k];
      Ug[k * ustride + g] = acc;
    }
  }
}

// value_e and its gradients from the formula values f[g] and partials fK[g], fU[k*ustride + g] at the Gauss
// points.  Outputs may be null.  With accumulate the results are added to what the arrays hold.
template <class T, int ELEM, int ORDER>
__host__ __device__ inline void response_element(int dpn, const T* X, const T* f, const T* fK, const T* fU,
                                                 long long ustride, T* val, T* dU, T* dK, T* dX) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM), NGP = elem_ngauss(ELEM, ORDER);
  T v = (T)0;
  if (dU) for (int i = 0; i < A * dpn; ++i) dU[i] = (T)0;
  if (dK) for (int a = 0; a < A; ++a) dK[a] = (T)0;
  if (dX) for (int i = 0; i < A * 3; ++i) dX[i] = (T)0;
  for (int g = 0; g < NGP; ++g) {
    ElemPoint<ELEM, ORDER, T> p;
    eval_point<ELEM, ORDER, T>(X, g, p);
    v += p.wd * f[g];
    for (int a = 0; a < A; ++a) {
      if (dK && fK) dK[a] += p.wd * p.N[a] * fK[g];
      if (dU && fU)
        for (int k = 0; k < dpn; ++k) dU[a * dpn + k] += p.wd * p.N[a] * fU[k * ustride + g];
      if (dX)
        for (int k = 0; k < D; ++k) dX[a * 3 + k] += p.wd * f[g] * p.gN[a][k];   // d(detJ)/dx = detJ grad N
    }
  }
  if (val) *val = v;
}

// lam_e^T d re / d de  (A)  and  lam_e^T d re / d x_e  (A x 3, unused coordinates zero); see the header comment.
template <class T, int ELEM, int ORDER, int PHYS>
__host__ __device__ inline void residual_adjoint_element(const T* X, const T* de, const T* ue, const T* le,
                                                         const Params<T>& P, T* dK, T* dX) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM), NGP = elem_ngauss(ELEM, ORDER);
  for (int a = 0; a < A; ++a) dK[a] = (T)0;
  for (int i = 0; i < A * 3; ++i) dX[i] = (T)0;
  for (int g = 0; g < NGP; ++g) {
    ElemPoint<ELEM, ORDER, T> p;
    eval_point<ELEM, ORDER, T>(X, g, p);
    T kg = (T)0;
    for (int a = 0; a < A; ++a) kg += p.N[a] * de[a];
    if constexpr (PHYS == ADJ_MECH) {
      // isotropic D of mechanical.py:60-82 as (lam, mu): 3-D c3, c4; 2-D plane stress
      const T E = P.v[0], nu = P.v[1];
      T lam, mu;
      if constexpr (D == 3) {
        const T c1 = E / (((T)1 + nu) * ((T)1 - (T)2 * nu));
        lam = c1 * nu;
        mu = c1 * (T)0.5 * ((T)1 - (T)2 * nu);
      } else {
        const T fpl = E / ((T)1 - nu * nu);
        lam = fpl * nu;
        mu = fpl * ((T)1 - nu) * (T)0.5;
      }
      T Gu[D][D], Gl[D][D], lg[D];
      for (int i = 0; i < D; ++i) {
        lg[i] = (T)0;
        for (int a = 0; a < A; ++a) lg[i] += p.N[a] * le[a * D + i];
        for (int j = 0; j < D; ++j) {
          T su = (T)0, sl = (T)0;
          for (int a = 0; a < A; ++a) {
            su += ue[a * D + i] * p.gN[a][j];
            sl += le[a * D + i] * p.gN[a][j];
          }
          Gu[i][j] = su;
          Gl[i][j] = sl;
        }
      }
      T tru = (T)0, trl = (T)0;
      for (int i = 0; i < D; ++i) {
        tru += Gu[i][i];
        trl += Gl[i][i];
      }
      T Su[D][D], Sl[D][D];
      for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
          Su[i][j] = mu * (Gu[i][j] + Gu[j][i]) + (i == j ? lam * tru : (T)0);
          Sl[i][j] = mu * (Gl[i][j] + Gl[j][i]) + (i == j ? lam * trl : (T)0);
        }
      T q = (T)0;
      for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) q += Su[i][j] * Gl[i][j];
      T M[D][D];
      for (int k = 0; k < D; ++k)
        for (int j = 0; j < D; ++j) {
          T acc = (T)0;
          for (int i = 0; i < D; ++i) acc += Gl[i][k] * Su[i][j] + Gu[i][k] * Sl[i][j];
          M[k][j] = acc;
        }
      T bl = (T)0;   // body force . (N lam)
      for (int i = 0; i < D; ++i) bl += P.v[2 + i] * lg[i];
      for (int b = 0; b < A; ++b) {
        dK[b] += p.wd * p.N[b] * q;
        for (int k = 0; k < D; ++k) {
          T mg = (T)0;
          for (int j = 0; j < D; ++j) mg += M[k][j] * p.gN[b][j];
          dX[b * 3 + k] += p.wd * (kg * (q * p.gN[b][k] - mg) - bl * p.gN[b][k]);
        }
      }
    } else {   // ADJ_THERMAL
      const T beta = P.v[5], cexp = P.v[6];
      T tg = (T)0, gt[D], gl[D];
      for (int a = 0; a < A; ++a) tg += p.N[a] * ue[a];
      for (int k = 0; k < D; ++k) {
        T st = (T)0, sl = (T)0;
        for (int a = 0; a < A; ++a) {
          st += ue[a] * p.gN[a][k];
          sl += le[a] * p.gN[a][k];
        }
        gt[k] = st;
        gl[k] = sl;
      }
      const T nl = (T)1 + ((beta != (T)0) ? beta * (T)pow((double)tg, (double)cexp) : (T)0);
      T q = (T)0;
      for (int k = 0; k < D; ++k) q += gl[k] * gt[k];
      const T kappa = kg * nl;
      for (int b = 0; b < A; ++b) {
        dK[b] += p.wd * p.N[b] * nl * q;
        T tb = (T)0, lb = (T)0;
        for (int k = 0; k < D; ++k) {
          tb += gt[k] * p.gN[b][k];
          lb += gl[k] * p.gN[b][k];
        }
        for (int k = 0; k < D; ++k) dX[b * 3 + k] += p.wd * kappa * (q * p.gN[b][k] - gl[k] * tb - gt[k] * lb);
      }
    }
  }
}


// ---- forward-mode route for the other losses ---------------------------------------------------------------
// phi = lam_e^T re(x_e, de) written once in a generic scalar type S; with S = Dual<T> one evaluation gives one
// directional derivative, and the A*D + A directions of an element are swept one after the other (no closed form
// to derive per constitutive law).  Used for the finite-strain losses (mechanical_neohooke.py:243-275,
// mechanical_saint_venant.py) and the implicit-Euler scalar losses (transient_thermal.py:42-73,
// phase_field.py:38-70); the linear-elastic and thermal closed forms above are checked against it in the tests.
enum : int { ADJ_NEOHOOKE = 2, ADJ_STVK = 4, ADJ_TTHERMAL = 5, ADJ_ALLENCAHN = 6 };

template <class T>
struct Dual {
  T v, d;
  __host__ __device__ Dual() : v((T)0), d((T)0) {}
  __host__ __device__ Dual(T a, T b) : v(a), d(b) {}
  __host__ __device__ Dual(double a) : v((T)a), d((T)0) {}
  __host__ __device__ Dual(float a) : v((T)a), d((T)0) {}
  __host__ __device__ Dual(int a) : v((T)a), d((T)0) {}
  __host__ __device__ friend Dual operator+(const Dual& a, const Dual& b) { return Dual(a.v + b.v, a.d + b.d); }
  __host__ __device__ friend Dual operator-(const Dual& a, const Dual& b) { return Dual(a.v - b.v, a.d - b.d); }
  __host__ __device__ friend Dual operator-(const Dual& a) { return Dual(-a.v, -a.d); }
  __host__ __device__ friend Dual operator*(const Dual& a, const Dual& b) {
    return Dual(a.v * b.v, a.v * b.d + a.d * b.v);
  }
  __host__ __device__ friend Dual operator/(const Dual& a, const Dual& b) {
    const T r = (T)1 / b.v, q = a.v * r;
    return Dual(q, (a.d - q * b.d) * r);
  }
  __host__ __device__ Dual& operator+=(const Dual& b) { v += b.v; d += b.d; return *this; }
  __host__ __device__ Dual& operator-=(const Dual& b) { v -= b.v; d -= b.d; return *this; }
  __host__ __device__ Dual& operator*=(const Dual& b) { *this = *this * b; return *this; }
};

__host__ __device__ inline double fol_log(double x) { return log(x); }
__host__ __device__ inline float fol_log(float x) { return logf(x); }
__host__ __device__ inline double fol_pow(double x, double a) { return pow(x, a); }
__host__ __device__ inline float fol_pow(float x, double a) { return powf(x, (float)a); }
template <class T>
__host__ __device__ inline Dual<T> fol_log(const Dual<T>& x) { return Dual<T>(fol_log(x.v), x.d / x.v); }
template <class T>
__host__ __device__ inline Dual<T> fol_pow(const Dual<T>& x, double a) {
  const T p1 = fol_pow(x.v, a - 1.0);
  return Dual<T>(p1 * x.v, (T)a * p1 * x.d);
}

template <class S, int D>
__host__ __device__ inline S det_small(const S (&M)[D][D]) {
  if constexpr (D == 2) return M[0][0] * M[1][1] - M[0][1] * M[1][0];
  else
    return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) + M[0][1] * (M[1][2] * M[2][0] - M[1][0] * M[2][2]) +
           M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}

template <class S, int D>
__host__ __device__ inline void inv_sym_small(const S (&C)[D][D], S (&iC)[D][D]) {
  const S r = S(1) / det_small<S, D>(C);
  if constexpr (D == 2) {
    iC[0][0] = C[1][1] * r; iC[0][1] = -C[0][1] * r; iC[1][0] = -C[1][0] * r; iC[1][1] = C[0][0] * r;
  } else {
    iC[0][0] = (C[1][1] * C[2][2] - C[1][2] * C[2][1]) * r;
    iC[0][1] = (C[0][2] * C[2][1] - C[0][1] * C[2][2]) * r;
    iC[0][2] = (C[0][1] * C[1][2] - C[0][2] * C[1][1]) * r;
    iC[1][0] = (C[1][2] * C[2][0] - C[1][0] * C[2][2]) * r;
    iC[1][1] = (C[0][0] * C[2][2] - C[0][2] * C[2][0]) * r;
    iC[1][2] = (C[0][2] * C[1][0] - C[0][0] * C[1][2]) * r;
    iC[2][0] = (C[1][0] * C[2][1] - C[1][1] * C[2][0]) * r;
    iC[2][1] = (C[0][1] * C[2][0] - C[0][0] * C[2][1]) * r;
    iC[2][2] = (C[0][0] * C[1][1] - C[0][1] * C[1][0]) * r;
  }
}

// phi = lam_e^T re for one element; X (A x 3) and de (A) in S, the dofs, the adjoint and the auxiliary nodal
// field (transient thermal: k0) in T.  PHYS uses the values of fol_physics.
template <class S, class T, int ELEM, int ORDER, int PHYS>
__host__ __device__ inline S element_phi(const S* X, const S* de, const T* ue, const T* le, const T* aux,
                                         const Params<T>& P) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM), NGP = elem_ngauss(ELEM, ORDER);
  constexpr bool SCALAR = (PHYS == ADJ_THERMAL || PHYS == ADJ_TTHERMAL || PHYS == ADJ_ALLENCAHN);
  constexpr bool TRANSPOSED = (PHYS == ADJ_TTHERMAL || PHYS == ADJ_ALLENCAHN);
  S phi = S(0);
  for (int g = 0; g < NGP; ++g) {
    double xi[3], w;
    gauss_point<ELEM, ORDER>(g, xi, w);
    S N[A], dN[A][D], gN[A][D];
    shape_functions<ELEM, S>(xi, N, dN);
    const S det = global_gradients<ELEM, S, TRANSPOSED>(X, dN, gN);
    const S wd = S(w) * det;
    S eg = S(0);
    for (int b = 0; b < A; ++b) eg += N[b] * de[b];
    if constexpr (SCALAR) {
      S fn = S(0), lg = S(0), kg = S(0), q = S(0);
      for (int b = 0; b < A; ++b) {
        fn += N[b] * S(ue[b]);
        lg += N[b] * S(le[b]);
        if (aux) kg += N[b] * S(aux[b]);
      }
      for (int k = 0; k < D; ++k) {
        S gf = S(0), gl = S(0);
        for (int b = 0; b < A; ++b) {
          gf += gN[b][k] * S(ue[b]);
          gl += gN[b][k] * S(le[b]);
        }
        q += gf * gl;
      }
      if constexpr (PHYS == ADJ_THERMAL) {
        const T beta = P.v[5];
        const S nl = S(1) + ((beta != (T)0) ? S(beta) * fol_pow(fn, (double)P.v[6]) : S(0));
        phi += wd * eg * nl * q;
      } else if constexpr (PHYS == ADJ_TTHERMAL) {
        const T beta = P.v[5], dt = P.v[10], rcp = P.v[8] * P.v[9];
        const S Kg = kg * (S(1) + ((beta != (T)0) ? S(beta) * fol_pow(fn, (double)P.v[6]) : S(0)));
        phi += wd * (S(rcp) * (fn - eg) * lg + S(dt) * Kg * q);
      } else {
        const T dt = P.v[10], ie2 = (T)1 / (P.v[11] * P.v[11]);
        phi += wd * (((fn - eg) + S(dt * ie2) * (fn * fn - S(1)) * fn) * lg + S(dt) * q);
      }
    } else {
      S Gu[D][D], Gl[D][D];
      S bl = S(0);
      for (int i = 0; i < D; ++i) {
        S lgi = S(0);
        for (int b = 0; b < A; ++b) lgi += N[b] * S(le[b * D + i]);
        bl += S(P.v[2 + i]) * lgi;
        for (int j = 0; j < D; ++j) {
          S su = S(0), sl = S(0);
          for (int b = 0; b < A; ++b) {
            su += gN[b][j] * S(ue[b * D + i]);
            sl += gN[b][j] * S(le[b * D + i]);
          }
          Gu[i][j] = su;
          Gl[i][j] = sl;
        }
      }
      const T nu = P.v[1];
      S q = S(0);
      if constexpr (PHYS == ADJ_MECH) {
        const T E = P.v[0];
        T lam, mu;
        if constexpr (D == 3) {
          const T c1 = E / (((T)1 + nu) * ((T)1 - (T)2 * nu));
          lam = c1 * nu;
          mu = c1 * (T)0.5 * ((T)1 - (T)2 * nu);
        } else {
          const T fpl = E / ((T)1 - nu * nu);
          lam = fpl * nu;
          mu = fpl * ((T)1 - nu) * (T)0.5;
        }
        S tru = S(0);
        for (int i = 0; i < D; ++i) tru += Gu[i][i];
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j)
            q += (S(mu) * (Gu[i][j] + Gu[j][i]) + (i == j ? S(lam) * tru : S(0))) * Gl[i][j];
        q = eg * q;
      } else {
        S F[D][D], C[D][D], Sm[D][D];
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j) F[i][j] = Gu[i][j] + (i == j ? S(1) : S(0));
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j) {
            S acc = S(0);
            for (int m = 0; m < D; ++m) acc += F[m][i] * F[m][j];
            C[i][j] = acc;
          }
        const S mu = eg / S((T)2 * ((T)1 + nu));
        if constexpr (PHYS == ADJ_NEOHOOKE) {   // neo_hooke.py:14-58, 64-109
          S iC[D][D];
          inv_sym_small<S, D>(C, iC);
          const S J = det_small<S, D>(F);
          S trC = S(0);
          for (int i = 0; i < D; ++i) trC += C[i][i];
          const S kk = eg / S((T)3 * ((T)1 - (T)2 * nu));
          const S p = S(0.5) * kk * (J - S(1) / J);
          const S Jm = (D == 2) ? S(1) / J : fol_pow(J, -2.0 / 3.0);
          for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j)
              Sm[i][j] = J * p * iC[i][j] + Jm * mu * ((i == j ? S(1) : S(0)) - trC * iC[i][j] / S((T)D));
        } else {                                // saint_venant.py:11-33
          const S lam = eg * S(nu / (((T)1 + nu) * ((T)1 - (T)2 * nu)));
          S trE = S(0);
          for (int i = 0; i < D; ++i) trE += S(0.5) * (C[i][i] - S(1));
          for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j)
              Sm[i][j] = (i == j ? lam * trE : S(0)) + mu * (C[i][j] - (i == j ? S(1) : S(0)));
        }
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j) {
            S ftl = S(0);
            for (int c = 0; c < D; ++c) ftl += F[c][i] * Gl[c][j];
            q += Sm[i][j] * ftl;
          }
      }
      phi += wd * (q - bl);
    }
  }
  return phi;
}

// ---- the same sensitivities at the cost of the POINT law ---------------------------------------------------
// phi_g depends on the geometry only through grad u, grad lam and w detJ, so the chain rule splits into the closed-form
// geometric part used above for linear elasticity (d detJ = detJ tr W, d grad v = -(grad v) W) and the derivative of the
// point function f(Gu, Gl) w.r.t. its two gradients:  with  Au = df/dGu,  Al = df/dGl,  M = Gu^T Au + Gl^T Al
//     d phi/d x_bk = w detJ [ (f - b.(N lam)) dN_b/dx_k - (M grad N_b)_k ],      d phi/d de_a = w detJ N_a df/d(N.de).
// Finite strain (f = S(F):(F^T Gl), linear in the modulus N.de): Al = F S in closed form, Au by dim^2 dual-number
// evaluations of the POINT function -- instead of A*dim + A sweeps over the whole element (same result to rounding;
// the element-sweep version stays as the cross-check of the tests).

// f at unit modulus for finite strain; also returns F and S (unit modulus) when asked (value evaluation)
template <class S, class T, int D, int PHYS>
__host__ __device__ inline S finite_strain_point_f(const S (&Gu)[D][D], const T (&Gl)[D][D], T nu, S (*Fout)[D],
                                                   S (*Sout)[D]) {
  S F[D][D], C[D][D], Sm[D][D];
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) F[i][j] = Gu[i][j] + (i == j ? S(1) : S(0));
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) {
      S acc = S(0);
      for (int m = 0; m < D; ++m) acc += F[m][i] * F[m][j];
      C[i][j] = acc;
    }
  const T mu = (T)1 / ((T)2 * ((T)1 + nu));
  if constexpr (PHYS == ADJ_NEOHOOKE) {
    S iC[D][D];
    inv_sym_small<S, D>(C, iC);
    const S J = det_small<S, D>(F);
    S trC = S(0);
    for (int i = 0; i < D; ++i) trC += C[i][i];
    const T kk = (T)1 / ((T)3 * ((T)1 - (T)2 * nu));
    const S p = S((T)0.5 * kk) * (J - S(1) / J);
    const S Jm = (D == 2) ? S(1) / J : fol_pow(J, -2.0 / 3.0);
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j)
        Sm[i][j] = J * p * iC[i][j] + Jm * S(mu) * ((i == j ? S(1) : S(0)) - trC * iC[i][j] / S((T)D));
  } else {
    const T lam = nu / (((T)1 + nu) * ((T)1 - (T)2 * nu));
    S trE = S(0);
    for (int i = 0; i < D; ++i) trE += S(0.5) * (C[i][i] - S(1));
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j)
        Sm[i][j] = (i == j ? S(lam) * trE : S(0)) + S(mu) * (C[i][j] - (i == j ? S(1) : S(0)));
  }
  S f = S(0);
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) {
      S ftl = S(0);
      for (int c = 0; c < D; ++c) ftl += F[c][i] * S(Gl[c][j]);
      f += Sm[i][j] * ftl;
      if (Fout) Fout[i][j] = F[i][j];
      if (Sout) Sout[i][j] = Sm[i][j];
    }
  return f;
}

template <class T, int ELEM, int ORDER, int PHYS>
__host__ __device__ inline void residual_adjoint_element_point(const T* X, const T* de, const T* ue, const T* le,
                                                               const T* aux, const Params<T>& P, T* dK, T* dX) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM), NGP = elem_ngauss(ELEM, ORDER);
  for (int a = 0; a < A; ++a) dK[a] = (T)0;
  for (int i = 0; i < A * 3; ++i) dX[i] = (T)0;
  for (int g = 0; g < NGP; ++g) {
    double xi[3], w;
    gauss_point<ELEM, ORDER>(g, xi, w);
    T N[A], dN[A][D];
    shape_functions<ELEM, T>(xi, N, dN);
    T eg = (T)0;
    for (int b = 0; b < A; ++b) eg += N[b] * de[b];
    if constexpr (PHYS == ADJ_NEOHOOKE || PHYS == ADJ_STVK) {
      T gN[A][D];
      const T wd = (T)w * global_gradients<ELEM, T, false>(X, dN, gN);
      T Gu[D][D], Gl[D][D], bl = (T)0;
      for (int i = 0; i < D; ++i) {
        T lgi = (T)0;
        for (int b = 0; b < A; ++b) lgi += N[b] * le[b * D + i];
        bl += P.v[2 + i] * lgi;
        for (int j = 0; j < D; ++j) {
          T su = (T)0, sl = (T)0;
          for (int b = 0; b < A; ++b) {
            su += gN[b][j] * ue[b * D + i];
            sl += gN[b][j] * le[b * D + i];
          }
          Gu[i][j] = su;
          Gl[i][j] = sl;
        }
      }
      T F[D][D], Sm[D][D];
      const T f1 = finite_strain_point_f<T, T, D, PHYS>(Gu, Gl, P.v[1], F, Sm);
      T Au[D][D], Al[D][D];
      for (int c = 0; c < D; ++c)                       // Al = F S (closed form)
        for (int j = 0; j < D; ++j) {
          T acc = (T)0;
          for (int i = 0; i < D; ++i) acc += F[c][i] * Sm[i][j];
          Al[c][j] = acc;
        }
      {                                                 // Au = df/dGu: dim^2 forward sweeps of the point function
        using S = Dual<T>;
        S Gd[D][D];
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j) Gd[i][j] = S(Gu[i][j], (T)0);
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j) {
            Gd[i][j].d = (T)1;
            Au[i][j] = finite_strain_point_f<S, T, D, PHYS>(Gd, Gl, P.v[1], nullptr, nullptr).d;
            Gd[i][j].d = (T)0;
          }
      }
      T M[D][D];
      for (int k = 0; k < D; ++k)
        for (int j = 0; j < D; ++j) {
          T acc = (T)0;
          for (int i = 0; i < D; ++i) acc += Gu[i][k] * Au[i][j] + Gl[i][k] * Al[i][j];
          M[k][j] = eg * acc;
        }
      const T val = eg * f1 - bl;
      for (int b = 0; b < A; ++b) {
        dK[b] += wd * N[b] * f1;
        for (int k = 0; k < D; ++k) {
          T mg = (T)0;
          for (int j = 0; j < D; ++j) mg += M[k][j] * gN[b][j];
          dX[b * 3 + k] += wd * (val * gN[b][k] - mg);
        }
      }
    } else {
      // implicit-Euler scalar losses: phi_g = w detJ [c1(f_n, f_c) (N lam) + c2(f_n) gl.gf] with the gradient convention
      // of transient_thermal.py:57-58 / phase_field.py:47-48, gT_a = inv(J) dN_a (J^-1 un-transposed).  For that
      // convention  d gT_a[k] / d x_bm = -inv[k][m] (dN_b . gT_a),  while d detJ / d x_bm = detJ (dN_b inv)[m].
      T J[D][D];
      for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
          T acc = (T)0;
          for (int a = 0; a < A; ++a) acc += X[a * 3 + i] * dN[a][j];
          J[i][j] = acc;
        }
      T inv[D][D];
      const T det = det_small<T, D>(J);
      {
        const T r = (T)1 / det;
        if constexpr (D == 2) {
          inv[0][0] = J[1][1] * r; inv[0][1] = -J[0][1] * r; inv[1][0] = -J[1][0] * r; inv[1][1] = J[0][0] * r;
        } else {
          inv[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * r;
          inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * r;
          inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * r;
          inv[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * r;
          inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * r;
          inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * r;
          inv[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * r;
          inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * r;
          inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r;
        }
      }
      const T wd = (T)w * det;
      T fn = (T)0, lg = (T)0, kg = (T)0, gf[D], gl[D];
      for (int b = 0; b < A; ++b) {
        fn += N[b] * ue[b];
        lg += N[b] * le[b];
        if (aux) kg += N[b] * aux[b];
      }
      for (int k = 0; k < D; ++k) {
        T sf = (T)0, sl = (T)0;
        for (int b = 0; b < A; ++b) {
          T gt = (T)0;                                   // gT_b[k] = sum_j dN_b[j] inv[k][j]
          for (int j = 0; j < D; ++j) gt += dN[b][j] * inv[k][j];
          sf += gt * ue[b];
          sl += gt * le[b];
        }
        gf[k] = sf;
        gl[k] = sl;
      }
      T q = (T)0;
      for (int k = 0; k < D; ++k) q += gf[k] * gl[k];
      const T dt = P.v[10];
      T c1, c2, dc1;                                     // dc1 = d c1 / d(N.de)
      if constexpr (PHYS == ADJ_TTHERMAL) {
        const T beta = P.v[5], rcp = P.v[8] * P.v[9];
        c1 = rcp * (fn - eg);
        dc1 = -rcp;
        c2 = dt * kg * ((T)1 + ((beta != (T)0) ? beta * fol_pow(fn, (double)P.v[6]) : (T)0));
      } else {
        const T ie2 = (T)1 / (P.v[11] * P.v[11]);
        c1 = (fn - eg) + dt * ie2 * (fn * fn - (T)1) * fn;
        dc1 = (T)-1;
        c2 = dt;
      }
      const T val = c1 * lg + c2 * q;
      T il[D], jf[D];                                    // (inv^T gl)_m, (inv^T gf)_m
      for (int m = 0; m < D; ++m) {
        T a1 = (T)0, a2 = (T)0;
        for (int k = 0; k < D; ++k) {
          a1 += gl[k] * inv[k][m];
          a2 += gf[k] * inv[k][m];
        }
        il[m] = a1;
        jf[m] = a2;
      }
      for (int b = 0; b < A; ++b) {
        dK[b] += wd * N[b] * dc1 * lg;
        T df = (T)0, dl = (T)0;                          // dN_b . gf, dN_b . gl
        for (int n = 0; n < D; ++n) {
          df += dN[b][n] * gf[n];
          dl += dN[b][n] * gl[n];
        }
        for (int m = 0; m < D; ++m) {
          T gstd = (T)0;                                 // standard gradient (dN_b inv)[m]: d detJ / d x_bm = detJ * it
          for (int j = 0; j < D; ++j) gstd += dN[b][j] * inv[j][m];
          dX[b * 3 + m] += wd * (val * gstd - c2 * (il[m] * df + jf[m] * dl));
        }
      }
    }
  }
}

// Element energy, the first return value of ComputeElement (fe_loss.py:149-176: ComputeElementsEnergies):
//   mechanical.py:116-117 / thermal.py:45-49   u^T (Ke u - Fe)                  (= phi with lam = u)
//   mechanical_neohooke.py:262, 271            sum_g w detJ psi,  psi of neo_hooke.py:14-58 / 64-109
//   mechanical_saint_venant.py                 sum_g w detJ (lam/2 tr(E)^2 + mu tr(E E))
//   transient_thermal.py:42-73                 sum_g w detJ [K(T_n)/2 |grad T_n|^2 + rho cp/(2 dt) (T_n - T_c)^2]
//   phase_field.py:38-70                       sum_g w detJ [|grad p_n|^2/2 + (p_n^2 - 1)^2/(4 eps^2) + (p_n - p_c)^2/(2 dt)]
template <class T, int ELEM, int ORDER, int PHYS>
__host__ __device__ inline T element_energy(const T* X, const T* de, const T* ue, const T* aux, const Params<T>& P) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM), NGP = elem_ngauss(ELEM, ORDER);
  if constexpr (PHYS == ADJ_MECH || PHYS == ADJ_THERMAL) {
    return element_phi<T, T, ELEM, ORDER, PHYS>(X, de, ue, ue, aux, P);
  } else {
    constexpr bool TRANSPOSED = (PHYS == ADJ_TTHERMAL || PHYS == ADJ_ALLENCAHN);
    T en = (T)0;
    for (int g = 0; g < NGP; ++g) {
      double xi[3], w;
      gauss_point<ELEM, ORDER>(g, xi, w);
      T N[A], dN[A][D], gN[A][D];
      shape_functions<ELEM, T>(xi, N, dN);
      const T wd = (T)w * global_gradients<ELEM, T, TRANSPOSED>(X, dN, gN);
      T eg = (T)0;
      for (int b = 0; b < A; ++b) eg += N[b] * de[b];
      if constexpr (TRANSPOSED) {
        T fn = (T)0, kg = (T)0, g2 = (T)0;
        for (int b = 0; b < A; ++b) {
          fn += N[b] * ue[b];
          if (aux) kg += N[b] * aux[b];
        }
        for (int k = 0; k < D; ++k) {
          T gf = (T)0;
          for (int b = 0; b < A; ++b) gf += gN[b][k] * ue[b];
          g2 += gf * gf;
        }
        const T dt = P.v[10];
        if constexpr (PHYS == ADJ_TTHERMAL) {
          const T beta = P.v[5], rcp = P.v[8] * P.v[9];
          const T Kg = kg * ((T)1 + ((beta != (T)0) ? beta * fol_pow(fn, (double)P.v[6]) : (T)0));
          en += (T)0.5 * Kg * wd * g2 + rcp * (T)0.5 / dt * wd * (fn - eg) * (fn - eg);
        } else {
          const T ie2 = (T)1 / (P.v[11] * P.v[11]);
          en += (T)0.5 * wd * g2 + wd * (T)0.25 * (fn * fn - (T)1) * (fn * fn - (T)1) * ie2 +
                (T)0.5 / dt * wd * (fn - eg) * (fn - eg);
        }
      } else {
        T F[D][D], C[D][D];
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j) {
            T acc = (i == j) ? (T)1 : (T)0;
            for (int b = 0; b < A; ++b) acc += gN[b][j] * ue[b * D + i];
            F[i][j] = acc;
          }
        T trC = (T)0;
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j) {
            T acc = (T)0;
            for (int m = 0; m < D; ++m) acc += F[m][i] * F[m][j];
            C[i][j] = acc;
            if (i == j) trC += acc;
          }
        const T nu = P.v[1];
        const T mu = eg / ((T)2 * ((T)1 + nu));
        if constexpr (PHYS == ADJ_NEOHOOKE) {
          const T J = det_small<T, D>(F);
          const T kk = eg / ((T)3 * ((T)1 - (T)2 * nu));
          const T Jm = (D == 2) ? (T)1 / J : fol_pow(J, -2.0 / 3.0);
          en += wd * ((kk * (T)0.25) * (J * J - (T)2 * fol_log(J) - (T)1) + (T)0.5 * mu * (Jm * trC - (T)D));
        } else {
          const T lam = eg * nu / (((T)1 + nu) * ((T)1 - (T)2 * nu));
          T trE = (T)0, ee = (T)0;
          for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) {
              const T Eij = (T)0.5 * (C[i][j] - (i == j ? (T)1 : (T)0));
              if (i == j) trE += Eij;
              ee += Eij * Eij;
            }
          en += wd * ((T)0.5 * lam * trE * trE + mu * ee);
        }
      }
    }
    return en;
  }
}

// The A*D + A directional derivatives of phi, one forward sweep each.
template <class T, int ELEM, int ORDER, int PHYS>
__host__ __device__ inline void residual_adjoint_element_dual(const T* X, const T* de, const T* ue, const T* le,
                                                              const T* aux, const Params<T>& P, T* dK, T* dX) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM);
  using S = Dual<T>;
  S Xs[A * 3], ds[A];
  for (int i = 0; i < A * 3; ++i) Xs[i] = S(X[i], (T)0);
  for (int b = 0; b < A; ++b) ds[b] = S(de[b], (T)0);
  for (int b = 0; b < A; ++b) {
    for (int k = 0; k < 3; ++k) {
      if (k >= D) {
        dX[b * 3 + k] = (T)0;
        continue;
      }
      Xs[b * 3 + k].d = (T)1;
      dX[b * 3 + k] = element_phi<S, T, ELEM, ORDER, PHYS>(Xs, ds, ue, le, aux, P).d;
      Xs[b * 3 + k].d = (T)0;
    }
    ds[b].d = (T)1;
    dK[b] = element_phi<S, T, ELEM, ORDER, PHYS>(Xs, ds, ue, le, aux, P).d;
    ds[b].d = (T)0;
  }
}

}  // namespace fol
