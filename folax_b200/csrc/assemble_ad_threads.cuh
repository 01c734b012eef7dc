// Element stage of the *_AD.py loss variants (mechanical_neohooke_AD.py:254-293, mechanical_saint_venant_AD.py:
// 271-310): residual = sum_g w detJ B(F)^T S_voigt(C(F)) - Fe, stiffness = jax.jacfwd(residual) in the reference.
// Here the same: the residual is written once in a generic scalar type and the stiffness comes from nd forward
// sweeps with dual numbers (adjoint.cuh: Dual) -- one element per thread.  These variants exist in the reference as
// cross-checks of the analytic classes, not as fast paths; the same holds here (the analytic kernels of
// assemble.cuh are the fast path).
//
// What the AD material models really compute (pinned on the reference's 19-digit goldens,
// tests/unit/test_neo_hooke_mechanical_loss_AD.py:38-190): they differentiate an energy written in the VOIGT
// vector of C (or E), and VoigtToTensor (utils.py:34-57) enters every shear component twice, so the Voigt shear
// stresses are 2x the tensor components; the 3-D Neo-Hooke energy is mu/2 (J^-2/3 tr C - 3) - mu ln J + lam/2 ln^2 J
// (neo_hooke.py:128-134; non-zero stress at F = I), the 2-D one mu/2 (tr C - 2) - mu ln J + lam/2 ln^2 J (:178-181),
// St-Venant lam/2 tr(E)^2 + mu tr(E E) (saint_venant.py:51-54).
// __host__ __device__: tests/host_shim runs the thread body on the CPU.
#pragma once
#include "adjoint.cuh"
#include "assemble_ad_args.cuh"

namespace fol {

enum : int { LAW_NEOHOOKE_AD = 7, LAW_STVK_AD = 8 };   // = FOL_NEOHOOKE_AD, FOL_STVENANT_AD

__host__ __device__ inline double fol_sqrt(double x) { return sqrt(x); }
__host__ __device__ inline float fol_sqrt(float x) { return sqrtf(x); }
template <class T>
__host__ __device__ inline Dual<T> fol_sqrt(const Dual<T>& x) {
  const T r = fol_sqrt(x.v);
  return Dual<T>(r, x.d / ((T)2 * r));
}

// residual (nd) and energy of one element in the scalar type S of the dofs
template <class S, class T, int ELEM, int ORDER, int LAW>
__host__ __device__ inline void element_residual_ad(const T* X, const T* de, const S* u, const Params<T>& P, S* re,
                                                    S& energy) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM), NGP = elem_ngauss(ELEM, ORDER);
  for (int i = 0; i < A * D; ++i) re[i] = S(0);
  energy = S(0);
  const T nu = P.v[1];
  for (int g = 0; g < NGP; ++g) {
    ElemPoint<ELEM, ORDER, T> p;
    eval_point<ELEM, ORDER, T>(X, g, p);
    T eg = (T)0;
    for (int b = 0; b < A; ++b) eg += p.N[b] * de[b];
    const T mu = eg / ((T)2 * ((T)1 + nu));
    const T lam = eg * nu / (((T)1 + nu) * ((T)1 - (T)2 * nu));
    S F[D][D], C[D][D], Sm[D][D], psi;
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j) {
        S acc = (i == j) ? S(1) : S(0);
        for (int b = 0; b < A; ++b) acc += u[b * D + i] * S(p.gN[b][j]);
        F[i][j] = acc;
      }
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j) {
        S acc = S(0);
        for (int m = 0; m < D; ++m) acc += F[m][i] * F[m][j];
        C[i][j] = acc;
      }
    if constexpr (LAW == LAW_NEOHOOKE_AD) {
      S iC[D][D];
      inv_sym_small<S, D>(C, iC);
      const S J = fol_sqrt(det_small<S, D>(C));
      const S lnJ = fol_log(J);
      S trC = S(0);
      for (int i = 0; i < D; ++i) trC += C[i][i];
      if constexpr (D == 3) {
        const S Jm = fol_pow(J, -2.0 / 3.0);
        psi = S((T)0.5 * mu) * (Jm * trC - S(3)) - S(mu) * lnJ + S((T)0.5 * lam) * lnJ * lnJ;
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j)
            Sm[i][j] = S(mu) * Jm * ((i == j ? S(1) : S(0)) - trC * iC[i][j] / S(3)) + (S(lam) * lnJ - S(mu)) * iC[i][j];
      } else {
        psi = S((T)0.5 * mu) * (trC - S(2)) - S(mu) * lnJ + S((T)0.5 * lam) * lnJ * lnJ;
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j)
            Sm[i][j] = S(mu) * ((i == j ? S(1) : S(0)) - iC[i][j]) + S(lam) * lnJ * iC[i][j];
      }
    } else {
      S trE = S(0), ee = S(0);
      for (int i = 0; i < D; ++i) trE += S(0.5) * (C[i][i] - S(1));
      for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
          const S Eij = S(0.5) * (C[i][j] - (i == j ? S(1) : S(0)));
          ee += Eij * Eij;
          Sm[i][j] = (i == j ? S(lam) * trE : S(0)) + S((T)2 * mu) * Eij;
        }
      psi = S((T)0.5 * lam) * trE * trE + S(mu) * ee;
    }
    const S wd = S(p.wd);
    energy += wd * psi;
    for (int a = 0; a < A; ++a)
      for (int c = 0; c < D; ++c) {
        S acc = S(0);
        for (int i = 0; i < D; ++i) {
          acc += Sm[i][i] * F[c][i] * S(p.gN[a][i]);
          for (int j = i + 1; j < D; ++j)   // Voigt shear rows carry 2 S_ij (see the header)
            acc += S(2) * Sm[i][j] * (F[c][i] * S(p.gN[a][j]) + F[c][j] * S(p.gN[a][i]));
        }
        re[a * D + c] += wd * acc - S(p.wd * p.N[a] * P.v[2 + c]);
      }
  }
}

template <class T, int ELEM, int ORDER, int LAW>
__host__ __device__ inline void assemble_ad_thread(long long e, const AdAsmArgs<T>& a) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM), ND = A * D;
  using S = Dual<T>;
  T X[A * 3], de[A];
  S u[ND];
  bool free_dof[ND];
  for (int b = 0; b < A; ++b) {
    const long long n = a.conn[e * A + b];
    for (int k = 0; k < 3; ++k) X[b * 3 + k] = a.xyz[n * 3 + k];
    de[b] = a.ctrl[n];
    for (int k = 0; k < D; ++k) {
      u[b * D + k] = S(a.u[n * D + k], (T)0);
      free_dof[b * D + k] = a.dir[n * D + k] == 0;
    }
  }
  T* ke = a.ke + e * (long long)(ND * ND);
  S re[ND], en;
  for (int j = 0; j < ND; ++j) {          // column j of K = d re / d u_j
    u[j].d = (T)1;
    element_residual_ad<S, T, ELEM, ORDER, LAW>(X, de, u, a.p, re, en);
    u[j].d = (T)0;
    for (int i = 0; i < ND; ++i) {
      const T kij = re[i].d;
      if (!a.transpose) ke[i * ND + j] = (free_dof[i] || i == j) ? kij : (T)0;
      else ke[j * ND + i] = (free_dof[j] || i == j) ? kij : (T)0;   // row j of K^T, then the row mask
    }
  }
  for (int i = 0; i < ND; ++i) a.re[e * ND + i] = free_dof[i] ? re[i].v : (T)0;
  if (a.energy) a.energy[e] = en.v;
}

}  // namespace fol
