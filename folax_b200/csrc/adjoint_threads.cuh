// Per-thread bodies of the adjoint-sensitivity kernels (adjoint.cu): gather one element, call the element
// routine of adjoint.cuh, write the element-major outputs that fol_residual_gather sums to the nodes in the
// residual's fixed order.  __host__ __device__ on purpose: tests/host_shim loops the same functions over the
// elements on the CPU.
#pragma once
#include "adjoint.cuh"

namespace fol {

template <class T>
struct InterpArgs {
  const int32_t* conn;
  const T* ctrl;   // (nn)
  const T* u;      // (nn * dpn)
  T* kg;           // (ne, NGP)
  T* ug;           // (dpn, ne, NGP): U[k] of the response formula is one contiguous (ne, NGP) block
  long long ne;
  int dpn;
};

template <class T>
struct ResponseArgs {
  const T* xyz;    // (nn, 3)
  const int32_t* conn;
  const T* f;      // (ne, NGP) formula values
  const T* fk;     // (ne, NGP) d f / d control, or null
  const T* fu;     // (dpn, ne, NGP) d f / d U[k], or null
  T* val;          // (ne) or null
  T* du;           // (ne, A*dpn) or null
  T* dk;           // (ne, A) or null
  T* dx;           // (ne, A*3) or null
  long long ne;
  int dpn;
};

template <class T>
struct AdjointArgs {
  const T* xyz;
  const int32_t* conn;
  const T* ctrl;
  const T* u;
  const T* lam;    // adjoint dof vector (ndof)
  const T* aux;    // auxiliary nodal field of the physics (transient thermal: k0), or null
  T* dk;           // (ne, A) or null
  T* dx;           // (ne, A*3) or null
  long long ne;
  int accumulate;  // add to dk / dx instead of overwriting (response part + residual part, fe_response.py:385, 515)
  Params<T> p;
  // batch of samples (grid.y = sample): per-sample strides, in values, of ctrl / (u, lam) / dk (dx is not batched)
  long long batch_node = 0, batch_dof = 0, batch_dk = 0;
  int batch_count = 0;
};

template <class T, int ELEM, int ORDER>
__host__ __device__ inline void gauss_interpolate_thread(long long e, const InterpArgs<T>& a) {
  constexpr int A = elem_nnode(ELEM), NGP = elem_ngauss(ELEM, ORDER);
  T de[A], ue[A * 3];
  for (int b = 0; b < A; ++b) {
    const long long n = a.conn[e * A + b];
    de[b] = a.ctrl[n];
    for (int k = 0; k < a.dpn; ++k) ue[b * a.dpn + k] = a.u[n * a.dpn + k];
  }
  gauss_interpolate_element<T, ELEM, ORDER>(a.dpn, de, ue, a.kg + e * NGP, a.ug + e * NGP, a.ne * (long long)NGP);
}

template <class T, int ELEM, int ORDER>
__host__ __device__ inline void response_thread(long long e, const ResponseArgs<T>& a) {
  constexpr int A = elem_nnode(ELEM), NGP = elem_ngauss(ELEM, ORDER);
  T X[A * 3];
  for (int b = 0; b < A; ++b) {
    const long long n = a.conn[e * A + b];
    for (int k = 0; k < 3; ++k) X[b * 3 + k] = a.xyz[n * 3 + k];
  }
  T v, dU[A * 3], dK[A], dX[A * 3];
  response_element<T, ELEM, ORDER>(a.dpn, X, a.f + e * NGP, a.fk ? a.fk + e * NGP : nullptr,
                                   a.fu ? a.fu + e * NGP : nullptr, a.ne * (long long)NGP, &v,
                                   a.du ? dU : nullptr, a.dk ? dK : nullptr, a.dx ? dX : nullptr);
  if (a.val) a.val[e] = v;
  if (a.du)
    for (int i = 0; i < A * a.dpn; ++i) a.du[e * (A * a.dpn) + i] = dU[i];
  if (a.dk)
    for (int b = 0; b < A; ++b) a.dk[e * A + b] = dK[b];
  if (a.dx)
    for (int i = 0; i < A * 3; ++i) a.dx[e * (A * 3) + i] = dX[i];
}

template <class T, int ELEM, int ORDER, int PHYS>
__host__ __device__ inline void residual_adjoint_thread(long long e, const AdjointArgs<T>& a) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM);
  constexpr int DPN = (PHYS == ADJ_THERMAL || PHYS == ADJ_TTHERMAL || PHYS == ADJ_ALLENCAHN) ? 1 : D;
  T X[A * 3], de[A], ue[A * DPN], le[A * DPN], ax[A];
  for (int b = 0; b < A; ++b) {
    const long long n = a.conn[e * A + b];
    for (int k = 0; k < 3; ++k) X[b * 3 + k] = a.xyz[n * 3 + k];
    de[b] = a.ctrl[n];
    ax[b] = a.aux ? a.aux[n] : (T)0;
    for (int k = 0; k < DPN; ++k) {
      ue[b * DPN + k] = a.u[n * DPN + k];
      le[b * DPN + k] = a.lam[n * DPN + k];
    }
  }
  T dK[A], dX[A * 3];
  if constexpr (PHYS == ADJ_MECH || PHYS == ADJ_THERMAL)
    residual_adjoint_element<T, ELEM, ORDER, PHYS>(X, de, ue, le, a.p, dK, dX);          // closed forms
  else   // point-law route (closed-form geometry + dim^2 dual sweeps of the point function where needed)
    residual_adjoint_element_point<T, ELEM, ORDER, PHYS>(X, de, ue, le, a.aux ? ax : nullptr, a.p, dK, dX);
  if (a.dk)
    for (int b = 0; b < A; ++b) {
      T* o = a.dk + e * A + b;
      *o = (a.accumulate ? *o : (T)0) + dK[b];
    }
  if (a.dx)
    for (int i = 0; i < A * 3; ++i) {
      T* o = a.dx + e * (A * 3) + i;
      *o = (a.accumulate ? *o : (T)0) + dX[i];
    }
}

// energy_elem[e] = element energy (ComputeElementsEnergies); the arrays of AdjointArgs, `dk` receives the energies
template <class T, int ELEM, int ORDER, int PHYS>
__host__ __device__ inline void element_energy_thread(long long e, const AdjointArgs<T>& a) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM);
  constexpr int DPN = (PHYS == ADJ_THERMAL || PHYS == ADJ_TTHERMAL || PHYS == ADJ_ALLENCAHN) ? 1 : D;
  T X[A * 3], de[A], ue[A * DPN], ax[A];
  for (int b = 0; b < A; ++b) {
    const long long n = a.conn[e * A + b];
    for (int k = 0; k < 3; ++k) X[b * 3 + k] = a.xyz[n * 3 + k];
    de[b] = a.ctrl[n];
    ax[b] = a.aux ? a.aux[n] : (T)0;
    for (int k = 0; k < DPN; ++k) ue[b * DPN + k] = a.u[n * DPN + k];
  }
  a.dk[e] = element_energy<T, ELEM, ORDER, PHYS>(X, de, ue, a.aux ? ax : nullptr, a.p);
}

// the whole-element forward-mode sweeps (A*dim + A directions): the cross-check of the tests for every physics
template <class T, int ELEM, int ORDER, int PHYS>
__host__ __device__ inline void residual_adjoint_dual_reference_thread(long long e, const AdjointArgs<T>& a) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM);
  constexpr int DPN = (PHYS == ADJ_THERMAL || PHYS == ADJ_TTHERMAL || PHYS == ADJ_ALLENCAHN) ? 1 : D;
  T X[A * 3], de[A], ue[A * DPN], le[A * DPN], ax[A];
  for (int b = 0; b < A; ++b) {
    const long long n = a.conn[e * A + b];
    for (int k = 0; k < 3; ++k) X[b * 3 + k] = a.xyz[n * 3 + k];
    de[b] = a.ctrl[n];
    ax[b] = a.aux ? a.aux[n] : (T)0;
    for (int k = 0; k < DPN; ++k) {
      ue[b * DPN + k] = a.u[n * DPN + k];
      le[b * DPN + k] = a.lam[n * DPN + k];
    }
  }
  residual_adjoint_element_dual<T, ELEM, ORDER, PHYS>(X, de, ue, le, a.aux ? ax : nullptr, a.p, a.dk + e * A,
                                                      a.dx + e * (A * 3));
}

}  // namespace fol
