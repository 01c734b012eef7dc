// Generic element-stage kernel of the residual + Jacobian assembly (all elements, all physics).
//
// Replaces, in one launch per mesh:  the gathers of fe_loss.py:232-248, ComputeElement of
// mechanical.py:98-117 / thermal.py:28-49 / mechanical_neohooke.py:243-275, the transpose switch
// and Dirichlet row mask of fe_loss.py:191-230 and the BCOO `data` write of fe_loss.py:299.
//
// Work split: one *lane group* per element, one lane per element node (row block of Ke).
//   phase 0  lane a gathers node a (coords, control, dofs, Dirichlet flags) into shared memory
//   phase 1  lane a evaluates the Gauss points g = a, a+A, ... : J, det J, grad N, the point
//            coefficient / constitutive update -> shared memory (no redundant geometry)
//   phase 2  lane a accumulates its DPN x ND row block of Ke over all Gauss points in
//            registers, forms re = Ke u - Fe, applies transpose + row mask while storing.
// Groups never straddle a warp, so __syncwarp() is the only barrier.
#pragma once
#include <math.h>

#include "common.cuh"
#include "elements.cuh"

namespace fol {

enum : int { MECH = 0, THERMAL = 1, NEOHOOKE = 2, J2 = 3, STVK = 4, TTHERMAL = 5, ALLENCAHN = 6 };

// implicit-Euler scalar losses: (current field, next field) in the (control, dof) slots
__host__ __device__ constexpr bool implicit_scalar(int phys) { return phys == TTHERMAL || phys == ALLENCAHN; }

// finite-strain total-Lagrangian laws share one kernel skeleton (F-weighted B, geometric stiffness)
__host__ __device__ constexpr bool finite_strain(int phys) { return phys == NEOHOOKE || phys == STVK; }

__host__ __device__ constexpr int phys_dpn(int phys, int elem) {
  return (phys == THERMAL || phys == TTHERMAL || phys == ALLENCAHN) ? 1 : elem_dim(elem);
}
__host__ __device__ constexpr int voigt_size(int dim) { return dim == 3 ? 6 : 3; }
// Gauss-point history width of the J2 model (plasticity.py:122-130)
__host__ __device__ constexpr int j2_state_size(int dim) { return dim == 3 ? 7 : 4; }

template <class T>
struct AsmArgs {
  const T* xyz;
  const int32_t* conn;
  const T* ctrl;
  const T* u;
  const uint8_t* dir;
  T* ke;
  T* re;
  const T* state_in;
  T* state_out;
  long long ne;
  int transpose;
  Params<T> p;
  const T* v = nullptr;   // matrix-free mode: re <- Ke'(u) v_e (element products of J v), ke is not written
  // matrix-free mode over a batch of samples (grid.y = sample): strides, in values, of ctrl / (u, v) / re per sample
  long long batch_node = 0, batch_dof = 0, batch_elem = 0;
  int batch_count = 0;    // samples (grid.y); 0 = unbatched
};

// per-Gauss-point constitutive data handed from phase 1 to phase 2
template <int PHYS, int DIM>
struct PointDataSize {
  // MECH/THERMAL: nothing beyond coef.  NEOHOOKE: F (DIM*DIM) + S (V) + C (V*V).
  // J2: sigma (V) + tangent (V*V).
  static constexpr int V = voigt_size(DIM);
  // implicit scalar losses: cNN, cBB, cNB, rN, rB, energy density, grad(next field)[DIM]
  static constexpr int value = finite_strain(PHYS) ? DIM * DIM + V + V * V
                               : (PHYS == J2 ? V + V * V : (implicit_scalar(PHYS) ? 6 + DIM : 0));
};

template <class T, int ELEM, int ORDER, int PHYS>
struct GroupSmem {
  static constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM), DPN = phys_dpn(PHYS, ELEM);
  static constexpr int ND = A * DPN, NGP = elem_ngauss(ELEM, ORDER);
  static constexpr int PD = PointDataSize<PHYS, D>::value;
  static constexpr int RAW = A * 3 + A + A + ND + ND + NGP * A * D + NGP + NGP + NGP * A + NGP * PD;
  static constexpr int PAD = (RAW % 2 == 0) ? 1 : 2;  // odd element count -> groups on distinct banks
  T X[A * 3];
  T de[A];
  T aux[A];
  T u[ND];
  T bc[ND];
  T gN[NGP][A][D];
  T coef[NGP];
  T wdet[NGP];
  T Nw[NGP][A];
  T pd[NGP * PD + PAD];
};

template <class T, int ND, class F>
__device__ __forceinline__ void store_row(T* __restrict__ dst, F&& val) {
  constexpr int RB = ND * (int)sizeof(T);
  if constexpr (sizeof(T) == 8 && RB % 16 == 0) {
#pragma unroll
    for (int c = 0; c < ND; c += 2) {
      double2 v;
      v.x = val(c);
      v.y = val(c + 1);
      __stcs(reinterpret_cast<double2*>(dst) + c / 2, v);
    }
  } else if constexpr (sizeof(T) == 4 && RB % 16 == 0) {
#pragma unroll
    for (int c = 0; c < ND; c += 4) {
      float4 v;
      v.x = val(c);
      v.y = val(c + 1);
      v.z = val(c + 2);
      v.w = val(c + 3);
      __stcs(reinterpret_cast<float4*>(dst) + c / 4, v);
    }
  } else if constexpr (sizeof(T) == 4 && RB % 8 == 0) {
#pragma unroll
    for (int c = 0; c < ND; c += 2) {
      float2 v;
      v.x = val(c);
      v.y = val(c + 1);
      __stcs(reinterpret_cast<float2*>(dst) + c / 2, v);
    }
  } else {
#pragma unroll
    for (int c = 0; c < ND; ++c) __stcs(dst + c, (T)val(c));
  }
}

// x^c with small non-negative integer exponents by repeated multiplication (0^0 = 1, like jnp power)
template <class T>
__device__ __forceinline__ T pow_ci(T x, T c) {
  const int ci = (int)c;
  if ((T)ci == c && ci >= 0 && ci <= 16) {
    T r = (T)1, b = x;
    for (int k = ci; k; k >>= 1) {
      if (k & 1) r *= b;
      b *= b;
    }
    return r;
  }
  return (T)pow((double)x, (double)c);
}

// same as store_row but into the shared-memory staging area (plain vector stores)
// element matrices up to 1152 bytes (Tet4 mechanics f64) are staged; see assemble_kernel
template <class T>
__host__ __device__ constexpr bool stage_output(int nd) { return (size_t)nd * nd * sizeof(T) <= 1152; }

template <class T, int ND, class F>
__device__ __forceinline__ void stage_row(T* __restrict__ dst, F&& val) {
  constexpr int RB = ND * (int)sizeof(T);
  if constexpr (sizeof(T) == 8 && RB % 16 == 0) {
#pragma unroll
    for (int c = 0; c < ND; c += 2) reinterpret_cast<double2*>(dst)[c / 2] = make_double2(val(c), val(c + 1));
  } else if constexpr (sizeof(T) == 4 && RB % 16 == 0) {
#pragma unroll
    for (int c = 0; c < ND; c += 4)
      reinterpret_cast<float4*>(dst)[c / 4] = make_float4(val(c), val(c + 1), val(c + 2), val(c + 3));
  } else {
#pragma unroll
    for (int c = 0; c < ND; ++c) dst[c] = (T)val(c);
  }
}

// ---- constitutive point laws (phase 1) ---------------------------------------------------

// Neo-Hooke, neo_hooke.py:14-58 (2-D) and :64-109 (3-D).  Writes F, S (Voigt), C (Voigt x Voigt)
// into pd and returns the strain-energy density.  Voigt order [xx,yy,zz,yz,xz,xy] / [xx,yy,xy]
// (utils.py:14-32, 103-130).
template <class T, int D>
__device__ __forceinline__ T neo_hooke_point(const T (&F)[D][D], T k, T mu, T* S, T* Cv) {
  constexpr int V = voigt_size(D);
  T C[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T acc = (T)0;
#pragma unroll
      for (int m = 0; m < D; ++m) acc += F[m][i] * F[m][j];
      C[i][j] = acc;
    }
  T iC[D][D];
  T J, trC;
  if constexpr (D == 2) {
    const T dC = C[0][0] * C[1][1] - C[0][1] * C[1][0];
    const T r = (T)1 / dC;
    iC[0][0] = C[1][1] * r; iC[0][1] = -C[0][1] * r; iC[1][0] = -C[1][0] * r; iC[1][1] = C[0][0] * r;
    J = F[0][0] * F[1][1] - F[0][1] * F[1][0];
    trC = C[0][0] + C[1][1];
  } else {
    const T c00 = C[1][1] * C[2][2] - C[1][2] * C[2][1];
    const T c01 = C[1][2] * C[2][0] - C[1][0] * C[2][2];
    const T c02 = C[1][0] * C[2][1] - C[1][1] * C[2][0];
    const T dC = C[0][0] * c00 + C[0][1] * c01 + C[0][2] * c02;
    const T r = (T)1 / dC;
    iC[0][0] = c00 * r; iC[1][0] = c01 * r; iC[2][0] = c02 * r;
    iC[0][1] = (C[0][2] * C[2][1] - C[0][1] * C[2][2]) * r;
    iC[1][1] = (C[0][0] * C[2][2] - C[0][2] * C[2][0]) * r;
    iC[2][1] = (C[0][1] * C[2][0] - C[0][0] * C[2][1]) * r;
    iC[0][2] = (C[0][1] * C[1][2] - C[0][2] * C[1][1]) * r;
    iC[1][2] = (C[0][2] * C[1][0] - C[0][0] * C[1][2]) * r;
    iC[2][2] = (C[0][0] * C[1][1] - C[0][1] * C[1][0]) * r;
    J = F[0][0] * (F[1][1] * F[2][2] - F[1][2] * F[2][1]) + F[0][1] * (F[1][2] * F[2][0] - F[1][0] * F[2][2]) +
        F[0][2] * (F[1][0] * F[2][1] - F[1][1] * F[2][0]);
    trC = C[0][0] + C[1][1] + C[2][2];
  }
  const T invd = (T)1 / (T)D;
  const T p = (T)0.5 * k * (J - (T)1 / J);
  const T dp = (T)0.5 * k * ((T)1 + (T)1 / (J * J));
  const T Jm = (D == 2) ? (T)1 / J : (T)pow((double)J, -2.0 / 3.0);
  const T psi = (k * (T)0.25) * (J * J - (T)2 * (T)log((double)J) - (T)1) + (T)0.5 * mu * (Jm * trC - (T)D);
  T Siso[D][D], Sm[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      Siso[i][j] = Jm * mu * ((i == j ? (T)1 : (T)0) - invd * trC * iC[i][j]);
      Sm[i][j] = J * p * iC[i][j] + Siso[i][j];
    }
  const T cvol1 = J * p + dp * J * J, cvol2 = (T)2 * J * p;
  const T ciso1 = (T)2 * invd * Jm * mu * trC, ciso2 = (T)2 * invd;
  constexpr int vi3[6] = {0, 1, 2, 1, 0, 0}, vj3[6] = {0, 1, 2, 2, 2, 1};
  constexpr int vi2[3] = {0, 1, 0}, vj2[3] = {0, 1, 1};
#pragma unroll
  for (int I = 0; I < V; ++I) {
    const int i = D == 3 ? vi3[I] : vi2[I], j = D == 3 ? vj3[I] : vj2[I];
    S[I] = Sm[i][j];
#pragma unroll
    for (int Jv = 0; Jv < V; ++Jv) {
      // the 2-D Voigt map mirrors the upper triangle (utils.py:107-116)
      const int I2 = (D == 2 && Jv < I) ? Jv : I, J2v = (D == 2 && Jv < I) ? I : Jv;
      const int ii = D == 3 ? i : vi2[I2], jj = D == 3 ? j : vj2[I2];
      const int kk = D == 3 ? vi3[Jv] : vi2[J2v], ll = D == 3 ? vj3[Jv] : vj2[J2v];
      const T ii_kl = iC[ii][jj] * iC[kk][ll];
      const T dsp = (T)0.5 * (iC[ii][kk] * iC[jj][ll] + iC[ii][ll] * iC[jj][kk]);
      Cv[I * V + Jv] = cvol1 * ii_kl - cvol2 * dsp + ciso1 * (dsp - invd * ii_kl) -
                       ciso2 * (iC[ii][jj] * Siso[kk][ll] + Siso[ii][jj] * iC[kk][ll]);
    }
  }
  return psi;
}

// Saint-Venant-Kirchhoff, saint_venant.py:11-33: E = (F^T F - I)/2, psi = lam/2 tr(E)^2 + mu tr(E E),
// S = lam tr(E) I + 2 mu E.  The reference builds the tangent from the UNsymmetrised fourth-order
// identity (utils.py fourth_order_identity_tensor), so its Voigt shear entries are 2 mu (kept as is).
template <class T, int D>
__device__ __forceinline__ T st_venant_point(const T (&F)[D][D], T lam, T mu, T* S, T* Cv) {
  constexpr int V = voigt_size(D);
  T E[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T acc = (i == j) ? (T)-1 : (T)0;
#pragma unroll
      for (int m = 0; m < D; ++m) acc += F[m][i] * F[m][j];
      E[i][j] = (T)0.5 * acc;
    }
  T tr = (T)0, ee = (T)0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    tr += E[i][i];
#pragma unroll
    for (int j = 0; j < D; ++j) ee += E[i][j] * E[j][i];
  }
  constexpr int vi3[6] = {0, 1, 2, 1, 0, 0}, vj3[6] = {0, 1, 2, 2, 2, 1};
  constexpr int vi2[3] = {0, 1, 0}, vj2[3] = {0, 1, 1};
#pragma unroll
  for (int I = 0; I < V; ++I) {
    const int i = D == 3 ? vi3[I] : vi2[I], j = D == 3 ? vj3[I] : vj2[I];
    S[I] = (i == j ? lam * tr : (T)0) + (T)2 * mu * E[i][j];
#pragma unroll
    for (int Jv = 0; Jv < V; ++Jv) {
      const int nrm = D;  // first D Voigt entries are the normal components
      T c = (T)0;
      if (I < nrm && Jv < nrm) c = lam + (I == Jv ? (T)2 * mu : (T)0);
      else if (I == Jv) c = (T)2 * mu;
      Cv[I * V + Jv] = c;
    }
  }
  return (T)0.5 * lam * tr * tr + mu * ee;
}

// row a of the F-weighted strain-displacement matrix, mechanical_neohooke.py:49-91:
// Ba[s][c] for Voigt row s and displacement component c of node a.
template <class T, int D>
__device__ __forceinline__ void neo_hooke_B(const T* F /* D*D row-major */, const T* g /* D */,
                                            T (&Ba)[voigt_size(D)][D]) {
#pragma unroll
  for (int c = 0; c < D; ++c) {
    if constexpr (D == 2) {
      Ba[0][c] = F[c * 2 + 0] * g[0];
      Ba[1][c] = F[c * 2 + 1] * g[1];
      Ba[2][c] = F[c * 2 + 1] * g[0] + F[c * 2 + 0] * g[1];
    } else {
      Ba[0][c] = F[c * 3 + 0] * g[0];
      Ba[1][c] = F[c * 3 + 1] * g[1];
      Ba[2][c] = F[c * 3 + 2] * g[2];
      Ba[3][c] = F[c * 3 + 1] * g[2] + F[c * 3 + 2] * g[1];
      Ba[4][c] = F[c * 3 + 0] * g[2] + F[c * 3 + 2] * g[0];
      Ba[5][c] = F[c * 3 + 0] * g[1] + F[c * 3 + 1] * g[0];
    }
  }
}

// row a of the linear strain-displacement matrix, mechanical.py:37-58
// (rows [xx,yy,zz,xy,yz,xz] / [xx,yy,xy]).
template <class T, int D>
__device__ __forceinline__ void linear_B(const T* g, T (&Ba)[voigt_size(D)][D]) {
#pragma unroll
  for (int s = 0; s < voigt_size(D); ++s)
#pragma unroll
    for (int c = 0; c < D; ++c) Ba[s][c] = (T)0;
  if constexpr (D == 2) {
    Ba[0][0] = g[0]; Ba[1][1] = g[1]; Ba[2][0] = g[1]; Ba[2][1] = g[0];
  } else {
    Ba[0][0] = g[0]; Ba[1][1] = g[1]; Ba[2][2] = g[2];
    Ba[3][0] = g[1]; Ba[3][1] = g[0];
    Ba[4][1] = g[2]; Ba[4][2] = g[1];
    Ba[5][0] = g[2]; Ba[5][2] = g[0];
  }
}

}  // namespace fol

#include "j2_point.cuh"

namespace fol {

// MATVEC = matrix-free mode (args.v): a separate instantiation, so the assembling kernels carry none of its
// registers or branches (as a run-time flag it cost them 10-20 %)
template <class T, int ELEM, int ORDER, int PHYS, int BLOCK, bool MATVEC = false>
__global__ void __launch_bounds__(BLOCK) assemble_kernel(const AsmArgs<T> args_in) {
  // batched matrix-free mode: sample blockIdx.y reads / writes its own slices (the assembling kernels see args_in as is)
  AsmArgs<T> shifted;
  const AsmArgs<T>* ap = &args_in;
  if constexpr (MATVEC) {
    shifted = args_in;
    const long long b = blockIdx.y;
    shifted.ctrl += b * args_in.batch_node;
    shifted.u += b * args_in.batch_dof;
    shifted.v += b * args_in.batch_dof;
    shifted.re += b * args_in.batch_elem;
    ap = &shifted;
  }
  const AsmArgs<T>& args = *ap;
  using SM = GroupSmem<T, ELEM, ORDER, PHYS>;
  constexpr int A = SM::A, D = SM::D, DPN = SM::DPN, ND = SM::ND, NGP = SM::NGP, PD = SM::PD;
  constexpr int GW = (A == 8) ? 8 : 4;  // lanes per element group (tri pads 3 -> 4)
  constexpr int GPB = BLOCK / GW;
  constexpr int V = voigt_size(D);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM* groups = reinterpret_cast<SM*>(smem_raw);

  const int grp = threadIdx.x / GW, a = threadIdx.x % GW;
  const long long e = (long long)blockIdx.x * GPB + grp;
  const bool active = (e < args.ne) && (a < A);
  SM& sm = groups[grp];
  const Params<T>& P = args.p;

  // ---- phase 0: gather (fe_loss.py:240-247; BC / mask vectors of :268-271 are one byte flag)
  if (active) {
    const long long n = args.conn[e * A + a];
#pragma unroll
    for (int k = 0; k < 3; ++k) sm.X[a * 3 + k] = __ldg(args.xyz + n * 3 + k);
    sm.de[a] = __ldg(args.ctrl + n);
    if constexpr (PHYS == TTHERMAL) sm.aux[a] = __ldg(args.state_in + n);   // nodal heterogeneity k0
#pragma unroll
    for (int k = 0; k < DPN; ++k) {
      sm.u[a * DPN + k] = __ldg(args.u + n * DPN + k);
      sm.bc[a * DPN + k] = args.dir[n * DPN + k] ? (T)0 : (T)1;
    }
  }
  __syncwarp();

  // ---- phase 1: Gauss points g = a, a+A, ...
  if (active) {
#pragma unroll 1
    for (int g = a; g < NGP; g += A) {
      double xi[3], w;
      gauss_point<ELEM, ORDER>(g, xi, w);
      T N[A], dN[A][D], gN[A][D];
      shape_functions<ELEM, T>(xi, N, dN);
      const T det = global_gradients<ELEM, T, implicit_scalar(PHYS)>(sm.X, dN, gN);
      const T wd = (T)w * det;
      T eg = (T)0;
#pragma unroll
      for (int b = 0; b < A; ++b) {
        eg += N[b] * sm.de[b];
        sm.Nw[g][b] = implicit_scalar(PHYS) ? N[b] : wd * N[b];
#pragma unroll
        for (int k = 0; k < D; ++k) sm.gN[g][b][k] = gN[b][k];
      }
      sm.wdet[g] = wd;
      if constexpr (PHYS == MECH) {
        sm.coef[g] = wd * eg;
      } else if constexpr (PHYS == THERMAL) {
        T tg = (T)0;
#pragma unroll
        for (int b = 0; b < A; ++b) tg += N[b] * sm.u[b];
        const T beta = P.v[5], cexp = P.v[6];
        const T nl = (beta != (T)0) ? beta * (T)pow((double)tg, (double)cexp) : (T)0;
        sm.coef[g] = wd * eg * ((T)1 + nl);
      } else if constexpr (implicit_scalar(PHYS)) {
        // eg = N . (current field); next field and its gradient at the point
        T fn = (T)0, gf[D];
#pragma unroll
        for (int k = 0; k < D; ++k) gf[k] = (T)0;
#pragma unroll
        for (int b = 0; b < A; ++b) {
          fn += N[b] * sm.u[b];
#pragma unroll
          for (int k = 0; k < D; ++k) gf[k] += gN[b][k] * sm.u[b];
        }
        T g2 = (T)0;
#pragma unroll
        for (int k = 0; k < D; ++k) g2 += gf[k] * gf[k];
        const T dt = P.v[10];
        T* pd = sm.pd + g * PD;
        if constexpr (PHYS == TTHERMAL) {   // transient_thermal.py:42-73
          const T beta = P.v[5], cexp = P.v[6], rcp = P.v[8] * P.v[9];
          T kg = (T)0;   // heterogeneity k0 (auxiliary nodal field) at the point
#pragma unroll
          for (int b = 0; b < A; ++b) kg += N[b] * sm.aux[b];
          const T Kg = kg * ((T)1 + ((beta != (T)0) ? beta * pow_ci<T>(fn, cexp) : (T)0));
          const T dk = (beta != (T)0) ? kg * beta * cexp * pow_ci<T>(fn, cexp - (T)1) : (T)0;
          pd[0] = rcp * wd;
          pd[1] = dt * wd * Kg;
          pd[2] = dt * wd * dk;
          pd[3] = rcp * wd * (fn - eg);
          pd[4] = dt * wd * Kg;
          pd[5] = (T)0.5 * Kg * wd * g2 + rcp * (T)0.5 / dt * wd * (fn - eg) * (fn - eg);
        } else {                            // phase_field.py:38-70
          const T ie2 = (T)1 / (P.v[11] * P.v[11]);
          pd[0] = wd * ((T)1 - dt * ie2 * ((T)3 * fn * fn - (T)1));
          pd[1] = dt * wd;
          pd[2] = (T)0;
          pd[3] = wd * ((fn - eg) + dt * ie2 * (fn * fn - (T)1) * fn);
          pd[4] = dt * wd;
          pd[5] = (T)0.5 * wd * g2 + wd * (T)0.25 * (fn * fn - (T)1) * (fn * fn - (T)1) * ie2 +
                  (T)0.5 / dt * wd * (fn - eg) * (fn - eg);
        }
#pragma unroll
        for (int k = 0; k < D; ++k) pd[6 + k] = gf[k];
        sm.coef[g] = wd;
      } else if constexpr (finite_strain(PHYS)) {
        T F[D][D];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) {
            T acc = (i == j) ? (T)1 : (T)0;
#pragma unroll
            for (int b = 0; b < A; ++b) acc += gN[b][j] * sm.u[b * D + i];
            F[i][j] = acc;
          }
        const T nu = P.v[1];
        const T kk = eg / ((T)3 * ((T)1 - (T)2 * nu)), mu = eg / ((T)2 * ((T)1 + nu));
        T* pd = sm.pd + g * PD;
        T S[V], Cv[V * V];
        if constexpr (PHYS == NEOHOOKE) neo_hooke_point<T, D>(F, kk, mu, S, Cv);
        else st_venant_point<T, D>(F, eg * nu / (((T)1 + nu) * ((T)1 - (T)2 * nu)), mu, S, Cv);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) pd[i * D + j] = F[i][j];
#pragma unroll
        for (int s = 0; s < V; ++s) pd[D * D + s] = S[s];
#pragma unroll
        for (int s = 0; s < V * V; ++s) pd[D * D + V + s] = Cv[s];
        sm.coef[g] = wd;
      } else {  // J2: strain = B u, return mapping + forward-mode tangent, history update
        T eps[V];
#pragma unroll
        for (int s = 0; s < V; ++s) eps[s] = (T)0;
#pragma unroll
        for (int b = 0; b < A; ++b) {
          T Bb[V][D];
          linear_B<T, D>(gN[b], Bb);
#pragma unroll
          for (int s = 0; s < V; ++s)
#pragma unroll
            for (int c = 0; c < D; ++c) eps[s] += Bb[s][c] * sm.u[b * D + c];
        }
        constexpr int NS = j2_state_size(D);
        T st[NS], st_new[NS];
        const long long sbase = (e * NGP + g) * NS;
#pragma unroll
        for (int s = 0; s < NS; ++s) st[s] = args.state_in[sbase + s];
        T* pd = sm.pd + g * PD;
        j2_point<T, D>(eps, st, P.v[0], P.v[1], P.v[5], P.v[6], P.v[7], pd, pd + V, st_new);
#pragma unroll
        for (int s = 0; s < NS; ++s)
          if (args.state_out != nullptr) args.state_out[sbase + s] = st_new[s];
        sm.coef[g] = wd;
      }
    }
  }
  __syncwarp();

  // small element matrices are staged in shared memory ([group][ND*ND], element-major like the output)
  // and leave as one bulk copy per warp; large ones (Hex8 mechanics) are stored directly
  constexpr bool STAGE = stage_output<T>(ND);
  T* stage_all = reinterpret_cast<T*>(smem_raw + sizeof(SM) * GPB);
  T* st = stage_all + (size_t)grp * (ND * ND);
  // matrix-free mode: the region after the groups holds, per group, v_e (ND) and the lanes' partial
  // products of Ke^T v_e (A x ND) instead of staged matrices
  constexpr bool matvec = MATVEC;
  T* mv = stage_all + (size_t)grp * (ND * (A + 1));
  T mv_diag[DPN];
#pragma unroll
  for (int i = 0; i < DPN; ++i) mv_diag[i] = (T)0;
  if constexpr (matvec) {
    if (active) {
      const long long n = args.conn[e * A + a];
#pragma unroll
      for (int k = 0; k < DPN; ++k) mv[a * DPN + k] = __ldg(args.v + n * DPN + k);
    }
    __syncwarp();
  }
  if (active) {
  // ---- phase 2: row block a of Ke, re = Ke u - Fe
  T K[A][DPN][DPN];
  T fint[DPN];
#pragma unroll
  for (int b = 0; b < A; ++b)
#pragma unroll
    for (int i = 0; i < DPN; ++i)
#pragma unroll
      for (int j = 0; j < DPN; ++j) K[b][i][j] = (T)0;
#pragma unroll
  for (int i = 0; i < DPN; ++i) fint[i] = (T)0;

  if constexpr (implicit_scalar(PHYS)) {
    // Ke_ab = sum_g cNN N_a N_b + cBB gN_a.gN_b + cNB N_a (g.gN_b);  re_a = sum_g rN N_a + rB gN_a.g
#pragma unroll 1
    for (int g = 0; g < NGP; ++g) {
      const T* pd = sm.pd + g * PD;
      const T Na = sm.Nw[g][a];
      T ga[D], gf[D], gag = (T)0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        ga[k] = sm.gN[g][a][k];
        gf[k] = pd[6 + k];
        gag += ga[k] * gf[k];
      }
      fint[0] += pd[3] * Na + pd[4] * gag;
#pragma unroll
      for (int b = 0; b < A; ++b) {
        T bb = (T)0, gb = (T)0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          bb += ga[k] * sm.gN[g][b][k];
          gb += gf[k] * sm.gN[g][b][k];
        }
        K[b][0][0] += pd[0] * Na * sm.Nw[g][b] + pd[1] * bb + pd[2] * Na * gb;
      }
    }
    if (a == 0 && args.state_out != nullptr) {   // element energy (the first return value of ComputeElement)
      T en = (T)0;
#pragma unroll 1
      for (int g = 0; g < NGP; ++g) en += sm.pd[g * PD + 5];
      args.state_out[e] = en;
    }
  } else if constexpr (PHYS == MECH || PHYS == THERMAL) {
    // K holds P_ab = sum_g coef g_a (x) g_b (MECH) or the scalar sum_g coef g_a.g_b (THERMAL)
#pragma unroll(NGP <= 8 ? NGP : 1)
    for (int g = 0; g < NGP; ++g) {
      const T cf = sm.coef[g];
      T ga[D];
#pragma unroll
      for (int i = 0; i < D; ++i) ga[i] = cf * sm.gN[g][a][i];
#pragma unroll
      for (int b = 0; b < A; ++b) {
        if constexpr (PHYS == MECH) {
#pragma unroll
          for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) K[b][i][j] += ga[i] * sm.gN[g][b][j];
        } else {
          T acc = K[b][0][0];
#pragma unroll
          for (int i = 0; i < D; ++i) acc += ga[i] * sm.gN[g][b][i];
          K[b][0][0] = acc;
        }
      }
    }
    if constexpr (PHYS == MECH) {
      // B^T D B of an isotropic D (mechanical.py:60-82) in terms of P: lam P + mu P^T + mu tr(P) I
      // with (lam, mu) = (D01, D_shear): 3-D c3, c4;  2-D plane stress E nu/(1-nu^2), E/(2(1+nu)).
      const T E = P.v[0], nu = P.v[1];
      T lam, mu;
      if constexpr (D == 3) {
        const T c1 = E / (((T)1 + nu) * ((T)1 - (T)2 * nu));
        lam = c1 * nu;
        mu = c1 * (T)0.5 * ((T)1 - (T)2 * nu);
      } else {
        const T f = E / ((T)1 - nu * nu);
        lam = f * nu;
        mu = f * ((T)1 - nu) * (T)0.5;
      }
#pragma unroll
      for (int b = 0; b < A; ++b) {
        T Pm[D][D];
        T tr = (T)0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
          tr += K[b][i][i];
#pragma unroll
          for (int j = 0; j < D; ++j) Pm[i][j] = K[b][i][j];
        }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) K[b][i][j] = lam * Pm[i][j] + mu * Pm[j][i] + (i == j ? mu * tr : (T)0);
      }
    }
    // re = Ke u (Fe subtracted below)
#pragma unroll
    for (int b = 0; b < A; ++b)
#pragma unroll
      for (int i = 0; i < DPN; ++i)
#pragma unroll
        for (int j = 0; j < DPN; ++j) fint[i] += K[b][i][j] * sm.u[b * DPN + j];
  } else {
    // NEOHOOKE: Ke_ab = sum_g wd (B_a^T C B_b + (g_a.S g_b) I), fint_a = sum_g wd B_a^T S
    // J2:       Ke_ab = sum_g wd  B_a^T Ct B_b,                 fint_a = sum_g wd B_a^T sigma
#pragma unroll 1
    for (int g = 0; g < NGP; ++g) {
      const T wd = sm.coef[g];
      const T* pd = sm.pd + g * PD;
      const T* Sv = finite_strain(PHYS) ? pd + D * D : pd;
      const T* Cv = Sv + V;
      T Ba[V][D];
      if constexpr (finite_strain(PHYS)) neo_hooke_B<T, D>(pd, sm.gN[g][a], Ba);
      else linear_B<T, D>(sm.gN[g][a], Ba);
      // BtC[c][t] = wd * sum_s Ba[s][c] C[s][t]
      T BtC[D][V];
#pragma unroll
      for (int c = 0; c < D; ++c)
#pragma unroll
        for (int t = 0; t < V; ++t) {
          T acc = (T)0;
#pragma unroll
          for (int s = 0; s < V; ++s) acc += Ba[s][c] * Cv[s * V + t];
          BtC[c][t] = wd * acc;
        }
#pragma unroll
      for (int c = 0; c < D; ++c) {
        T acc = (T)0;
#pragma unroll
        for (int s = 0; s < V; ++s) acc += Ba[s][c] * Sv[s];
        fint[c] += wd * acc;
      }
      T Sg[D];  // S_mat g_a (geometric stiffness, mechanical_neohooke.py:107-241)
      if constexpr (finite_strain(PHYS)) {
        constexpr int vmap3[3][3] = {{0, 5, 4}, {5, 1, 3}, {4, 3, 2}};
        constexpr int vmap2[2][2] = {{0, 2}, {2, 1}};
#pragma unroll
        for (int i = 0; i < D; ++i) {
          T acc = (T)0;
#pragma unroll
          for (int j = 0; j < D; ++j) acc += Sv[D == 3 ? vmap3[i][j] : vmap2[i][j]] * sm.gN[g][a][j];
          Sg[i] = wd * acc;
        }
      }
#pragma unroll
      for (int b = 0; b < A; ++b) {
        T Bb[V][D];
        if constexpr (finite_strain(PHYS)) neo_hooke_B<T, D>(pd, sm.gN[g][b], Bb);
        else linear_B<T, D>(sm.gN[g][b], Bb);
        T geo = (T)0;
        if constexpr (finite_strain(PHYS)) {
#pragma unroll
          for (int i = 0; i < D; ++i) geo += Sg[i] * sm.gN[g][b][i];
        }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) {
            T acc = (i == j) ? geo : (T)0;
#pragma unroll
            for (int t = 0; t < V; ++t) acc += BtC[i][t] * Bb[t][j];
            K[b][i][j] += acc;
          }
      }
    }
  }

  // body force Fe_a = b * sum_g w detJ N_a (mechanical.py:110; the thermal generic path has none)
  if constexpr (PHYS != THERMAL && !implicit_scalar(PHYS)) {
    T nw = (T)0;
#pragma unroll 1
    for (int g = 0; g < NGP; ++g) nw += sm.Nw[g][a];
#pragma unroll
    for (int i = 0; i < DPN; ++i) fint[i] -= P.v[2 + i] * nw;
  }

  if constexpr (matvec) {
    // ---- matrix-free product with the masked element matrix (fe_loss.py:191-230 applied to Ke or Ke^T):
    //      y_r = free row ? sum_c Ke(^T)[r][c] v_c : Ke[r][r] v_r
    if (!args.transpose) {
#pragma unroll
      for (int i = 0; i < DPN; ++i) {
        const int r = a * DPN + i;
        T acc = (T)0;
#pragma unroll
        for (int c = 0; c < ND; ++c) acc += K[c / DPN][i][c % DPN] * mv[c];
        args.re[e * ND + r] = (sm.bc[r] != (T)0) ? acc : K[a][i][i] * mv[r];
      }
    } else {
      // lane a holds rows (a, i) of Ke = columns of Ke^T: its share of (Ke^T v)_(b,j) is sum_i K[b][i][j] v_(a,i);
      // the shares meet in shared memory and are summed in lane order after the warp barrier below
#pragma unroll
      for (int b = 0; b < A; ++b)
#pragma unroll
        for (int j = 0; j < DPN; ++j) {
          T acc = (T)0;
#pragma unroll
          for (int i = 0; i < DPN; ++i) acc += K[b][i][j] * mv[a * DPN + i];
          mv[ND + a * ND + b * DPN + j] = acc;
        }
#pragma unroll
      for (int i = 0; i < DPN; ++i) mv_diag[i] = K[a][i][i];
    }
  } else {
  // ---- store: transpose switch + Dirichlet row mask (fe_loss.py:191-230), data of :299
#pragma unroll
  for (int i = 0; i < DPN; ++i) {
    const int r = a * DPN + i;
    args.re[e * ND + r] = sm.bc[r] * fint[i];
  }
  T* ke = args.ke + e * (long long)(ND * ND);
  if (!args.transpose) {
#pragma unroll
    for (int i = 0; i < DPN; ++i) {
      const int r = a * DPN + i;
      const bool freerow = sm.bc[r] != (T)0;
      auto val = [&](int c) -> T {
        const T v = K[c / DPN][i][c % DPN];
        return (freerow || c == r) ? v : (T)0;
      };
      if constexpr (STAGE) stage_row<T, ND>(st + r * ND, val);
      else store_row<T, ND>(ke + r * ND, val);
    }
  } else {
    // lane a holds rows a*DPN+i of Ke = columns of Ke^T
#pragma unroll
    for (int b = 0; b < A; ++b)
#pragma unroll
      for (int j = 0; j < DPN; ++j) {
        const int r = b * DPN + j;  // row of Ke^T
        const bool freerow = sm.bc[r] != (T)0;
#pragma unroll
        for (int i = 0; i < DPN; ++i) {
          const int c = a * DPN + i;
          (STAGE ? st : ke)[r * ND + c] = (freerow || c == r) ? K[b][i][j] : (T)0;
        }
      }
  }
  }  // !matvec
  }  // active

  if constexpr (matvec) {
    if (args.transpose) {
      __syncwarp();
      if (active) {
#pragma unroll
        for (int i = 0; i < DPN; ++i) {
          const int r = a * DPN + i;
          T acc = (T)0;
#pragma unroll
          for (int b = 0; b < A; ++b) acc += mv[ND + b * ND + r];
          args.re[e * ND + r] = (sm.bc[r] != (T)0) ? acc : mv_diag[i] * mv[r];
        }
      }
    }
    return;
  }

  // ---- the warp's elements are consecutive: their staged matrices leave as ONE contiguous bulk
  // async copy (cp.async.bulk shared -> global, TMA engine) instead of scattered 16-byte stores
  if constexpr (STAGE) {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncwarp();
    constexpr int GPWARP = 32 / GW;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long e_w = (long long)blockIdx.x * GPB + (long long)warp * GPWARP;
    long long cnt = args.ne - e_w;
    cnt = cnt < 0 ? 0 : (cnt > GPWARP ? GPWARP : cnt);
    const T* src = stage_all + (size_t)warp * GPWARP * (ND * ND);
    T* dst = args.ke + e_w * (long long)(ND * ND);
    const unsigned bytes = (unsigned)(cnt * ND * ND * sizeof(T));
    const bool bulk_ok = (bytes % 16u == 0u) && ((reinterpret_cast<unsigned long long>(dst) & 15ull) == 0ull);
    if (cnt > 0) {
      if (bulk_ok) {
        if (lane == 0) {
          const unsigned saddr = (unsigned)__cvta_generic_to_shared(src);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(saddr), "r"(bytes)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");  // smem must outlive the copy
        }
      } else {  // ragged tail / unaligned output: plain coalesced copy
        for (int idx = lane; idx < (int)(cnt * ND * ND); idx += 32) dst[idx] = src[idx];
      }
    }
  }
}

template <class T, int ELEM, int ORDER, int PHYS>
int launch_assemble(cudaStream_t s, const AsmArgs<T>& args) {
  using SM = GroupSmem<T, ELEM, ORDER, PHYS>;
  constexpr int GW = (SM::A == 8) ? 8 : 4;
  // 128 threads unless the per-group staging (high-order rules with constitutive point data)
  // would not leave room for two resident blocks per SM
  constexpr size_t kBudget = 100 * 1024;
  constexpr size_t kPerGroup = sizeof(SM) + (stage_output<T>(SM::ND) ? sizeof(T) * SM::ND * SM::ND : 0);  // + staged Ke'
  constexpr int BLOCK = kPerGroup * (128 / GW) <= kBudget ? 128 : (kPerGroup * (64 / GW) <= kBudget ? 64 : 32);
  constexpr int GPB = BLOCK / GW;
  // matrix-free mode keeps v_e and the partial products (ND * (A + 1) values per group) where the staged Ke' would be
  constexpr size_t kPerGroupMv = sizeof(SM) + sizeof(T) * SM::ND * (SM::A + 1);
  constexpr size_t kMaxSmem = (kPerGroup > kPerGroupMv ? kPerGroup : kPerGroupMv) * GPB;
  const size_t smem = (args.v ? kPerGroupMv : kPerGroup) * GPB;
  const long long grid = cdiv(args.ne, GPB);
  if (args.v) {
    auto kern = assemble_kernel<T, ELEM, ORDER, PHYS, BLOCK, true>;
    static PerDeviceOnce configured;
    if (configured.need()) {
      FOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
      configured.done();
    }
    if (grid == 0) return FOL_OK;
    kern<<<dim3((unsigned)grid, (unsigned)(args.batch_count > 0 ? args.batch_count : 1)), BLOCK, smem, s>>>(args);
    return check_launch("assemble_kernel (matrix-free)");
  }
  auto kern = assemble_kernel<T, ELEM, ORDER, PHYS, BLOCK, false>;
  static PerDeviceOnce configured;
  if (configured.need()) {
    FOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
    configured.done();
  }
  if (grid == 0) return FOL_OK;
  kern<<<(unsigned)grid, BLOCK, smem, s>>>(args);
  return check_launch("assemble_kernel");
}

template <class T, int PHYS>
int dispatch_assemble(cudaStream_t s, int element, int num_gp, const AsmArgs<T>& args) {
#define FOL_CASE(E, O) \
  if (element == E && num_gp == O) return launch_assemble<T, E, O, PHYS>(s, args);
  FOL_CASE(HEX, 1) FOL_CASE(HEX, 2) FOL_CASE(HEX, 3)
  FOL_CASE(QUAD, 1) FOL_CASE(QUAD, 2) FOL_CASE(QUAD, 3)
  FOL_CASE(TET, 1) FOL_CASE(TET, 2) FOL_CASE(TET, 3)
  FOL_CASE(TRI, 1) FOL_CASE(TRI, 2) FOL_CASE(TRI, 3)
#undef FOL_CASE
  return fail(FOL_ERR_UNSUPPORTED, "unsupported element / num_gp");
}

}  // namespace fol
