// Batched physics loss + its VJP (fe_loss.py:250-262 and the JAX-AD gradient of it).
//
// For every sample b the kernel produces, in ONE pass over the sample's nodal fields,
//   grad_u[b] = assembled UN-masked residual R(u_b)   (= dE_b/du_b because the reference
//               stop_gradient's the element residual, mechanical.py:116, thermal.py:45-49)
//   grad_k[b] = dE_b/dK_b  (thermal / neo-hooke; identically zero for mechanical)
//   E_b       = sum_e energy_e
// Matrix-free (Ke is never formed), node-centric: thread (node n, sample block) walks the
// node->element adjacency in fixed order, so the float sums are deterministic and need no atomics.
// Geometry factors (grad N, w detJ per Gauss point) are shared by all samples and come from the
// L2-resident geometry cache; each thread reuses them across S samples held in registers.
#pragma once
#include "assemble.cuh"

namespace fol {

template <class T>
struct EnergyArgs {
  const T* geom;           // [ne][NGP][A*D + 1]
  const int32_t* conn;
  const int32_t* adj_ptr;
  const int32_t* adj;
  const T* ctrl;           // (nb, nn)
  const T* u;              // (nb, ndof)
  T* grad_u;               // (nb, ndof)
  T* grad_k;               // (nb, nn) or null
  T* partial;              // (nb, gridDim.x) block partial energies
  long long ne, nn, nb;
  Params<T> p;
};

template <class T, int ELEM, int ORDER>
__global__ void geometry_cache_kernel(const T* __restrict__ xyz, const int32_t* __restrict__ conn, long long ne,
                                      T* __restrict__ geom) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM), NGP = elem_ngauss(ELEM, ORDER), W = A * D + 1;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ne * NGP) return;
  const long long e = t / NGP;
  const int g = (int)(t % NGP);
  T X[A * 3];
#pragma unroll
  for (int a = 0; a < A; ++a) {
    const long long n = conn[e * A + a];
#pragma unroll
    for (int k = 0; k < 3; ++k) X[a * 3 + k] = xyz[n * 3 + k];
  }
  double xi[3], w;
  gauss_point<ELEM, ORDER>(g, xi, w);
  T N[A], dN[A][D], gN[A][D];
  shape_functions<ELEM, T>(xi, N, dN);
  const T det = global_gradients<ELEM, T>(X, dN, gN);
  T* out = geom + t * W;
#pragma unroll
  for (int a = 0; a < A; ++a)
#pragma unroll
    for (int k = 0; k < D; ++k) out[a * D + k] = gN[a][k];
  out[A * D] = (T)w * det;
}

template <class T, int ELEM, int ORDER, int PHYS, int S, int BLOCK>
__global__ void __launch_bounds__(BLOCK) energy_grads_kernel(const EnergyArgs<T> args) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM), DPN = phys_dpn(PHYS, ELEM), ND = A * DPN;
  constexpr int NGP = elem_ngauss(ELEM, ORDER), W = A * D + 1, V = voigt_size(D);
  const long long n = (long long)blockIdx.x * BLOCK + threadIdx.x;
  const long long b0 = (long long)blockIdx.y * S;
  const long long ndof = args.nn * DPN;
  const Params<T>& P = args.p;

  T R[S][DPN], dK[S], en[S];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    dK[s] = (T)0;
    en[s] = (T)0;
#pragma unroll
    for (int k = 0; k < DPN; ++k) R[s][k] = (T)0;
  }

  if (n < args.nn) {
    T lam = (T)0, mu = (T)0;
    if constexpr (PHYS == MECH) {
      const T E = P.v[0], nu = P.v[1];
      if constexpr (D == 3) {
        const T c1 = E / (((T)1 + nu) * ((T)1 - (T)2 * nu));
        lam = c1 * nu;
        mu = c1 * (T)0.5 * ((T)1 - (T)2 * nu);
      } else {
        const T f = E / ((T)1 - nu * nu);
        lam = f * nu;
        mu = f * ((T)1 - nu) * (T)0.5;
      }
    }
    const int beg = args.adj_ptr[n], end = args.adj_ptr[n + 1];
    for (int it = beg; it < end; ++it) {
      const int ea = args.adj[it];
      const long long e = ea / A;
      const int a = ea % A;
      long long nodes[A];
#pragma unroll
      for (int b = 0; b < A; ++b) nodes[b] = args.conn[e * A + b];
      T ue[S][ND], de[S][A];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const long long bb = (b0 + s < args.nb) ? b0 + s : args.nb - 1;
#pragma unroll
        for (int b = 0; b < A; ++b) {
          de[s][b] = __ldg(args.ctrl + bb * args.nn + nodes[b]);
#pragma unroll
          for (int k = 0; k < DPN; ++k) ue[s][b * DPN + k] = __ldg(args.u + bb * ndof + nodes[b] * DPN + k);
        }
      }
#pragma unroll 1
      for (int g = 0; g < NGP; ++g) {
        const T* gm = args.geom + (e * NGP + g) * W;
        T gN[A][D];
#pragma unroll
        for (int b = 0; b < A; ++b)
#pragma unroll
          for (int k = 0; k < D; ++k) gN[b][k] = __ldg(gm + b * D + k);
        const T wd = __ldg(gm + A * D);
        double xi[3], w;
        gauss_point<ELEM, ORDER>(g, xi, w);
        T N[A], dN[A][D];
        shape_functions<ELEM, T>(xi, N, dN);
        // the local node `a` is a runtime index: select its row once
        T ga[D], Na = (T)0;
#pragma unroll
        for (int k = 0; k < D; ++k) ga[k] = (T)0;
#pragma unroll
        for (int b = 0; b < A; ++b)
          if (b == a) {
            Na = N[b];
#pragma unroll
            for (int k = 0; k < D; ++k) ga[k] = gN[b][k];
          }
#pragma unroll
        for (int s = 0; s < S; ++s) {
          if constexpr (PHYS == NEOHOOKE) break;
          T eg = (T)0;
#pragma unroll
          for (int b = 0; b < A; ++b) eg += N[b] * de[s][b];
          if constexpr (PHYS == THERMAL) {
            T tg = (T)0, gT[D];
#pragma unroll
            for (int k = 0; k < D; ++k) gT[k] = (T)0;
#pragma unroll
            for (int b = 0; b < A; ++b) {
              tg += N[b] * ue[s][b];
#pragma unroll
              for (int k = 0; k < D; ++k) gT[k] += gN[b][k] * ue[s][b];
            }
            const T beta = P.v[5], cexp = P.v[6];
            const T nl = (T)1 + ((beta != (T)0) ? beta * (T)pow((double)tg, (double)cexp) : (T)0);
            T flux = (T)0, g2 = (T)0;
#pragma unroll
            for (int k = 0; k < D; ++k) {
              flux += ga[k] * gT[k];
              g2 += gT[k] * gT[k];
            }
            R[s][0] += wd * eg * nl * flux;
            dK[s] += wd * Na * nl * g2;
          } else if constexpr (PHYS == MECH) {
            // H[i][j] = du_i/dx_j ; sigma = lam tr(eps) I + mu (H + H^T) scaled by wd*E_g
            T H[D][D];
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
              for (int j = 0; j < D; ++j) {
                T acc = (T)0;
#pragma unroll
                for (int b = 0; b < A; ++b) acc += gN[b][j] * ue[s][b * D + i];
                H[i][j] = acc;
              }
            T tr = (T)0;
#pragma unroll
            for (int i = 0; i < D; ++i) tr += H[i][i];
            const T cf = wd * eg;
#pragma unroll
            for (int i = 0; i < D; ++i) {
              T acc = lam * tr * ga[i];
#pragma unroll
              for (int j = 0; j < D; ++j) acc += mu * (H[i][j] + H[j][i]) * ga[j];
              R[s][i] += cf * acc;
            }
          }
        }
        if constexpr (PHYS == NEOHOOKE) {
          // energy = sum_g wd psi (true strain energy, mechanical_neohooke.py:262, 271): psi, S are
          // linear in E_g, so evaluate the law at unit modulus and scale.  dE/du = F_int (no Fe).
          const T nu = P.v[1];
          const T k1 = (T)1 / ((T)3 * ((T)1 - (T)2 * nu)), mu1 = (T)1 / ((T)2 * ((T)1 + nu));
#pragma unroll
          for (int s = 0; s < S; ++s) {
            T eg = (T)0;
#pragma unroll
            for (int b = 0; b < A; ++b) eg += N[b] * de[s][b];
            T F[D][D];
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
              for (int j = 0; j < D; ++j) {
                T acc = (i == j) ? (T)1 : (T)0;
#pragma unroll
                for (int b = 0; b < A; ++b) acc += gN[b][j] * ue[s][b * D + i];
                F[i][j] = acc;
              }
            T Sv[V], Cv[V * V];
            const T psi1 = neo_hooke_point<T, D>(F, k1, mu1, Sv, Cv);
            T Ba[V][D];
            neo_hooke_B<T, D>(&F[0][0], ga, Ba);
#pragma unroll
            for (int c = 0; c < D; ++c) {
              T acc = (T)0;
#pragma unroll
              for (int v = 0; v < V; ++v) acc += Ba[v][c] * Sv[v];
              R[s][c] += wd * eg * acc;
            }
            dK[s] += wd * Na * psi1;
            if (a == 0) en[s] += wd * eg * psi1;  // each element's energy is counted at its node 0
          }
        }
        if constexpr (PHYS == MECH) {
          // body force: Fe_a = b * w detJ N_a (same for every sample)
#pragma unroll
          for (int s = 0; s < S; ++s)
#pragma unroll
            for (int i = 0; i < D; ++i) R[s][i] -= P.v[2 + i] * wd * Na;
        }
      }
    }
    // write gradients, accumulate this node's share of the energy E_b = u_b . R_b
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const long long bb = b0 + s;
      if (bb < args.nb) {
#pragma unroll
        for (int k = 0; k < DPN; ++k) {
          if constexpr (PHYS != NEOHOOKE) {  // E_b = u_b . R_b (mechanical.py:116-117, thermal.py:45-49)
            const T uk = __ldg(args.u + bb * ndof + n * DPN + k);
            en[s] += uk * R[s][k];
          }
          args.grad_u[bb * ndof + n * DPN + k] = R[s][k];
        }
        if (args.grad_k) args.grad_k[bb * args.nn + n] = dK[s];
      }
    }
  }

  // deterministic block reduction of the per-sample energies (fixed tree)
  __shared__ T red[S][BLOCK / 32];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    T v = en[s];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[s][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < S) {
    T v = (T)0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; ++w) v += red[threadIdx.x][w];
    const long long bb = b0 + threadIdx.x;
    if (bb < args.nb) args.partial[bb * gridDim.x + blockIdx.x] = v;
  }
}

// E_b = sum of block partials (fixed order), then L = mean E_b^p, (min, max, mean), scale_b
template <class T>
__global__ void loss_reduce_kernel(const T* __restrict__ partial, long long nb, int nblocks, double exponent,
                                   T* __restrict__ energy, T* __restrict__ out4, T* __restrict__ scale) {
  // single block; thread-strided over samples, then a fixed-order tree
  __shared__ double s_sum[256], s_min[256], s_max[256];
  double lsum = 0.0, lmin = INFINITY, lmax = -INFINITY;
  for (long long b = threadIdx.x; b < nb; b += blockDim.x) {
    double E;
    if (nblocks > 0) {
      T acc = (T)0;
      for (int k = 0; k < nblocks; ++k) acc += partial[b * nblocks + k];
      energy[b] = acc;
      E = (double)acc;
    } else {
      E = (double)energy[b];
    }
    const double Ep = (exponent == 1.0) ? E : pow(E, exponent);
    const double dE = (exponent == 1.0) ? 1.0 : exponent * pow(E, exponent - 1.0);
    scale[b] = (T)(dE / (double)nb);
    lsum += Ep;
    lmin = fmin(lmin, Ep);
    lmax = fmax(lmax, Ep);
  }
  s_sum[threadIdx.x] = lsum;
  s_min[threadIdx.x] = lmin;
  s_max[threadIdx.x] = lmax;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
      s_min[threadIdx.x] = fmin(s_min[threadIdx.x], s_min[threadIdx.x + o]);
      s_max[threadIdx.x] = fmax(s_max[threadIdx.x], s_max[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mean = s_sum[0] / (double)nb;
    out4[0] = (T)mean;
    out4[1] = (T)s_min[0];
    out4[2] = (T)s_max[0];
    out4[3] = (T)mean;
  }
}

template <class T>
__global__ void scale_grads_kernel(long long nb, long long ndof, long long nn, const T* __restrict__ scale,
                                   T upstream, const uint8_t* __restrict__ dir, T* __restrict__ grad_u,
                                   T* __restrict__ grad_k) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long b = blockIdx.y;
  const T sc = upstream * scale[b];
  if (i < ndof) grad_u[b * ndof + i] = dir[i] ? (T)0 : sc * grad_u[b * ndof + i];
  if (grad_k && i < nn) grad_k[b * nn + i] = sc * grad_k[b * nn + i];
}

}  // namespace fol
