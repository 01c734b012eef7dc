// Batched physics loss + its VJP (fe_loss.py:250-262 and the JAX-AD gradient of it).
//
// For every sample b:
//   grad_u[b] = assembled UN-masked residual R(u_b)   (= dE_b/du_b because the reference
//               stop_gradient's the element residual, mechanical.py:116, thermal.py:45-49;
//               for Neo-Hooke dE/du = F_int, mechanical_neohooke.py:262-275)
//   grad_k[b] = dE_b/dK_b  (thermal / neo-hooke; identically zero for mechanical)
//   E_b       = sum_e energy_e
// Matrix-free (Ke is never formed), atomics-free and single-pass: ONE kernel, no HBM scratch.
// The mesh plan splits the nodes into compact tiles (Morton-ordered, <= 128 nodes) and lists, per
// tile, the elements touching it.  A CTA owns (tile, a strided set of samples):
//   phase A  thread (tile element, S samples): each listed element is evaluated once per sample from
//            the geometry cache (grad N, w detJ per Gauss point; SoA, shared by all samples, L1/L2
//            resident) and its vectors re_e, dK_e (psi_e) go to shared memory;
//   phase B  thread (tile node, S samples): fixed-order sum over the node's adjacency (deterministic),
//            coalesced write of grad_u / grad_k, block reduction of the sample's energy share.
// Elements on tile borders are evaluated by each tile that needs them (~1.2x for compact tiles)
// instead of being exchanged through memory.
#pragma once
#include "assemble.cuh"

namespace fol {

template <class T>
struct EnergyArgs {
  const T* geom;              // [NGP][A*D + 1][ne]
  const int32_t* conn;
  const int32_t* adj_ptr;     // node -> adjacency range (same order as fol_node_adjacency)
  const int32_t* adj_local;   // per adjacency entry: (element index within the node's tile)*A + a
  const int32_t* tile_node_ptr;
  const int32_t* tile_nodes;  // node ids grouped by tile
  const int32_t* tile_elem_ptr;
  const int32_t* tile_elems;  // element ids grouped by tile
  const int32_t* tile_conn;   // (len(tile_elems), A): element nodes in tile-local numbering
  const int32_t* tile_lnode_ptr;
  const int32_t* tile_lnodes; // per tile: the nodes its elements touch (owned nodes first, tile order)
  const T* ctrl;              // (nb, nn)
  const T* u;                 // (nb, ndof)
  T* grad_u;                  // (nb, ndof)
  T* grad_k;                  // (nb, nn) or null
  T* partial;                 // (nb, ntiles [* warps]) per-tile energy shares
  const T* dir_values;        // (ndof) Dirichlet value per dof, NaN where free: overwrites u while staging
                              //        (fe_loss.py:91-92, 255), or null
  const uint8_t* dir_flag;    // (ndof) 1 where grad_u is to be written as zero (the cotangent is cut at the
                              //        overwritten entries), or null
  T out_scale;                // grad_u / grad_k are written multiplied by this (1/nb when the exponent is 1)
  long long ne, nn, nb;
  int ntiles, ecap, lcap;     // tiles, max elements per tile, max local nodes per tile (shared-memory rows)
  Params<T> p;
  long long mesh_flags = 0;   // FOL_MESH_* bits of include/folax_b200.h (facts about the mesh the host plan established)
};

// integer powers are evaluated by repeated multiplication, like lax.integer_pow does for the
// reference's `T**c` with a Python-int exponent (thermal.py:34)
template <class T>
__device__ __forceinline__ T pow_c(T x, T c) {
  const int ci = (int)c;
  if ((T)ci == c && ci >= 0 && ci <= 15) {   // warp-uniform: c is a kernel parameter
    const T x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;
    T r = (ci & 1) ? x : (T)1;
    r = (ci & 2) ? r * x2 : r;
    r = (ci & 4) ? r * x4 : r;
    r = (ci & 8) ? r * x8 : r;
    return r;
  }
  return (T)pow((double)x, (double)c);
}

// the element energy is a sum of Gauss-point densities (a true potential) rather than u . re
__host__ __device__ constexpr bool point_energy(int phys) { return finite_strain(phys) || implicit_scalar(phys); }
// geometry-cache rows per Gauss point: grad N (A*D), w detJ [, the nodal heterogeneity k0 at the point]
__host__ __device__ constexpr int geom_width(int phys, int elem) {
  return elem_nnode(elem) * elem_dim(elem) + 1 + (phys == TTHERMAL ? 1 : 0);
}

// TRANSPOSED: the gradient convention of transient_thermal.py:57-58 / phase_field.py:47-48 (elements.cuh);
// AUX: one more row, N . aux (aux = nodal heterogeneity k0 of the transient thermal loss, mesh-resident)
template <class T, int ELEM, int ORDER, bool TRANSPOSED, bool AUX>
__global__ void geometry_cache_kernel(const T* __restrict__ xyz, const int32_t* __restrict__ conn, long long ne,
                                      const T* __restrict__ aux, T* __restrict__ geom) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM), W = A * D + 1 + (AUX ? 1 : 0);
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int g = blockIdx.y;
  if (e >= ne) return;
  T X[A * 3];
  T ax = (T)0;
  double xi[3], w;
  gauss_point<ELEM, ORDER>(g, xi, w);
  T N[A], dN[A][D], gN[A][D];
  shape_functions<ELEM, T>(xi, N, dN);
#pragma unroll
  for (int a = 0; a < A; ++a) {
    const long long n = conn[e * A + a];
#pragma unroll
    for (int k = 0; k < 3; ++k) X[a * 3 + k] = xyz[n * 3 + k];
    if constexpr (AUX) ax += N[a] * aux[n];
  }
  const T det = global_gradients<ELEM, T, TRANSPOSED>(X, dN, gN);
  T* out = geom + (long long)g * W * ne + e;
#pragma unroll
  for (int a = 0; a < A; ++a)
#pragma unroll
    for (int k = 0; k < D; ++k) out[(long long)(a * D + k) * ne] = gN[a][k];
  out[(long long)(A * D) * ne] = (T)w * det;
  if constexpr (AUX) out[(long long)(A * D + 1) * ne] = ax;
}

// scratch row count per element: re (ND) [+ dK (A)] [+ psi (1)]
__host__ __device__ constexpr int energy_kw(int phys, int elem) {
  return elem_nnode(elem) * phys_dpn(phys, elem) + (phys == MECH ? 0 : elem_nnode(elem)) + (point_energy(phys) ? 1 : 0);
}

// Element vectors of element e for S samples starting at sample b0 (samples past nb are clamped):
// re[s][nd], dK[s][a] (thermal / neo-hooke), en[s] (neo-hooke strain energy).
// element gather (fe_loss.py:155-164) from the tile's staged nodal rows: st = [C = DPN + 1][lcap] of one sample
template <class T, int A, int DPN>
__device__ __forceinline__ void gather_staged(const T* st, int lcap, const int (&ln)[A], T (&ue)[1][A * DPN],
                                              T (&de)[1][A]) {
#pragma unroll
  for (int a = 0; a < A; ++a) {
#pragma unroll
    for (int k = 0; k < DPN; ++k) ue[0][a * DPN + k] = st[k * lcap + ln[a]];
    de[0][a] = st[DPN * lcap + ln[a]];
  }
}

template <class T>
__device__ __forceinline__ void cp_async_elem(T* sdst, const T* gsrc) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(sdst);
  if constexpr (sizeof(T) == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(saddr), "l"(gsrc) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(saddr), "l"(gsrc) : "memory");
}

// geometry factors of one element fit in registers for the small elements / rules
__host__ __device__ constexpr bool geom_in_regs(int elem, int order, int phys = MECH) {
  return elem_ngauss(elem, order) * geom_width(phys, elem) <= 40;
}

template <class T, int ELEM, int ORDER, int PHYS, int S, bool REGS>
__device__ __forceinline__ void element_vectors(const EnergyArgs<T>& args, long long e, const int (&nodes)[elem_nnode(ELEM)],
                                                const T* greg,
                                                const T (&ue)[S][elem_nnode(ELEM) * phys_dpn(PHYS, ELEM)],
                                                const T (&de)[S][elem_nnode(ELEM)],
                                                T (&re)[S][elem_nnode(ELEM) * phys_dpn(PHYS, ELEM)],
                                                T (&dK)[S][elem_nnode(ELEM)], T (&en)[S]) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM), DPN = phys_dpn(PHYS, ELEM), ND = A * DPN;
  constexpr int NGP = elem_ngauss(ELEM, ORDER), V = voigt_size(D), W = geom_width(PHYS, ELEM);
  const Params<T>& P = args.p;
  (void)nodes;

#pragma unroll
  for (int s = 0; s < S; ++s) {
    en[s] = (T)0;
#pragma unroll
    for (int b = 0; b < A; ++b) {
      dK[s][b] = (T)0;
#pragma unroll
      for (int k = 0; k < DPN; ++k) re[s][b * DPN + k] = (T)0;
    }
  }
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

  // small rules are unrolled so the reference shape values N_a(xi_g) fold to constants
  const T* gm = args.geom + e;
  const long long ne = args.ne;
#pragma unroll(NGP <= 4 ? NGP : 1)
  for (int g = 0; g < NGP; ++g) {
    T gN[A][D], wd, kq = (T)0;   // kq: heterogeneity k0 at the point (transient thermal)
    if constexpr (REGS) {
#pragma unroll
      for (int b = 0; b < A; ++b)
#pragma unroll
        for (int k = 0; k < D; ++k) gN[b][k] = greg[g * W + b * D + k];
      wd = greg[g * W + A * D];
      if constexpr (PHYS == TTHERMAL) kq = greg[g * W + A * D + 1];
    } else {
#pragma unroll
      for (int b = 0; b < A; ++b)
#pragma unroll
        for (int k = 0; k < D; ++k) {
          gN[b][k] = __ldg(gm);
          gm += ne;
        }
      wd = __ldg(gm);
      gm += ne;
      if constexpr (PHYS == TTHERMAL) {
        kq = __ldg(gm);
        gm += ne;
      }
    }
    (void)kq;
    double xi[3], w;
    gauss_point<ELEM, ORDER>(g, xi, w);
    T N[A], dN[A][D];
    shape_functions<ELEM, T>(xi, N, dN);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      T eg = (T)0;
#pragma unroll
      for (int b = 0; b < A; ++b) eg += N[b] * de[s][b];
      if constexpr (PHYS == THERMAL) {
        // kappa_g = (N.K)(1 + beta (N.T)^c); re_a += w detJ kappa_g grad N_a . grad T
        // dE/dK_a += N_a (1 + beta T_g^c) |grad T|^2 w detJ        (thermal.py:28-49)
        T tg = (T)0, gT[D];
#pragma unroll
        for (int k = 0; k < D; ++k) gT[k] = (T)0;
#pragma unroll
        for (int b = 0; b < A; ++b) {
          tg += N[b] * ue[s][b];
#pragma unroll
          for (int k = 0; k < D; ++k) gT[k] += gN[b][k] * ue[s][b];
        }
        const T beta = P.v[5];
        const T nl = (T)1 + ((beta != (T)0) ? beta * pow_c<T>(tg, P.v[6]) : (T)0);
        T g2 = (T)0;
#pragma unroll
        for (int k = 0; k < D; ++k) g2 += gT[k] * gT[k];
        const T cf = wd * eg * nl, ck = wd * nl * g2;
#pragma unroll
        for (int b = 0; b < A; ++b) {
          T flux = (T)0;
#pragma unroll
          for (int k = 0; k < D; ++k) flux += gN[b][k] * gT[k];
          re[s][b] += cf * flux;
          dK[s][b] += ck * N[b];
        }
      } else if constexpr (implicit_scalar(PHYS)) {
        // implicit-Euler scalar potentials (transient_thermal.py:42-73, phase_field.py:38-70), fully differentiated:
        // ue = next field, de = current field; re <- dE/d(next), dK <- dE/d(current), en <- E
        T fn = (T)0, gf[D];
#pragma unroll
        for (int k = 0; k < D; ++k) gf[k] = (T)0;
#pragma unroll
        for (int b = 0; b < A; ++b) {
          fn += N[b] * ue[s][b];
#pragma unroll
          for (int k = 0; k < D; ++k) gf[k] += gN[b][k] * ue[s][b];
        }
        T g2 = (T)0;
#pragma unroll
        for (int k = 0; k < D; ++k) g2 += gf[k] * gf[k];
        const T dt = P.v[10], df = fn - eg;
        T cflux, cn, rate;
        if constexpr (PHYS == TTHERMAL) {
          const T beta = P.v[5], cexp = P.v[6];
          const T Kg = kq * ((T)1 + ((beta != (T)0) ? beta * pow_ci<T>(fn, cexp) : (T)0));
          const T dk = (beta != (T)0) ? kq * beta * cexp * pow_ci<T>(fn, cexp - (T)1) : (T)0;
          rate = P.v[8] * P.v[9] / dt * wd * df;
          cflux = Kg * wd;
          cn = (T)0.5 * dk * wd * g2 + rate;
          en[s] += (T)0.5 * Kg * wd * g2 + (T)0.5 * rate * df;
        } else {
          const T ie2 = (T)1 / (P.v[11] * P.v[11]), f2 = fn * fn - (T)1;
          rate = wd / dt * df;
          cflux = wd;
          cn = wd * ie2 * f2 * fn + rate;
          en[s] += (T)0.5 * wd * g2 + wd * ie2 * (T)0.25 * f2 * f2 + (T)0.5 * rate * df;
        }
#pragma unroll
        for (int b = 0; b < A; ++b) {
          T flux = (T)0;
#pragma unroll
          for (int k = 0; k < D; ++k) flux += gN[b][k] * gf[k];
          re[s][b] += cflux * flux + cn * N[b];
          dK[s][b] -= rate * N[b];
        }
      } else if constexpr (PHYS == MECH) {
        // H = grad u; sigma = E_g (lam tr(H) I + mu (H + H^T)); re_a += w detJ sigma grad N_a - Fe_a
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
        T sig[D][D];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) sig[i][j] = cf * (mu * (H[i][j] + H[j][i]) + (i == j ? lam * tr : (T)0));
#pragma unroll
        for (int b = 0; b < A; ++b)
#pragma unroll
          for (int i = 0; i < D; ++i) {
            T acc = -P.v[2 + i] * wd * N[b];
#pragma unroll
            for (int j = 0; j < D; ++j) acc += sig[i][j] * gN[b][j];
            re[s][b * D + i] += acc;
          }
      } else {  // NEOHOOKE: psi and S are linear in E_g -> evaluate at unit modulus and scale
        const T nu = P.v[1];
        const T k1 = (T)1 / ((T)3 * ((T)1 - (T)2 * nu)), mu1 = (T)1 / ((T)2 * ((T)1 + nu));
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
        const T psi1 = (PHYS == NEOHOOKE) ? neo_hooke_point<T, D>(F, k1, mu1, Sv, Cv)
                                          : st_venant_point<T, D>(F, nu / (((T)1 + nu) * ((T)1 - (T)2 * nu)), mu1, Sv, Cv);
        en[s] += wd * eg * psi1;
#pragma unroll
        for (int b = 0; b < A; ++b) {
          T Ba[V][D];
          neo_hooke_B<T, D>(&F[0][0], gN[b], Ba);
#pragma unroll
          for (int c = 0; c < D; ++c) {
            T acc = (T)0;
#pragma unroll
            for (int v = 0; v < V; ++v) acc += Ba[v][c] * Sv[v];
            re[s][b * D + c] += wd * eg * acc;
          }
          dK[s][b] += wd * N[b] * psi1;
        }
      }
    }
  }
}

template <class T, int ELEM, int ORDER, int PHYS, int S, int BLOCK>
__global__ void __launch_bounds__(BLOCK, 2)
energy_tile_kernel(const EnergyArgs<T> args) {
  // S = samples per CTA pass.  Shared memory: sv [S][KW][ecap] element vectors of the pass,
  // stage [2][S][C][lcap] nodal rows (dofs + control) of the tile's nodes, double-buffered: the rows
  // of pass p+1 arrive by cp.async while pass p computes, so no thread ever waits on a gather.
  constexpr int A = elem_nnode(ELEM), DPN = phys_dpn(PHYS, ELEM), ND = A * DPN, KW = energy_kw(PHYS, ELEM);
  constexpr int C = DPN + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ecap = args.ecap, lcap = args.lcap;
  T* sv = reinterpret_cast<T*>(smem_raw);
  T* stage = sv + (size_t)S * KW * ecap;
  __shared__ T red[S][BLOCK / 32];
  const int t = blockIdx.x;
  const int e_beg = __ldg(args.tile_elem_ptr + t), n_el = __ldg(args.tile_elem_ptr + t + 1) - e_beg;
  const int n_beg = __ldg(args.tile_node_ptr + t), n_nd = __ldg(args.tile_node_ptr + t + 1) - n_beg;
  const int l_beg = __ldg(args.tile_lnode_ptr + t), n_ln = __ldg(args.tile_lnode_ptr + t + 1) - l_beg;
  const long long ndof = args.nn * DPN;

  // phase B role: node threadIdx.x of the tile (= local node threadIdx.x) and its adjacency range
  const bool has_node = (int)threadIdx.x < n_nd;
  const int n = has_node ? __ldg(args.tile_nodes + n_beg + threadIdx.x) : 0;
  const int a_beg = has_node ? __ldg(args.adj_ptr + n) : 0, a_end = has_node ? __ldg(args.adj_ptr + n + 1) : 0;

  // staging role: local node l = threadIdx.x (+ BLOCK ...) copies its S*C values per pass
  auto stage_pass = [&](int buf, long long b0) {
    T* dst0 = stage + (size_t)buf * S * C * lcap;
    for (int l = threadIdx.x; l < n_ln; l += BLOCK) {
      const long long gn = __ldg(args.tile_lnodes + l_beg + l);
#pragma unroll
      for (int sidx = 0; sidx < S; ++sidx) {
        const long long bb = (b0 + sidx < args.nb) ? b0 + sidx : args.nb - 1;
        T* dst = dst0 + (sidx * C) * lcap + l;
#pragma unroll
        for (int k = 0; k < DPN; ++k) cp_async_elem<T>(dst + k * lcap, args.u + bb * ndof + gn * DPN + k);
        cp_async_elem<T>(dst + DPN * lcap, args.ctrl + bb * args.nn + gn);
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };

  // fast path: one thread per tile element; geometry factors and local node ids live in registers
  constexpr bool REGS = geom_in_regs(ELEM, ORDER, PHYS);
  constexpr int GW = REGS ? elem_ngauss(ELEM, ORDER) * geom_width(PHYS, ELEM) : 1;
  const bool fast = REGS && n_el <= BLOCK;
  T greg[GW];
  int my_ln[A];
  long long my_el = -1;
  if constexpr (REGS) {
    if (fast && (int)threadIdx.x < n_el) {
      my_el = __ldg(args.tile_elems + e_beg + threadIdx.x);
#pragma unroll
      for (int b = 0; b < A; ++b) my_ln[b] = __ldg(args.tile_conn + (long long)(e_beg + threadIdx.x) * A + b);
#pragma unroll
      for (int k = 0; k < GW; ++k) greg[k] = __ldg(args.geom + (long long)k * args.ne + my_el);
    }
  }

  auto store_vectors = [&](int sidx, int j, const T (&re)[1][ND], const T (&dK)[1][A], const T (&en1)[1]) {
    T* out = sv + (sidx * KW) * ecap + j;
#pragma unroll
    for (int k = 0; k < ND; ++k) { *out = re[0][k]; out += ecap; }
    if constexpr (PHYS != MECH) {
#pragma unroll
      for (int b = 0; b < A; ++b) { *out = dK[0][b]; out += ecap; }
    }
    if constexpr (point_energy(PHYS)) *out = en1[0];
  };

  long long b0 = (long long)blockIdx.y * S;
  const long long bstep = (long long)gridDim.y * S;
  int buf = 0;
  if (b0 < args.nb) stage_pass(0, b0);
  for (; b0 < args.nb; b0 += bstep, buf ^= 1) {
    const int ns = (args.nb - b0 < S) ? (int)(args.nb - b0) : S;
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    if (args.dir_values) {   // Dirichlet overwrite of the rows this thread staged (its own copies have landed)
      T* dst0 = stage + (size_t)buf * S * C * lcap;
      for (int l = threadIdx.x; l < n_ln; l += BLOCK) {
        const long long gn = __ldg(args.tile_lnodes + l_beg + l);
#pragma unroll
        for (int k = 0; k < DPN; ++k) {
          const T dv = __ldg(args.dir_values + gn * DPN + k);
          if (dv == dv) {
#pragma unroll
            for (int sidx = 0; sidx < S; ++sidx) dst0[(sidx * C + k) * lcap + l] = dv;
          }
        }
      }
    }
    __syncthreads();   // this pass's rows are visible; everyone is done with the previous pass's sv / rows
    if (b0 + bstep < args.nb) stage_pass(buf ^ 1, b0 + bstep);
    const T* st0 = stage + (size_t)buf * S * C * lcap;

    // ---- phase A: every element of the tile, once per sample
    if constexpr (REGS) {
      if (fast && my_el >= 0) {
        for (int sidx = 0; sidx < ns; ++sidx) {
          T ue[1][ND], de[1][A], re[1][ND], dK[1][A], en1[1];
          gather_staged<T, A, DPN>(st0 + (sidx * C) * lcap, lcap, my_ln, ue, de);
          element_vectors<T, ELEM, ORDER, PHYS, 1, true>(args, my_el, my_ln, greg, ue, de, re, dK, en1);
          store_vectors(sidx, threadIdx.x, re, dK, en1);
        }
      }
    }
    if (!fast) {
      const int items = n_el * ns;
      int j = threadIdx.x, sidx = 0;
      while (j >= n_el && sidx < ns) { j -= n_el; ++sidx; }
      for (int it = threadIdx.x; it < items; it += BLOCK) {
        const long long e = __ldg(args.tile_elems + e_beg + j);
        int ln[A];
#pragma unroll
        for (int b = 0; b < A; ++b) ln[b] = __ldg(args.tile_conn + (long long)(e_beg + j) * A + b);
        T ue[1][ND], de[1][A], re[1][ND], dK[1][A], en1[1];
        gather_staged<T, A, DPN>(st0 + (sidx * C) * lcap, lcap, ln, ue, de);
        element_vectors<T, ELEM, ORDER, PHYS, 1, false>(args, e, ln, nullptr, ue, de, re, dK, en1);
        store_vectors(sidx, j, re, dK, en1);
        j += BLOCK;
        while (j >= n_el) { j -= n_el; ++sidx; }
      }
    }
    __syncthreads();
    // ---- phase B: fixed-order adjacency sums for this tile's nodes
    T en[S];
#pragma unroll
    for (int s = 0; s < S; ++s) en[s] = (T)0;
    if (has_node) {
      T R[S][DPN], dk[S];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        dk[s] = (T)0;
#pragma unroll
        for (int k = 0; k < DPN; ++k) R[s][k] = (T)0;
      }
      for (int it = a_beg; it < a_end; ++it) {
        const int ja = __ldg(args.adj_local + it);
        const int jl = ja / A, a = ja - jl * A;
        const T* in = sv + jl + (a * DPN) * ecap;
        const T* ink = sv + jl + (ND + a) * ecap;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          if (s < ns) {
#pragma unroll
            for (int k = 0; k < DPN; ++k) R[s][k] += in[(s * KW + k) * ecap];
            if constexpr (PHYS != MECH) dk[s] += ink[(s * KW) * ecap];
            // each element's strain energy is counted once: at its local node 0, by the tile owning it
            if constexpr (point_energy(PHYS)) en[s] += (a == 0) ? sv[jl + (s * KW + ND + A) * ecap] : (T)0;
          }
        }
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        if (s < ns) {
          const long long bb = b0 + s;
#pragma unroll
          for (int k = 0; k < DPN; ++k) {
            // E_b = u_b . R_b (mechanical.py:116-117, thermal.py:45-49); u from the staged rows
            if constexpr (!point_energy(PHYS)) en[s] += st0[(s * C + k) * lcap + threadIdx.x] * R[s][k];
            const bool cut = args.dir_flag && args.dir_flag[(long long)n * DPN + k];
            args.grad_u[bb * ndof + (long long)n * DPN + k] = cut ? (T)0 : args.out_scale * R[s][k];
          }
          if constexpr (PHYS != MECH) {
            if (args.grad_k) args.grad_k[bb * args.nn + n] = args.out_scale * dk[s];
          }
        }
      }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      T v = en[s];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) red[s][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if ((int)threadIdx.x < ns) {
      T v = (T)0;
#pragma unroll
      for (int w = 0; w < BLOCK / 32; ++w) v += red[threadIdx.x][w];
      args.partial[(b0 + threadIdx.x) * args.ntiles + t] = v;
    }
  }
}

// E_b = sum of block partials (fixed order); optionally L = mean E_b^p, (min, max, mean), scale_b
template <class T>
__global__ void energy_sum_kernel(const T* __restrict__ partial, long long nb, int nblocks, T* __restrict__ energy) {
  // one warp per sample: lane-strided partial sums, then a fixed xor tree
  const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= nb) return;
  T acc = (T)0;
  for (int k = threadIdx.x & 31; k < nblocks; k += 32) acc += partial[b * nblocks + k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) energy[b] = acc;
}

template <class T>
__global__ void loss_reduce_kernel(long long nb, double exponent, const T* __restrict__ energy, T* __restrict__ out4,
                                   T* __restrict__ scale) {
  __shared__ double s_sum[256], s_min[256], s_max[256];
  double lsum = 0.0, lmin = INFINITY, lmax = -INFINITY;
  for (long long b = threadIdx.x; b < nb; b += blockDim.x) {
    const double E = (double)energy[b];
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

// backward of the batched loss: grad_u[b,:] *= sc_b (zero at Dirichlet dofs), grad_k[b,:] *= sc_b with
// sc_b = upstream * (*upstream_dev) * (prescaled ? 1 : scale[b]).  When the forward pass already wrote the
// gradients with their final scale (exponent 1) and the upstream cotangent is 1, there is nothing to do.
template <class T>
__global__ void scale_grads_kernel(long long nb, long long ndof, long long nn, const T* __restrict__ scale,
                                   T upstream, const T* __restrict__ upstream_dev, int prescaled,
                                   const uint8_t* __restrict__ dir, T* __restrict__ grad_u,
                                   T* __restrict__ grad_k) {
  const T up = upstream * (upstream_dev ? *upstream_dev : (T)1);
  if (prescaled && up == (T)1) return;
  const long long m = ndof > nn ? ndof : nn;
  const long long chunks = (m + blockDim.x - 1) / blockDim.x;
  for (long long w = blockIdx.x; w < chunks * nb; w += gridDim.x) {
    const long long b = w / chunks, i = (w - b * chunks) * blockDim.x + threadIdx.x;
    const T sc = up * (prescaled ? (T)1 : scale[b]);
    if (i < ndof) grad_u[b * ndof + i] = dir[i] ? (T)0 : sc * grad_u[b * ndof + i];
    if (grad_k && i < nn) grad_k[b * nn + i] = sc * grad_k[b * nn + i];
  }
}

}  // namespace fol
