// Pipelined form of the fused batched loss + VJP kernel (energy.cuh) for tile plans whose per-tile element
// list fits one thread per element (plan["ecap"] <= BLOCK) and whose geometry factors fit in registers.
//
// Same math, same fixed summation order and same results as energy_tile_kernel; what changes is the schedule:
//   * one barrier per pass instead of three: the element vectors are double-buffered in shared memory, so a
//     thread finishes the node sums of pass p-1 (phase B) and goes straight on to the element evaluations of
//     pass p (phase A); warps drift apart inside the interval instead of idling at barriers;
//   * the node's adjacency (shared-memory offsets) is held in registers, every shared-memory stride is a
//     compile-time constant, the conductivity law 1 + beta T^c is a template parameter (NL);
//   * per-warp energy shares go straight to the `partial` scratch (summed in fixed order by
//     energy_sum_kernel), so there is no block reduction.
#pragma once
#include "energy.cuh"

namespace fol {

// NL: 0 -> beta == 0; 1..4 -> integer exponent c; -1 -> generic pow (thermal.py:34)
template <class T, int NL>
__device__ __forceinline__ T conductivity_factor(T tg, T beta, T c) {
  if constexpr (NL == 0) return (T)1;
  else if constexpr (NL == 1) return (T)1 + beta * tg;
  else if constexpr (NL == 2) return (T)1 + beta * (tg * tg);
  else if constexpr (NL == 3) return (T)1 + beta * (tg * (tg * tg));
  else if constexpr (NL == 4) { const T t2 = tg * tg; return (T)1 + beta * (t2 * t2); }
  else return (T)1 + ((beta != (T)0) ? beta * pow_c<T>(tg, c) : (T)0);
}

// thermal element vectors of ONE sample with the geometry factors in registers (thermal.py:28-49)
template <class T, int ELEM, int ORDER, int NL>
__device__ __forceinline__ void thermal_vectors(const T* greg, const T (&Te)[elem_nnode(ELEM)],
                                                const T (&Ke)[elem_nnode(ELEM)], T beta, T cexp,
                                                T (&re)[elem_nnode(ELEM)], T (&dK)[elem_nnode(ELEM)]) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM), NGP = elem_ngauss(ELEM, ORDER), W = A * D + 1;
#pragma unroll
  for (int b = 0; b < A; ++b) re[b] = dK[b] = (T)0;
#pragma unroll
  for (int g = 0; g < NGP; ++g) {
    double xi[3], w;
    gauss_point<ELEM, ORDER>(g, xi, w);
    T N[A], dN[A][D];
    shape_functions<ELEM, T>(xi, N, dN);
    const T* gN = greg + g * W;
    const T wd = gN[A * D];
    T eg = (T)0, tg = (T)0, gT[D];
#pragma unroll
    for (int k = 0; k < D; ++k) gT[k] = (T)0;
#pragma unroll
    for (int b = 0; b < A; ++b) {
      eg += N[b] * Ke[b];
      tg += N[b] * Te[b];
#pragma unroll
      for (int k = 0; k < D; ++k) gT[k] += gN[b * D + k] * Te[b];
    }
    const T nl = conductivity_factor<T, NL>(tg, beta, cexp);
    T g2 = (T)0;
#pragma unroll
    for (int k = 0; k < D; ++k) g2 += gT[k] * gT[k];
    const T cf = wd * eg * nl, ck = wd * nl * g2;
#pragma unroll
    for (int b = 0; b < A; ++b) {
      T flux = (T)0;
#pragma unroll
      for (int k = 0; k < D; ++k) flux += gN[b * D + k] * gT[k];
      re[b] += cf * flux;
      dK[b] += ck * N[b];
    }
  }
}

template <class T, int ELEM, int ORDER, int PHYS, int NL, int S, int BLOCK, int MINB, int LCAP>
__global__ void __launch_bounds__(BLOCK, MINB) energy_tile2_kernel(const EnergyArgs<T> args) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM), DPN = phys_dpn(PHYS, ELEM), ND = A * DPN;
  constexpr int KW = energy_kw(PHYS, ELEM), C = DPN + 1, NW = BLOCK / 32;
  constexpr int GW = elem_ngauss(ELEM, ORDER) * geom_width(PHYS, ELEM);
  constexpr int MAXADJ = 8;                 // adjacency entries held in registers; longer lists continue from global
  constexpr int SVB = S * KW * BLOCK;       // one element-vector buffer
  static_assert(geom_in_regs(ELEM, ORDER, PHYS), "energy_tile2_kernel keeps the geometry factors in registers");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int lcap = LCAP;                // row stride of the staged nodal rows (>= plan lcap, checked by the host)
  T* sv = reinterpret_cast<T*>(smem_raw);   // [2][S][KW][BLOCK]
  T* stage = sv + 2 * SVB;                  // [2][S][C][LCAP]
  const int t = blockIdx.x, tid = threadIdx.x;
  const int e_beg = __ldg(args.tile_elem_ptr + t), n_el = __ldg(args.tile_elem_ptr + t + 1) - e_beg;
  const int n_beg = __ldg(args.tile_node_ptr + t), n_nd = __ldg(args.tile_node_ptr + t + 1) - n_beg;
  const int l_beg = __ldg(args.tile_lnode_ptr + t), n_ln = __ldg(args.tile_lnode_ptr + t + 1) - l_beg;
  const long long ndof = args.nn * DPN;
  const int npart = args.ntiles * NW;

  // ---- node role (phase B): the tile's nodes are dealt to the warps in equal contiguous chunks, so every warp
  //      carries the same share of phase B (no warp idles at the barrier); lnode = the node's tile-local index
  const int per_warp = (n_nd + NW - 1) / NW;
  const int lnode = (tid >> 5) * per_warp + (tid & 31);
  const bool has_node = (tid & 31) < per_warp && lnode < n_nd;
  const int n = has_node ? __ldg(args.tile_nodes + n_beg + lnode) : 0;
  const int a_beg = has_node ? __ldg(args.adj_ptr + n) : 0, a_end = has_node ? __ldg(args.adj_ptr + n + 1) : 0;
  const int cnt = a_end - a_beg;
  // off: row of the entry's first dof; the dK row is off + ND * BLOCK for one dof per node, else kept in offk
  constexpr int NK = (DPN == 1 || PHYS == MECH) ? 1 : MAXADJ;
  int off[MAXADJ], offk[NK];
  unsigned first_mask = 0;                  // entries whose local node is 0: they carry the element's strain energy
#pragma unroll
  for (int i = 0; i < MAXADJ; ++i) {
    off[i] = 0;
    if (i < NK) offk[i] = 0;
    if (i < cnt) {
      const int ja = __ldg(args.adj_local + a_beg + i);
      const int jl = ja / A, a = ja - jl * A;
      off[i] = jl + a * DPN * BLOCK;
      if constexpr (NK == MAXADJ) offk[i] = jl + (ND + a) * BLOCK;
      first_mask |= (a == 0 ? 1u : 0u) << i;
    }
  }

  // ---- element role (phase A): element tid of the tile; geometry factors and local node ids in registers
  const bool has_el = tid < n_el;
  T greg[GW];
  int my_ln[A];
  long long my_el = 0;
  if (has_el) {
    my_el = __ldg(args.tile_elems + e_beg + tid);
#pragma unroll
    for (int b = 0; b < A; ++b) my_ln[b] = __ldg(args.tile_conn + (long long)(e_beg + tid) * A + b);
#pragma unroll
    for (int k = 0; k < GW; ++k) greg[k] = __ldg(args.geom + (long long)k * args.ne + my_el);
  } else {
#pragma unroll
    for (int b = 0; b < A; ++b) my_ln[b] = 0;
#pragma unroll
    for (int k = 0; k < GW; ++k) greg[k] = (T)0;
  }

  // staging role: this thread copies the rows of local nodes tid and tid + BLOCK (LCAP <= 2 BLOCK); their global
  // ids stay in registers so a pass issues its copies without waiting on an index load
  static_assert(LCAP <= 2 * BLOCK, "two staged local nodes per thread");
  int gnode[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) gnode[j] = (tid + j * BLOCK < n_ln) ? __ldg(args.tile_lnodes + l_beg + tid + j * BLOCK) : -1;
  // Dirichlet values (NaN = free) of the rows this thread stages
  T dval[2][DPN];
  bool any_dir = false;
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int k = 0; k < DPN; ++k) {
      dval[j][k] = (args.dir_values && gnode[j] >= 0) ? __ldg(args.dir_values + (long long)gnode[j] * DPN + k)
                                                      : (T)NAN;
      any_dir |= (dval[j][k] == dval[j][k]);
    }
  // Dirichlet overwrite (fe_loss.py:91-92, 255) of the rows this thread staged, once its own copies have landed
  auto patch_pass = [&](int buf) {
    if (!any_dir) return;
    T* dst0 = stage + (size_t)buf * S * C * lcap + tid;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int k = 0; k < DPN; ++k) {
        if (dval[j][k] == dval[j][k]) {
#pragma unroll
          for (int sidx = 0; sidx < S; ++sidx) dst0[(sidx * C + k) * lcap + j * BLOCK] = dval[j][k];
        }
      }
  };
  auto stage_pass = [&](int buf, long long b0) {
    T* dst0 = stage + (size_t)buf * S * C * lcap + tid;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (gnode[j] >= 0) {
#pragma unroll
        for (int sidx = 0; sidx < S; ++sidx) {
          const long long bb = (b0 + sidx < args.nb) ? b0 + sidx : args.nb - 1;
          T* dst = dst0 + (sidx * C) * lcap + j * BLOCK;
#pragma unroll
          for (int k = 0; k < DPN; ++k) cp_async_elem<T>(dst + k * lcap, args.u + bb * ndof + (long long)gnode[j] * DPN + k);
          cp_async_elem<T>(dst + DPN * lcap, args.ctrl + bb * args.nn + gnode[j]);
        }
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };

  T ukeep[S][DPN];                          // this thread's node dofs of the pass whose phase B is pending
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int k = 0; k < DPN; ++k) ukeep[s][k] = (T)0;

  unsigned cut_mask = 0;                    // dofs of this node whose cotangent is cut (Dirichlet)
  if (has_node && args.dir_flag) {
#pragma unroll
    for (int k = 0; k < DPN; ++k) cut_mask |= (args.dir_flag[(long long)n * DPN + k] ? 1u : 0u) << k;
  }
  T* const gu_node = args.grad_u + (long long)n * DPN;
  T* const gk_node = args.grad_k ? args.grad_k + n : nullptr;

  // phase B of one pass: fixed-order adjacency sums from the element vectors in `svb`
  auto phase_b = [&](const T* svb, long long b0, int ns) {
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
#pragma unroll
      for (int i = 0; i < MAXADJ; ++i) {
        if (i >= cnt) break;
#pragma unroll
        for (int s = 0; s < S; ++s) {
#pragma unroll
          for (int k = 0; k < DPN; ++k) R[s][k] += svb[off[i] + (s * KW + k) * BLOCK];
          if constexpr (PHYS != MECH) {
            if constexpr (NK == MAXADJ) dk[s] += svb[offk[i] + (s * KW) * BLOCK];
            else dk[s] += svb[off[i] + (s * KW + ND) * BLOCK];
          }
          if constexpr (point_energy(PHYS)) {
            // local node 0: off[i] is the element's column itself
            if ((first_mask >> i) & 1u) en[s] += svb[off[i] + (s * KW + ND + A) * BLOCK];
          }
        }
      }
      for (int it = a_beg + MAXADJ; it < a_end; ++it) {
        const int ja = __ldg(args.adj_local + it);
        const int jl = ja / A, a = ja - jl * A;
#pragma unroll
        for (int s = 0; s < S; ++s) {
#pragma unroll
          for (int k = 0; k < DPN; ++k) R[s][k] += svb[jl + (s * KW + a * DPN + k) * BLOCK];
          if constexpr (PHYS != MECH) dk[s] += svb[jl + (s * KW + ND + a) * BLOCK];
          if constexpr (point_energy(PHYS)) en[s] += (a == 0) ? svb[jl + (s * KW + ND + A) * BLOCK] : (T)0;
        }
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        if (s < ns) {
          T* gu = gu_node + (b0 + s) * ndof;
#pragma unroll
          for (int k = 0; k < DPN; ++k) {
            // E_b = u_b . R_b (mechanical.py:116-117, thermal.py:45-49)
            if constexpr (!point_energy(PHYS)) en[s] += ukeep[s][k] * R[s][k];
            gu[k] = ((cut_mask >> k) & 1u) ? (T)0 : args.out_scale * R[s][k];
          }
          if constexpr (PHYS != MECH) {
            if (gk_node) gk_node[(b0 + s) * args.nn] = args.out_scale * dk[s];
          }
        }
      }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      T v = en[s];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((tid & 31) == 0 && s < ns) args.partial[(b0 + s) * npart + t * NW + (tid >> 5)] = v;
    }
  };

  // phase A of one pass: this thread's element, once per sample (samples past nb are clamped copies: computed,
  // never written out -- phase B stores only s < ns)
  auto phase_a = [&](const T* st0, T* out0) {
    if (!has_el) return;
#pragma unroll 1
    for (int sidx = 0; sidx < S; ++sidx) {
      const T* st = st0 + (sidx * C) * lcap;
      T* out = out0 + (sidx * KW) * BLOCK;
      if constexpr (PHYS == THERMAL) {
        T Te[A], Ke[A], re[A], dK[A];
#pragma unroll
        for (int b = 0; b < A; ++b) {
          Te[b] = st[my_ln[b]];
          Ke[b] = st[lcap + my_ln[b]];
        }
        thermal_vectors<T, ELEM, ORDER, NL>(greg, Te, Ke, args.p.v[5], args.p.v[6], re, dK);
#pragma unroll
        for (int b = 0; b < A; ++b) {
          out[b * BLOCK] = re[b];
          out[(A + b) * BLOCK] = dK[b];
        }
      } else {
        T ue[1][ND], de[1][A], re[1][ND], dK[1][A], en1[1];
        gather_staged<T, A, DPN>(st, lcap, my_ln, ue, de);
        element_vectors<T, ELEM, ORDER, PHYS, 1, true>(args, my_el, my_ln, greg, ue, de, re, dK, en1);
#pragma unroll
        for (int k = 0; k < ND; ++k) out[k * BLOCK] = re[0][k];
        if constexpr (PHYS != MECH) {
#pragma unroll
          for (int b = 0; b < A; ++b) out[(ND + b) * BLOCK] = dK[0][b];
        }
        if constexpr (point_energy(PHYS)) out[(ND + A) * BLOCK] = en1[0];
      }
    }
  };

  long long b0 = (long long)blockIdx.y * S;
  const long long bstep = (long long)gridDim.y * S;
  if (b0 >= args.nb) return;
  stage_pass(0, b0);
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  patch_pass(0);
  __syncthreads();
  // Between two barriers a warp owes phase B of the previous pass (latency-bound: shared-memory sums, shuffles,
  // stores) and phase A of this pass (FP64-bound).  Warps sharing a scheduler take them in opposite order, so one
  // warp's FP64 stream covers the other's latencies instead of all warps idling / contending together.
  const int wid = tid >> 5;
  const bool a_first = (((wid >> 2) ^ wid ^ (int)blockIdx.x ^ (int)blockIdx.y) & 1) != 0;
  int buf = 0;
  long long bprev = -1;
  for (; b0 < args.nb; b0 += bstep, buf ^= 1) {
    if (b0 + bstep < args.nb) stage_pass(buf ^ 1, b0 + bstep);
    const T* st0 = stage + (size_t)buf * S * C * lcap;
    const int nsprev = (args.nb - bprev < S) ? (int)(args.nb - bprev) : S;
    if (a_first) {
      phase_a(st0, sv + buf * SVB + tid);
      if (bprev >= 0) phase_b(sv + (buf ^ 1) * SVB, bprev, nsprev);
    } else {
      if (bprev >= 0) phase_b(sv + (buf ^ 1) * SVB, bprev, nsprev);
      phase_a(st0, sv + buf * SVB + tid);
    }
    if (has_node) {
#pragma unroll
      for (int s = 0; s < S; ++s)
#pragma unroll
        for (int k = 0; k < DPN; ++k) ukeep[s][k] = st0[(s * C + k) * lcap + lnode];
    }
    bprev = b0;
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    patch_pass(buf ^ 1);
    __syncthreads();   // sv[buf] complete and the next pass's rows visible; sv[buf^1] / stage[buf] free again
  }
  phase_b(sv + (buf ^ 1) * SVB, bprev, (args.nb - bprev < S) ? (int)(args.nb - bprev) : S);
}

}  // namespace fol
