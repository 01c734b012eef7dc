// Sample-vectorised form of the pipelined batched loss + VJP kernel for the north-star FOL configuration
// (BASELINE.json configs[2]): ThermalLoss on Quad4 with the 2x2 rule (thermal.py:28-49, fe_loss.py:250-262 and the
// JAX-AD gradient of it).
//
// Same tile plan, same fixed summation order, same results as energy_tile2_kernel (energy2.cuh); what changes is the
// shared-memory layout: the S = 16 / sizeof(T) samples of a pass (2 doubles or 4 floats) sit NEXT to each other, so
//   * phase A (thread = tile element) reads the 4 + 4 nodal values of ALL its samples with 8 LDS.128 and stores the
//     8 element-vector rows with 8 STS.128 -- 1/S of the shared-memory instructions and address arithmetic of the
//     sample-major layout -- and evaluates the S samples as independent instruction streams (ILP S on the FP chains);
//   * phase B (thread = tile node) takes one LDS.128 per adjacency entry and row for all samples.
// energy_tile2_kernel measured 342 issued instructions per element-sample with 161 FP64 among them, FP64 pipe 51 %
// busy, stalls on shared-memory latency and fixed-latency dependencies (profiles/r1/energy_tile2_ncu_summary.txt);
// the float32 variant was issue-bound at 0.16 of HBM.  Double-buffered element vectors, ONE barrier per pass, the
// two phases of neighbouring warps in opposite order, cp.async staging with the Dirichlet overwrite fused -- as there.
#pragma once
#include "energy2.cuh"

namespace fol {

template <class T> struct SampleVec;
template <> struct SampleVec<double> { using type = double2; };
template <> struct SampleVec<float> { using type = float4; };

__device__ __forceinline__ void unpack(const double2& v, double (&o)[2]) { o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ void unpack(const float4& v, float (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ double2 pack(const double (&o)[2]) { return make_double2(o[0], o[1]); }
__device__ __forceinline__ float4 pack(const float (&o)[4]) { return make_float4(o[0], o[1], o[2], o[3]); }

// Thermal element vectors of ONE sample on an AFFINE Quad4 (a parallelogram: J is the same at every Gauss point), from
// J^-1 (row-major, jinv[j*2+k] = d xi_j / d x_k) and w detJ instead of the 36 cached gradient values:
// grad N_b(g) = dN_b(xi_g) . J^-1 with dN_b(xi_g) compile-time constants, so
//   grad T_g = J^-T (sum_b dN_b(g) T_b),   re_b += dN_b(g) . (cf_g J^-1 grad T_g)
// Same sums as thermal_vectors up to rounding (the general kernel reads gradients that were rounded once more).
template <class T, int NL>
__device__ __forceinline__ void thermal_vectors_affine(const T (&jinv)[4], T wd, const T (&Te)[4], const T (&Ke)[4],
                                                       T beta, T cexp, T (&re)[4], T (&dK)[4]) {
#pragma unroll
  for (int b = 0; b < 4; ++b) re[b] = dK[b] = (T)0;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    double xi[3], w;
    gauss_point<QUAD, 2>(g, xi, w);
    T N[4], dN[4][2];
    shape_functions<QUAD, T>(xi, N, dN);
    T eg = (T)0, tg = (T)0, t0 = (T)0, t1 = (T)0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      eg += N[b] * Ke[b];
      tg += N[b] * Te[b];
      t0 += dN[b][0] * Te[b];
      t1 += dN[b][1] * Te[b];
    }
    const T gx = t0 * jinv[0] + t1 * jinv[2], gy = t0 * jinv[1] + t1 * jinv[3];   // grad T = J^-T (dN^T T)
    const T nl = conductivity_factor<T, NL>(tg, beta, cexp);
    const T g2 = gx * gx + gy * gy;
    const T cf = wd * eg * nl, ck = wd * nl * g2;
    const T w0 = cf * (jinv[0] * gx + jinv[1] * gy), w1 = cf * (jinv[2] * gx + jinv[3] * gy);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      re[b] += dN[b][0] * w0 + dN[b][1] * w1;
      dK[b] += ck * N[b];
    }
  }
}

template <class T, int NL, int BLOCK, int MINB, int LCAP, bool AFFINE>
// Register caps follow the PER-SCHEDULER register file (16 K registers for the warps a scheduler hosts): two 192-thread
// CTAs put 3 warps on a scheduler (168 registers), two 256-thread CTAs 4 warps (128 registers, 16 warps / SM).
__global__ void __maxnreg__(BLOCK * MINB <= 384 ? 168 : 128) energy_qt_kernel(const EnergyArgs<T> args) {
  constexpr int S = 16 / (int)sizeof(T);    // samples per pass = one 16-byte vector
  constexpr int A = 4, KW = 8, NW = BLOCK / 32, GW = AFFINE ? 5 : 4 * 9;
  constexpr int MAXADJ = 8;                 // adjacency entries held in registers; longer lists continue from global
  constexpr int SVB = KW * BLOCK;           // one element-vector buffer (in sample vectors)
  using V = typename SampleVec<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V* sv = reinterpret_cast<V*>(smem_raw);   // [2][KW][BLOCK]: rows re_0..3, dK_0..3 of the tile's elements
  V* stage = sv + 2 * SVB;                  // [2][2][LCAP]:   rows T, K of the tile's local nodes
  const int t = blockIdx.x, tid = threadIdx.x;
  const int e_beg = __ldg(args.tile_elem_ptr + t), n_el = __ldg(args.tile_elem_ptr + t + 1) - e_beg;
  const int n_beg = __ldg(args.tile_node_ptr + t), n_nd = __ldg(args.tile_node_ptr + t + 1) - n_beg;
  const int l_beg = __ldg(args.tile_lnode_ptr + t), n_ln = __ldg(args.tile_lnode_ptr + t + 1) - l_beg;
  const long long nn = args.nn;
  const int npart = args.ntiles * NW;

  // ---- node role (phase B): the tile's nodes dealt to the warps in equal contiguous chunks
  const int per_warp = (n_nd + NW - 1) / NW;
  const int lnode = (tid >> 5) * per_warp + (tid & 31);
  const bool has_node = (tid & 31) < per_warp && lnode < n_nd;
  const int n = has_node ? __ldg(args.tile_nodes + n_beg + lnode) : 0;
  const int a_beg = has_node ? __ldg(args.adj_ptr + n) : 0, a_end = has_node ? __ldg(args.adj_ptr + n + 1) : 0;
  const int cnt = a_end - a_beg;
  int cnt_warp = cnt > MAXADJ ? MAXADJ : cnt;   // longest register-held list among the warp's nodes
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt_warp = max(cnt_warp, __shfl_xor_sync(0xffffffffu, cnt_warp, o));
  int off[MAXADJ];                          // re row of the entry; its dK row is 4 * BLOCK further
#pragma unroll
  for (int i = 0; i < MAXADJ; ++i) {
    off[i] = 0;
    if (i < cnt) {
      const int ja = __ldg(args.adj_local + a_beg + i);
      const int jl = ja / A, a = ja - jl * A;
      off[i] = jl + a * BLOCK;
    }
  }

  // ---- element role (phase A): element tid of the tile; geometry factors and local node ids in registers
  const bool has_el = tid < n_el;
  T greg[GW];
  int my_ln[A];
  if (has_el) {
    const long long my_el = __ldg(args.tile_elems + e_beg + tid);
#pragma unroll
    for (int b = 0; b < A; ++b) my_ln[b] = __ldg(args.tile_conn + (long long)(e_beg + tid) * A + b);
    if constexpr (AFFINE) {
      // J^-1 from the cached gradients of nodes 0 and 1 at Gauss point 0: [dN_0; dN_1] J^-1 = [grad N_0; grad N_1]
      double xi[3], w;
      gauss_point<QUAD, 2>(0, xi, w);
      T N[4], dN[4][2];
      shape_functions<QUAD, T>(xi, N, dN);
      const T idet = (T)1 / (dN[0][0] * dN[1][1] - dN[0][1] * dN[1][0]);
      T gn[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) gn[k] = __ldg(args.geom + (long long)k * args.ne + my_el);   // gN_0x, gN_0y, gN_1x, gN_1y
      greg[0] = idet * (dN[1][1] * gn[0] - dN[0][1] * gn[2]);
      greg[1] = idet * (dN[1][1] * gn[1] - dN[0][1] * gn[3]);
      greg[2] = idet * (dN[0][0] * gn[2] - dN[1][0] * gn[0]);
      greg[3] = idet * (dN[0][0] * gn[3] - dN[1][0] * gn[1]);
      greg[4] = __ldg(args.geom + (long long)8 * args.ne + my_el);                               // w detJ
    } else {
#pragma unroll
      for (int k = 0; k < GW; ++k) greg[k] = __ldg(args.geom + (long long)k * args.ne + my_el);
    }
  } else {
#pragma unroll
    for (int b = 0; b < A; ++b) my_ln[b] = 0;
#pragma unroll
    for (int k = 0; k < GW; ++k) greg[k] = (T)0;
  }

  // ---- staging role: this thread copies the rows of local nodes tid and tid + BLOCK (LCAP <= 2 BLOCK)
  static_assert(LCAP <= 2 * BLOCK, "two staged local nodes per thread");
  int gnode[2];
  T dval[2];                                // Dirichlet value (NaN = free) of the rows this thread stages
  bool any_dir = false;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    gnode[j] = (tid + j * BLOCK < n_ln) ? __ldg(args.tile_lnodes + l_beg + tid + j * BLOCK) : -1;
    dval[j] = (args.dir_values && gnode[j] >= 0) ? __ldg(args.dir_values + gnode[j]) : (T)NAN;
    any_dir |= (dval[j] == dval[j]);
  }
  auto stage_pass = [&](int buf, long long b0) {
    T* dst0 = reinterpret_cast<T*>(stage + (size_t)buf * 2 * LCAP);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (gnode[j] >= 0) {
        T* du = dst0 + (size_t)(tid + j * BLOCK) * S;
        T* dk = du + (size_t)LCAP * S;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const long long bb = (b0 + s < args.nb) ? b0 + s : args.nb - 1;
          cp_async_elem<T>(du + s, args.u + bb * nn + gnode[j]);
          cp_async_elem<T>(dk + s, args.ctrl + bb * nn + gnode[j]);
        }
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  // Dirichlet overwrite (fe_loss.py:91-92, 255) of the rows this thread staged, once its own copies have landed
  auto patch_pass = [&](int buf) {
    if (!any_dir) return;
    T* dst0 = reinterpret_cast<T*>(stage + (size_t)buf * 2 * LCAP);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (dval[j] == dval[j]) {
#pragma unroll
        for (int s = 0; s < S; ++s) dst0[(size_t)(tid + j * BLOCK) * S + s] = dval[j];
      }
    }
  };

  bool cut = false;                         // the node's cotangent is cut (Dirichlet)
  if (has_node && args.dir_flag) cut = args.dir_flag[n] != 0;
  T* const gu_node = args.grad_u + n;
  T* const gk_node = args.grad_k ? args.grad_k + n : nullptr;
  T ukeep[S];                               // this thread's node values of the pass whose phase B is pending
#pragma unroll
  for (int s = 0; s < S; ++s) ukeep[s] = (T)0;

  // phase B of one pass: fixed-order adjacency sums from the element vectors in `svb`
  auto phase_b = [&](const V* svb, long long b0, int ns) {
    T en[S];
#pragma unroll
    for (int s = 0; s < S; ++s) en[s] = (T)0;
    if (has_node) {
      T R[S], dk[S];
#pragma unroll
      for (int s = 0; s < S; ++s) R[s] = dk[s] = (T)0;
      // all loads of a group of four entries are issued before the first add (entries past the node's count read
      // row 0 of element 0 -- a valid address -- and are discarded); the trip count is warp-uniform
#pragma unroll
      for (int i0 = 0; i0 < MAXADJ; i0 += 4) {
        if (i0 >= cnt_warp) break;
        V rv[4], kv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          rv[i] = svb[off[i0 + i]];
          kv[i] = svb[off[i0 + i] + 4 * BLOCK];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i0 + i < cnt) {                  // fixed order: entry i0, i0 + 1, ...
            T r[S], k[S];
            unpack(rv[i], r);
            unpack(kv[i], k);
#pragma unroll
            for (int s = 0; s < S; ++s) {
              R[s] += r[s];
              dk[s] += k[s];
            }
          }
        }
      }
      for (int it = a_beg + MAXADJ; it < a_end; ++it) {
        const int ja = __ldg(args.adj_local + it);
        const int jl = ja / A, a = ja - jl * A;
        T r[S], k[S];
        unpack(svb[jl + a * BLOCK], r);
        unpack(svb[jl + (4 + a) * BLOCK], k);
#pragma unroll
        for (int s = 0; s < S; ++s) {
          R[s] += r[s];
          dk[s] += k[s];
        }
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        if (s < ns) {
          en[s] = ukeep[s] * R[s];          // E_b = T_b . R_b (thermal.py:45-49)
          gu_node[(b0 + s) * nn] = cut ? (T)0 : args.out_scale * R[s];
          if (gk_node) gk_node[(b0 + s) * nn] = args.out_scale * dk[s];
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

  // phase A of one pass: this thread's element for the S samples of the pass (independent instruction streams)
  auto phase_a = [&](const V* st, V* out) {
    if (!has_el) return;
    T Te[S][A], Ke[S][A], re[S][A], dK[S][A];
#pragma unroll
    for (int b = 0; b < A; ++b) {
      T tv[S], kv[S];
      unpack(st[my_ln[b]], tv);
      unpack(st[LCAP + my_ln[b]], kv);
#pragma unroll
      for (int s = 0; s < S; ++s) {
        Te[s][b] = tv[s];
        Ke[s][b] = kv[s];
      }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      if constexpr (AFFINE) {
        const T jinv[4] = {greg[0], greg[1], greg[2], greg[3]};
        thermal_vectors_affine<T, NL>(jinv, greg[4], Te[s], Ke[s], args.p.v[5], args.p.v[6], re[s], dK[s]);
      } else {
        thermal_vectors<T, QUAD, 2, NL>(greg, Te[s], Ke[s], args.p.v[5], args.p.v[6], re[s], dK[s]);
      }
    }
#pragma unroll
    for (int b = 0; b < A; ++b) {
      T rv[S], kv[S];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        rv[s] = re[s][b];
        kv[s] = dK[s][b];
      }
      out[b * BLOCK] = pack(rv);
      out[(4 + b) * BLOCK] = pack(kv);
    }
  };

  long long b0 = (long long)blockIdx.y * S;
  const long long bstep = (long long)gridDim.y * S;
  if (b0 >= args.nb) return;
  stage_pass(0, b0);
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  patch_pass(0);
  __syncthreads();
  // warps sharing a scheduler take the two phases in opposite order (one warp's FP stream covers the other's
  // shared-memory latencies)
  const int wid = tid >> 5;
  const bool a_first = (((wid >> 2) ^ wid ^ (int)blockIdx.x ^ (int)blockIdx.y) & 1) != 0;
  int buf = 0;
  long long bprev = -1;
  for (; b0 < args.nb; b0 += bstep, buf ^= 1) {
    if (b0 + bstep < args.nb) stage_pass(buf ^ 1, b0 + bstep);
    const V* st0 = stage + (size_t)buf * 2 * LCAP;
    const int nsprev = (args.nb - bprev < S) ? (int)(args.nb - bprev) : S;
    if (a_first) {
      phase_a(st0, sv + buf * SVB + tid);
      if (bprev >= 0) phase_b(sv + (buf ^ 1) * SVB, bprev, nsprev);
    } else {
      if (bprev >= 0) phase_b(sv + (buf ^ 1) * SVB, bprev, nsprev);
      phase_a(st0, sv + buf * SVB + tid);
    }
    if (has_node) unpack(st0[lnode], ukeep);
    bprev = b0;
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    patch_pass(buf ^ 1);
    __syncthreads();   // sv[buf] complete and the next pass's rows visible; sv[buf^1] / stage[buf] free again
  }
  phase_b(sv + (buf ^ 1) * SVB, bprev, (args.nb - bprev < S) ? (int)(args.nb - bprev) : S);
}

}  // namespace fol
