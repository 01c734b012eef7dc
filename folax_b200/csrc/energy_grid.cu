// Batched thermal loss + VJP on a STRUCTURED Quad4 grid (BASELINE.json configs[2]: 256 x 256 quads, 1024 samples):
// ThermalLoss2DQuad.ComputeBatchLoss and its JAX-AD gradient (thermal.py:28-49, fe_loss.py:250-262) for meshes whose
// nodes are numbered row-major (node(c, r) = r (nx + 1) + c), whose elements are [n, n + 1, n + nx + 2, n + nx + 1]
// (fol/tools/usefull_functions.py:213-258 builds exactly that) and whose elements are all the same parallelogram, so
// that J^-1 and w detJ are two launch constants.  The host plan establishes those facts (energy_plan.grid_structure);
// any other mesh keeps the tile kernels (energy_qt.cuh / energy2.cuh), which gather through the connectivity.
//
// What the structure buys: no connectivity, no geometry cache, no adjacency lists, no CTA barrier per pass and almost
// no address arithmetic -- the float64 kernel is left with the FP64 pipe as its bound.
//   * A CTA owns `rows` node rows of ONE sample (float32: of TWO samples, packed per lane -- FMUL2 / FADD2 / FFMA2)
//     over a panel of 32 W element columns (W <= 8 consumer warps) and marches UP the grid: lane l of warp w owns node
//     column c0 + 32 w + l and, per step, the element between rows e, e + 1 and columns c, c + 1.
//   * One PRODUCER warp streams the node rows of T, K (and the Dirichlet values) into a shared-memory ring with 1-D bulk
//     copies (cp.async.bulk + mbarrier complete_tx: one instruction per array and row for the whole CTA; sources
//     aligned down to 16 bytes, the lanes read at the row's shift); the consumer warps wait on the row's `full`
//     barrier, take their four corner values with four shared-memory loads (the bottom corners are the previous step's
//     top corners) and release the slot through its `empty` barrier.
//   * The element-vector entries leave through registers: the two left corners accumulate in the lane, the two right
//     corners are summed per lane over consecutive rows and handed to lane l + 1 by one shuffle: every node sum has a
//     fixed order, no atomics, deterministic.
//   * The element arithmetic is sum-factorised for the bilinear element with the 2 x 2 rule and the edge values of a
//     node row are computed once for the element rows below and above it (115 FP instructions per element on
//     axis-aligned grids, 123 with the node sums; `grid_edge`, `grid_element`), every operation with its rounding
//     written out, so the results do not depend on chunk height, batch size or code path.
//   * The node column shared by two warps gets its two halves through shared memory once per chunk (`combine`), the
//     only CTA barrier of the kernel.  The chunk recomputes the element row below it (1 / rows extra work) instead of
//     exchanging partial sums with the chunk below; panels overlap by one element column in the same way (nx > 256).
//   * Dirichlet handling (overwrite of T while reading, cut of the cotangent at the store) runs only in the warps
//     whose columns hold a Dirichlet node (`col_dir`, a per-column flag from the host plan).
#include "energy_grid_common.cuh"

namespace fol {

template <class T>
struct GridArgs {
  const T* ctrl;              // (nb, nn)
  const T* u;                 // (nb, nn)
  T* grad_u;                  // (nb, nn)
  T* grad_k;                  // (nb, nn) or null
  T* partial;                 // (nb, npart) energy shares, one per warp
  const T* dir_values;        // (nn) NaN where free, or null: overwrites u while reading (fe_loss.py:91-92, 255)
  const uint8_t* dir_flag;    // (nn) 1 where grad_u is written as zero, or null
  const uint8_t* col_dir;     // (nx + 1) 1 where the node column holds a Dirichlet node, or null (= every column may)
  T out_scale, beta, cexp, wd;
  T jinv[4];                  // row-major d xi_j / d x_k of the one element shape
  int nx, ny;                 // elements per direction
  int rows, nchunks, npanels, npart, W;
  int row_bytes;              // bytes of one staged row (16-byte multiple)
  long long nn, nb;
};

namespace {

// 1 + beta T^c (thermal.py:34); integer powers by repeated multiplication like lax.integer_pow
template <class T, int NL>
__device__ __forceinline__ T grid_conductivity(T tg, T beta, T c) {
  const T one = bcd<T>(1.0);
  if constexpr (NL == 0) return one;
  else if constexpr (NL == 1) return op_fma(beta, tg, one);
  else if constexpr (NL == 2) return op_fma(beta, op_mul(tg, tg), one);
  else if constexpr (NL == 3) return op_fma(beta, op_mul(tg, op_mul(tg, tg)), one);
  else if constexpr (NL == 4) { const T t2 = op_mul(tg, tg); return op_fma(beta, op_mul(t2, t2), one); }
  else {
    if constexpr (std::is_same<T, float2>::value)
      return op_fma(beta, make_float2(pow_c<float>(tg.x, c.x), pow_c<float>(tg.y, c.y)), one);
    else return op_fma(beta, pow_c<T>(tg, c), one);
  }
}

// Values a node row contributes to the elements below and above it, for the lane's column pair (own, right): T and K
// on the edge at xi = -s / +s (1-D Lagrange weights a = (1 - s)/2, b = (1 + s)/2, s = 1/sqrt(3)) and the difference of
// T along the edge.  Computed ONCE per node row and lane: the top edge of an element row is the bottom edge of the next.
template <class T>
struct GridEdge {
  T m, p, d, km, kp;
};
template <class T>
__device__ __forceinline__ GridEdge<T> grid_edge(T t_own, T t_right, T k_own, T k_right) {
  constexpr double s = FOL_S3;
  const T a = bcd<T>(0.5 * (1.0 - s)), b = bcd<T>(0.5 * (1.0 + s));
  GridEdge<T> e;
  e.m = lin2(b, t_own, a, t_right);
  e.p = lin2(a, t_own, b, t_right);
  e.d = op_sub(t_right, t_own);
  e.km = lin2(b, k_own, a, k_right);
  e.kp = lin2(a, k_own, b, k_right);
  return e;
}

// Element vectors re = dE/dT_e and dK = dE/dK_e of the thermal Quad4 with the 2 x 2 rule on a parallelogram, local
// nodes 0 (-,-), 1 (+,-), 2 (+,+), 3 (-,+) and Gauss points in the same order (quadrilateral_2d_4.py:54-58), from the
// edge values of its bottom (B) and top (U) node rows: values at the four points by the same weights across eta;
// dT/dxi depends on eta only, dT/deta on xi only (so do their products with J^-1 on an axis-aligned grid); the weighted
// fluxes go back to the nodes through the same weights.  Same sums as thermal_vectors_affine (energy_qt.cuh) in
// another association: equal to rounding.  e_el = T_e . re (thermal.py:45-49) = sum_g K_g (w detJ nl_g |grad T_g|^2).
// 115 instructions on an axis-aligned grid (DIAG), 139 on a sheared one.  T: the lane type (one or two samples).
template <class T, int NL, bool DIAG, bool GK>
__device__ __forceinline__ void grid_element(const GridEdge<T>& B, const GridEdge<T>& U, const T (&ji)[4], T wd, T beta,
                                             T cexp, T (&re)[4], T (&dK)[4], T& e_el) {
  constexpr double s = FOL_S3;
  const T a = bcd<T>(0.5 * (1.0 - s)), b = bcd<T>(0.5 * (1.0 + s)), ah = bcd<T>(0.25 * (1.0 - s)),
          bh = bcd<T>(0.25 * (1.0 + s)), half = bcd<T>(0.5);
  const T t0e[2] = {lin2(bh, B.d, ah, U.d), lin2(ah, B.d, bh, U.d)};        // dT/dxi at eta = -s, +s
  const T t1x[2] = {op_sub(U.m, B.m), op_sub(U.p, B.p)};                    // 2 dT/deta at xi = -s, +s
  const T tg[4] = {lin2(b, B.m, a, U.m), lin2(b, B.p, a, U.p), lin2(a, B.p, b, U.p), lin2(a, B.m, b, U.m)};
  const T eg[4] = {lin2(b, B.km, a, U.km), lin2(b, B.kp, a, U.kp), lin2(a, B.kp, b, U.kp), lin2(a, B.km, b, U.km)};
  constexpr int ETA[4] = {0, 0, 1, 1}, XI[4] = {0, 1, 1, 0};                // eta / xi index of Gauss point g
  T w0[4], w1[4], ck[4];
  if constexpr (DIAG) {
    // grad T = (j00 dT/dxi, j11 dT/deta): two values each; so are the flux factors j00 gx, j11 gy and the squares
    const T j3h = op_mul(half, ji[3]);
    T gx[2], gy[2], fx[2], fy[2], gx2[2], gy2[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      gx[q] = op_mul(t0e[q], ji[0]);
      gy[q] = op_mul(t1x[q], j3h);
      fx[q] = op_mul(ji[0], gx[q]);
      fy[q] = op_mul(ji[3], gy[q]);
      gx2[q] = op_mul(gx[q], gx[q]);
      gy2[q] = op_mul(gy[q], gy[q]);
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const T wn = op_mul(wd, grid_conductivity<T, NL>(tg[g], beta, cexp));
      const T cf = op_mul(wn, eg[g]);
      ck[g] = op_mul(wn, op_add(gx2[ETA[g]], gy2[XI[g]]));
      w0[g] = op_mul(cf, fx[ETA[g]]);
      w1[g] = op_mul(cf, fy[XI[g]]);
    }
  } else {
    const T j2h = op_mul(half, ji[2]), j3h = op_mul(half, ji[3]);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const T gx = lin2(t0e[ETA[g]], ji[0], t1x[XI[g]], j2h);                // grad T = J^-T (dN^T T)
      const T gy = lin2(t0e[ETA[g]], ji[1], t1x[XI[g]], j3h);
      const T wn = op_mul(wd, grid_conductivity<T, NL>(tg[g], beta, cexp));
      const T cf = op_mul(wn, eg[g]);
      ck[g] = op_mul(wn, lin2(gx, gx, gy, gy));
      w0[g] = op_mul(cf, lin2(ji[0], gx, ji[1], gy));
      w1[g] = op_mul(cf, lin2(ji[2], gx, ji[3], gy));
    }
  }
  const T s0lo = op_add(w0[0], w0[1]), s0hi = op_add(w0[2], w0[3]);
  const T s1l = op_add(w1[0], w1[3]), s1r = op_add(w1[1], w1[2]);
  const T S0b = lin2(bh, s0lo, ah, s0hi), S0t = lin2(ah, s0lo, bh, s0hi);  // sum_g dN/dxi weights, bottom / top nodes
  const T S1l = lin2(bh, s1l, ah, s1r), S1r = lin2(ah, s1l, bh, s1r);      // sum_g dN/deta weights, left / right nodes
  re[0] = op_neg(op_add(S0b, S1l));
  re[1] = op_sub(S0b, S1r);
  re[2] = op_add(S0t, S1r);
  re[3] = op_sub(S1l, S0t);
  if constexpr (GK) {
    const T Cml = lin2(b, ck[0], a, ck[1]), Cmr = lin2(a, ck[0], b, ck[1]);  // eta = -s row reduced to the left / right nodes
    const T Cpl = lin2(b, ck[3], a, ck[2]), Cpr = lin2(a, ck[3], b, ck[2]);  // eta = +s row
    dK[0] = lin2(b, Cml, a, Cpl);
    dK[1] = lin2(b, Cmr, a, Cpr);
    dK[2] = lin2(a, Cmr, b, Cpr);
    dK[3] = lin2(a, Cml, b, Cpl);
  } else {
    dK[0] = dK[1] = dK[2] = dK[3] = bcd<T>(0.0);
  }
  e_el = op_fma(eg[3], ck[3], op_fma(eg[2], ck[2], op_fma(eg[1], ck[1], op_mul(eg[0], ck[0]))));
}

}  // namespace

template <class S, int NS, int NL, bool DIAG, bool GK>
// 9 warps per CTA (8 consumers + the producer).  Registers are per SCHEDULER (16 K each): 96 registers let a scheduler
// host 5 float64 warps (two CTAs = 18 warps per SM), 72 registers 7 float32 warps (three CTAs = 27 warps; measured
// with two samples per lane: 0.337 ms at 72 registers, 0.382 ms at 96; float64 at 72 registers spills and gains nothing)
__global__ void __launch_bounds__(288) __maxnreg__(sizeof(S) == 8 ? FOL_GRID_REGS64 : (NS == 2 ? FOL_GRID_REGS32P : 72))
    energy_grid_kernel(const GridArgs<S> args) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using V = typename LaneT<S, NS>::type;                     // lane value: one sample, or two float32 samples
  using Pair = NodePair<V>;
  constexpr int PER = 16 / (int)sizeof(S);
  constexpr int NROW = 2 * NS + 1;                           // ring rows per slot: T (per sample), K (per sample), D
  const int W = args.W, tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  // shared memory: ring [kRing][NROW][row_bytes] | left [W][rows] | right [W][rows] | barriers
  unsigned char* const ring = smem_raw;
  const int slot_bytes = NROW * args.row_bytes;
  Pair* const left = reinterpret_cast<Pair*>(smem_raw + (size_t)kRing * slot_bytes);
  Pair* const right = left + (size_t)W * args.rows;
  unsigned long long* const bars = reinterpret_cast<unsigned long long*>(right + (size_t)W * args.rows);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kRing);

  // item = (sample group, chunk of node rows, panel of element columns)
  long long item = blockIdx.x;
  const int panel = (int)(item % args.npanels);
  item /= args.npanels;
  const int chunk = (int)(item % args.nchunks);
  long long smp[NS];                                         // the lane's samples (an odd batch repeats its last one)
  smp[0] = (item / args.nchunks) * NS;
  if constexpr (NS == 2) smp[1] = smp[0] + 1 < args.nb ? smp[0] + 1 : smp[0];
  const bool second = NS == 2 && smp[0] + 1 < args.nb;       // the second sample is a real one
  const int nx = args.nx, ny = args.ny, NXn = nx + 1;
  const int sp = panel * (32 * W - 1);                       // first element column of the panel
  const int ncols = min(32 * W + 1, NXn - sp);               // node columns the panel stages
  const int r0 = chunk * args.rows, r1 = min(r0 + args.rows, ny + 1);   // owned node rows [r0, r1)
  const int e_beg = max(r0 - 1, 0), e_end = min(r1, ny);     // element rows [e_beg, e_end): one recomputed row below
  const int nstage = e_end - e_beg + 1;                      // node rows e_beg .. e_end
  const bool has_dirv = args.dir_values != nullptr;

  if (tid == 0) {
    for (int s = 0; s < kRing; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, W);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (w == W) {
    // ---- producer: lane 0 streams the node rows e_beg .. e_end into the ring
    if (l == 0) {
      const long long total = args.nb * args.nn;
      for (int i = 0; i < nstage; ++i) {
        const int slot = i % kRing;
        if (i >= kRing) mbar_wait(empty0 + 8 * slot, ((i / kRing) - 1) & 1);
        unsigned char* const base = ring + (size_t)slot * slot_bytes;
        const uint32_t bar = full0 + 8 * slot;
        const long long first_d = (long long)(e_beg + i) * NXn + sp;
        RowCopy cu[NS];                                      // u and ctrl: same shape, same offsets
#pragma unroll
        for (int j = 0; j < NS; ++j) cu[j] = plan_row<S>(total, smp[j] * args.nn + first_d, ncols);
        const RowCopy cd = plan_row<S>(args.nn, first_d, ncols);
        // plain tail stores (the last row of the last sample only) first, then the arrive that publishes them and
        // arms the transaction count, then the bulk copies that complete it
        uint32_t bytes = has_dirv ? cd.bytes : 0u;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          copy_tail<S>(cu[j], args.u, reinterpret_cast<S*>(base + j * args.row_bytes));
          copy_tail<S>(cu[j], args.ctrl, reinterpret_cast<S*>(base + (NS + j) * args.row_bytes));
          bytes += 2 * cu[j].bytes;
        }
        if (has_dirv) copy_tail<S>(cd, args.dir_values, reinterpret_cast<S*>(base + 2 * NS * args.row_bytes));
        mbar_expect_tx(bar, bytes);
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          copy_bulk<S>(cu[j], args.u, reinterpret_cast<S*>(base + j * args.row_bytes), bar);
          copy_bulk<S>(cu[j], args.ctrl, reinterpret_cast<S*>(base + (NS + j) * args.row_bytes), bar);
        }
        if (has_dirv) copy_bulk<S>(cd, args.dir_values, reinterpret_cast<S*>(base + 2 * NS * args.row_bytes), bar);
      }
    }
  } else {
    // ---- consumers
    const int c = sp + 32 * w + l;                           // own node column = left column of this lane's element
    const bool el_valid = c < nx;
    const bool all_valid = __all_sync(0xffffffffu, el_valid);
    const bool write_own = c <= nx && (l > 0 || w == 0) && (c > sp || panel == 0);   // lane 0 of warps >= 1: `combine`
    const bool count = el_valid && (c > sp || panel == 0);   // the panel's first element column belongs to the panel left of it
    const bool keep_left = l == 0 && w > 0, keep_right = l == 31;
    const int cc = min(c, nx);
    const int i_own = min(32 * w + l, ncols - 1);            // staged entry of the own column; the right one is i_own + 1
    // Dirichlet work only in the warps whose 33 columns hold a Dirichlet node
    bool wdir = has_dirv || args.dir_flag != nullptr;
    if (wdir && args.col_dir) {
      const bool mine = args.col_dir[cc] != 0 || (l == 31 && args.col_dir[min(c + 1, nx)] != 0);
      wdir = __any_sync(0xffffffffu, mine);
    }
    const bool wdirv = wdir && has_dirv, wcut = wdir && args.dir_flag != nullptr;
    const bool has_gk = GK && args.grad_k != nullptr;

    const V ji[4] = {bc<V>(args.jinv[0]), bc<V>(args.jinv[1]), bc<V>(args.jinv[2]), bc<V>(args.jinv[3])};
    const V wd = bc<V>(args.wd), beta = bc<V>(args.beta), cexp = bc<V>(args.cexp), scale = bc<V>(args.out_scale);
    const V zero = bcd<V>(0.0);
    // shared-memory addresses of this lane's entries in ring slot 0 (first T row; the other rows follow at row_bytes)
    const uint32_t ring_bytes = (uint32_t)(kRing * slot_bytes), rb = (uint32_t)args.row_bytes;
    const uint32_t a_own = smem_u32(ring) + (uint32_t)i_own * (uint32_t)sizeof(S);
    uint32_t slot_off = 0, parity = 0, slot_bar = 0;         // ring position of the next row to take
    const long long row_d = (long long)e_beg * NXn + sp;
    uint32_t sh_u[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) sh_u[j] = (uint32_t)((smp[j] * args.nn + row_d) & (PER - 1)) * (uint32_t)sizeof(S);
    uint32_t sh_d = (uint32_t)(row_d & (PER - 1)) * (uint32_t)sizeof(S);
    const uint32_t sh_step = (uint32_t)(NXn & (PER - 1)) * (uint32_t)sizeof(S);
    // global element offset of node (cc, row) within the sample, as 32 bits (nn < 2^31 is checked by the host)
    unsigned node = (unsigned)e_beg * (unsigned)NXn + (unsigned)cc;
    S* gu0[NS];
    S* gk0[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      gu0[j] = args.grad_u + smp[j] * args.nn;
      gk0[j] = has_gk ? args.grad_k + smp[j] * args.nn : gu0[j];
    }
    const uint8_t* const fl0 = wcut ? args.dir_flag : args.col_dir;   // never read unless wcut
    uint32_t a_left = smem_u32(left + (size_t)w * args.rows), a_right = smem_u32(right + (size_t)w * args.rows);

    auto lds = [](uint32_t a) {
      S v;
      if constexpr (sizeof(S) == 8) asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(a));
      else asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(a));
      return v;
    };
    // value of ring row `row0` (+ 1 for the second sample) at byte address a
    auto ldv = [&](uint32_t a, const uint32_t (&sh)[NS]) {
      if constexpr (NS == 2) return make_float2(lds(a + sh[0]), lds(a + rb + sh[1]));
      else return lds(a + sh[0]);
    };
    auto sts_pair = [](uint32_t a, V x, V y) {
      if constexpr (sizeof(V) == 8 && NS == 1) asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(a), "d"(x), "d"(y) : "memory");
      else if constexpr (NS == 2)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(a), "f"(x.x), "f"(x.y), "f"(y.x), "f"(y.y) : "memory");
      else asm volatile("st.shared.v2.f32 [%0], {%1, %2};\n" ::"r"(a), "f"(x), "f"(y) : "memory");
    };
    auto shfl_up = [](V v) {
      if constexpr (NS == 2) return make_float2(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1));
      else return __shfl_up_sync(0xffffffffu, v, 1);
    };
    // the four values of the next staged node row: own / right column of T and K
    auto take = [&](V& t0, V& k0, V& t1, V& k1) {
      mbar_wait(full0 + slot_bar, parity);
      const uint32_t aT = a_own + slot_off, aK = aT + NS * rb;
      t0 = ldv(aT, sh_u);
      t1 = ldv(aT + (uint32_t)sizeof(S), sh_u);
      k0 = ldv(aK, sh_u);
      k1 = ldv(aK + (uint32_t)sizeof(S), sh_u);
      if (wdirv) {                                           // Dirichlet overwrite (fe_loss.py:91-92, 255)
        const uint32_t aD = aT + 2u * NS * rb + sh_d;
        const S d0 = lds(aD), d1 = lds(aD + (uint32_t)sizeof(S));
        if (d0 == d0) t0 = bc<V>(d0);
        if (d1 == d1) t1 = bc<V>(d1);
      }
      __syncwarp();
      if (l == 0) mbar_arrive(empty0 + slot_bar);
      slot_off += (uint32_t)slot_bytes;
      slot_bar += 8;
      if (slot_off == ring_bytes) {
        slot_off = 0;
        slot_bar = 0;
        parity ^= 1;
      }
#pragma unroll
      for (int j = 0; j < NS; ++j) sh_u[j] = (sh_u[j] + sh_step) & 15u;
      sh_d = (sh_d + sh_step) & 15u;
    };

    V en = zero;
    // node row (at offset `node`) is complete once the element rows below and above it are in: left half (own lane:
    // below + above) + right half of the lane to the left; the column two warps share waits for `combine`
    uint8_t cut_now = 0;                                     // dir_flag of node (cc, current row), loaded one row ahead
    if (wcut) cut_now = fl0[node + (e_beg < r0 ? (unsigned)NXn : 0u)];
    auto finish_row = [&](V leftR, V leftK, V rpR, V rpK, bool more) {
      const V inR = shfl_up(rpR), inK = shfl_up(rpK);
      uint8_t cut_next = 0;
      if (wcut && more) cut_next = fl0[node + (unsigned)NXn];
      if (keep_left) sts_pair(a_left, leftR, leftK);
      if (keep_right) sts_pair(a_right, rpR, rpK);
      a_left += (uint32_t)sizeof(Pair);
      a_right += (uint32_t)sizeof(Pair);
      V R = (l > 0) ? op_add(leftR, inR) : leftR;
      const V K = (l > 0) ? op_add(leftK, inK) : leftK;
      if (wcut && cut_now != 0) R = zero;
      if (write_own) {
        const V oR = op_mul(scale, R), oK = op_mul(scale, K);
        if constexpr (NS == 2) {
          gu0[0][node] = oR.x;
          if (has_gk) gk0[0][node] = oK.x;
          if (second) {
            gu0[1][node] = oR.y;
            if (has_gk) gk0[1][node] = oK.y;
          }
        } else {
          gu0[0][node] = oR;
          if (has_gk) gk0[0][node] = oK;
        }
      }
      cut_now = cut_next;
    };
    // one element row between the edge values of the node rows below (B) and above (U); carries of the row below in
    // (oc, rc)
    using Edge = GridEdge<V>;
    V ocR = zero, ocK = zero, rcR = zero, rcK = zero;
    auto element = [&](const Edge& B, const Edge& U, V (&re)[4], V (&dK)[4], V& e_el) {
      grid_element<V, NL, DIAG, GK>(B, U, ji, wd, beta, cexp, re, dK, e_el);
      if (!all_valid) {                                      // warp-uniform: a ragged last warp only
        if (!el_valid) {
#pragma unroll
          for (int q = 0; q < 4; ++q) re[q] = dK[q] = zero;
          e_el = zero;
        }
      }
    };
    auto step = [&](const Edge& B, const Edge& U) {
      V re[4], dK[4], e_el;
      element(B, U, re, dK, e_el);
      en = op_add(en, e_el);
      finish_row(op_add(ocR, re[0]), op_add(ocK, dK[0]), op_add(rcR, re[1]), op_add(rcK, dK[1]), true);
      node += (unsigned)NXn;
      ocR = re[3]; ocK = dK[3]; rcR = re[2]; rcK = dK[2];
    };
    auto take_edge = [&]() {
      V t0, k0, t1, k1;
      take(t0, k0, t1, k1);
      return grid_edge<V>(t0, t1, k0, k1);
    };

    Edge EA = take_edge(), EB;                               // two register sets of edge values, used alternately
    int e = e_beg;
    if (e_beg < r0) {
      // the recomputed element row below the chunk: only its shares of node row r0 (the carries) are kept
      EB = take_edge();
      V re[4], dK[4], e_el;
      element(EA, EB, re, dK, e_el);
      ocR = re[3]; ocK = dK[3]; rcR = re[2]; rcK = dK[2];
      node += (unsigned)NXn;
      EA = EB;
      ++e;
    }
#if FOL_GRID_PREFETCH
    // variant (measured SLOWER: 0.692 vs 0.660 ms float64, 0.365 vs 0.338 ms float32, profiles/r2/
    // energy_grid_variants.txt): the next node row is taken between an element's arithmetic and its node sums, so that
    // the barrier wait and the shared-memory loads of row e + 2 overlap the shuffles / stores that close row e
    if (e < e_end) {
      EB = take_edge();
      for (;;) {
        V re[4], dK[4], e_el;
        element(EA, EB, re, dK, e_el);
        const bool more = e + 1 < e_end;
        Edge EN = EB;
        if (more) EN = take_edge();
        en = op_add(en, e_el);
        finish_row(op_add(ocR, re[0]), op_add(ocK, dK[0]), op_add(rcR, re[1]), op_add(rcK, dK[1]), true);
        node += (unsigned)NXn;
        ocR = re[3]; ocK = dK[3]; rcR = re[2]; rcK = dK[2];
        if (!more) break;
        EA = EB;
        EB = EN;
        ++e;
      }
    }
#else
    for (; e + 1 < e_end; e += 2) {
      EB = take_edge();
      step(EA, EB);
      EA = take_edge();
      step(EB, EA);
    }
    if (e < e_end) {
      EB = take_edge();
      step(EA, EB);
    }
#endif
    if (r1 == ny + 1) finish_row(ocR, ocK, rcR, rcK, false);  // the top node row of the grid closes with the carries alone

    // energy shares of this warp
    if (!count) en = zero;
    const long long pslot = ((long long)panel * args.nchunks + chunk) * W + w;
    if constexpr (NS == 2) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        en.x += __shfl_xor_sync(0xffffffffu, en.x, o);
        en.y += __shfl_xor_sync(0xffffffffu, en.y, o);
      }
      if (l == 0) {
        args.partial[smp[0] * args.npart + pslot] = en.x;
        if (second) args.partial[smp[1] * args.npart + pslot] = en.y;
      }
    } else {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) en += __shfl_xor_sync(0xffffffffu, en, o);
      if (l == 0) args.partial[smp[0] * args.npart + pslot] = en;
    }
  }

  // combine: node columns sp + 32 b (b = 1..W) got their left half from lane 31 of warp b - 1 and their right half from
  // lane 0 of warp b; column sp + 32 W closes here only when it is the grid's last column
  __syncthreads();
  const int nrow = r1 - r0;
  for (int idx = tid; idx < W * nrow; idx += blockDim.x) {
    const int bnd = idx / nrow + 1, i = idx - (bnd - 1) * nrow;
    const int col = sp + 32 * bnd;
    if (col > nx || (bnd == W && col != nx)) continue;
    const Pair lo = right[(size_t)(bnd - 1) * args.rows + i];
    V R = lo.t, K = lo.k;
    if (bnd < W) {
      const Pair hi = left[(size_t)bnd * args.rows + i];
      R = op_add(hi.t, R);                                   // same order as finish_row: own (left) half + incoming
      K = op_add(hi.k, K);
    }
    const long long node = (long long)(r0 + i) * NXn + col;
    const bool cut = args.dir_flag ? args.dir_flag[node] != 0 : false;
    const V oR = op_mul(bc<V>(args.out_scale), R), oK = op_mul(bc<V>(args.out_scale), K);
    if constexpr (NS == 2) {
      args.grad_u[smp[0] * args.nn + node] = cut ? 0.f : oR.x;
      if (GK && args.grad_k) args.grad_k[smp[0] * args.nn + node] = oK.x;
      if (second) {
        args.grad_u[smp[1] * args.nn + node] = cut ? 0.f : oR.y;
        if (GK && args.grad_k) args.grad_k[smp[1] * args.nn + node] = oK.y;
      }
    } else {
      args.grad_u[smp[0] * args.nn + node] = cut ? (S)0 : oR;
      if (GK && args.grad_k) args.grad_k[smp[0] * args.nn + node] = oK;
    }
  }
}

namespace {

template <class T, int NS>
size_t grid_smem(int W, long long nx, int rows) {
  return (size_t)kRing * (2 * NS + 1) * grid_row_bytes<T>(W, nx) + (size_t)2 * W * rows * 2 * NS * sizeof(T) + 2 * kRing * 8;
}

template <class T, int NS, int NL, bool DIAG, bool GK>
int launch_grid(cudaStream_t s, GridArgs<T> a, T* energy) {
  auto kern = energy_grid_kernel<T, NS, NL, DIAG, GK>;
  const GridShape g = grid_shape(a.nx);
  const int threads = 32 * (g.W + 1);
  static PerDeviceOnce configured;
  if (configured.need()) {
    FOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grid_smem<T, NS>(8, 256, 128)));
    FOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured.done();
  }
  // rows per chunk: whole waves of resident CTAs against the recomputed row and the pipeline fill of every chunk
  static const int forced = energy2_env_int("FOL_ENERGY_GRID_ROWS", 0);
  int best_rows = 0;
  double best = -1.0;
  int sms = 148, dev = 0;
  FOL_CUDA(cudaGetDevice(&dev));
  FOL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int nrows_total = a.ny + 1;
  const long long groups = cdiv(a.nb, NS);
  for (int rows = kMinRows; rows <= 128; ++rows) {
    if (rows > nrows_total && rows != kMinRows) break;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, grid_smem<T, NS>(g.W, a.nx, rows)) !=
            cudaSuccess || per_sm < 1)
      continue;
    const long long nchunks = cdiv(nrows_total, rows);
    const long long items = nchunks * g.npanels * groups, slots = (long long)sms * per_sm;
    const double waves = (double)cdiv(items, slots);
    const double work = (double)(a.ny + (nchunks - 1)) + 3.0 * nchunks;     // element rows computed + fill, per sample
    const double eff = ((double)a.ny / work) * ((double)items / (waves * slots));
    if (eff > best + 1e-9) {
      best = eff;
      best_rows = rows;
    }
  }
  if (forced >= kMinRows && forced <= 128) best_rows = forced;
  if (best_rows == 0) return fail(FOL_ERR_CUDA, "fol_energy_and_grads_grid: the kernel does not fit on this device");
  a.rows = best_rows;
  a.nchunks = (int)cdiv(nrows_total, best_rows);
  a.npanels = g.npanels;
  a.W = g.W;
  a.row_bytes = grid_row_bytes<T>(g.W, a.nx);
  a.npart = a.nchunks * a.npanels * g.W;
  const long long items = (long long)a.nchunks * a.npanels * groups;
  FOL_REQUIRE(items < (1LL << 31), "fol_energy_and_grads_grid: too many work items for one launch");
  kern<<<(unsigned)items, threads, grid_smem<T, NS>(g.W, a.nx, best_rows), s>>>(a);
  int rc = check_launch("energy_grid_kernel");
  if (rc) return rc;
  energy_sum_kernel<T><<<(unsigned)cdiv(a.nb, 8), 256, 0, s>>>(a.partial, a.nb, a.npart, energy);
  return check_launch("energy_sum_kernel");
}

template <class T, int NL, bool DIAG, bool GK>
int launch_grid_ns(cudaStream_t s, const GridArgs<T>& a, T* energy) {
  if constexpr (sizeof(T) == 4) {
    // float32: two samples per lane on the packed FP32 instructions -- also for a single sample (evaluated twice,
    // stored once), so that a sample's result never depends on the batch it came in (the packed and the scalar
    // instruction streams were measured 1 ulp apart in a third of the entries)
    static const int pair = energy2_env_int("FOL_ENERGY_GRID_PAIR", 1);
    if (pair) return launch_grid<T, 2, NL, DIAG, GK>(s, a, energy);
  }
  return launch_grid<T, 1, NL, DIAG, GK>(s, a, energy);
}

template <class T>
int dispatch_grid(cudaStream_t s, const GridArgs<T>& a, T* energy) {
  const T beta = a.beta, c = a.cexp;
  const int ci = (int)c;
  const int nl = (beta == (T)0) ? 0 : (((T)ci == c && ci >= 1 && ci <= 4) ? ci : -1);
  const bool diag = a.jinv[1] == (T)0 && a.jinv[2] == (T)0;
  const bool gk = a.grad_k != nullptr;
#define FOL_GRID(NLV)                                                                   \
  if (nl == NLV) {                                                                      \
    if (diag) return gk ? launch_grid_ns<T, NLV, true, true>(s, a, energy) : launch_grid_ns<T, NLV, true, false>(s, a, energy);   \
    return gk ? launch_grid_ns<T, NLV, false, true>(s, a, energy) : launch_grid_ns<T, NLV, false, false>(s, a, energy);           \
  }
#ifdef FOL_GRID_ONLY_NL4      /* tuning builds: one conductivity law, a sixth of the compile time */
  FOL_GRID(4)
#else
  FOL_GRID(0) FOL_GRID(1) FOL_GRID(2) FOL_GRID(3) FOL_GRID(4) FOL_GRID(-1)
#endif
#undef FOL_GRID
  return fail(FOL_ERR_INVALID, "fol_energy_and_grads_grid: bad conductivity law");
}

}  // namespace
}  // namespace fol

using namespace fol;

extern "C" {

int64_t fol_energy_grid_work_size(int64_t nx, int64_t ny, int64_t nb) {
  if (nx < 1 || ny < 1 || nb < 0) return 0;
  const GridShape g = grid_shape(nx);
  return nb * (cdiv(ny + 1, kMinRows) + 1) * g.npanels * g.W + 16;
}

int fol_energy_and_grads_grid(fol_stream_t s, int dtype, int64_t nx, int64_t ny, int64_t nb, const double* jinv_host,
                              double w_detj, const void* ctrl, const void* u, const void* dir_values,
                              const uint8_t* dir_flag, const uint8_t* col_dir, double out_scale,
                              const double* params_host, void* grad_u, void* grad_k, void* energy, void* work) {
  FOL_REQUIRE(nx >= 1 && ny >= 1 && nb >= 0 && (nx + 1) * (ny + 1) < (1LL << 31), "fol_energy_and_grads_grid: bad grid size");
  FOL_REQUIRE(jinv_host && params_host && ctrl && u && grad_u && energy && work, "fol_energy_and_grads_grid: null pointer");
  FOL_REQUIRE(dtype == FOL_F64 || dtype == FOL_F32, "fol_energy_and_grads_grid: unknown dtype");
  FOL_REQUIRE(((uintptr_t)ctrl | (uintptr_t)u | (uintptr_t)dir_values) % 16 == 0,
              "fol_energy_and_grads_grid: ctrl, u and dir_values must be 16-byte aligned (bulk copies)");
  if (nb == 0) return FOL_OK;
  auto run = [&](auto* tag) {
    using T = std::remove_pointer_t<decltype(tag)>;
    GridArgs<T> a;
    a.ctrl = (const T*)ctrl;
    a.u = (const T*)u;
    a.grad_u = (T*)grad_u;
    a.grad_k = (T*)grad_k;
    a.partial = (T*)work;
    a.dir_values = (const T*)dir_values;
    a.dir_flag = dir_flag;
    a.col_dir = col_dir;
    a.out_scale = (T)out_scale;
    a.beta = (T)params_host[5];
    a.cexp = (T)params_host[6];
    a.wd = (T)w_detj;
    for (int i = 0; i < 4; ++i) a.jinv[i] = (T)jinv_host[i];
    a.nx = (int)nx;
    a.ny = (int)ny;
    a.nn = (nx + 1) * (ny + 1);
    a.nb = nb;
    a.rows = a.nchunks = a.npanels = a.npart = a.W = a.row_bytes = 0;
    return dispatch_grid<T>((cudaStream_t)s, a, (T*)energy);
  };
  if (dtype == FOL_F64) return run((double*)nullptr);
  return run((float*)nullptr);
}

}  // extern "C"
