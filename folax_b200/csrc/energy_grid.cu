// Batched thermal loss + VJP on a STRUCTURED Quad4 grid (BASELINE.json configs[2]: 256 x 256 quads, 1024 samples):
// ThermalLoss2DQuad.ComputeBatchLoss and its JAX-AD gradient (thermal.py:28-49, fe_loss.py:250-262) for meshes whose
// nodes are numbered row-major (node(c, r) = r (nx + 1) + c), whose elements are [n, n + 1, n + nx + 2, n + nx + 1]
// (fol/tools/usefull_functions.py:213-258 builds exactly that) and whose elements are all the same parallelogram, so
// that J^-1 and w detJ are two launch constants.  The host plan establishes those facts (energy_plan.grid_structure);
// any other mesh keeps the tile kernels (energy_qt.cuh / energy2.cuh), which gather through the connectivity.
//
// What the structure buys: no connectivity, no geometry cache, no adjacency lists, no CTA barrier per pass and almost
// no address arithmetic -- the float64 kernel is left with the FP64 pipe as its bound.
//   * A CTA owns `rows` node rows of ONE sample over a panel of 32 W element columns (W <= 8 consumer warps) and
//     marches UP the grid: lane l of warp w owns node column c0 + 32 w + l and, per step, the element between rows
//     e, e + 1 and columns c, c + 1.
//   * One PRODUCER warp streams the node rows of T, K (and the Dirichlet values) into a shared-memory ring with 1-D bulk
//     copies (cp.async.bulk + mbarrier complete_tx: one instruction per array and row for the whole CTA; sources
//     aligned down to 16 bytes, the lanes read at the row's shift); the consumer warps wait on the row's `full`
//     barrier, take their four corner values with four shared-memory loads (the bottom corners are the previous step's
//     top corners) and release the slot through its `empty` barrier.
//   * The element-vector entries leave through registers: the two left corners accumulate in the lane, the two right
//     corners are summed per lane over consecutive rows and handed to lane l + 1 by one shuffle: every node sum has a
//     fixed order, no atomics, deterministic.
//   * The element arithmetic is sum-factorised for the bilinear element with the 2 x 2 rule (131 FP instructions per
//     element on axis-aligned grids; `grid_element`).
//   * The node column shared by two warps gets its two halves through shared memory once per chunk (`combine`), the
//     only CTA barrier of the kernel.  The chunk recomputes the element row below it (1 / rows extra work) instead of
//     exchanging partial sums with the chunk below; panels overlap by one element column in the same way (nx > 256).
//   * Dirichlet handling (overwrite of T while reading, cut of the cotangent at the store) runs only in the warps
//     whose columns hold a Dirichlet node (`col_dir`, a per-column flag from the host plan).
#include <type_traits>

#include "energy2.cuh"
#include "energy2_launch.cuh"

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

constexpr int kRing = 8;      // node rows in flight per CTA

template <class T>
struct alignas(2 * sizeof(T)) NodePair {
  T t, k;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "GRID_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra GRID_DONE_%=;\n\t"
      "bra GRID_WAIT_%=;\n\t"
      "GRID_DONE_%=:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <class T>
__device__ __forceinline__ T shfl_up1(T v) {
  return __shfl_up_sync(0xffffffffu, v, 1);
}

// Arithmetic with the rounding written out: every product, sum and fused multiply-add below is the instruction it
// names (no compiler contraction), so the element vectors are bit-identical in every inlined copy of grid_element --
// the gradients do not depend on the chunk height or on which code path (recomputed row or owned row) evaluated an
// element.
__device__ __forceinline__ double op_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float op_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double op_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float op_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double op_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float op_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double op_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float op_fma(float a, float b, float c) { return fmaf(a, b, c); }
// a x + b y
template <class T>
__device__ __forceinline__ T lin2(T a, T x, T b, T y) {
  return op_fma(a, x, op_mul(b, y));
}

// 1 + beta T^c (thermal.py:34); integer powers by repeated multiplication like lax.integer_pow
template <class T, int NL>
__device__ __forceinline__ T grid_conductivity(T tg, T beta, T c) {
  if constexpr (NL == 0) return (T)1;
  else if constexpr (NL == 1) return op_fma(beta, tg, (T)1);
  else if constexpr (NL == 2) return op_fma(beta, op_mul(tg, tg), (T)1);
  else if constexpr (NL == 3) return op_fma(beta, op_mul(tg, op_mul(tg, tg)), (T)1);
  else if constexpr (NL == 4) { const T t2 = op_mul(tg, tg); return op_fma(beta, op_mul(t2, t2), (T)1); }
  else return op_fma(beta, pow_c<T>(tg, c), (T)1);
}

// Element vectors re = dE/dT_e and dK = dE/dK_e of the thermal Quad4 with the 2 x 2 rule on a parallelogram, local
// nodes 0 (-,-), 1 (+,-), 2 (+,+), 3 (-,+) and Gauss points in the same order (quadrilateral_2d_4.py:54-58), written
// through the 1-D Lagrange weights a = (1 - s)/2, b = (1 + s)/2 at the abscissae -+s, s = 1/sqrt(3):
//   values on the bottom / top edge at xi = -+s, then at the four points; dT/dxi depends on eta only, dT/deta on xi only
//   (so do their products with J^-1 on an axis-aligned grid); the weighted fluxes go back to the nodes through the same
//   weights.  Same sums as thermal_vectors_affine (energy_qt.cuh) in another association: equal to rounding.
// e_el = T_e . re (thermal.py:45-49).  124 instructions on an axis-aligned grid (DIAG), 148 on a sheared one.
template <class T, int NL, bool DIAG, bool GK>
__device__ __forceinline__ void grid_element(const T (&Tn)[4], const T (&Kn)[4], const T (&ji)[4], T wd, T beta, T cexp,
                                             T (&re)[4], T (&dK)[4], T& e_el) {
  constexpr double s = FOL_S3;
  const T a = (T)(0.5 * (1.0 - s)), b = (T)(0.5 * (1.0 + s)), ah = (T)(0.25 * (1.0 - s)), bh = (T)(0.25 * (1.0 + s));
  const T Bm = lin2(b, Tn[0], a, Tn[1]), Bp = lin2(a, Tn[0], b, Tn[1]);     // T on the bottom edge at xi = -s, +s
  const T Um = lin2(b, Tn[3], a, Tn[2]), Up = lin2(a, Tn[3], b, Tn[2]);     // ... on the top edge
  const T dB = op_sub(Tn[1], Tn[0]), dU = op_sub(Tn[2], Tn[3]);
  const T t0e[2] = {lin2(bh, dB, ah, dU), lin2(ah, dB, bh, dU)};            // dT/dxi at eta = -s, +s
  const T t1x[2] = {op_sub(Um, Bm), op_sub(Up, Bp)};                        // 2 dT/deta at xi = -s, +s
  const T KBm = lin2(b, Kn[0], a, Kn[1]), KBp = lin2(a, Kn[0], b, Kn[1]);
  const T KUm = lin2(b, Kn[3], a, Kn[2]), KUp = lin2(a, Kn[3], b, Kn[2]);
  const T tg[4] = {lin2(b, Bm, a, Um), lin2(b, Bp, a, Up), lin2(a, Bp, b, Up), lin2(a, Bm, b, Um)};
  const T eg[4] = {lin2(b, KBm, a, KUm), lin2(b, KBp, a, KUp), lin2(a, KBp, b, KUp), lin2(a, KBm, b, KUm)};
  constexpr int ETA[4] = {0, 0, 1, 1}, XI[4] = {0, 1, 1, 0};                // eta / xi index of Gauss point g
  T w0[4], w1[4], ck[4];
  if constexpr (DIAG) {
    // grad T = (j00 dT/dxi, j11 dT/deta): two values each; so are the flux factors j00 gx, j11 gy and the squares
    const T j3h = op_mul((T)0.5, ji[3]);
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
      if constexpr (GK) ck[g] = op_mul(wn, op_add(gx2[ETA[g]], gy2[XI[g]]));
      w0[g] = op_mul(cf, fx[ETA[g]]);
      w1[g] = op_mul(cf, fy[XI[g]]);
    }
  } else {
    const T j2h = op_mul((T)0.5, ji[2]), j3h = op_mul((T)0.5, ji[3]);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const T gx = lin2(t0e[ETA[g]], ji[0], t1x[XI[g]], j2h);                // grad T = J^-T (dN^T T)
      const T gy = lin2(t0e[ETA[g]], ji[1], t1x[XI[g]], j3h);
      const T wn = op_mul(wd, grid_conductivity<T, NL>(tg[g], beta, cexp));
      const T cf = op_mul(wn, eg[g]);
      if constexpr (GK) ck[g] = op_mul(wn, lin2(gx, gx, gy, gy));
      w0[g] = op_mul(cf, lin2(ji[0], gx, ji[1], gy));
      w1[g] = op_mul(cf, lin2(ji[2], gx, ji[3], gy));
    }
  }
  const T s0lo = op_add(w0[0], w0[1]), s0hi = op_add(w0[2], w0[3]);
  const T s1l = op_add(w1[0], w1[3]), s1r = op_add(w1[1], w1[2]);
  const T S0b = lin2(bh, s0lo, ah, s0hi), S0t = lin2(ah, s0lo, bh, s0hi);  // sum_g dN/dxi weights, bottom / top nodes
  const T S1l = lin2(bh, s1l, ah, s1r), S1r = lin2(ah, s1l, bh, s1r);      // sum_g dN/deta weights, left / right nodes
  re[0] = -op_add(S0b, S1l);
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
    dK[0] = dK[1] = dK[2] = dK[3] = (T)0;
  }
  e_el = op_fma(Tn[3], re[3], op_fma(Tn[2], re[2], op_fma(Tn[1], re[1], op_mul(Tn[0], re[0]))));
}

}  // namespace

// One row of `n` values starting at element `first` of an array of `total` values (base 16-byte aligned) goes into a
// ring row as ONE bulk copy from the 16-byte block that holds the first value to the last WHOLE block of the array;
// the (at most 16 / sizeof(T) - 1) values of a partial last block of the array follow by plain stores.  The lanes read
// value j of the row at dst[shift + j], shift = first % (16 / sizeof(T)).
struct RowCopy {
  long long a0;          // first value of the first block
  long long tail_beg, tail_end;   // values copied by plain stores
  uint32_t bytes;        // bulk bytes
};
template <class T>
__device__ __forceinline__ RowCopy plan_row(long long total, long long first, int n) {
  constexpr long long PER = 16 / sizeof(T);
  RowCopy r;
  r.a0 = first & ~(PER - 1);
  long long a1 = (first + n + PER - 1) & ~(PER - 1);                         // one past the last block
  const long long whole = total & ~(PER - 1);                                // one past the last whole block of the array
  r.tail_beg = r.tail_end = 0;
  if (a1 > whole) {
    r.tail_beg = whole > first ? whole : first;
    r.tail_end = first + n;
    a1 = whole;
  }
  r.bytes = a1 > r.a0 ? (uint32_t)((a1 - r.a0) * sizeof(T)) : 0u;
  return r;
}
template <class T>
__device__ __forceinline__ void copy_tail(const RowCopy& r, const T* base, T* dst) {
  for (long long j = r.tail_beg; j < r.tail_end; ++j) dst[j - r.a0] = base[j];
}
template <class T>
__device__ __forceinline__ void copy_bulk(const RowCopy& r, const T* base, T* dst, uint32_t bar) {
  if (r.bytes) bulk_g2s(smem_u32(dst), base + r.a0, r.bytes, bar);
}

template <class T, int NL, bool DIAG, bool GK>
// 9 warps per CTA (8 consumers + the producer).  Registers are per SCHEDULER (16 K each): 96 registers let a scheduler
// host 5 float64 warps (two CTAs = 18 warps per SM), 72 registers 7 float32 warps (three CTAs = 27 warps)
__global__ void __launch_bounds__(288) __maxnreg__(sizeof(T) == 8 ? 96 : 72) energy_grid_kernel(const GridArgs<T> args) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Pair = NodePair<T>;
  constexpr int PER = 16 / (int)sizeof(T);
  const int W = args.W, tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  // shared memory: ring [kRing][3 rows: T, K, D][row_bytes] | left [W][rows] | right [W][rows] | barriers
  unsigned char* const ring = smem_raw;
  const int slot_bytes = 3 * args.row_bytes;
  Pair* const left = reinterpret_cast<Pair*>(smem_raw + (size_t)kRing * slot_bytes);
  Pair* const right = left + (size_t)W * args.rows;
  unsigned long long* const bars = reinterpret_cast<unsigned long long*>(right + (size_t)W * args.rows);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kRing);

  // item = (sample, chunk of node rows, panel of element columns)
  long long item = blockIdx.x;
  const int panel = (int)(item % args.npanels);
  item /= args.npanels;
  const int chunk = (int)(item % args.nchunks);
  const long long smp = item / args.nchunks;
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
        const long long first = smp * args.nn + (long long)(e_beg + i) * NXn + sp;
        const long long first_d = (long long)(e_beg + i) * NXn + sp;
        T* const dT = reinterpret_cast<T*>(ring + (size_t)slot * slot_bytes);
        T* const dK = reinterpret_cast<T*>(ring + (size_t)slot * slot_bytes + args.row_bytes);
        T* const dD = reinterpret_cast<T*>(ring + (size_t)slot * slot_bytes + 2 * args.row_bytes);
        const uint32_t bar = full0 + 8 * slot;
        const RowCopy cu = plan_row<T>(total, first, ncols);           // u and ctrl: same shape, same offsets
        const RowCopy cd = plan_row<T>(args.nn, first_d, ncols);
        // plain tail stores (the last row of the last sample only) first, then the arrive that publishes them and
        // arms the transaction count, then the bulk copies that complete it
        copy_tail<T>(cu, args.u, dT);
        copy_tail<T>(cu, args.ctrl, dK);
        if (has_dirv) copy_tail<T>(cd, args.dir_values, dD);
        mbar_expect_tx(bar, 2 * cu.bytes + (has_dirv ? cd.bytes : 0u));
        copy_bulk<T>(cu, args.u, dT, bar);
        copy_bulk<T>(cu, args.ctrl, dK, bar);
        if (has_dirv) copy_bulk<T>(cd, args.dir_values, dD, bar);
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

    const T ji[4] = {args.jinv[0], args.jinv[1], args.jinv[2], args.jinv[3]};
    const T wd = args.wd, beta = args.beta, cexp = args.cexp, scale = args.out_scale;
    // shared-memory addresses of this lane's entries in ring slot 0 (T row; the K and D rows follow at row_bytes)
    const uint32_t ring_bytes = (uint32_t)(kRing * slot_bytes);
    const uint32_t a_own = smem_u32(ring) + (uint32_t)i_own * (uint32_t)sizeof(T);
    uint32_t slot_off = 0, parity = 0, slot_bar = 0;         // ring position of the next row to take
    uint32_t sh_u = (uint32_t)((smp * args.nn + (long long)e_beg * NXn + sp) & (PER - 1)) * (uint32_t)sizeof(T);
    uint32_t sh_d = (uint32_t)(((long long)e_beg * NXn + sp) & (PER - 1)) * (uint32_t)sizeof(T);
    const uint32_t sh_step = (uint32_t)(NXn & (PER - 1)) * (uint32_t)sizeof(T);
    // global element offset of node (cc, row) within the sample, as 32 bits (nn < 2^31 is checked by the host)
    unsigned node = (unsigned)e_beg * (unsigned)NXn + (unsigned)cc;
    T* const gu0 = args.grad_u + smp * args.nn;
    T* const gk0 = has_gk ? args.grad_k + smp * args.nn : gu0;
    const uint8_t* const fl0 = wcut ? args.dir_flag : args.col_dir;   // never read unless wcut
    uint32_t a_left = smem_u32(left + (size_t)w * args.rows), a_right = smem_u32(right + (size_t)w * args.rows);

    auto lds = [](uint32_t a) {
      T v;
      if constexpr (sizeof(T) == 8) asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(a));
      else asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(a));
      return v;
    };
    auto sts_pair = [](uint32_t a, T x, T y) {
      if constexpr (sizeof(T) == 8) asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(a), "d"(x), "d"(y) : "memory");
      else asm volatile("st.shared.v2.f32 [%0], {%1, %2};\n" ::"r"(a), "f"(x), "f"(y) : "memory");
    };
    // the four values of the next staged node row: own / right column of T and K
    auto take = [&](T& t0, T& k0, T& t1, T& k1) {
      mbar_wait(full0 + slot_bar, parity);
      const uint32_t aT = a_own + slot_off + sh_u, aK = aT + (uint32_t)args.row_bytes;
      t0 = lds(aT);
      t1 = lds(aT + (uint32_t)sizeof(T));
      k0 = lds(aK);
      k1 = lds(aK + (uint32_t)sizeof(T));
      if (wdirv) {                                           // Dirichlet overwrite (fe_loss.py:91-92, 255)
        const uint32_t aD = a_own + slot_off + 2u * (uint32_t)args.row_bytes + sh_d;
        const T d0 = lds(aD), d1 = lds(aD + (uint32_t)sizeof(T));
        t0 = (d0 == d0) ? d0 : t0;
        t1 = (d1 == d1) ? d1 : t1;
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
      sh_u = (sh_u + sh_step) & 15u;
      sh_d = (sh_d + sh_step) & 15u;
    };

    T en = (T)0;
    // node row (at offset `node`) is complete once the element rows below and above it are in: left half (own lane:
    // below + above) + right half of the lane to the left; the column two warps share waits for `combine`
    uint8_t cut_now = 0;                                     // dir_flag of node (cc, current row), loaded one row ahead
    if (wcut) cut_now = fl0[node + (e_beg < r0 ? (unsigned)NXn : 0u)];
    auto finish_row = [&](T leftR, T leftK, T rpR, T rpK, bool more) {
      const T inR = shfl_up1(rpR), inK = shfl_up1(rpK);
      uint8_t cut_next = 0;
      if (wcut && more) cut_next = fl0[node + (unsigned)NXn];
      if (keep_left) sts_pair(a_left, leftR, leftK);
      if (keep_right) sts_pair(a_right, rpR, rpK);
      a_left += 2 * (uint32_t)sizeof(T);
      a_right += 2 * (uint32_t)sizeof(T);
      T R = (l > 0) ? op_add(leftR, inR) : leftR;
      const T K = (l > 0) ? op_add(leftK, inK) : leftK;
      if (wcut && cut_now != 0) R = (T)0;
      if (write_own) {
        gu0[node] = op_mul(scale, R);
        if (has_gk) gk0[node] = op_mul(scale, K);
      }
      cut_now = cut_next;
    };
    // one element row: corners (b0, b1) below, (t0, t1) above, carries of the row below in (oc, rc)
    T ocR = (T)0, ocK = (T)0, rcR = (T)0, rcK = (T)0;
    auto element = [&](T Tb0, T Kb0, T Tb1, T Kb1, T Tt0, T Kt0, T Tt1, T Kt1, T (&re)[4], T (&dK)[4], T& e_el) {
      const T Tn[4] = {Tb0, Tb1, Tt1, Tt0}, Kn[4] = {Kb0, Kb1, Kt1, Kt0};
      grid_element<T, NL, DIAG, GK>(Tn, Kn, ji, wd, beta, cexp, re, dK, e_el);
      if (!all_valid) {                                      // warp-uniform: a ragged last warp only
        if (!el_valid) {
#pragma unroll
          for (int q = 0; q < 4; ++q) re[q] = dK[q] = (T)0;
          e_el = (T)0;
        }
      }
    };
    auto step = [&](T Tb0, T Kb0, T Tb1, T Kb1, T Tt0, T Kt0, T Tt1, T Kt1) {
      T re[4], dK[4], e_el;
      element(Tb0, Kb0, Tb1, Kb1, Tt0, Kt0, Tt1, Kt1, re, dK, e_el);
      en = op_add(en, e_el);
      finish_row(op_add(ocR, re[0]), op_add(ocK, dK[0]), op_add(rcR, re[1]), op_add(rcK, dK[1]), true);
      node += (unsigned)NXn;
      ocR = re[3]; ocK = dK[3]; rcR = re[2]; rcK = dK[2];
    };

    T A0, AK0, A1, AK1, B0, BK0, B1, BK1;                    // two register sets of corner values, used alternately
    take(A0, AK0, A1, AK1);
    int e = e_beg;
    if (e_beg < r0) {
      // the recomputed element row below the chunk: only its shares of node row r0 (the carries) are kept
      take(B0, BK0, B1, BK1);
      T re[4], dK[4], e_el;
      element(A0, AK0, A1, AK1, B0, BK0, B1, BK1, re, dK, e_el);
      ocR = re[3]; ocK = dK[3]; rcR = re[2]; rcK = dK[2];
      node += (unsigned)NXn;
      A0 = B0; AK0 = BK0; A1 = B1; AK1 = BK1;
      ++e;
    }
    for (; e + 1 < e_end; e += 2) {
      take(B0, BK0, B1, BK1);
      step(A0, AK0, A1, AK1, B0, BK0, B1, BK1);
      take(A0, AK0, A1, AK1);
      step(B0, BK0, B1, BK1, A0, AK0, A1, AK1);
    }
    if (e < e_end) {
      take(B0, BK0, B1, BK1);
      step(A0, AK0, A1, AK1, B0, BK0, B1, BK1);
    }
    if (r1 == ny + 1) finish_row(ocR, ocK, rcR, rcK, false);  // the top node row of the grid closes with the carries alone

    // energy share of this warp
    if (!count) en = (T)0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) en += __shfl_xor_sync(0xffffffffu, en, o);
    if (l == 0) args.partial[smp * args.npart + ((long long)panel * args.nchunks + chunk) * W + w] = en;
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
    T R = lo.t, K = lo.k;
    if (bnd < W) {
      const Pair hi = left[(size_t)bnd * args.rows + i];
      R = op_add(hi.t, R);                                   // same order as finish_row: own (left) half + incoming
      K = op_add(hi.k, K);
    }
    const long long node = (long long)(r0 + i) * NXn + col;
    const bool cut = args.dir_flag ? args.dir_flag[node] != 0 : false;
    args.grad_u[smp * args.nn + node] = cut ? (T)0 : op_mul(args.out_scale, R);
    if (GK && args.grad_k) args.grad_k[smp * args.nn + node] = op_mul(args.out_scale, K);
  }
}

namespace {

struct GridShape {
  int W, npanels;
};
inline GridShape grid_shape(long long nx) {
  GridShape g;
  g.W = (int)(nx <= 256 ? cdiv(nx, 32) : 8);
  g.npanels = (int)(nx <= 256 ? 1 : cdiv(nx - 1, 32 * g.W - 1));
  return g;
}
constexpr int kMinRows = 8;

template <class T>
int grid_row_bytes(int W, long long nx) {
  const long long ncols = (32LL * W + 1 < nx + 1) ? 32LL * W + 1 : nx + 1;
  return (int)(((ncols + 16 / sizeof(T)) * sizeof(T) + 15) / 16 * 16);   // + shift + one entry read past the last column
}
template <class T>
size_t grid_smem(int W, long long nx, int rows) {
  return (size_t)kRing * 3 * grid_row_bytes<T>(W, nx) + (size_t)2 * W * rows * 2 * sizeof(T) + 2 * kRing * 8;
}

template <class T, int NL, bool DIAG, bool GK>
int launch_grid(cudaStream_t s, GridArgs<T> a, T* energy) {
  auto kern = energy_grid_kernel<T, NL, DIAG, GK>;
  const GridShape g = grid_shape(a.nx);
  const int threads = 32 * (g.W + 1);
  static PerDeviceOnce configured;
  if (configured.need()) {
    FOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grid_smem<T>(8, 256, 128)));
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
  for (int rows = kMinRows; rows <= 128; ++rows) {
    if (rows > nrows_total && rows != kMinRows) break;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, grid_smem<T>(g.W, a.nx, rows)) != cudaSuccess ||
        per_sm < 1)
      continue;
    const long long nchunks = cdiv(nrows_total, rows);
    const long long items = nchunks * g.npanels * a.nb, slots = (long long)sms * per_sm;
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
  const long long items = (long long)a.nchunks * a.npanels * a.nb;
  FOL_REQUIRE(items < (1LL << 31), "fol_energy_and_grads_grid: too many work items for one launch");
  kern<<<(unsigned)items, threads, grid_smem<T>(g.W, a.nx, best_rows), s>>>(a);
  int rc = check_launch("energy_grid_kernel");
  if (rc) return rc;
  energy_sum_kernel<T><<<(unsigned)cdiv(a.nb, 8), 256, 0, s>>>(a.partial, a.nb, a.npart, energy);
  return check_launch("energy_sum_kernel");
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
    if (diag) return gk ? launch_grid<T, NLV, true, true>(s, a, energy) : launch_grid<T, NLV, true, false>(s, a, energy);   \
    return gk ? launch_grid<T, NLV, false, true>(s, a, energy) : launch_grid<T, NLV, false, false>(s, a, energy);           \
  }
  FOL_GRID(0) FOL_GRID(1) FOL_GRID(2) FOL_GRID(3) FOL_GRID(4) FOL_GRID(-1)
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
