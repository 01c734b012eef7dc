// Batched thermal loss + VJP on a STRUCTURED Quad4 grid (BASELINE.json configs[2]: 256 x 256 quads, 1024 samples):
// ThermalLoss2DQuad.ComputeBatchLoss and its JAX-AD gradient (thermal.py:28-49, fe_loss.py:250-262) for meshes whose
// nodes are numbered row-major (node(c, r) = r (nx + 1) + c), whose elements are [n, n + 1, n + nx + 2, n + nx + 1]
// (fol/tools/usefull_functions.py:213-258 builds exactly that) and whose elements are all the same parallelogram, so
// that J^-1 and w detJ are two launch constants.  The host plan establishes those facts (energy_plan.grid_structure);
// any other mesh keeps the tile kernels (energy_qt.cuh / energy2.cuh), which gather through the connectivity.
//
// What the structure buys: no connectivity, no geometry cache, no adjacency lists and no CTA barrier per pass.
//   * A warp marches UP the grid: lane l owns node column c0 + l of ONE sample and, per step, the element between rows
//     e, e + 1 and columns c, c + 1.  The node rows are streamed by cp.async into a per-warp ring (kDepth rows ahead,
//     one 256-byte line per array and row), so a node value is read from HBM once per chunk and the element takes its
//     four corners from two 16-byte shared-memory loads; the bottom corners are the previous step's top corners.
//   * The four element-vector entries leave through registers: the two left corners accumulate in the lane, the two
//     right corners are summed per lane over consecutive rows and handed to lane l + 1 by one shuffle -- every node sum
//     has the fixed order ((below-left + above-left... see `march`), no atomics, deterministic.
//   * The element arithmetic is sum-factorised for the bilinear element with the 2 x 2 rule (~150 FP instructions per
//     element instead of ~190; `grid_element`).
//   * Warp boundaries: lane 31 also stages the column right of it (33 entries per row); the node column shared by two
//     warps gets its two halves through shared memory once per chunk (`combine`), the only CTA barrier of the kernel.
//   * A CTA (W <= 8 warps = one panel of 32 W element columns) owns `rows` node rows of one sample and recomputes the
//     element row below them (1 / rows extra work) instead of exchanging partial sums with the chunk below.
// Roofline: FP64 pipe in float64 (~300 flops per element), issue slots in float32.
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
  const T* dir_values;        // (nn) NaN where free, or null: overwrites u while staging (fe_loss.py:91-92, 255)
  const uint8_t* dir_flag;    // (nn) 1 where grad_u is written as zero, or null
  T out_scale, beta, cexp, wd;
  T jinv[4];                  // row-major d xi_j / d x_k of the one element shape
  int nx, ny;                 // elements per direction
  int rows, nchunks, npanels, npart;
  long long nn, nb;
};

namespace {

constexpr int kDepth = 4;     // node rows in flight per warp (a power of two)
constexpr int kEnt = 34;      // staged entries per row: the warp's 32 columns + the one right of lane 31 (+ padding)

template <class T>
struct alignas(2 * sizeof(T)) NodePair {
  T t, k;
};

template <class T>
__device__ __forceinline__ T shfl_up1(T v) {
  return __shfl_up_sync(0xffffffffu, v, 1);
}

// Element vectors re = dE/dT_e and dK = dE/dK_e of the thermal Quad4 with the 2 x 2 rule on a parallelogram, local
// nodes 0 (-,-), 1 (+,-), 2 (+,+), 3 (-,+) and Gauss points in the same order (quadrilateral_2d_4.py:54-58), written
// through the 1-D Lagrange weights a = (1 - s)/2, b = (1 + s)/2 at the abscissae -+s, s = 1/sqrt(3):
//   values on the bottom / top edge at xi = -+s, then at the four points; dT/dxi depends on eta only, dT/deta on xi only;
//   the weighted fluxes go back to the nodes through the same weights.  Same sums as thermal_vectors_affine
//   (energy_qt.cuh) in another association: equal to rounding.
template <class T, int NL, bool DIAG, bool GK>
__device__ __forceinline__ void grid_element(const T (&Tn)[4], const T (&Kn)[4], const T (&ji)[4], T wd, T beta, T cexp,
                                             T (&re)[4], T (&dK)[4]) {
  constexpr double s = FOL_S3;
  const T a = (T)(0.5 * (1.0 - s)), b = (T)(0.5 * (1.0 + s)), ah = (T)(0.25 * (1.0 - s)), bh = (T)(0.25 * (1.0 + s));
  const T Bm = b * Tn[0] + a * Tn[1], Bp = a * Tn[0] + b * Tn[1];      // T on the bottom edge at xi = -s, +s
  const T Um = b * Tn[3] + a * Tn[2], Up = a * Tn[3] + b * Tn[2];      // ... on the top edge
  const T dB = Tn[1] - Tn[0], dU = Tn[2] - Tn[3];
  const T t0m = bh * dB + ah * dU, t0p = ah * dB + bh * dU;            // dT/dxi at eta = -s, +s
  const T t1m = (T)0.5 * (Um - Bm), t1p = (T)0.5 * (Up - Bp);          // dT/deta at xi = -s, +s
  const T KBm = b * Kn[0] + a * Kn[1], KBp = a * Kn[0] + b * Kn[1];
  const T KUm = b * Kn[3] + a * Kn[2], KUp = a * Kn[3] + b * Kn[2];
  const T tg[4] = {b * Bm + a * Um, b * Bp + a * Up, a * Bp + b * Up, a * Bm + b * Um};
  const T eg[4] = {b * KBm + a * KUm, b * KBp + a * KUp, a * KBp + b * KUp, a * KBm + b * KUm};
  const T t0[4] = {t0m, t0m, t0p, t0p}, t1[4] = {t1m, t1p, t1p, t1m};
  T w0[4], w1[4], ck[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    T gx, gy;
    if constexpr (DIAG) {
      gx = t0[g] * ji[0];
      gy = t1[g] * ji[3];
    } else {
      gx = t0[g] * ji[0] + t1[g] * ji[2];                              // grad T = J^-T (dN^T T)
      gy = t0[g] * ji[1] + t1[g] * ji[3];
    }
    const T wn = wd * conductivity_factor<T, NL>(tg[g], beta, cexp);
    const T cf = wn * eg[g];
    if constexpr (GK) ck[g] = wn * (gx * gx + gy * gy);
    if constexpr (DIAG) {
      w0[g] = cf * (ji[0] * gx);
      w1[g] = cf * (ji[3] * gy);
    } else {
      w0[g] = cf * (ji[0] * gx + ji[1] * gy);
      w1[g] = cf * (ji[2] * gx + ji[3] * gy);
    }
  }
  const T s0lo = w0[0] + w0[1], s0hi = w0[2] + w0[3], s1l = w1[0] + w1[3], s1r = w1[1] + w1[2];
  const T S0b = bh * s0lo + ah * s0hi, S0t = ah * s0lo + bh * s0hi;    // sum_g dN/dxi weights, bottom / top nodes
  const T S1l = bh * s1l + ah * s1r, S1r = ah * s1l + bh * s1r;        // sum_g dN/deta weights, left / right nodes
  re[0] = -(S0b + S1l);
  re[1] = S0b - S1r;
  re[2] = S0t + S1r;
  re[3] = S1l - S0t;
  if constexpr (GK) {
    const T Cml = b * ck[0] + a * ck[1], Cmr = a * ck[0] + b * ck[1];  // eta = -s row reduced to the left / right nodes
    const T Cpl = b * ck[3] + a * ck[2], Cpr = a * ck[3] + b * ck[2];  // eta = +s row
    dK[0] = b * Cml + a * Cpl;
    dK[1] = b * Cmr + a * Cpr;
    dK[2] = a * Cmr + b * Cpr;
    dK[3] = a * Cml + b * Cpl;
  } else {
    dK[0] = dK[1] = dK[2] = dK[3] = (T)0;
  }
}

}  // namespace

template <class T, int NL, bool DIAG, bool GK>
__global__ void __launch_bounds__(256, 2) energy_grid_kernel(const GridArgs<T> args) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int W = blockDim.x >> 5, tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  using Pair = NodePair<T>;
  Pair* ring = reinterpret_cast<Pair*>(smem_raw) + (size_t)w * kDepth * kEnt;             // [W][kDepth][kEnt]
  T* ringd = reinterpret_cast<T*>(reinterpret_cast<Pair*>(smem_raw) + (size_t)W * kDepth * kEnt) +
             (size_t)w * kDepth * kEnt;                                                    // [W][kDepth][kEnt]
  Pair* left = reinterpret_cast<Pair*>(reinterpret_cast<T*>(reinterpret_cast<Pair*>(smem_raw) + (size_t)W * kDepth * kEnt) +
                                       (size_t)W * kDepth * kEnt);                         // [W][rows]: lane 0's half
  Pair* right = left + (size_t)W * args.rows;                                              // [W][rows]: lane 31's right column

  // item = (sample, chunk of node rows, panel of element columns)
  long long item = blockIdx.x;
  const int panel = (int)(item % args.npanels);
  item /= args.npanels;
  const int chunk = (int)(item % args.nchunks);
  const long long smp = item / args.nchunks;
  const int nx = args.nx, ny = args.ny, NXn = nx + 1;
  const int sp = panel * (32 * W - 1);                       // first element column of the panel
  const int c = sp + 32 * w + l;                             // own node column = left column of this lane's element
  const int r0 = chunk * args.rows, r1 = min(r0 + args.rows, ny + 1);   // owned node rows [r0, r1)
  const int e_beg = max(r0 - 1, 0), e_end = min(r1, ny);     // element rows [e_beg, e_end): one recomputed row below
  const bool el_valid = c < nx;
  const bool write_own = c <= nx && (l > 0 || w == 0) && (c > sp || panel == 0);   // lane 0 of warps >= 1: `combine`
  const bool count = el_valid && (c > sp || panel == 0);     // the panel's first element column belongs to the panel left of it
  const int cc = min(c, nx), c32 = min(sp + 32 * w + 32, nx);
  const T* const u_col = args.u + smp * args.nn + cc;
  const T* const k_col = args.ctrl + smp * args.nn + cc;
  const bool has_dir = args.dir_values != nullptr;

  auto stage = [&](int rr) {                                 // node row rr -> ring slot; always one commit group
    if (rr <= e_end) {
      const int slot = (rr - e_beg) & (kDepth - 1);
      Pair* dst = ring + slot * kEnt;
      const long long ro = (long long)rr * NXn;
      cp_async_elem<T>(&dst[l].t, u_col + ro);
      cp_async_elem<T>(&dst[l].k, k_col + ro);
      if (has_dir) cp_async_elem<T>(ringd + slot * kEnt + l, args.dir_values + ro + cc);
      if (l == 0) {
        cp_async_elem<T>(&dst[32].t, u_col + ro + (c32 - cc));
        cp_async_elem<T>(&dst[32].k, k_col + ro + (c32 - cc));
        if (has_dir) cp_async_elem<T>(ringd + slot * kEnt + 32, args.dir_values + ro + c32);
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  // own and right-neighbour values of node row rr once it has landed; then the slot is refilled kDepth rows ahead
  auto take = [&](int rr, T& t0, T& k0, T& t1, T& k1) {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(kDepth - 1) : "memory");
    const int slot = (rr - e_beg) & (kDepth - 1);
    Pair* src = ring + slot * kEnt;
    if (has_dir) {                                           // Dirichlet overwrite of the entries this lane staged
      const T dv = ringd[slot * kEnt + l];
      if (dv == dv) src[l].t = dv;
      if (l == 0) {
        const T d32 = ringd[slot * kEnt + 32];
        if (d32 == d32) src[32].t = d32;
      }
    }
    __syncwarp();
    const Pair p0 = src[l], p1 = src[l + 1];
    t0 = p0.t; k0 = p0.k; t1 = p1.t; k1 = p1.k;
    __syncwarp();
    stage(rr + kDepth);
  };

#pragma unroll
  for (int d = 0; d < kDepth; ++d) stage(e_beg + d);

  const T ji[4] = {args.jinv[0], args.jinv[1], args.jinv[2], args.jinv[3]};
  T Tb0, Kb0, Tb1, Kb1;                                      // bottom corners: own column, right column
  take(e_beg, Tb0, Kb0, Tb1, Kb1);
  T ocR = (T)0, ocK = (T)0;                                  // own column, current row: share of the element row below
  T rcR = (T)0, rcK = (T)0;                                  // right column, current row: share of the element row below
  T en = (T)0;
  T* const gu = args.grad_u + smp * args.nn + cc;
  T* const gk = (GK && args.grad_k) ? args.grad_k + smp * args.nn + cc : nullptr;

  // node row `row` is complete once the element rows below and above it are in: left half (own lane) + right half of
  // the lane to the left.  Fixed order per node: ((below-left + above-left) + (below-right + above-right)) in terms of
  // the elements around it -- wait for `combine` on the column two warps share.
  auto finish_row = [&](int row, T leftR, T leftK, T rpR, T rpK, bool cut) {
    const T inR = shfl_up1(rpR), inK = shfl_up1(rpK);
    if (row < r0) return;                                    // the recomputed row below the chunk (warp-uniform)
    const int i = row - r0;
    if (l == 0 && w > 0) left[(size_t)w * args.rows + i] = Pair{leftR, leftK};
    if (l == 31) right[(size_t)w * args.rows + i] = Pair{rpR, rpK};
    if (write_own) {
      const T R = (l > 0) ? leftR + inR : leftR, K = (l > 0) ? leftK + inK : leftK;
      const long long ro = (long long)row * NXn;
      gu[ro] = cut ? (T)0 : args.out_scale * R;
      if (GK && gk) gk[ro] = args.out_scale * K;
    }
  };

  for (int e = e_beg; e < e_end; ++e) {
    const bool cut = args.dir_flag ? args.dir_flag[(long long)e * NXn + cc] != 0 : false;   // of node (c, e)
    T Tt0, Kt0, Tt1, Kt1;
    take(e + 1, Tt0, Kt0, Tt1, Kt1);
    const T Tn[4] = {Tb0, Tb1, Tt1, Tt0}, Kn[4] = {Kb0, Kb1, Kt1, Kt0};
    T re[4], dK[4];
    grid_element<T, NL, DIAG, GK>(Tn, Kn, ji, args.wd, args.beta, args.cexp, re, dK);
    if (!el_valid) {
#pragma unroll
      for (int q = 0; q < 4; ++q) re[q] = dK[q] = (T)0;
    }
    if (count && e >= r0) en += (Tn[0] * re[0] + Tn[1] * re[1]) + (Tn[2] * re[2] + Tn[3] * re[3]);   // thermal.py:45-49
    finish_row(e, ocR + re[0], ocK + dK[0], rcR + re[1], rcK + dK[1], cut);
    ocR = re[3]; ocK = dK[3]; rcR = re[2]; rcK = dK[2];
    Tb0 = Tt0; Kb0 = Kt0; Tb1 = Tt1; Kb1 = Kt1;
  }
  if (r1 == ny + 1) {                                        // the top node row of the grid closes with the carries alone
    const bool cut = args.dir_flag ? args.dir_flag[(long long)ny * NXn + cc] != 0 : false;
    finish_row(ny, ocR, ocK, rcR, rcK, cut);
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");

  // energy share of this warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) en += __shfl_xor_sync(0xffffffffu, en, o);
  if (l == 0) args.partial[smp * args.npart + ((long long)panel * args.nchunks + chunk) * W + w] = en;

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
      R = hi.t + R;                                          // same order as finish_row: own (left) half + incoming
      K = hi.k + K;
    }
    const long long node = (long long)(r0 + i) * NXn + col;
    const bool cut = args.dir_flag ? args.dir_flag[node] != 0 : false;
    args.grad_u[smp * args.nn + node] = cut ? (T)0 : args.out_scale * R;
    if (GK && args.grad_k) args.grad_k[smp * args.nn + node] = args.out_scale * K;
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
size_t grid_smem(int W, int rows) {
  return (size_t)W * kDepth * kEnt * 3 * sizeof(T) + (size_t)2 * W * rows * 2 * sizeof(T);
}

template <class T, int NL, bool DIAG, bool GK>
int launch_grid(cudaStream_t s, GridArgs<T> a, T* energy) {
  auto kern = energy_grid_kernel<T, NL, DIAG, GK>;
  const GridShape g = grid_shape(a.nx);
  static PerDeviceOnce configured;
  if (configured.need()) {
    FOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grid_smem<T>(8, 128)));
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
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * g.W, grid_smem<T>(g.W, rows)) != cudaSuccess ||
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
  a.npart = a.nchunks * a.npanels * g.W;
  const long long items = (long long)a.nchunks * a.npanels * a.nb;
  FOL_REQUIRE(items < (1LL << 31), "fol_energy_and_grads_grid: too many work items for one launch");
  kern<<<(unsigned)items, 32 * g.W, grid_smem<T>(g.W, best_rows), s>>>(a);
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
                              const uint8_t* dir_flag, double out_scale, const double* params_host, void* grad_u,
                              void* grad_k, void* energy, void* work) {
  FOL_REQUIRE(nx >= 1 && ny >= 1 && nb >= 0 && nx < (1 << 24) && ny < (1 << 24), "fol_energy_and_grads_grid: bad grid size");
  FOL_REQUIRE(jinv_host && params_host && ctrl && u && grad_u && energy && work, "fol_energy_and_grads_grid: null pointer");
  FOL_REQUIRE(dtype == FOL_F64 || dtype == FOL_F32, "fol_energy_and_grads_grid: unknown dtype");
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
    a.out_scale = (T)out_scale;
    a.beta = (T)params_host[5];
    a.cexp = (T)params_host[6];
    a.wd = (T)w_detj;
    for (int i = 0; i < 4; ++i) a.jinv[i] = (T)jinv_host[i];
    a.nx = (int)nx;
    a.ny = (int)ny;
    a.nn = (nx + 1) * (ny + 1);
    a.nb = nb;
    a.rows = a.nchunks = a.npanels = a.npart = 0;
    return dispatch_grid<T>((cudaStream_t)s, a, (T*)energy);
  };
  if (dtype == FOL_F64) return run((double*)nullptr);
  return run((float*)nullptr);
}

}  // extern "C"
