// Tuned element-stage kernel for the north-star workload: 3-D Hex8 small-strain elasticity, float64,
// 2x2x2 Gauss rule (MechanicalLoss3DHexa, mechanical.py:98-117 + fe_loss.py:191-230, 299).
//
// Same results as the generic kernel (assemble.cuh); different machine mapping:
//   * persistent warps (16 per SM), each iteration owns a tile of 4 consecutive elements;
//   * nodal gathers for tile i+1 (and connectivity for tile i+2) are in flight while tile i computes;
//   * phase 1: lane (element, Gauss point) -> J, det J, grad N, coefficient: 32 independent
//     geometry evaluations per warp, nothing computed twice;
//   * phase 2: per element the 24x24 matrix P = sum_g s_g v_g v_g^T (v = grad N flattened) is a
//     24x24x8 GEMM: 18 DMMA m8n8k4 (FP64 tensor path) fed straight from the staged gradients, the
//     accumulator fragment of lane (a, k) being exactly the 3x3 blocks (a,2k), (a,2k+1);
//     Ke = lam P + mu P^T + mu tr(P) I, re = Ke u - Fe, Dirichlet row mask in registers;
//   * Ke rows are staged in shared memory and leave the SM as ONE contiguous 4608-byte bulk
//     async copy per element (cp.async.bulk shared->global, the TMA engine): full-line writes, no
//     store instructions on the LSU path; the copy of element i drains while element i+1 is computed.
// Ke is symmetric only to rounding here (the scale s_g rides on the A operand of the DMMA), so calls
// with transpose_jacobian=True are served by the generic kernel, which transposes exactly.
#include <cstdlib>

#include "assemble.cuh"
#include "assemble_hex_common.cuh"

namespace fol {

namespace {

#ifndef FOL_HEX_WARPS
#define FOL_HEX_WARPS 6
#endif
#ifndef FOL_HEX_SLOTS
#define FOL_HEX_SLOTS 1
#endif
constexpr int kWarps = FOL_HEX_WARPS;   // warps per CTA, each fully independent (2 CTAs = 12 warps / SM)
constexpr int kSlots = FOL_HEX_SLOTS;   // Ke staging slots per warp
#ifndef FOL_HEX_HALVES
#define FOL_HEX_HALVES 0
#endif
constexpr bool kHalves = FOL_HEX_HALVES != 0;   // release / refill the staging slot in two halves
#ifndef FOL_HEX_SYM
#define FOL_HEX_SYM 1
#endif
constexpr bool kSym = FOL_HEX_SYM != 0;         // tiles of P below the diagonal mirrored by shuffles instead of computed
using hexk::kTile;
using namespace hexk;

// LAYOUT 0: the round-1 layout (one Ke staging slot, (dN/dz, coefficient) pairs, double-buffered X / u / de).
// LAYOUT 1: compact -- dN/dz and the coefficient stored separately (the coefficient is per Gauss point, not per node),
//           X / de single-buffered (phase 1 consumes them; the next gather is issued right after it), byte Dirichlet
//           flags -- which makes room for a SECOND Ke staging slot at the same 12 warps / SM: the bulk copy of element
//           i can take until element i + 2 needs the slot, instead of stalling the warp when the write queue is deep
//           Measured: no gain (profiles/r2/hex_kernel_experiments.md) -- what the stores cost under sustained load is
//           board power, not slot waits.  Kept selectable (FOL_HEX_LAYOUT=0 / 1) for A/B runs.
// LAYOUT 2 (LAYOUT 3, the default, = LAYOUT 2 + a leaner hand-off, below): the compact layout with ONE staging slot: 13.6 KB per warp, i.e. 16 warps per SM (two CTAs of
//           8 warps at 128 registers, no spills) instead of 12.  The kernel's phase 2 is a chain of dependent DMMA /
//           DFMA / shuffle instructions (stall reasons `wait` and `math_pipe_throttle`), so a third more warps per
//           scheduler is what shortens it: same box, alternating processes, 128^3: 2.56 -> 1.92 ms per step right after
//           the warm-up, 2.56 -> 2.10 ms sustained; outputs bit-identical to layout 0 (profiles/r2/hex_layout_ab.jsonl).
template <int LAYOUT>
struct __align__(128) WarpSmemT;

template <>
struct __align__(128) WarpSmemT<0> {
  static constexpr int kStage = kSlots;
  double stage[kSlots][576];         // Ke staging slots (bulk-copy sources)
  // [element][gauss][node ^ swz(gauss)]: (dN/dx, dN/dy) and (dN/dz, w detJ E_g).  The XOR swizzle
  // swz(g) = ((g & 3) << 1) | (g >> 2) makes both the phase-1 stores (lane = row) and the DMMA
  // fragment loads (lane = (node, gauss mod 4)) bank-conflict free without padding.
  double2 gxy[kTile][8][8];
  double2 gzs[kTile][8][8];
  // nodal data of the tile, SoA over the 32 (element, node) lanes: conflict-free cp.async targets;
  // double-buffered (the next tile lands while this one computes)
  double X[2][3][32];
  double u[2][3][32];
  double de[2][32];
  double wd[kTile][8];               // w detJ per Gauss point (body force)
  float bc[kTile][24];               // 1 = free dof, 0 = Dirichlet dof
};

template <int SLOTS>
struct __align__(128) CompactSmem {
  static constexpr int kStage = SLOTS;
  double stage[SLOTS][576];          // Ke staging slots
  double2 gxy[kTile][8][8];          // [element][gauss][node ^ swz(gauss)]: (dN/dx, dN/dy)
  double gz[kTile][8][8];            // [element][gauss][node ^ swz(gauss) ^ 2 (element & 1)]: dN/dz (8-byte accesses are
                                     //   served per half-warp = two elements: the element bit keeps them on distinct banks)
  double coef[kTile * 8];            // w detJ E_g per (element, gauss)
  double X[3][32];                   // single-buffered
  double u[2][3][32];                // double-buffered: phase 2 of this tile still reads it while the next lands
  double de[32];                     // single-buffered
  double wd[kTile][8];
  uint8_t bc[kTile][24];             // 1 = free dof, 0 = Dirichlet dof
};
template <>
struct __align__(128) WarpSmemT<1> : CompactSmem<2> {};
// LAYOUT 2: the compact layout with ONE staging slot, which fits 16 warps per SM (2 CTAs of 8 warps, 128 registers, no
//           spills) instead of 12: a third more warps to cover the DMMA / FP64 dependency stalls of phase 2.
template <>
struct __align__(128) WarpSmemT<2> : CompactSmem<1> {};
// LAYOUT 3: layout 2 with a leaner hand-off to the copy engine (A/B: FOL_HEX_LAYOUT=3): the Dirichlet flags of the
//           lane's three rows are read once per element and reused for the row mask and the residual (no branches
//           around the three residual stores), and every lane executes the bulk-copy wait (no divergent region).
//           Same box, alternating processes: 2.10-2.11 -> 2.07 ms sustained, 1.92 ms right after the warm-up either
//           way, outputs bit-identical (profiles/r2/hex_lean_ab.jsonl).  The default.
template <>
struct __align__(128) WarpSmemT<3> : CompactSmem<1> {};
constexpr int kWarpsDense = 8;
static_assert(2 * (sizeof(WarpSmemT<1>) * kWarps + 1024) <= 227 * 1024, "compact layout: two CTAs per SM must fit");
static_assert(2 * (sizeof(WarpSmemT<2>) * kWarpsDense + 1024) <= 227 * 1024, "dense layout: two CTAs of 8 warps per SM must fit");
template <int LAYOUT>
constexpr int warps_of() { return LAYOUT >= 2 ? kWarpsDense : kWarps; }
template <int LAYOUT>
constexpr int min_ctas_of() { return 2; }   // every layout is sized for two CTAs per SM

}  // namespace

namespace {
__device__ __forceinline__ double sm_x(const WarpSmemT<0>& sm, int buf, int i, int k) { return sm.X[buf][i][k]; }
template <int S>
__device__ __forceinline__ double sm_x(const CompactSmem<S>& sm, int, int i, int k) { return sm.X[i][k]; }
__device__ __forceinline__ double sm_de(const WarpSmemT<0>& sm, int buf, int k) { return sm.de[buf][k]; }
template <int S>
__device__ __forceinline__ double sm_de(const CompactSmem<S>& sm, int, int k) { return sm.de[k]; }
}  // namespace

template <bool FUSE, int LAYOUT>
__global__ void __launch_bounds__(warps_of<LAYOUT>() * 32, min_ctas_of<LAYOUT>())
assemble_hex_mech_f64_kernel(const AsmArgs<double> args, const long long ntiles, const int has_body, const HaloFuse hf) {
  using WarpSmem = WarpSmemT<LAYOUT>;
  constexpr int kWarps = warps_of<LAYOUT>();
  constexpr int kStage = WarpSmem::kStage;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpSmem& sm = reinterpret_cast<WarpSmem*>(smem_raw)[warp];
  const long long nwarps = (long long)gridDim.x * kWarps;
  // vt: position in the visiting order (interface layers first when FUSE), tile: the tile itself
  long long vt = (long long)blockIdx.x * kWarps + warp;
  if (vt >= ntiles) return;
  auto real_tile = [&](long long v) -> long long {
    if constexpr (FUSE) return v < ntiles ? halo_real_tile(hf, v, ntiles) : v;
    else return v;
  };
  bool push_done = !FUSE;

  const double E = args.p.v[0], nu = args.p.v[1];
  const double c1 = E / ((1.0 + nu) * (1.0 - 2.0 * nu));
  const double lam = c1 * nu, mu = c1 * 0.5 * (1.0 - 2.0 * nu);

  // lane roles
  const int el_p = lane >> 3, sub = lane & 7;   // phases 0/1: (element in tile, node | gauss point)
  const int ra = lane >> 2, kq = lane & 3;      // phase 2: (row node a, column pair k)
  const int swz_p = ((sub & 3) << 1) | (sub >> 2);

  // Gauss point `sub` of the 2x2x2 rule (hexahedra_3d_8.py:23-33): xi = sgn(sub) / sqrt(3), w = 1.
  // Trilinear shape data factorised per axis: f?[0] = 1 - xi_?, f?[1] = 1 + xi_? and the pair
  // products below give N_a = fx*fyz, dN_a/dxi = sx(a)*fyz, ... with compile-time node signs.
  const double px = sgn_x(sub) * FOL_S3, py = sgn_y(sub) * FOL_S3, pz = sgn_z(sub) * FOL_S3;
  const double fx[2] = {1.0 - px, 1.0 + px};
  double fyz[2][2], fxz[2][2], fxy[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      fyz[i][j] = 0.125 * (i ? 1.0 + py : 1.0 - py) * (j ? 1.0 + pz : 1.0 - pz);
      fxz[i][j] = 0.125 * (i ? 1.0 + px : 1.0 - px) * (j ? 1.0 + pz : 1.0 - pz);
      fxy[i][j] = 0.125 * (i ? 1.0 + px : 1.0 - px) * (j ? 1.0 + py : 1.0 - py);
    }

  // software pipeline of the gathers: node ids two tiles ahead, nodal data one tile ahead.
  // The id comes back through a predicated load straight into its 32-bit register and is widened only where it is
  // used, one tile later (hold_back below): any earlier dependent instruction -- a select, a sign extension --
  // would make the warp sit out the full DRAM latency of the connectivity load (21 % of the stall samples before).
  auto node_of = [&](long long v) -> int {
    const long long t = real_tile(v);
    const long long e = t * kTile + el_p;
    const int ok = (t < ntiles && e < args.ne) ? 1 : 0;
    const int32_t* src = args.conn + (ok ? e * 8 + sub : 0);
    int n;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\tmov.s32 %0, 0;\n\t@p ld.global.nc.s32 %0, [%1];\n\t}\n"
        : "=r"(n)
        : "l"(src), "r"(ok));
    return n;
  };
  auto hold_back = [](int& a, unsigned& b, unsigned& c, unsigned& d) {
    asm volatile("" : "+r"(a), "+r"(b), "+r"(c), "+r"(d));
  };
  // asynchronous gather of one node straight into shared memory (no registers held across the tile)
  auto gather_async = [&](int buf, long long n) {
    const double* px = args.xyz + n * 3;
    const double* pu = args.u + n * 3;
    if constexpr (LAYOUT == 0) {
      cp_async8(&sm.X[buf][0][lane], px); cp_async8(&sm.X[buf][1][lane], px + 1); cp_async8(&sm.X[buf][2][lane], px + 2);
      cp_async8(&sm.de[buf][lane], args.ctrl + n);
    } else {
      cp_async8(&sm.X[0][lane], px); cp_async8(&sm.X[1][lane], px + 1); cp_async8(&sm.X[2][lane], px + 2);
      cp_async8(&sm.de[lane], args.ctrl + n);
    }
    cp_async8(&sm.u[buf][0][lane], pu); cp_async8(&sm.u[buf][1][lane], pu + 1); cp_async8(&sm.u[buf][2][lane], pu + 2);
    cp_async_commit();
  };

  // Dirichlet flags of the next tile's node: three byte loads kept in three registers, consumed one
  // tile later (packing them right away would stall on the load latency)
  int n_next = node_of(vt + nwarps);
  const long long n_first = node_of(vt);
  gather_async(0, n_first);
  const uint8_t* pf0 = args.dir + n_first * 3;
  unsigned f0 = __ldg(pf0), f1 = __ldg(pf0 + 1), f2 = __ldg(pf0 + 2);
  int buf = 0;

  for (; vt < ntiles; vt += nwarps, buf ^= 1) {
    const long long e0 = real_tile(vt) * kTile;
    hold_back(n_next, f0, f1, f2);   // loaded one tile ago; nothing may consume them before this point

    // ---- phase 0: this tile's nodal data has landed in shared memory; start the next gather
    sm.bc[el_p][sub * 3 + 0] = f0 ? 0 : 1;
    sm.bc[el_p][sub * 3 + 1] = f1 ? 0 : 1;
    sm.bc[el_p][sub * 3 + 2] = f2 ? 0 : 1;
    cp_async_wait_all();
    __syncwarp();
    // the next tile's gather (and its Dirichlet flags / the node ids two tiles ahead): in flight during this tile.
    // LAYOUT 1 single-buffers X / de, which phase 1 still reads: there it is issued right after phase 1.
    auto issue_next = [&]() {
      gather_async(buf ^ 1, (long long)n_next);
      const uint8_t* pf = args.dir + (long long)n_next * 3;
      f0 = __ldg(pf); f1 = __ldg(pf + 1); f2 = __ldg(pf + 2);
      n_next = node_of(vt + 2 * nwarps);
    };
    if constexpr (LAYOUT == 0) issue_next();

    // ---- phase 1: lane (element, Gauss point): J, det J, grad N, coefficient (geometry.py:88-97)
    {
      double J[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double j0 = 0.0, j1 = 0.0, j2 = 0.0;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          const int bx = ((a & 3) == 1 || (a & 3) == 2), by = (a >> 1) & 1, bz = (a >> 2) & 1;
          const double x = sm_x(sm, buf, i, el_p * 8 + a);
          j0 += (bx ? x : -x) * fyz[by][bz];
          j1 += (by ? x : -x) * fxz[bx][bz];
          j2 += (bz ? x : -x) * fxy[bx][by];
        }
        J[i][0] = j0; J[i][1] = j1; J[i][2] = j2;
      }
      const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
      const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
      const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
      const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
      const double rd = 1.0 / det;
      double inv[3][3];
      inv[0][0] = c00 * rd; inv[1][0] = c01 * rd; inv[2][0] = c02 * rd;
      inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * rd;
      inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * rd;
      inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * rd;
      inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * rd;
      inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * rd;
      inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * rd;
      double eg = 0.0;
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int bx = ((a & 3) == 1 || (a & 3) == 2), by = (a >> 1) & 1, bz = (a >> 2) & 1;
        eg += fx[bx] * fyz[by][bz] * sm_de(sm, buf, el_p * 8 + a);
      }
      const double wd = det;  // Gauss weight is 1
      const double coef = wd * eg;
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int bx = ((a & 3) == 1 || (a & 3) == 2), by = (a >> 1) & 1, bz = (a >> 2) & 1;
        const double d0 = bx ? fyz[by][bz] : -fyz[by][bz];
        const double d1 = by ? fxz[bx][bz] : -fxz[bx][bz];
        const double d2 = bz ? fxy[bx][by] : -fxy[bx][by];
        double g[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) g[k] = d0 * inv[0][k] + d1 * inv[1][k] + d2 * inv[2][k];
        sm.gxy[el_p][sub][a ^ swz_p] = make_double2(g[0], g[1]);
        if constexpr (LAYOUT == 0) sm.gzs[el_p][sub][a ^ swz_p] = make_double2(g[2], coef);
        else sm.gz[el_p][sub][a ^ swz_p ^ ((el_p & 1) << 1)] = g[2];
      }
      if constexpr (LAYOUT != 0) sm.coef[lane] = coef;
      sm.wd[el_p][sub] = wd;
    }
    __syncwarp();
    if constexpr (LAYOUT != 0) issue_next();   // X / de of this tile are consumed

    // ---- phase 2: one element at a time, lane (a, k)
#pragma unroll 1
    for (int el = 0; el < kTile; ++el) {
      const long long e = e0 + el;
      if (e >= args.ne) break;
      // P = sum_g s_g v_g v_g^T on the FP64 tensor path: rows/cols ordered m = 8*dim + node
      double c[3][3][2];
#pragma unroll
      for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int s = 0; s < 3; ++s) c[t][s][0] = c[t][s][1] = 0.0;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const double2 xy = sm.gxy[el][4 * kk + kq][ra ^ ((kq << 1) | kk)];
        double2 zs;
        if constexpr (LAYOUT == 0) {
          zs = sm.gzs[el][4 * kk + kq][ra ^ ((kq << 1) | kk)];
        } else {
          zs.x = sm.gz[el][4 * kk + kq][ra ^ ((kq << 1) | kk) ^ ((el & 1) << 1)];
          zs.y = sm.coef[el * 8 + 4 * kk + kq];
        }
        const double bf[3] = {xy.x, xy.y, zs.x};
        const double af[3] = {zs.y * xy.x, zs.y * xy.y, zs.y * zs.x};
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int s = (kSym ? t : 0); s < 3; ++s) dmma884(c[t][s][0], c[t][s][1], af[t], bf[s]);
      }
      if constexpr (kSym) {
        // P is symmetric: the tiles below the diagonal are the transposes of the ones above, P[(a,t),(b,s)] =
        // P[(b,s),(a,t)].  Lane (a, k) wants tile(s,t)[b = 2k + h][a], which lane (a' = 2k + h, k' = a >> 1) holds in
        // register h' = a & 1: four 64-bit shuffles and two selects per tile instead of two DMMAs (12 instead of 18 per
        // element: a sixth less FP64 work, which is what the power cap meters under sustained load).
        const int src0 = ((2 * kq) << 2) | (ra >> 1), src1 = ((2 * kq + 1) << 2) | (ra >> 1);
        const bool odd = (ra & 1) != 0;
#pragma unroll
        for (int t = 1; t < 3; ++t)
#pragma unroll
          for (int s = 0; s < t; ++s) {
            const double a0 = __shfl_sync(0xffffffffu, c[s][t][0], src0), a1 = __shfl_sync(0xffffffffu, c[s][t][1], src0);
            const double b0 = __shfl_sync(0xffffffffu, c[s][t][0], src1), b1 = __shfl_sync(0xffffffffu, c[s][t][1], src1);
            c[t][s][0] = odd ? a1 : a0;
            c[t][s][1] = odd ? b1 : b0;
          }
      }
      // Ke blocks (a, 2k) and (a, 2k+1): lam P + mu P^T + mu tr(P) I  (B^T D B of an isotropic D)
      double K[2][3][3];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const double tr = c[0][0][h] + c[1][1][h] + c[2][2][h];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j)
            K[h][i][j] = lam * c[i][j][h] + mu * c[j][i][h] + (i == j ? mu * tr : 0.0);
      }
      // (done before touching the staging slot: the previous element's copy keeps draining meanwhile)
      // re = Ke u - Fe: partial over this lane's 6 columns, then butterfly over the 4 k-lanes
      double r[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 3; ++j) acc += K[h][i][j] * sm.u[buf][j][el * 8 + 2 * kq + h];
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        r[i] = acc;
      }
      if (has_body & 1) {  // Fe_a = b * sum_g w detJ N_a(g)   (mechanical.py:110)
        double nw = 0.0;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const double gx = 1.0 + sgn_x(ra) * sgn_x(g) * FOL_S3, gy = 1.0 + sgn_y(ra) * sgn_y(g) * FOL_S3;
          const double gz = 1.0 + sgn_z(ra) * sgn_z(g) * FOL_S3;
          nw += sm.wd[el][g] * (0.125 * gx * gy * gz);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) r[i] -= args.p.v[2 + i] * nw;
      }
      // stage the rows and hand them to the bulk-copy engine.  The Dirichlet row mask
      // (fe_loss.py:191-207) only matters for elements touching a fixed dof: warp-uniform test.
      const bool fx[3] = {sm.bc[el][ra * 3 + 0] == 0, sm.bc[el][ra * 3 + 1] == 0, sm.bc[el][ra * 3 + 2] == 0};
      const bool fixed_rows = fx[0] | fx[1] | fx[2];
      const bool any_fixed = __any_sync(0xffffffffu, fixed_rows);
      double* const slot = sm.stage[kStage > 1 ? (el & (kStage - 1)) : 0];
      auto write_rows = [&]() {
        if (!any_fixed) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            double2* dst = reinterpret_cast<double2*>(slot + (ra * 3 + i) * 24 + kq * 6);
            dst[0] = make_double2(K[0][i][0], K[0][i][1]);
            dst[1] = make_double2(K[0][i][2], K[1][i][0]);
            dst[2] = make_double2(K[1][i][1], K[1][i][2]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int row = ra * 3 + i;
            const bool freerow = LAYOUT == 3 ? !fx[i] : sm.bc[el][row] != 0;
            double v[6];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const int col = (2 * kq + h) * 3 + j;
                v[h * 3 + j] = (freerow || col == row) ? K[h][i][j] : 0.0;
              }
            double2* dst = reinterpret_cast<double2*>(slot + row * 24 + kq * 6);
            dst[0] = make_double2(v[0], v[1]);
            dst[1] = make_double2(v[2], v[3]);
            dst[2] = make_double2(v[4], v[5]);
          }
        }
      };
      auto store_re = [&]() {
        if constexpr (LAYOUT == 3) {
          if (kq == 0) {
#pragma unroll
            for (int i = 0; i < 3; ++i) args.re[e * 24 + ra * 3 + i] = fx[i] ? 0.0 : r[i];
          }
        } else if (kq == 0) {
#pragma unroll
          for (int i = 0; i < 3; ++i)
            args.re[e * 24 + ra * 3 + i] = (any_fixed && sm.bc[el][ra * 3 + i] == 0) ? 0.0 : r[i];
        }
      };
      if constexpr (kHalves) {
        // the slot is released and refilled in two halves (rows 0-11, 12-23), each its own bulk copy: a half only
        // waits for the copy issued two copies earlier, so the engine always has one copy in flight
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        if (ra < 4) write_rows();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) bulk_store(args.ke + e * 576, slot, 288 * sizeof(double));
        store_re();
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        if (ra >= 4) write_rows();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) bulk_store(args.ke + e * 576 + 288, slot + 288, 288 * sizeof(double));
      } else {
        // the copy that last used this slot has drained it (lanes other than 0 have no copies of their own: for them
        // the wait returns at once, and LAYOUT 3 lets them execute it instead of branching around it)
        if (LAYOUT == 3 || lane == 0) bulk_wait_read<kStage - 1>();
        __syncwarp();
        write_rows();
        fence_async_smem();
        __syncwarp();
        if (lane == 0 && !(has_body & 2)) bulk_store(args.ke + e * 576, slot, 576 * sizeof(double));
        store_re();
      }
    }
    __syncwarp();  // everyone is done with X / u / gradients of this tile
    if constexpr (FUSE) {
      if (vt < hf.tiles_lo + hf.tiles_hi) halo_tile_done(hf, lane);
      if (!push_done) push_done = halo_try_push(hf, lane);
    }
  }
  if constexpr (FUSE) {
    if (!push_done) halo_drain(hf, lane);
  }
  if (lane == 0) bulk_wait_read<0>();  // shared memory must outlive the last copies
}

std::atomic<int> g_grid_margin{0};

template <bool FUSE, int LAYOUT>
static int launch_hex(cudaStream_t s, const AsmArgs<double>& args, const HaloFuse* hf) {
  constexpr int kWarps = warps_of<LAYOUT>();
  static PerDeviceGrid per_device;
  const size_t smem = sizeof(WarpSmemT<LAYOUT>) * kWarps;
  int grid = 0;
  FOL_CUDA(per_device.get(assemble_hex_mech_f64_kernel<FUSE, LAYOUT>, kWarps * 32, smem, &grid));
  if (args.ne == 0) return FOL_OK;
  const long long ntiles = cdiv(args.ne, kTile);
  const long long want = cdiv(ntiles, kWarps);
  int has_body = (args.p.v[2] != 0.0 || args.p.v[3] != 0.0 || args.p.v[4] != 0.0) ? 1 : 0;
  // diagnostic only (scripts/fused_ab.py): FOL_HEX_DIAG=nostore runs the kernel without its Ke bulk stores, which
  // separates the compute / latency time of the kernel from its HBM write stream
  static const bool nostore = [] { const char* v = std::getenv("FOL_HEX_DIAG"); return v && std::string(v) == "nostore"; }();
  if (nostore) has_body |= 2;
  // persistent grid, optionally leaving room for communication kernels that must run concurrently
  int g = grid - g_grid_margin.load();
  if (g < 1) g = 1;
  const unsigned blocks = (unsigned)(want < g ? want : g);
  assemble_hex_mech_f64_kernel<FUSE, LAYOUT><<<blocks, kWarps * 32, smem, s>>>(args, ntiles, has_body, hf ? *hf : HaloFuse{});
  return check_launch("assemble_hex_mech_f64_kernel");
}

int assemble_hex_mech_f64(cudaStream_t s, const AsmArgs<double>& args, const HaloFuse* hf) {
  // default: layout 3 (16 warps / SM, lean hand-off).  FOL_HEX_LAYOUT=0 / 1 select the 12-warp layouts, 2 the 16-warp
  // layout with the original hand-off, for A/B runs
  // (profiles/r2/hex_kernel_experiments.md); all three produce the same bits.
  static const int layout = [] { const char* v = std::getenv("FOL_HEX_LAYOUT"); return v ? std::atoi(v) : 3; }();
  if (layout == 0) return hf ? launch_hex<true, 0>(s, args, hf) : launch_hex<false, 0>(s, args, hf);
  if (layout == 2) return hf ? launch_hex<true, 2>(s, args, hf) : launch_hex<false, 2>(s, args, hf);
  if (layout == 3) return hf ? launch_hex<true, 3>(s, args, hf) : launch_hex<false, 3>(s, args, hf);
  return hf ? launch_hex<true, 1>(s, args, hf) : launch_hex<false, 1>(s, args, hf);
}

}  // namespace fol
