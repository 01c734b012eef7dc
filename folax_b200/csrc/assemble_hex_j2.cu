// Tuned element-stage kernel for BASELINE.json configs[4]: 3-D Hex8 J2 elastoplasticity with Gauss-point history,
// float64, 2x2x2 rule (ElastoplasticityLoss3DHexa; mechanical_elastoplasticity.py:32-94, 153-235,
// plasticity.py:136-325, fe_loss.py:191-230, 299).
//
// Same machine mapping as assemble_hex.cu (persistent warps, tiles of 4 elements, cp.async gathers one tile ahead,
// FP64 tensor-path MMAs, one 4608-byte bulk copy per element matrix), with the point law of j2_point.cuh in phase 1:
//   * phase 1, lane (element, Gauss point): geometry, strain = B u, radial return + scalar forward-mode tangent
//     (j2_radial: ~60 registers, no local memory), staged per point: d = dev(e_trial), w detJ sigma and the three
//     tangent scalars.  The history of the tile arrives by cp.async one tile ahead and the new history leaves as ONE
//     1792-byte bulk copy per tile (the (ne, 8, 7) layout is contiguous per tile);
//   * phase 2, lane (row node a, column pair k): the point tangent is C_el - a Dev - b d (x) (w . d), i.e. an isotropic
//     matrix with point-wise moduli (lam + a/3, 2G - a) minus a rank-one term, so
//       Ke_ab = P1 + tr(P2) I + offdiag(P2^T) - sum_g (w detJ b) p_a (x) pt_b,
//     P1 = sum_g w detJ (lam + a_g/3) g_a (x) g_b,  P2 = sum_g w detJ (2G - a_g) g_a (x) g_b,  p_a = d . g_a,
//     pt_a = (w . d) . g_a: three 24x24x8 products on the FP64 tensor path (18 DMMA m8n8k4 each); P1 and the rank-one
//     family accumulate straight into the Ke fragment.  Elements without a plastic point (warp-uniform test) run ONE
//     family and apply the constant moduli afterwards: they cost what the elastic kernel costs.
//     f_int = sum_g w detJ B^T sigma from the staged stresses (butterfly over the 4 k-lanes).
// The reference's strain convention is kept: engineering shears enter the strain TENSOR unhalved
// (mechanical_elastoplasticity.py:50-55), so the shear stiffness is 2G and C_el = lam 1(x)1 + 2G I_6.
// transpose_jacobian=True and the matrix-free mode are served by the generic kernel (assemble.cuh).
#include <cstdlib>

#include "assemble.cuh"
#include "assemble_hex_common.cuh"

namespace fol {

namespace {

using namespace hexk;

#ifndef FOL_J2_WARPS
#define FOL_J2_WARPS 5
#endif
constexpr int kWarpsJ2 = FOL_J2_WARPS;   // warps per CTA of layout 0, each fully independent (2 CTAs = 10 warps / SM, 19.9 KB of staging per warp)

// LAYOUT 0: (dN/dz, w detJ) pairs and float Dirichlet flags, 5 warps per CTA (10 warps / SM).
// LAYOUT 1: compact -- dN/dz alone (w detJ is per Gauss point, not per node: 32 doubles instead of 256) and byte flags:
//           17.9 KB per warp, which fits 6 warps per CTA (12 warps / SM at the same two CTAs, 162 registers).  The default:
//           same box, alternating processes, 128^3, all points plastic: 4.55 -> 3.87 ms per step, outputs bit-identical
//           (profiles/r2/hex_layout_ab.jsonl) -- like the elastic kernel, phase 2 is latency-bound, not throughput-bound.
template <int LAYOUT>
struct GzStore;
template <>
struct GzStore<0> { double2 v[kTile][8][8]; };
template <>
struct GzStore<1> { double v[kTile][8][8]; };
template <>
struct GzStore<2> { double v[kTile][8][8]; };   // LAYOUT 2 = layout 1 + the leaner hand-off (A/B: FOL_J2_LAYOUT=2)

template <int LAYOUT>
struct __align__(128) WarpSmemJ2T {
  // Ke staging slot (bulk-copy source).  Its first 224 doubles double as the staging of the tile's NEW history in the
  // global (element, point, 7) layout (one 1792-byte bulk copy per tile): the history copy is issued before the
  // element loop, whose first staging write waits for it, and the last Ke copy of a tile is waited for before the
  // next tile's history is written.
  double stage[576];
  double2 gxy[kTile][8][8];          // [element][gauss][node ^ swz(gauss)]: (dN/dx, dN/dy), see assemble_hex.cu
  // LAYOUT 0: gz2 = (dN/dz, w detJ) pairs; LAYOUT 1: gz1[element][gauss][node ^ swz(gauss) ^ 2 (element & 1)] = dN/dz
  // (8-byte accesses are served per half-warp = two elements in phase 1: the element bit keeps them on distinct banks)
  // and wdj[element * 8 + gauss] = w detJ
  GzStore<LAYOUT> gz;
  double wdj[32];
  // nodal data and history of the tile, SoA over the 32 lanes.  SINGLE buffers: phase 1 consumes them completely, so
  // the gather of the next tile is issued right after phase 1 and lands behind phase 2
  double X[3][32];
  double u[3][32];
  double st[7][32];
  double dv[6][32];                  // dev of the trial elastic strain          } per (element, Gauss point),
  double ws[6][32];                  // w detJ sigma                             } SoA: conflict-free phase-1 stores,
  double wl[32], wm[32], wb[32];     // w detJ (lam + a/3), w detJ (2G - a), w detJ b   } broadcast phase-2 loads
  uint8_t bc[kTile][24];             // 1 = free dof, 0 = Dirichlet dof
};
constexpr int kWarpsJ2Dense = 6;
template <int LAYOUT>
constexpr int j2_warps_of() { return LAYOUT >= 1 ? kWarpsJ2Dense : kWarpsJ2; }

// two CTAs per SM must fit the 227 KB of shared memory (1 KB per CTA is reserved by the system)
static_assert(2 * (sizeof(WarpSmemJ2T<0>) * kWarpsJ2 + 1024) <= 227 * 1024, "WarpSmemJ2: two CTAs per SM do not fit");
static_assert(2 * (sizeof(WarpSmemJ2T<1>) * kWarpsJ2Dense + 1024) <= 227 * 1024, "compact WarpSmemJ2: two CTAs of 6 warps do not fit");

}  // namespace

template <bool FUSE, int LAYOUT>
__global__ void __launch_bounds__(j2_warps_of<LAYOUT>() * 32, 2)
assemble_hex_j2_f64_kernel(const AsmArgs<double> args, const long long ntiles, const int has_body, const HaloFuse hf) {
  using WarpSmemJ2 = WarpSmemJ2T<LAYOUT>;
  constexpr int kWarpsJ2 = j2_warps_of<LAYOUT>();
  extern __shared__ __align__(128) unsigned char smem_raw_j2[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpSmemJ2& sm = reinterpret_cast<WarpSmemJ2*>(smem_raw_j2)[warp];
  const long long nwarps = (long long)gridDim.x * kWarpsJ2;
  // vt: position in the visiting order (interface layers first when FUSE, see assemble_hex_common.cuh)
  long long vt = (long long)blockIdx.x * kWarpsJ2 + warp;
  if (vt >= ntiles) return;
  auto real_tile = [&](long long v) -> long long {
    if constexpr (FUSE) return v < ntiles ? halo_real_tile(hf, v, ntiles) : v;
    else return v;
  };
  bool push_done = !FUSE;

  const double E = args.p.v[0], nu = args.p.v[1], y0 = args.p.v[5], h1 = args.p.v[6], h2 = args.p.v[7];
  const double lam = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu)), G = E / (2.0 * (1.0 + nu));

  const int el_p = lane >> 3, sub = lane & 7;   // phases 0/1: (element in tile, node | Gauss point)
  const int ra = lane >> 2, kq = lane & 3;      // phase 2: (row node a, column pair k)
  const int swz_p = ((sub & 3) << 1) | (sub >> 2);

  // Gauss point `sub` of the 2x2x2 rule (hexahedra_3d_8.py:23-33), trilinear shape data factorised per axis
  const double px = sgn_x(sub) * FOL_S3, py = sgn_y(sub) * FOL_S3, pz = sgn_z(sub) * FOL_S3;
  double fyz[2][2], fxz[2][2], fxy[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      fyz[i][j] = 0.125 * (i ? 1.0 + py : 1.0 - py) * (j ? 1.0 + pz : 1.0 - pz);
      fxz[i][j] = 0.125 * (i ? 1.0 + px : 1.0 - px) * (j ? 1.0 + pz : 1.0 - pz);
      fxy[i][j] = 0.125 * (i ? 1.0 + px : 1.0 - px) * (j ? 1.0 + py : 1.0 - py);
    }

  // software pipeline of the gathers (node ids two tiles ahead, nodal data and history one tile ahead); see
  // assemble_hex.cu for why the id is held back in its register
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
  auto gather_async = [&](long long n, long long v) {
    const long long t = real_tile(v);
    const double* gx = args.xyz + n * 3;
    const double* gu = args.u + n * 3;
    cp_async8(&sm.X[0][lane], gx); cp_async8(&sm.X[1][lane], gx + 1); cp_async8(&sm.X[2][lane], gx + 2);
    cp_async8(&sm.u[0][lane], gu); cp_async8(&sm.u[1][lane], gu + 1); cp_async8(&sm.u[2][lane], gu + 2);
    // history of (element, Gauss point) = this lane; tiles past the end read element 0 (never used)
    const long long e = t * kTile + el_p;
    const double* gs = args.state_in + ((t < ntiles && e < args.ne) ? (e * 8 + sub) * 7 : 0);
#pragma unroll
    for (int s = 0; s < 7; ++s) cp_async8(&sm.st[s][lane], gs + s);
    cp_async_commit();
  };
  int n_next = node_of(vt + nwarps);
  const long long n_first = node_of(vt);
  gather_async(n_first, vt);
  const uint8_t* pf0 = args.dir + n_first * 3;
  unsigned f0 = __ldg(pf0), f1 = __ldg(pf0 + 1), f2 = __ldg(pf0 + 2);

  for (; vt < ntiles; vt += nwarps) {
    const long long e0 = real_tile(vt) * kTile;
    hold_back(n_next, f0, f1, f2);

    // ---- phase 0: this tile's nodal data and history have landed
    sm.bc[el_p][sub * 3 + 0] = f0 ? 0 : 1;
    sm.bc[el_p][sub * 3 + 1] = f1 ? 0 : 1;
    sm.bc[el_p][sub * 3 + 2] = f2 ? 0 : 1;
    cp_async_wait_all();
    __syncwarp();
    // ---- phase 1: lane (element, Gauss point): geometry (geometry.py:88-97), strain = B u, return mapping
    unsigned plastic_mask;
    {
      double J[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double j0 = 0.0, j1 = 0.0, j2 = 0.0;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          const int bx = ((a & 3) == 1 || (a & 3) == 2), by = (a >> 1) & 1, bz = (a >> 2) & 1;
          const double x = sm.X[i][el_p * 8 + a];
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
      const double wd = det;  // Gauss weight is 1
      double H[3][3];         // displacement gradient H[i][k] = sum_a u_a[i] dN_a/dx_k
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) H[i][k] = 0.0;
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
        if constexpr (LAYOUT == 0) sm.gz.v[el_p][sub][a ^ swz_p] = make_double2(g[2], wd);
        else sm.gz.v[el_p][sub][a ^ swz_p ^ ((el_p & 1) << 1)] = g[2];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const double ui = sm.u[i][el_p * 8 + a];
#pragma unroll
          for (int k = 0; k < 3; ++k) H[i][k] += ui * g[k];
        }
      }
      // rows of the linear B matrix (mechanical.py:46-58): [xx, yy, zz, xy, yz, xz], engineering shears
      const double e_tot[6] = {H[0][0], H[1][1], H[2][2], H[0][1] + H[1][0], H[1][2] + H[2][1], H[0][2] + H[2][0]};
      double ep[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) ep[k] = sm.st[k][lane];
      const double xi = sm.st[6][lane];
      J2Point<double> p;
      j2_radial<double>(e_tot, ep, xi, lam, G, y0, h1, h2, p);
      if (lane == 0) bulk_wait_read<0>();   // the previous tile's last Ke copy has left the staging slot
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        sm.dv[k][lane] = p.d[k];
        sm.ws[k][lane] = wd * p.sig[k];
        sm.stage[lane * 7 + k] = ep[k] + p.c * p.d[k];
      }
      sm.stage[lane * 7 + 6] = xi + p.dl;
      sm.wl[lane] = wd * (lam + p.a * (1.0 / 3.0));
      sm.wm[lane] = wd * (2.0 * G - p.a);
      sm.wb[lane] = wd * p.b;
      sm.wdj[lane] = wd;
      plastic_mask = __ballot_sync(0xffffffffu, (p.a != 0.0) | (p.b != 0.0));
    }
    fence_async_smem();
    __syncwarp();
    // X / u / history of this tile are consumed: the next tile's gather lands behind phase 2
    gather_async((long long)n_next, vt + nwarps);
    {
      const uint8_t* pf = args.dir + (long long)n_next * 3;
      f0 = __ldg(pf); f1 = __ldg(pf + 1); f2 = __ldg(pf + 2);
    }
    n_next = node_of(vt + 2 * nwarps);
    {  // new history of the tile: one contiguous bulk copy (mechanical_elastoplasticity.py:217)
      long long cnt = args.ne - e0;
      cnt = cnt > kTile ? kTile : cnt;
      if (lane == 0) bulk_store(args.state_out + e0 * 56, sm.stage, (unsigned)(cnt * 56 * sizeof(double)));
    }

    // ---- phase 2: one element at a time, lane (a, k)
#pragma unroll 1
    for (int el = 0; el < kTile; ++el) {
      const long long e = e0 + el;
      if (e >= args.ne) break;
      const bool plastic = ((plastic_mask >> (el * 8)) & 0xffu) != 0u;   // warp-uniform
      // K[t][s][h]: Ke blocks (a, 2k + h), entry (t, s); cs: the second family (plastic elements only)
      double K[3][3][2], cs[3][3][2];
#pragma unroll
      for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int s = 0; s < 3; ++s) K[t][s][0] = K[t][s][1] = cs[t][s][0] = cs[t][s][1] = 0.0;
      double bf[2][3], r[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int gp = 4 * kk + kq, ix = el * 8 + gp;
        const double2 xy = sm.gxy[el][gp][ra ^ ((kq << 1) | kk)];
        double2 zs;
        if constexpr (LAYOUT == 0) {
          zs = sm.gz.v[el][gp][ra ^ ((kq << 1) | kk)];
        } else {
          zs.x = sm.gz.v[el][gp][ra ^ ((kq << 1) | kk) ^ ((el & 1) << 1)];
          zs.y = sm.wdj[ix];
        }
        bf[kk][0] = xy.x; bf[kk][1] = xy.y; bf[kk][2] = zs.x;
        // f_int share of this lane's two Gauss points: (w detJ sigma) . grad N_a
        const double s0 = sm.ws[0][ix], s1 = sm.ws[1][ix], s2 = sm.ws[2][ix];
        const double s3 = sm.ws[3][ix], s4 = sm.ws[4][ix], s5 = sm.ws[5][ix];
        r[0] += s0 * xy.x + s3 * xy.y + s5 * zs.x;
        r[1] += s3 * xy.x + s1 * xy.y + s4 * zs.x;
        r[2] += s5 * xy.x + s4 * xy.y + s2 * zs.x;
        if (!plastic) {
          // all eight points elastic: one family P = sum_g w detJ g_a (x) g_b, moduli applied afterwards
          const double af[3] = {zs.y * xy.x, zs.y * xy.y, zs.y * zs.x};
#pragma unroll
          for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int s = t; s < 3; ++s) dmma884(K[t][s][0], K[t][s][1], af[t], bf[kk][s]);
        } else {
          // the moduli differ from point to point: family 1 weighted by w detJ (lam + a/3) accumulates straight into
          // Ke, family 2 weighted by w detJ (2G - a) enters transposed within the 3x3 blocks
          const double wl = sm.wl[ix], wm = sm.wm[ix];
          const double al[3] = {wl * xy.x, wl * xy.y, wl * zs.x};
          const double am[3] = {wm * xy.x, wm * xy.y, wm * zs.x};
#pragma unroll
          for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int s = t; s < 3; ++s) {
              dmma884(K[t][s][0], K[t][s][1], al[t], bf[kk][s]);
              dmma884(cs[t][s][0], cs[t][s][1], am[t], bf[kk][s]);
            }
        }
      }
      {
        // both weighted sums of g_a (x) g_b are symmetric, P[(a,t),(b,s)] = P[(b,s),(a,t)]: the tiles below the diagonal
        // are mirrored from the ones above by shuffles instead of being computed (see assemble_hex.cu)
        const int src0 = ((2 * kq) << 2) | (ra >> 1), src1 = ((2 * kq + 1) << 2) | (ra >> 1);
        const bool odd = (ra & 1) != 0;
#pragma unroll
        for (int t = 1; t < 3; ++t)
#pragma unroll
          for (int s = 0; s < t; ++s) {
            const double a0 = __shfl_sync(0xffffffffu, K[s][t][0], src0), a1 = __shfl_sync(0xffffffffu, K[s][t][1], src0);
            const double b0 = __shfl_sync(0xffffffffu, K[s][t][0], src1), b1 = __shfl_sync(0xffffffffu, K[s][t][1], src1);
            K[t][s][0] = odd ? a1 : a0;
            K[t][s][1] = odd ? b1 : b0;
            if (plastic) {
              const double c0 = __shfl_sync(0xffffffffu, cs[s][t][0], src0), c1 = __shfl_sync(0xffffffffu, cs[s][t][1], src0);
              const double d0 = __shfl_sync(0xffffffffu, cs[s][t][0], src1), d1 = __shfl_sync(0xffffffffu, cs[s][t][1], src1);
              cs[t][s][0] = odd ? c1 : c0;
              cs[t][s][1] = odd ? d1 : d0;
            }
          }
      }
      if (!plastic) {
        // C_el = lam 1(x)1 + 2G I_6:  Ke = lam P + 2G [tr(P) I + offdiag(P^T)]
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const double p00 = K[0][0][h], p01 = K[0][1][h], p02 = K[0][2][h], p10 = K[1][0][h], p11 = K[1][1][h],
                       p12 = K[1][2][h], p20 = K[2][0][h], p21 = K[2][1][h], p22 = K[2][2][h];
          const double tr = 2.0 * G * (p00 + p11 + p22);
          K[0][0][h] = lam * p00 + tr; K[0][1][h] = lam * p01 + 2.0 * G * p10; K[0][2][h] = lam * p02 + 2.0 * G * p20;
          K[1][0][h] = lam * p10 + 2.0 * G * p01; K[1][1][h] = lam * p11 + tr; K[1][2][h] = lam * p12 + 2.0 * G * p21;
          K[2][0][h] = lam * p20 + 2.0 * G * p02; K[2][1][h] = lam * p21 + 2.0 * G * p12; K[2][2][h] = lam * p22 + tr;
        }
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const double tr = cs[0][0][h] + cs[1][1][h] + cs[2][2][h];
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) K[i][j][h] += (i == j) ? tr : cs[j][i][h];
        }
        // rank-one family: Ke -= sum_g (w detJ b) p (x) pt, accumulated on the tensor path into the Ke fragment
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int ix = el * 8 + 4 * kk + kq;
          const double d0 = sm.dv[0][ix], d1 = sm.dv[1][ix], d2 = sm.dv[2][ix];
          const double d3 = sm.dv[3][ix], d4 = sm.dv[4][ix], d5 = sm.dv[5][ix];
          const double wb = -sm.wb[ix];
          const double gx = bf[kk][0], gy = bf[kk][1], gz = bf[kk][2];
          const double sx = d3 * gy + d5 * gz, sy = d3 * gx + d4 * gz, sz = d5 * gx + d4 * gy;   // shear parts
          const double nx = d0 * gx, ny = d1 * gy, nz = d2 * gz;
          const double af[3] = {wb * (nx + sx), wb * (ny + sy), wb * (nz + sz)};
          const double pt[3] = {nx + 2.0 * sx, ny + 2.0 * sy, nz + 2.0 * sz};
#pragma unroll
          for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int s = 0; s < 3; ++s) dmma884(K[t][s][0], K[t][s][1], af[t], pt[s]);
        }
      }
      // f_int of node a: butterfly over the 4 k-lanes (each holds two of the eight Gauss points)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        r[i] += __shfl_xor_sync(0xffffffffu, r[i], 1);
        r[i] += __shfl_xor_sync(0xffffffffu, r[i], 2);
      }
      if (has_body & 1) {  // Fe_a = b * sum_g w detJ N_a(g)
        double nw = 0.0;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const double gx = 1.0 + sgn_x(ra) * sgn_x(g) * FOL_S3, gy = 1.0 + sgn_y(ra) * sgn_y(g) * FOL_S3;
          const double gz = 1.0 + sgn_z(ra) * sgn_z(g) * FOL_S3;
          nw += sm.wdj[el * 8 + g] * (0.125 * gx * gy * gz);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) r[i] -= args.p.v[2 + i] * nw;
      }
      // Dirichlet row mask (fe_loss.py:191-207): only for elements touching a fixed dof (warp-uniform test)
      const bool fx[3] = {sm.bc[el][ra * 3 + 0] == 0, sm.bc[el][ra * 3 + 1] == 0, sm.bc[el][ra * 3 + 2] == 0};
      const bool fixed_rows = fx[0] | fx[1] | fx[2];
      const bool any_fixed = __any_sync(0xffffffffu, fixed_rows);
      // the copies that last used the staging slot / the history buffer are done (LAYOUT 2: every lane executes the
      // wait -- lanes other than 0 have no copies of their own -- instead of branching around it)
      if (LAYOUT == 2 || lane == 0) bulk_wait_read<0>();
      __syncwarp();
      if (!any_fixed) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double2* dst = reinterpret_cast<double2*>(sm.stage + (ra * 3 + i) * 24 + kq * 6);
          dst[0] = make_double2(K[i][0][0], K[i][1][0]);
          dst[1] = make_double2(K[i][2][0], K[i][0][1]);
          dst[2] = make_double2(K[i][1][1], K[i][2][1]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int row = ra * 3 + i;
          const bool freerow = LAYOUT == 2 ? !fx[i] : sm.bc[el][row] != 0;
          double v[6];
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const int col = (2 * kq + h) * 3 + j;
              v[h * 3 + j] = (freerow || col == row) ? K[i][j][h] : 0.0;
            }
          double2* dst = reinterpret_cast<double2*>(sm.stage + row * 24 + kq * 6);
          dst[0] = make_double2(v[0], v[1]);
          dst[1] = make_double2(v[2], v[3]);
          dst[2] = make_double2(v[4], v[5]);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) bulk_store(args.ke + e * 576, sm.stage, 576 * sizeof(double));
      if (kq == 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          if constexpr (LAYOUT == 2) args.re[e * 24 + ra * 3 + i] = fx[i] ? 0.0 : r[i];
          else args.re[e * 24 + ra * 3 + i] = (any_fixed && sm.bc[el][ra * 3 + i] == 0) ? 0.0 : r[i];
        }
      }
    }
    __syncwarp();  // everyone is done with X / u / history / gradients of this tile
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

template <bool FUSE, int LAYOUT>
static int launch_hex_j2(cudaStream_t s, const AsmArgs<double>& args, const HaloFuse* hf) {
  constexpr int kW = j2_warps_of<LAYOUT>();
  static PerDeviceGrid per_device;
  const size_t smem = sizeof(WarpSmemJ2T<LAYOUT>) * kW;
  int g = 0;
  FOL_CUDA(per_device.get(assemble_hex_j2_f64_kernel<FUSE, LAYOUT>, kW * 32, smem, &g));
  if (args.ne == 0) return FOL_OK;
  const long long ntiles = cdiv(args.ne, kTile);
  const long long want = cdiv(ntiles, kW);
  const int has_body = (args.p.v[2] != 0.0 || args.p.v[3] != 0.0 || args.p.v[4] != 0.0) ? 1 : 0;
  const unsigned blocks = (unsigned)(want < g ? want : g);
  assemble_hex_j2_f64_kernel<FUSE, LAYOUT><<<blocks, kW * 32, smem, s>>>(args, ntiles, has_body, hf ? *hf : HaloFuse{});
  return check_launch("assemble_hex_j2_f64_kernel");
}

int assemble_hex_j2_f64(cudaStream_t s, const AsmArgs<double>& args, const HaloFuse* hf) {
  // default: layout 2 (compact, 12 warps / SM, lean hand-off: 3.91 -> 3.87 ms sustained on the same box);
  // FOL_J2_LAYOUT=0 selects the 10-warp layout, 1 the 12-warp layout with the original hand-off, for A/B runs
  static const int layout = [] { const char* v = std::getenv("FOL_J2_LAYOUT"); return v ? std::atoi(v) : 2; }();
  if (layout == 0) return hf ? launch_hex_j2<true, 0>(s, args, hf) : launch_hex_j2<false, 0>(s, args, hf);
  if (layout == 1) return hf ? launch_hex_j2<true, 1>(s, args, hf) : launch_hex_j2<false, 1>(s, args, hf);
  return hf ? launch_hex_j2<true, 2>(s, args, hf) : launch_hex_j2<false, 2>(s, args, hf);
}

}  // namespace fol
