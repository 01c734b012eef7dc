// Tuned element-stage kernel for 3-D Hex8 heat conduction, float64, 2x2x2 Gauss rule (ThermalLoss3DHexa:
// thermal.py:28-49 + fe_loss.py:191-230, 299): Se = sum_g w detJ (N.K)(1 + beta (N.T)^c) grad N grad N^T, re = Se T.
//
// Same results as the generic kernel (assemble.cuh), which runs this case at 0.13 of the HBM roofline -- its row-block
// accumulation re-reads the staged gradients from shared memory 8 times (90 % of the shared-memory wavefront peak,
// profiles/r1/sweep_other_configs_ncu.txt) and evaluates the conductivity law with pow().  Mapping of the tuned Hex8
// elasticity kernel (assemble_hex.cu), with a much smaller phase 2:
//   * persistent warps, tiles of 4 consecutive elements; nodal gathers of tile i+1 (connectivity of tile i+2) are in
//     flight (8-byte cp.async into SoA rows) while tile i computes;
//   * phase 1: lane (element, Gauss point) -> J, det J, grad N, coefficient -- 32 independent geometry evaluations per
//     warp, nothing computed twice;
//   * phase 2: per element the 8x8 matrix sum_{g,d} (s_g grad_d N_a)(grad_d N_b) is an 8x8x24 GEMM: 6 DMMA m8n8k4 fed
//     straight from the staged gradients; lane (a, k) ends up with Se[a][2k], Se[a][2k+1], applies the Dirichlet row
//     mask in registers and stores 16 bytes -- the warp writes the element's 512 contiguous bytes in one instruction;
//     re = Se T by a 4-lane butterfly.
// Se is symmetric only to rounding (the scale rides on the A operand), so transpose_jacobian=True goes to the generic
// kernel, which transposes exactly.  Algorithmic bytes: 512 (Se) + 64 (re) + 32 (connectivity) + 41 (nodal data, each
// node shared by 8 elements) = 649 B per element.
#include "assemble.cuh"
#include "assemble_hex_common.cuh"
#include "energy.cuh"

namespace fol {

namespace {

using hexk::kTile;
using namespace hexk;

constexpr int kWarpsTh = 8;   // warps per CTA, each fully independent (2 CTAs = 16 warps / SM)

struct __align__(128) ThermalWarpSmem {
  // [element][gauss][node ^ swz(gauss)]: (dN/dx, dN/dy) and (dN/dz, w detJ kappa_g); same XOR swizzle as assemble_hex.cu
  double2 gxy[kTile][8][8];
  double2 gzs[kTile][8][8];
  // nodal data of the tile, SoA over (element, node), double-buffered.  An element's 8 values are followed by 2 pad
  // values: phase 1 reads them as four 16-byte pairs, the same pair in all lanes of an element -- 4 addresses per
  // warp, 80 bytes apart, i.e. on distinct banks (64 bytes apart would be a 2-way conflict on every load)
  double X[2][3][kTile * 10];
  double T[2][kTile * 10];
  double de[2][kTile * 10];
  float bc[kTile][8];                // 1 = free dof, 0 = Dirichlet dof
};

}  // namespace

__global__ void __launch_bounds__(kWarpsTh * 32, 2)
assemble_hex_thermal_f64_kernel(const AsmArgs<double> args, const long long ntiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ThermalWarpSmem& sm = reinterpret_cast<ThermalWarpSmem*>(smem_raw)[warp];
  const long long nwarps = (long long)gridDim.x * kWarpsTh;
  long long vt = (long long)blockIdx.x * kWarpsTh + warp;
  if (vt >= ntiles) return;
  const double beta = args.p.v[5], cexp = args.p.v[6];

  // lane roles
  const int el_p = lane >> 3, sub = lane & 7;   // phases 0/1: (element in tile, node | gauss point)
  const int ra = lane >> 2, kq = lane & 3;      // phase 2: (row node a, column pair k)
  const int swz_p = ((sub & 3) << 1) | (sub >> 2);

  // Gauss point `sub` of the 2x2x2 rule (hexahedra_3d_8.py:23-33): xi = sgn(sub) / sqrt(3), w = 1; trilinear shape
  // data factorised per axis as in assemble_hex.cu
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

  // software pipeline of the gathers: node ids two tiles ahead, nodal data one tile ahead (see assemble_hex.cu for
  // why the id is loaded by a predicated load and not touched until it is used)
  auto node_of = [&](long long t) -> int {
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
  auto hold_back = [](int& a, unsigned& b) { asm volatile("" : "+r"(a), "+r"(b)); };
  auto gather_async = [&](int buf, long long n) {
    const double* pxyz = args.xyz + n * 3;
    const int slot = el_p * 10 + sub;
    cp_async8(&sm.X[buf][0][slot], pxyz);
    cp_async8(&sm.X[buf][1][slot], pxyz + 1);
    cp_async8(&sm.X[buf][2][slot], pxyz + 2);
    cp_async8(&sm.T[buf][slot], args.u + n);
    cp_async8(&sm.de[buf][slot], args.ctrl + n);
    cp_async_commit();
  };

  int n_next = node_of(vt + nwarps);
  const long long n_first = node_of(vt);
  gather_async(0, n_first);
  unsigned f0 = __ldg(args.dir + n_first);
  int buf = 0;

  for (; vt < ntiles; vt += nwarps, buf ^= 1) {
    const long long e0 = vt * kTile;
    hold_back(n_next, f0);   // loaded one tile ago; nothing may consume them before this point

    // ---- phase 0: this tile's nodal data has landed in shared memory; start the next gather
    sm.bc[el_p][sub] = f0 ? 0.f : 1.f;
    cp_async_wait_all();
    __syncwarp();
    gather_async(buf ^ 1, (long long)n_next);
    f0 = __ldg(args.dir + (long long)n_next);
    n_next = node_of(vt + 2 * nwarps);

    // ---- phase 1: lane (element, Gauss point): J, det J, grad N, coefficient (geometry.py:88-97, thermal.py:31-38)
    {
      double J[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double j0 = 0.0, j1 = 0.0, j2 = 0.0;
        const double2* xr = reinterpret_cast<const double2*>(&sm.X[buf][i][el_p * 10]);
#pragma unroll
        for (int a2 = 0; a2 < 4; ++a2) {
          const double2 xp = xr[a2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int a = 2 * a2 + h;
            const int bx = ((a & 3) == 1 || (a & 3) == 2), by = (a >> 1) & 1, bz = (a >> 2) & 1;
            const double x = h ? xp.y : xp.x;
            j0 += (bx ? x : -x) * fyz[by][bz];
            j1 += (by ? x : -x) * fxz[bx][bz];
            j2 += (bz ? x : -x) * fxy[bx][by];
          }
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
      double eg = 0.0, tg = 0.0;
      const double2* dr = reinterpret_cast<const double2*>(&sm.de[buf][el_p * 10]);
      const double2* tr = reinterpret_cast<const double2*>(&sm.T[buf][el_p * 10]);
#pragma unroll
      for (int a2 = 0; a2 < 4; ++a2) {
        const double2 dp = dr[a2], tp = tr[a2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int a = 2 * a2 + h;
          const int bx = ((a & 3) == 1 || (a & 3) == 2), by = (a >> 1) & 1, bz = (a >> 2) & 1;
          const double N = fx[bx] * fyz[by][bz];
          eg += N * (h ? dp.y : dp.x);
          tg += N * (h ? tp.y : tp.x);
        }
      }
      const double nl = (beta != 0.0) ? beta * pow_c<double>(tg, cexp) : 0.0;   // thermal.py:34
      const double coef = det * eg * (1.0 + nl);                                  // Gauss weight is 1
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
        sm.gzs[el_p][sub][a ^ swz_p] = make_double2(g[2], coef);
      }
    }
    __syncwarp();

    // ---- phase 2: one element at a time, lane (a, k): Se[a][2k], Se[a][2k + 1]
#pragma unroll
    for (int el = 0; el < kTile; ++el) {
      const long long e = e0 + el;
      if (e >= args.ne) break;
      double c0 = 0.0, c1 = 0.0;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const double2 xy = sm.gxy[el][4 * kk + kq][ra ^ ((kq << 1) | kk)];
        const double2 zs = sm.gzs[el][4 * kk + kq][ra ^ ((kq << 1) | kk)];
        dmma884(c0, c1, zs.y * xy.x, xy.x);
        dmma884(c0, c1, zs.y * xy.y, xy.y);
        dmma884(c0, c1, zs.y * zs.x, zs.x);
      }
      // re = Se T: partial over this lane's two columns, then butterfly over the 4 k-lanes
      const double2 tq = *reinterpret_cast<const double2*>(&sm.T[buf][el * 10 + 2 * kq]);
      double r = c0 * tq.x + c1 * tq.y;
      r += __shfl_xor_sync(0xffffffffu, r, 1);
      r += __shfl_xor_sync(0xffffffffu, r, 2);
      // Dirichlet row mask (fe_loss.py:191-207): a fixed row keeps its diagonal entry only, its residual is zero
      const bool freerow = sm.bc[el][ra] != 0.f;
      double2 out;
      out.x = (freerow || 2 * kq == ra) ? c0 : 0.0;
      out.y = (freerow || 2 * kq + 1 == ra) ? c1 : 0.0;
      __stcs(reinterpret_cast<double2*>(args.ke + e * 64 + ra * 8 + 2 * kq), out);
      if (kq == 0) args.re[e * 8 + ra] = freerow ? r : 0.0;
    }
    __syncwarp();  // everyone is done with the gradients / nodal data of this tile
  }
  cp_async_wait_all();
}

int assemble_hex_thermal_f64(cudaStream_t s, const AsmArgs<double>& args) {
  static PerDeviceGrid per_device;
  const size_t smem = sizeof(ThermalWarpSmem) * kWarpsTh;
  int grid = 0;
  FOL_CUDA(per_device.get(assemble_hex_thermal_f64_kernel, kWarpsTh * 32, smem, &grid));
  if (args.ne == 0) return FOL_OK;
  if ((reinterpret_cast<uintptr_t>(args.ke) & 15) != 0) return 1;   // 16-byte stores: the generic kernel takes it
  const long long ntiles = cdiv(args.ne, kTile);
  const long long want = cdiv(ntiles, kWarpsTh);
  const unsigned blocks = (unsigned)(want < grid ? want : grid);
  assemble_hex_thermal_f64_kernel<<<blocks, kWarpsTh * 32, smem, s>>>(args, ntiles);
  return check_launch("assemble_hex_thermal_f64_kernel");
}

}  // namespace fol
