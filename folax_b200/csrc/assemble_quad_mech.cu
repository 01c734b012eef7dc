// Tuned element-stage kernel for 2-D Quad4 small-strain elasticity (plane stress), float64, 2x2 Gauss rule
// (MechanicalLoss2DQuad, mechanical.py:98-117 + fe_loss.py:191-230, 299; BASELINE.json configs[0]'s element).
//
// Same results as the generic kernel (assemble.cuh), which runs this case at 0.35 of the HBM roofline, bound by the
// shared-memory pipe (87 % of the wavefront peak: every lane re-reads the staged gradients of all nodes and points).
// Mapping of the tuned Hex8 kernels (assemble_hex.cu, assemble_hex_thermal.cu), one dimension down:
//   * persistent warps, tiles of 8 consecutive elements; nodal gathers of tile i+1 (connectivity of tile i+2) are in
//     flight (8-byte cp.async into SoA rows) while tile i computes;
//   * phase 1: lane (element, Gauss point) -> J, det J, grad N, coefficient w detJ (N.K): 32 independent geometry
//     evaluations per warp, nothing computed twice;
//   * phase 2: per element the 8x8 matrix P = sum_g s_g v_g v_g^T (v = grad N flattened, index 2 node + dim) is an
//     8x8x4 GEMM: ONE DMMA m8n8k4 fed from the staged gradients; lane (m, k) holds row m = (a, i) and the block column
//     b = k: P_ab[i][0..1]; the transposed entries and the trace needed by Ke_ab = lam P_ab + mu P_ab^T + mu tr(P_ab) I
//     (B^T D B of the isotropic plane-stress D) sit in the lane of the other dimension (m ^ 1): two 64-bit shuffles;
//     re = Ke u - Fe by a 4-lane butterfly, Dirichlet row mask in registers; the warp stores the element's 512
//     contiguous bytes with one 16-byte store per lane.
// Ke is symmetric only to rounding (the scale rides on the A operand), so transpose_jacobian=True goes to the generic
// kernel, which transposes exactly.  Algorithmic bytes (scripts/sweep_bench.py): 512 (Ke) + 16 (connectivity) + 48
// (nodal data, one node per element) = 576 B per element.
#include "assemble.cuh"
#include "assemble_hex_common.cuh"

namespace fol {

namespace {

using namespace hexk;

constexpr int kWarpsQ = 8;    // warps per CTA, each fully independent
constexpr int kTileQ = 8;     // elements per warp iteration: 8 elements x 4 Gauss points = 32 lanes in phase 1

struct __align__(128) QuadWarpSmem {
  // [element][gauss][m ^ swz(element, gauss)], m = 2 node + dim: grad N.  swz = ((gauss >> 1) << 2) | (element & 3)
  // keeps both the phase-1 stores (lane = (element, gauss), same m in all lanes) and the phase-2 operand loads
  // (lane = (m, gauss)) free of bank conflicts -- unswizzled, the stores were 8-way conflicted (64-byte lane stride)
  double v[kTileQ][4][8];
  double coef[kTileQ][4];
  double wd[kTileQ][4];              // w detJ per Gauss point (body force)
  // nodal data of the tile, SoA over (element, node) = 32 lanes, double-buffered
  double X[2][2][32];
  double2 u[2][32];                  // (ux, uy) of a node: one 16-byte cp.async, one 16-byte load in phase 2
  double de[2][32];
  float bc[kTileQ][8];               // 1 = free dof, 0 = Dirichlet dof
};

}  // namespace

__global__ void __launch_bounds__(kWarpsQ * 32, 3)
assemble_quad_mech_f64_kernel(const AsmArgs<double> args, const long long ntiles, const int has_body) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  QuadWarpSmem& sm = reinterpret_cast<QuadWarpSmem*>(smem_raw)[warp];
  const long long nwarps = (long long)gridDim.x * kWarpsQ;
  long long vt = (long long)blockIdx.x * kWarpsQ + warp;
  if (vt >= ntiles) return;

  // plane stress (mechanical.py:60-70): D = E/(1-nu^2) [[1, nu, 0], [nu, 1, 0], [0, 0, (1-nu)/2]] = lam, mu form
  const double E = args.p.v[0], nu = args.p.v[1];
  const double f = E / (1.0 - nu * nu);
  const double lam = f * nu, mu = f * (1.0 - nu) * 0.5;

  // lane roles
  const int el_p = lane >> 2, sub = lane & 3;   // phases 0/1: (element in tile, node | gauss point)
  const int ra = lane >> 2, kq = lane & 3;      // phase 2: row m = ra = (node a = ra >> 1, dim i = ra & 1), block column b = kq
  const int ia = ra & 1;

  // Gauss point `sub` of the 2x2 rule (quadrilateral_2d_4.py:54-58): xi = sgn(sub) / sqrt(3), w = 1
  const double px = sgn_x(sub) * FOL_S3, py = sgn_y(sub) * FOL_S3;
  double N[4], dNx[4], dNy[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double fx = 1.0 + sgn_x(a) * px, fy = 1.0 + sgn_y(a) * py;
    N[a] = 0.25 * fx * fy;
    dNx[a] = 0.25 * sgn_x(a) * fy;
    dNy[a] = 0.25 * sgn_y(a) * fx;
  }

  auto node_of = [&](long long t) -> int {
    const long long e = t * kTileQ + el_p;
    const int ok = (t < ntiles && e < args.ne) ? 1 : 0;
    const int32_t* src = args.conn + (ok ? e * 4 + sub : 0);
    int n;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\tmov.s32 %0, 0;\n\t@p ld.global.nc.s32 %0, [%1];\n\t}\n"
        : "=r"(n)
        : "l"(src), "r"(ok));
    return n;
  };
  auto hold_back = [](int& a, unsigned& b, unsigned& c) { asm volatile("" : "+r"(a), "+r"(b), "+r"(c)); };
  auto gather_async = [&](int buf, long long n) {
    const double* pxy = args.xyz + n * 3;       // coordinates are stored (nn, 3) for every element type
    const double* pu = args.u + n * 2;
    cp_async8(&sm.X[buf][0][lane], pxy);
    cp_async8(&sm.X[buf][1][lane], pxy + 1);
    cp_async16(&sm.u[buf][lane], pu);            // 16-byte aligned: two dofs per node
    cp_async8(&sm.de[buf][lane], args.ctrl + n);
    cp_async_commit();
  };

  int n_next = node_of(vt + nwarps);
  const long long n_first = node_of(vt);
  gather_async(0, n_first);
  unsigned f0 = __ldg(args.dir + n_first * 2), f1 = __ldg(args.dir + n_first * 2 + 1);
  int buf = 0;

  for (; vt < ntiles; vt += nwarps, buf ^= 1) {
    const long long e0 = vt * kTileQ;
    hold_back(n_next, f0, f1);   // loaded one tile ago; nothing may consume them before this point

    // ---- phase 0: this tile's nodal data has landed in shared memory; start the next gather
    sm.bc[el_p][sub * 2 + 0] = f0 ? 0.f : 1.f;
    sm.bc[el_p][sub * 2 + 1] = f1 ? 0.f : 1.f;
    cp_async_wait_all();
    __syncwarp();
    gather_async(buf ^ 1, (long long)n_next);
    f0 = __ldg(args.dir + (long long)n_next * 2);
    f1 = __ldg(args.dir + (long long)n_next * 2 + 1);
    n_next = node_of(vt + 2 * nwarps);

    // ---- phase 1: lane (element, Gauss point): J, det J, grad N, coefficient (geometry.py:88-97)
    {
      double J00 = 0.0, J01 = 0.0, J10 = 0.0, J11 = 0.0, eg = 0.0;   // J[i][j] = d x_i / d xi_j
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const double x = sm.X[buf][0][el_p * 4 + a], y = sm.X[buf][1][el_p * 4 + a];
        J00 += x * dNx[a]; J01 += x * dNy[a];
        J10 += y * dNx[a]; J11 += y * dNy[a];
        eg += N[a] * sm.de[buf][el_p * 4 + a];
      }
      const double det = J00 * J11 - J01 * J10;
      const double rd = 1.0 / det;
      // grad N_a = dN_a . J^-1:  d/dx = (dNx J11 - dNy J10) / det,  d/dy = (dNy J00 - dNx J01) / det
      const int swz_p = ((sub >> 1) << 2) | (el_p & 3);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        sm.v[el_p][sub][(2 * a + 0) ^ swz_p] = (dNx[a] * J11 - dNy[a] * J10) * rd;
        sm.v[el_p][sub][(2 * a + 1) ^ swz_p] = (dNy[a] * J00 - dNx[a] * J01) * rd;
      }
      sm.coef[el_p][sub] = det * eg;   // Gauss weight is 1
      sm.wd[el_p][sub] = det;
    }
    __syncwarp();

    // ---- phase 2: one element at a time, lane (m, k)
#pragma unroll
    for (int el = 0; el < kTileQ; ++el) {
      const long long e = e0 + el;
      if (e >= args.ne) break;
      double c0 = 0.0, c1 = 0.0;                         // P[m][2k], P[m][2k+1] = P_ab[i][0], P_ab[i][1], b = k
      {
        const double vm = sm.v[el][kq][ra ^ (((kq >> 1) << 2) | (el & 3))];   // A[m = ra][g = kq] (scaled), B[g = kq][n = ra]
        dmma884(c0, c1, sm.coef[el][kq] * vm, vm);
      }
      // the other row of the 2x2 block P_ab sits in lane m ^ 1 (same node a, other dimension): lane ^ 4
      const double o0 = __shfl_xor_sync(0xffffffffu, c0, 4), o1 = __shfl_xor_sync(0xffffffffu, c1, 4);
      // i = 0: P = [[c0, c1], [o0, o1]];  i = 1: P = [[o0, o1], [c0, c1]]
      const double tr = ia ? o0 + c1 : c0 + o1;
      const double pt0 = ia ? o1 : c0;                   // P_ab^T[i][0] = P_ab[0][i]
      const double pt1 = ia ? c1 : o0;                   // P_ab^T[i][1] = P_ab[1][i]
      double k0 = lam * c0 + mu * pt0, k1 = lam * c1 + mu * pt1;
      if (ia) k1 += mu * tr; else k0 += mu * tr;
      // re = Ke u - Fe: partial over this lane's two columns (node b = kq), then butterfly over the 4 k-lanes
      const double2 ub = sm.u[buf][el * 4 + kq];
      double r = k0 * ub.x + k1 * ub.y;
      r += __shfl_xor_sync(0xffffffffu, r, 1);
      r += __shfl_xor_sync(0xffffffffu, r, 2);
      if (has_body) {                                    // Fe_(a,i) = b_i sum_g w detJ N_a(g)   (mechanical.py:110)
        const int a = ra >> 1;
        double nw = 0.0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const double gx = 1.0 + sgn_x(a) * sgn_x(g) * FOL_S3, gy = 1.0 + sgn_y(a) * sgn_y(g) * FOL_S3;
          nw += sm.wd[el][g] * (0.25 * gx * gy);
        }
        r -= args.p.v[2 + ia] * nw;
      }
      // Dirichlet row mask (fe_loss.py:191-207): a fixed row keeps its diagonal entry only, its residual is zero
      const bool freerow = sm.bc[el][ra] != 0.f;
      double2 out;
      out.x = (freerow || 2 * kq == ra) ? k0 : 0.0;
      out.y = (freerow || 2 * kq + 1 == ra) ? k1 : 0.0;
      __stcs(reinterpret_cast<double2*>(args.ke + e * 64 + ra * 8 + 2 * kq), out);
      if (kq == 0) args.re[e * 8 + ra] = freerow ? r : 0.0;
    }
    __syncwarp();  // everyone is done with the gradients / nodal data of this tile
  }
  cp_async_wait_all();
}

int assemble_quad_mech_f64(cudaStream_t s, const AsmArgs<double>& args) {
  static PerDeviceGrid per_device;
  const size_t smem = sizeof(QuadWarpSmem) * kWarpsQ;
  int grid = 0;
  FOL_CUDA(per_device.get(assemble_quad_mech_f64_kernel, kWarpsQ * 32, smem, &grid));
  if (args.ne == 0) return FOL_OK;
  if (((reinterpret_cast<uintptr_t>(args.ke) | reinterpret_cast<uintptr_t>(args.u)) & 15) != 0)
    return 1;                        // 16-byte stores / dof gathers: the generic kernel takes unaligned buffers
  const long long ntiles = cdiv(args.ne, kTileQ);
  const long long want = cdiv(ntiles, kWarpsQ);
  const unsigned blocks = (unsigned)(want < grid ? want : grid);
  const int has_body = (args.p.v[2] != 0.0 || args.p.v[3] != 0.0) ? 1 : 0;
  assemble_quad_mech_f64_kernel<<<blocks, kWarpsQ * 32, smem, s>>>(args, ntiles, has_body);
  return check_launch("assemble_quad_mech_f64_kernel");
}

}  // namespace fol
