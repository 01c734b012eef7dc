// Tuned element-stage kernel for 3-D Hex8 small-strain elasticity in FLOAT32, 2x2x2 Gauss rule (MechanicalLoss3DHexa
// with loss_settings["dtype"] = "float32"; mechanical.py:98-117 + fe_loss.py:191-230, 299).  north_star names float32
// (1e-5) as a first-class precision next to float64; the generic kernel reached 0.40 of the HBM roofline there.
//
// Same machine mapping as assemble_hex.cu -- persistent warps, tiles of 4 consecutive elements, cp.async gathers one
// tile ahead, one bulk (TMA-engine) copy per element matrix -- with the float64 DMMA product replaced by 3xTF32
// tensor-core products (each fp32 operand split into a tf32 head and a tf32 tail; hi*hi + hi*lo + lo*hi accumulated in
// fp32 recovers fp32-grade products, which one TF32 pass -- 10 mantissa bits -- would not):
//   * phase 1, lane (element, Gauss point): J, det J, grad N, coefficient in float32; staged per (element, point, node)
//     as ONE float4 (dN/dx, dN/dy, dN/dz, w detJ E_g), XOR-swizzled like the float64 kernel (16-byte cells, 128-byte rows);
//   * phase 2, lane (row node a, column pair k): P = sum_g s_g v_g v_g^T as 6 m16n8k8 tile positions x 3 split terms =
//     18 mma.sync per element, fed from two LDS.128 per lane; the accumulator fragment is the lane's 3x3 blocks
//     (a, 2k), (a, 2k+1); Ke = lam P + mu P^T + mu tr(P) I, re = Ke u - Fe (butterfly over the 4 k-lanes), Dirichlet mask in registers;
//   * the 2304-byte element matrix is staged in one of two shared-memory slots and leaves as one cp.async.bulk.
// Algorithmic bytes: 2376 B/element (2304 Ke + 32 conn + 28 nodal in + 12 residual out).
#include "assemble.cuh"
#include "assemble_hex_common.cuh"

namespace fol {

namespace {

using namespace hexk;

constexpr int kWarpsF = 8;           // warps per CTA, each fully independent (2 CTAs = 16 warps / SM)

struct __align__(128) WarpSmemF {
  float stage[2][576];               // two Ke staging slots (bulk-copy sources), 2304 B each: the copy of element i may
                                     // drain until element i + 2 needs its slot
  float4 g[kTile][8][8];             // [element][gauss][node ^ swz(gauss)]: (dN/dx, dN/dy, dN/dz, w detJ E_g)
  float X[2][3][32];                 // nodal data of the tile, SoA over the 32 (element, node) lanes, double-buffered
  float u[2][3][32];
  float de[2][32];
  float wd[kTile][8];                // w detJ per Gauss point (body force)
  uint8_t bc[kTile][24];             // 1 = free dof, 0 = Dirichlet dof
};
static_assert(2 * (sizeof(WarpSmemF) * kWarpsF + 1024) <= 227 * 1024, "two CTAs per SM must fit");

__device__ __forceinline__ void cp_async4(void* sdst, const void* gsrc) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(saddr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void bulk_store_f(float* gdst, const float* ssrc, unsigned bytes) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(saddr), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}

// float -> (hi, lo) tf32 pair: x = hi + lo to ~2^-21 relative (3xTF32: hi*hi + hi*lo + lo*hi recovers fp32-grade products
// on the tensor path; the lo*lo term is below fp32 rounding)
// (round-to-nearest, ties away, done on the bit pattern: add half an ulp of the 10-bit mantissa and clear the 13 low
// bits -- what cvt.rna.tf32.f32 computes for finite values, in 2 integer instructions instead of the ~6 the compiler
// emits for the cvt on sm_100a, which showed as a quarter of the kernel's instructions)
__device__ __forceinline__ unsigned round_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void split_tf32(float x, unsigned& hi, unsigned& lo) {
  hi = round_tf32(x);
  lo = round_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

}  // namespace

__global__ void __launch_bounds__(kWarpsF * 32, 2)
assemble_hex_mech_f32_kernel(const AsmArgs<float> args, const long long ntiles, const int has_body) {
  extern __shared__ __align__(128) unsigned char smem_raw_f32[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpSmemF& sm = reinterpret_cast<WarpSmemF*>(smem_raw_f32)[warp];
  const long long nwarps = (long long)gridDim.x * kWarpsF;
  long long tile = (long long)blockIdx.x * kWarpsF + warp;
  if (tile >= ntiles) return;

  const float E = args.p.v[0], nu = args.p.v[1];
  const float c1 = E / ((1.0f + nu) * (1.0f - 2.0f * nu));
  const float lam = c1 * nu, mu = c1 * 0.5f * (1.0f - 2.0f * nu);

  const int el_p = lane >> 3, sub = lane & 7;   // phases 0/1: (element in tile, node | Gauss point)
  const int ra = lane >> 2, kq = lane & 3;      // phase 2: (row node a, column pair k)
  const int swz_p = ((sub & 3) << 1) | (sub >> 2);

  // Gauss point `sub` of the 2x2x2 rule (hexahedra_3d_8.py:23-33): the shape data are evaluated in double at compile
  // time / once per thread and rounded to float, as the reference's float32 tables are
  const double px = sgn_x(sub) * FOL_S3, py = sgn_y(sub) * FOL_S3, pz = sgn_z(sub) * FOL_S3;
  const float fx[2] = {(float)(1.0 - px), (float)(1.0 + px)};
  float fyz[2][2], fxz[2][2], fxy[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      fyz[i][j] = (float)(0.125 * (i ? 1.0 + py : 1.0 - py) * (j ? 1.0 + pz : 1.0 - pz));
      fxz[i][j] = (float)(0.125 * (i ? 1.0 + px : 1.0 - px) * (j ? 1.0 + pz : 1.0 - pz));
      fxy[i][j] = (float)(0.125 * (i ? 1.0 + px : 1.0 - px) * (j ? 1.0 + py : 1.0 - py));
    }

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
  auto hold_back = [](int& a, unsigned& b, unsigned& c, unsigned& d) {
    asm volatile("" : "+r"(a), "+r"(b), "+r"(c), "+r"(d));
  };
  auto gather_async = [&](int buf, long long n) {
    const float* gx = args.xyz + n * 3;
    const float* gu = args.u + n * 3;
    cp_async4(&sm.X[buf][0][lane], gx); cp_async4(&sm.X[buf][1][lane], gx + 1); cp_async4(&sm.X[buf][2][lane], gx + 2);
    cp_async4(&sm.u[buf][0][lane], gu); cp_async4(&sm.u[buf][1][lane], gu + 1); cp_async4(&sm.u[buf][2][lane], gu + 2);
    cp_async4(&sm.de[buf][lane], args.ctrl + n);
    cp_async_commit();
  };
  int n_next = node_of(tile + nwarps);
  const long long n_first = node_of(tile);
  gather_async(0, n_first);
  const uint8_t* pf0 = args.dir + n_first * 3;
  unsigned f0 = __ldg(pf0), f1 = __ldg(pf0 + 1), f2 = __ldg(pf0 + 2);
  int buf = 0;

  for (; tile < ntiles; tile += nwarps, buf ^= 1) {
    const long long e0 = tile * kTile;
    hold_back(n_next, f0, f1, f2);

    // ---- phase 0: this tile's nodal data has landed; start the next gather
    sm.bc[el_p][sub * 3 + 0] = f0 ? 0 : 1;
    sm.bc[el_p][sub * 3 + 1] = f1 ? 0 : 1;
    sm.bc[el_p][sub * 3 + 2] = f2 ? 0 : 1;
    cp_async_wait_all();
    __syncwarp();
    gather_async(buf ^ 1, (long long)n_next);
    {
      const uint8_t* pf = args.dir + (long long)n_next * 3;
      f0 = __ldg(pf); f1 = __ldg(pf + 1); f2 = __ldg(pf + 2);
    }
    n_next = node_of(tile + 2 * nwarps);

    // ---- phase 1: lane (element, Gauss point): J, det J, grad N, coefficient (geometry.py:88-97)
    {
      float J[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float j0 = 0.f, j1 = 0.f, j2 = 0.f;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          const int bx = ((a & 3) == 1 || (a & 3) == 2), by = (a >> 1) & 1, bz = (a >> 2) & 1;
          const float x = sm.X[buf][i][el_p * 8 + a];
          j0 += (bx ? x : -x) * fyz[by][bz];
          j1 += (by ? x : -x) * fxz[bx][bz];
          j2 += (bz ? x : -x) * fxy[bx][by];
        }
        J[i][0] = j0; J[i][1] = j1; J[i][2] = j2;
      }
      const float c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
      const float c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
      const float c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
      const float det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
      const float rd = 1.0f / det;
      float inv[3][3];
      inv[0][0] = c00 * rd; inv[1][0] = c01 * rd; inv[2][0] = c02 * rd;
      inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * rd;
      inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * rd;
      inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * rd;
      inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * rd;
      inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * rd;
      inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * rd;
      float eg = 0.f;
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int bx = ((a & 3) == 1 || (a & 3) == 2), by = (a >> 1) & 1, bz = (a >> 2) & 1;
        eg += fx[bx] * fyz[by][bz] * sm.de[buf][el_p * 8 + a];
      }
      const float wd = det;  // Gauss weight is 1
      const float coef = wd * eg;
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int bx = ((a & 3) == 1 || (a & 3) == 2), by = (a >> 1) & 1, bz = (a >> 2) & 1;
        const float d0 = bx ? fyz[by][bz] : -fyz[by][bz];
        const float d1 = by ? fxz[bx][bz] : -fxz[bx][bz];
        const float d2 = bz ? fxy[bx][by] : -fxy[bx][by];
        float g[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) g[k] = d0 * inv[0][k] + d1 * inv[1][k] + d2 * inv[2][k];
        sm.g[el_p][sub][a ^ swz_p] = make_float4(g[0], g[1], g[2], coef);
      }
      sm.wd[el_p][sub] = wd;
    }
    __syncwarp();

    // ---- phase 2: one element at a time, lane (a, k)
#pragma unroll 1
    for (int el = 0; el < kTile; ++el) {
      const long long e = e0 + el;
      if (e >= args.ne) break;
      // P = sum_g s_g v_g v_g^T (24 x 24 x 8) on the tensor path with fp32-grade accuracy (3xTF32): rows ordered
      // m = 8 * dim + node, so an m16n8k8 tile covers dims (0, 1) of the 8 row nodes against one dim of the 8 column
      // nodes and a second tile covers dim 2 (upper half unused); its accumulator fragment of lane (a, k) is exactly the
      // 3x3 blocks (a, 2k), (a, 2k + 1).  The lane feeds ONLY its own node a at Gauss points k and k + 4 (two LDS.128).
      float c[3][3][2];
      {
        const float4 g0 = sm.g[el][kq][ra ^ (kq << 1)];                 // swz(g) = ((g & 3) << 1) | (g >> 2)
        const float4 g1 = sm.g[el][4 + kq][ra ^ ((kq << 1) | 1)];
        const float bv[3][2] = {{g0.x, g1.x}, {g0.y, g1.y}, {g0.z, g1.z}};
        unsigned ahi[3][2], alo[3][2], bhi[3][2], blo[3][2];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          split_tf32(g0.w * bv[t][0], ahi[t][0], alo[t][0]);
          split_tf32(g1.w * bv[t][1], ahi[t][1], alo[t][1]);
          split_tf32(bv[t][0], bhi[t][0], blo[t][0]);
          split_tf32(bv[t][1], bhi[t][1], blo[t][1]);
        }
        // A fragments: a0 (row a, k), a1 (row a + 8, k), a2 (row a, k + 4), a3 (row a + 8, k + 4)
        const unsigned A01h[4] = {ahi[0][0], ahi[1][0], ahi[0][1], ahi[1][1]}, A01l[4] = {alo[0][0], alo[1][0], alo[0][1], alo[1][1]};
        const unsigned A2h[4] = {ahi[2][0], 0u, ahi[2][1], 0u}, A2l[4] = {alo[2][0], 0u, alo[2][1], 0u};
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const unsigned Bh[2] = {bhi[s][0], bhi[s][1]}, Bl[2] = {blo[s][0], blo[s][1]};
          float d01[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
          mma_tf32(d01, A01l, Bh);      // small terms first
          mma_tf32(d01, A01h, Bl);
          mma_tf32(d01, A01h, Bh);
          mma_tf32(d2, A2l, Bh);
          mma_tf32(d2, A2h, Bl);
          mma_tf32(d2, A2h, Bh);
          c[0][s][0] = d01[0]; c[0][s][1] = d01[1];
          c[1][s][0] = d01[2]; c[1][s][1] = d01[3];
          c[2][s][0] = d2[0];  c[2][s][1] = d2[1];
        }
      }
      // Ke blocks (a, 2k) and (a, 2k+1): lam P + mu P^T + mu tr(P) I  (B^T D B of an isotropic D)
      float K[2][3][3];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float tr = c[0][0][h] + c[1][1][h] + c[2][2][h];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j)
            K[h][i][j] = lam * c[i][j][h] + mu * c[j][i][h] + (i == j ? mu * tr : 0.f);
      }
      // re = Ke u - Fe: partial over this lane's 6 columns, then butterfly over the 4 k-lanes
      float r[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float acc = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 3; ++j) acc += K[h][i][j] * sm.u[buf][j][el * 8 + 2 * kq + h];
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        r[i] = acc;
      }
      if (has_body) {  // Fe_a = b * sum_g w detJ N_a(g)   (mechanical.py:110)
        float nw = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const double gx = 1.0 + sgn_x(ra) * sgn_x(g) * FOL_S3, gy = 1.0 + sgn_y(ra) * sgn_y(g) * FOL_S3;
          const double gz = 1.0 + sgn_z(ra) * sgn_z(g) * FOL_S3;
          nw += sm.wd[el][g] * (float)(0.125 * gx * gy * gz);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) r[i] -= args.p.v[2 + i] * nw;
      }
      // Dirichlet row mask (fe_loss.py:191-207): only for elements touching a fixed dof (warp-uniform test)
      const bool fixed_rows = (sm.bc[el][ra * 3 + 0] == 0) | (sm.bc[el][ra * 3 + 1] == 0) | (sm.bc[el][ra * 3 + 2] == 0);
      const bool any_fixed = __any_sync(0xffffffffu, fixed_rows);
      float* const slot = sm.stage[el & 1];
      if (lane == 0) bulk_wait_read<1>();   // the copy that last used THIS slot (two elements ago) has drained it
      __syncwarp();
      if (!any_fixed) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float2* dst = reinterpret_cast<float2*>(slot + (ra * 3 + i) * 24 + kq * 6);
          dst[0] = make_float2(K[0][i][0], K[0][i][1]);
          dst[1] = make_float2(K[0][i][2], K[1][i][0]);
          dst[2] = make_float2(K[1][i][1], K[1][i][2]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int row = ra * 3 + i;
          const bool freerow = sm.bc[el][row] != 0;
          float v[6];
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const int col = (2 * kq + h) * 3 + j;
              v[h * 3 + j] = (freerow || col == row) ? K[h][i][j] : 0.f;
            }
          float2* dst = reinterpret_cast<float2*>(slot + row * 24 + kq * 6);
          dst[0] = make_float2(v[0], v[1]);
          dst[1] = make_float2(v[2], v[3]);
          dst[2] = make_float2(v[4], v[5]);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) bulk_store_f(args.ke + e * 576, slot, 576 * sizeof(float));
      if (kq == 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
          args.re[e * 24 + ra * 3 + i] = (any_fixed && sm.bc[el][ra * 3 + i] == 0) ? 0.f : r[i];
      }
    }
    __syncwarp();  // everyone is done with X / u / gradients of this tile
  }
  if (lane == 0) bulk_wait_read<0>();  // shared memory must outlive the last copies
}

int assemble_hex_mech_f32(cudaStream_t s, const AsmArgs<float>& args) {
  static PerDeviceGrid per_device;
  const size_t smem = sizeof(WarpSmemF) * kWarpsF;
  int grid = 0;
  FOL_CUDA(per_device.get(assemble_hex_mech_f32_kernel, kWarpsF * 32, smem, &grid));
  if (args.ne == 0) return FOL_OK;
  // the bulk copies need 16-byte aligned element matrices: 2304 B per element keeps the alignment of the base
  if ((reinterpret_cast<unsigned long long>(args.ke) & 15ull) != 0ull) return 1;   // caller falls back to the generic kernel
  const long long ntiles = cdiv(args.ne, kTile);
  const long long want = cdiv(ntiles, kWarpsF);
  const int has_body = (args.p.v[2] != 0.f || args.p.v[3] != 0.f || args.p.v[4] != 0.f) ? 1 : 0;
  assemble_hex_mech_f32_kernel<<<(unsigned)(want < grid ? want : grid), kWarpsF * 32, smem, s>>>(args, ntiles, has_body);
  return check_launch("assemble_hex_mech_f32_kernel");
}

}  // namespace fol
