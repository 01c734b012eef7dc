// Batched elasticity loss + VJP on a STRUCTURED Quad4 grid: MechanicalLoss2DQuad.ComputeBatchLoss and its JAX-AD
// gradient (mechanical.py:98-117: E_e = u_e . stop_gradient(Se u_e - Fe), so dE/du is the assembled, un-masked residual
// and dE/dK = 0; fe_loss.py:250-262) for the meshes energy_grid.cu serves -- row-major node numbers, elements
// [n, n + 1, n + nx + 2, n + nx + 1], ONE parallelogram shape (fol/tools/usefull_functions.py:213-258).  Plane stress,
// 2 x 2 rule, two dofs per node stored interleaved.
//
// Same machine mapping as energy_grid_kernel (energy_grid.cu; shared pieces in energy_grid_common.cuh): one producer
// warp streams node rows into a shared-memory ring with bulk copies, W <= 8 consumer warps march up the grid with one
// node column per lane, the pair that travels between lanes / through the shared-column arrays is (dE/du_x, dE/du_y)
// where the thermal kernel carries (dE/dT, dE/dK).  A u row holds 2 values per node, so the ring rows are twice as long;
// the Dirichlet values / flags are per dof.  Element: the displacement gradient at the four Gauss points from the edge
// values of the two node rows (sum-factorised as there), sigma = D eps with the plane-stress D of mechanical.py:60-70
// scaled by w detJ (N.K), back to the nodes through the same 1-D weights; the body force is a constant per corner on a
// uniform grid.  Every operation with its rounding written out (results independent of chunk height and batch).
#include "energy_grid_common.cuh"

namespace fol {

template <class T>
struct GridMechArgs {
  const T* ctrl;              // (nb, nn)
  const T* u;                 // (nb, 2 nn): (ux, uy) per node
  T* grad_u;                  // (nb, 2 nn)
  T* partial;                 // (nb, npart) energy shares, one per warp
  const T* dir_values;        // (2 nn) NaN where free, or null: overwrites u while reading (fe_loss.py:91-92, 255)
  const uint8_t* dir_flag;    // (2 nn) 1 where grad_u is written as zero, or null
  const uint8_t* col_dir;     // (nx + 1) 1 where the node column holds a Dirichlet dof, or null (= every column may)
  T out_scale, wd;
  T d11, d12, d33;            // plane-stress D of the unit control: E / (1 - nu^2) [1, nu, (1 - nu) / 2]
  T body[2];
  T jinv[4];                  // row-major d xi_j / d x_k of the one element shape
  int nx, ny;
  int rows, nchunks, npanels, npart, W;
  int row_bytes;              // bytes of one staged row (2 values per node; 16-byte multiple)
  long long nn, nb;
};

namespace {

// Edge values of a node row for the lane's column pair (own, right): both displacement components and K on the edge at
// xi = -s / +s, and the differences along the edge (see GridEdge in energy_grid.cu)
template <class T>
struct MechEdge {
  T mx, px, dx, my, py, dy, km, kp;
};
template <class T>
__device__ __forceinline__ MechEdge<T> mech_edge(T x_own, T y_own, T x_right, T y_right, T k_own, T k_right) {
  constexpr double s = FOL_S3;
  const T a = bcd<T>(0.5 * (1.0 - s)), b = bcd<T>(0.5 * (1.0 + s));
  MechEdge<T> e;
  e.mx = lin2(b, x_own, a, x_right);
  e.px = lin2(a, x_own, b, x_right);
  e.dx = op_sub(x_right, x_own);
  e.my = lin2(b, y_own, a, y_right);
  e.py = lin2(a, y_own, b, y_right);
  e.dy = op_sub(y_right, y_own);
  e.km = lin2(b, k_own, a, k_right);
  e.kp = lin2(a, k_own, b, k_right);
  return e;
}

// Element vectors rx, ry = (Se u_e) per corner and component, and u_e . Se u_e, of the plane-stress Quad4 with the 2 x 2
// rule on a parallelogram (local nodes and Gauss points ordered as in grid_element): displacement gradient H at the
// points from the edge values (d/dxi depends on eta only, d/deta on xi only), eps = sym H, sigma = D eps scaled by
// w detJ (N.K), nodal forces sum_g (sigma J^-T) . dN through the same 1-D weights.
template <class T, bool DIAG>
__device__ __forceinline__ void mech_element(const MechEdge<T>& B, const MechEdge<T>& U, const T (&ji)[4], T wd, T d11,
                                             T d12, T d33, T (&rx)[4], T (&ry)[4], T& e_el) {
  constexpr double s = FOL_S3;
  const T a = bcd<T>(0.5 * (1.0 - s)), b = bcd<T>(0.5 * (1.0 + s)), ah = bcd<T>(0.25 * (1.0 - s)),
          bh = bcd<T>(0.25 * (1.0 + s)), half = bcd<T>(0.5);
  const T xe[2] = {lin2(bh, B.dx, ah, U.dx), lin2(ah, B.dx, bh, U.dx)};     // d ux / d xi at eta = -s, +s
  const T ye[2] = {lin2(bh, B.dy, ah, U.dy), lin2(ah, B.dy, bh, U.dy)};     // d uy / d xi
  const T xn[2] = {op_sub(U.mx, B.mx), op_sub(U.px, B.px)};                 // 2 d ux / d eta at xi = -s, +s
  const T yn[2] = {op_sub(U.my, B.my), op_sub(U.py, B.py)};                 // 2 d uy / d eta
  const T eg[4] = {lin2(b, B.km, a, U.km), lin2(b, B.kp, a, U.kp), lin2(a, B.kp, b, U.kp), lin2(a, B.km, b, U.km)};
  constexpr int ETA[4] = {0, 0, 1, 1}, XI[4] = {0, 1, 1, 0};                // eta / xi index of Gauss point g
  const T j2h = op_mul(half, ji[2]), j3h = op_mul(half, ji[3]);
  T qx0[4], qx1[4], qy0[4], qy1[4];                                         // per point: (sigma J^-T) rows, x / y component
  T en = bcd<T>(0.0);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    T hxx, hxy, hyx, hyy;                                                   // H[c][k] = d u_c / d x_k
    if constexpr (DIAG) {
      hxx = op_mul(xe[ETA[g]], ji[0]);
      hxy = op_mul(xn[XI[g]], j3h);
      hyx = op_mul(ye[ETA[g]], ji[0]);
      hyy = op_mul(yn[XI[g]], j3h);
    } else {
      hxx = lin2(xe[ETA[g]], ji[0], xn[XI[g]], j2h);
      hxy = lin2(xe[ETA[g]], ji[1], xn[XI[g]], j3h);
      hyx = lin2(ye[ETA[g]], ji[0], yn[XI[g]], j2h);
      hyy = lin2(ye[ETA[g]], ji[1], yn[XI[g]], j3h);
    }
    const T gam = op_add(hxy, hyx);
    const T w = op_mul(wd, eg[g]);
    const T sxx = op_mul(w, lin2(d11, hxx, d12, hyy)), syy = op_mul(w, lin2(d12, hxx, d11, hyy));
    const T sxy = op_mul(w, op_mul(d33, gam));
    en = op_fma(sxx, hxx, op_fma(syy, hyy, op_fma(sxy, gam, en)));
    if constexpr (DIAG) {
      qx0[g] = op_mul(ji[0], sxx);
      qx1[g] = op_mul(ji[3], sxy);
      qy0[g] = op_mul(ji[0], sxy);
      qy1[g] = op_mul(ji[3], syy);
    } else {
      qx0[g] = lin2(ji[0], sxx, ji[1], sxy);
      qx1[g] = lin2(ji[2], sxx, ji[3], sxy);
      qy0[g] = lin2(ji[0], sxy, ji[1], syy);
      qy1[g] = lin2(ji[2], sxy, ji[3], syy);
    }
  }
  auto project = [&](const T (&w0)[4], const T (&w1)[4], T (&re)[4]) {
    const T s0lo = op_add(w0[0], w0[1]), s0hi = op_add(w0[2], w0[3]);
    const T s1l = op_add(w1[0], w1[3]), s1r = op_add(w1[1], w1[2]);
    const T S0b = lin2(bh, s0lo, ah, s0hi), S0t = lin2(ah, s0lo, bh, s0hi);
    const T S1l = lin2(bh, s1l, ah, s1r), S1r = lin2(ah, s1l, bh, s1r);
    re[0] = op_neg(op_add(S0b, S1l));
    re[1] = op_sub(S0b, S1r);
    re[2] = op_add(S0t, S1r);
    re[3] = op_sub(S1l, S0t);
  };
  project(qx0, qx1, rx);
  project(qy0, qy1, ry);
  e_el = en;
}

}  // namespace

template <class S, int NS, bool DIAG, bool BODY>
// same CTA shape as energy_grid_kernel (energy_grid.cu); 96 registers = 5 warps per scheduler = two 9-warp CTAs per SM
__global__ void __launch_bounds__(288) __maxnreg__(sizeof(S) == 8 ? 96 : (NS == 2 ? 96 : 72))
    energy_grid_mech_kernel(const GridMechArgs<S> args) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using V = typename LaneT<S, NS>::type;                     // lane value: one sample, or two float32 samples
  using Pair = NodePair<V>;
  constexpr int PER = 16 / (int)sizeof(S);
  constexpr int NROW = 2 * NS + 1;                           // ring rows per slot: u (per sample), K (per sample), D
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
        const long long first_n = (long long)(e_beg + i) * NXn + sp;     // first node of the row segment
        RowCopy cu[NS], ck[NS];                              // u: two dofs per node; ctrl: one value per node
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          cu[j] = plan_row<S>(2 * total, 2 * (smp[j] * args.nn + first_n), 2 * ncols);
          ck[j] = plan_row<S>(total, smp[j] * args.nn + first_n, ncols);
        }
        const RowCopy cd = plan_row<S>(2 * args.nn, 2 * first_n, 2 * ncols);
        // plain tail stores (the last row of the last sample only) first, then the arrive that publishes them and
        // arms the transaction count, then the bulk copies that complete it
        uint32_t bytes = has_dirv ? cd.bytes : 0u;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          copy_tail<S>(cu[j], args.u, reinterpret_cast<S*>(base + j * args.row_bytes));
          copy_tail<S>(ck[j], args.ctrl, reinterpret_cast<S*>(base + (NS + j) * args.row_bytes));
          bytes += cu[j].bytes + ck[j].bytes;
        }
        if (has_dirv) copy_tail<S>(cd, args.dir_values, reinterpret_cast<S*>(base + 2 * NS * args.row_bytes));
        mbar_expect_tx(bar, bytes);
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          copy_bulk<S>(cu[j], args.u, reinterpret_cast<S*>(base + j * args.row_bytes), bar);
          copy_bulk<S>(ck[j], args.ctrl, reinterpret_cast<S*>(base + (NS + j) * args.row_bytes), bar);
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

    const V ji[4] = {bc<V>(args.jinv[0]), bc<V>(args.jinv[1]), bc<V>(args.jinv[2]), bc<V>(args.jinv[3])};
    const V wd = bc<V>(args.wd), scale = bc<V>(args.out_scale);
    const V d11 = bc<V>(args.d11), d12 = bc<V>(args.d12), d33 = bc<V>(args.d33);
    const V fbx = bc<V>(args.body[0] * args.wd), fby = bc<V>(args.body[1] * args.wd);   // Fe of one corner: b w detJ
    const V zero = bcd<V>(0.0);
    // shared-memory addresses of this lane's entries in ring slot 0: u rows hold (ux, uy) per node, K rows one value
    const uint32_t ring_bytes = (uint32_t)(kRing * slot_bytes), rb = (uint32_t)args.row_bytes;
    const uint32_t a_own_u = smem_u32(ring) + 2u * (uint32_t)i_own * (uint32_t)sizeof(S);
    const uint32_t a_own_k = smem_u32(ring) + NS * rb + (uint32_t)i_own * (uint32_t)sizeof(S);
    uint32_t slot_off = 0, parity = 0, slot_bar = 0;         // ring position of the next row to take
    const long long row_n = (long long)e_beg * NXn + sp;     // first node of the first staged row segment
    uint32_t sh_u[NS], sh_k[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      sh_u[j] = (uint32_t)((2 * (smp[j] * args.nn + row_n)) & (PER - 1)) * (uint32_t)sizeof(S);
      sh_k[j] = (uint32_t)((smp[j] * args.nn + row_n) & (PER - 1)) * (uint32_t)sizeof(S);
    }
    uint32_t sh_d = (uint32_t)((2 * row_n) & (PER - 1)) * (uint32_t)sizeof(S);
    const uint32_t sh_step_u = (uint32_t)((2 * NXn) & (PER - 1)) * (uint32_t)sizeof(S);
    const uint32_t sh_step_k = (uint32_t)(NXn & (PER - 1)) * (uint32_t)sizeof(S);
    // global offset of node (cc, row) within the sample, as 32 bits (2 nn < 2^31 is checked by the host)
    unsigned node = (unsigned)e_beg * (unsigned)NXn + (unsigned)cc;
    S* gu0[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) gu0[j] = args.grad_u + smp[j] * 2 * args.nn;
    const uint8_t* const fl0 = wcut ? args.dir_flag : args.col_dir;   // never read unless wcut
    uint32_t a_left = smem_u32(left + (size_t)w * args.rows), a_right = smem_u32(right + (size_t)w * args.rows);

    auto lds = [](uint32_t a) {
      S v;
      if constexpr (sizeof(S) == 8) asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(a));
      else asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(a));
      return v;
    };
    auto ldv = [&](uint32_t a, const uint32_t (&sh)[NS]) {   // ring row at a (+ one row for the second sample)
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
    // the values of the next staged node row: (ux, uy, K) of the own and of the right column, as edge values
    using Edge = MechEdge<V>;
    auto take_edge = [&]() {
      mbar_wait(full0 + slot_bar, parity);
      const uint32_t aU = a_own_u + slot_off, aK = a_own_k + slot_off;
      constexpr uint32_t SZ = (uint32_t)sizeof(S);
      V x0 = ldv(aU, sh_u), y0 = ldv(aU + SZ, sh_u), x1 = ldv(aU + 2 * SZ, sh_u), y1 = ldv(aU + 3 * SZ, sh_u);
      const V k0 = ldv(aK, sh_k), k1 = ldv(aK + SZ, sh_k);
      if (wdirv) {                                           // Dirichlet overwrite (fe_loss.py:91-92, 255), per dof
        const uint32_t aD = a_own_u + slot_off + 2u * NS * rb + sh_d;
        const S dx0 = lds(aD), dy0 = lds(aD + SZ), dx1 = lds(aD + 2 * SZ), dy1 = lds(aD + 3 * SZ);
        if (dx0 == dx0) x0 = bc<V>(dx0);
        if (dy0 == dy0) y0 = bc<V>(dy0);
        if (dx1 == dx1) x1 = bc<V>(dx1);
        if (dy1 == dy1) y1 = bc<V>(dy1);
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
      for (int j = 0; j < NS; ++j) {
        sh_u[j] = (sh_u[j] + sh_step_u) & 15u;
        sh_k[j] = (sh_k[j] + sh_step_k) & 15u;
      }
      sh_d = (sh_d + sh_step_u) & 15u;
      return mech_edge<V>(x0, y0, x1, y1, k0, k1);
    };

    V en = zero;
    // node row (at offset `node`) is complete once the element rows below and above it are in: left half (own lane:
    // below + above) + right half of the lane to the left; the column two warps share waits for `combine`.
    // X / Y: the two components of dE/du at the node (the pair travels like (R, K) in the thermal kernel).
    unsigned cut_now = 0;                                    // dir_flag of the node's two dofs, loaded one row ahead
    auto load_cut = [&](unsigned n) {
      return (unsigned)fl0[2 * (size_t)n] | ((unsigned)fl0[2 * (size_t)n + 1] << 8);
    };
    if (wcut) cut_now = load_cut(node + (e_beg < r0 ? (unsigned)NXn : 0u));
    auto finish_row = [&](V leftX, V leftY, V rpX, V rpY, bool more) {
      const V inX = shfl_up(rpX), inY = shfl_up(rpY);
      unsigned cut_next = 0;
      if (wcut && more) cut_next = load_cut(node + (unsigned)NXn);
      if (keep_left) sts_pair(a_left, leftX, leftY);
      if (keep_right) sts_pair(a_right, rpX, rpY);
      a_left += (uint32_t)sizeof(Pair);
      a_right += (uint32_t)sizeof(Pair);
      V X = (l > 0) ? op_add(leftX, inX) : leftX;
      V Y = (l > 0) ? op_add(leftY, inY) : leftY;
      if (wcut) {
        if (cut_now & 0xffu) X = zero;
        if (cut_now & 0xff00u) Y = zero;
      }
      if (write_own) {
        const V oX = op_mul(scale, X), oY = op_mul(scale, Y);
        if constexpr (NS == 2) {
          *reinterpret_cast<float2*>(gu0[0] + 2 * (size_t)node) = make_float2(oX.x, oY.x);
          if (second) *reinterpret_cast<float2*>(gu0[1] + 2 * (size_t)node) = make_float2(oX.y, oY.y);
        } else if constexpr (sizeof(S) == 8) {
          *reinterpret_cast<double2*>(gu0[0] + 2 * (size_t)node) = make_double2(oX, oY);
        } else {
          *reinterpret_cast<float2*>(gu0[0] + 2 * (size_t)node) = make_float2(oX, oY);
        }
      }
      cut_now = cut_next;
    };
    // one element row between the edge values of the node rows below (B) and above (U); carries of the row below in
    // (oc, rc)
    V ocX = zero, ocY = zero, rcX = zero, rcY = zero;
    auto element = [&](const Edge& B, const Edge& U, V (&rx)[4], V (&ry)[4], V& e_el) {
      mech_element<V, DIAG>(B, U, ji, wd, d11, d12, d33, rx, ry, e_el);
      if constexpr (BODY) {                                  // re = Se u - Fe, Fe_(a,c) = b_c w detJ at every corner
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          rx[q] = op_sub(rx[q], fbx);
          ry[q] = op_sub(ry[q], fby);
        }
        // E = u . (Se u - Fe): the corner values of a component sum to (m + p) of the two edges
        const V sx = op_add(op_add(B.mx, B.px), op_add(U.mx, U.px)), sy = op_add(op_add(B.my, B.py), op_add(U.my, U.py));
        e_el = op_sub(e_el, op_fma(sx, fbx, op_mul(sy, fby)));
      }
      if (!all_valid) {                                      // warp-uniform: a ragged last warp only
        if (!el_valid) {
#pragma unroll
          for (int q = 0; q < 4; ++q) rx[q] = ry[q] = zero;
          e_el = zero;
        }
      }
    };
    auto step = [&](const Edge& B, const Edge& U) {
      V rx[4], ry[4], e_el;
      element(B, U, rx, ry, e_el);
      en = op_add(en, e_el);
      finish_row(op_add(ocX, rx[0]), op_add(ocY, ry[0]), op_add(rcX, rx[1]), op_add(rcY, ry[1]), true);
      node += (unsigned)NXn;
      ocX = rx[3]; ocY = ry[3]; rcX = rx[2]; rcY = ry[2];
    };

    Edge EA = take_edge(), EB;                               // two register sets of edge values, used alternately
    int e = e_beg;
    if (e_beg < r0) {
      // the recomputed element row below the chunk: only its shares of node row r0 (the carries) are kept
      EB = take_edge();
      V rx[4], ry[4], e_el;
      element(EA, EB, rx, ry, e_el);
      ocX = rx[3]; ocY = ry[3]; rcX = rx[2]; rcY = ry[2];
      node += (unsigned)NXn;
      EA = EB;
      ++e;
    }
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
    if (r1 == ny + 1) finish_row(ocX, ocY, rcX, rcY, false);  // the top node row of the grid closes with the carries alone

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
    V X = lo.t, Y = lo.k;
    if (bnd < W) {
      const Pair hi = left[(size_t)bnd * args.rows + i];
      X = op_add(hi.t, X);                                   // same order as finish_row: own (left) half + incoming
      Y = op_add(hi.k, Y);
    }
    const long long bnode = (long long)(r0 + i) * NXn + col;
    const bool cutx = args.dir_flag ? args.dir_flag[2 * bnode] != 0 : false;
    const bool cuty = args.dir_flag ? args.dir_flag[2 * bnode + 1] != 0 : false;
    const V oX = op_mul(bc<V>(args.out_scale), X), oY = op_mul(bc<V>(args.out_scale), Y);
    if constexpr (NS == 2) {
      S* g0 = args.grad_u + smp[0] * 2 * args.nn + 2 * bnode;
      g0[0] = cutx ? 0.f : oX.x;
      g0[1] = cuty ? 0.f : oY.x;
      if (second) {
        S* g1 = args.grad_u + smp[1] * 2 * args.nn + 2 * bnode;
        g1[0] = cutx ? 0.f : oX.y;
        g1[1] = cuty ? 0.f : oY.y;
      }
    } else {
      S* g0 = args.grad_u + smp[0] * 2 * args.nn + 2 * bnode;
      g0[0] = cutx ? (S)0 : oX;
      g0[1] = cuty ? (S)0 : oY;
    }
  }
}

namespace {

template <class T>
int mech_row_bytes(int W, long long nx) {
  const long long ncols = (32LL * W + 1 < nx + 1) ? 32LL * W + 1 : nx + 1;
  return (int)(((2 * ncols + 16 / sizeof(T) + 1) * sizeof(T) + 15) / 16 * 16);   // + shift + the entry pair read past the last column
}
template <class T, int NS>
size_t mech_smem(int W, long long nx, int rows) {
  return (size_t)kRing * (2 * NS + 1) * mech_row_bytes<T>(W, nx) + (size_t)2 * W * rows * 2 * NS * sizeof(T) + 2 * kRing * 8;
}

template <class T, int NS, bool DIAG, bool BODY>
int launch_grid_mech(cudaStream_t s, GridMechArgs<T> a, T* energy) {
  auto kern = energy_grid_mech_kernel<T, NS, DIAG, BODY>;
  const GridShape g = grid_shape(a.nx);
  const int threads = 32 * (g.W + 1);
  static PerDeviceOnce configured;
  if (configured.need()) {
    FOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mech_smem<T, NS>(8, 256, 128)));
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
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, mech_smem<T, NS>(g.W, a.nx, rows)) !=
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
  if (best_rows == 0) return fail(FOL_ERR_CUDA, "fol_energy_and_grads_grid_mech: the kernel does not fit on this device");
  a.rows = best_rows;
  a.nchunks = (int)cdiv(nrows_total, best_rows);
  a.npanels = g.npanels;
  a.W = g.W;
  a.row_bytes = mech_row_bytes<T>(g.W, a.nx);
  a.npart = a.nchunks * a.npanels * g.W;
  const long long items = (long long)a.nchunks * a.npanels * groups;
  FOL_REQUIRE(items < (1LL << 31), "fol_energy_and_grads_grid_mech: too many work items for one launch");
  kern<<<(unsigned)items, threads, mech_smem<T, NS>(g.W, a.nx, best_rows), s>>>(a);
  int rc = check_launch("energy_grid_mech_kernel");
  if (rc) return rc;
  energy_sum_kernel<T><<<(unsigned)cdiv(a.nb, 8), 256, 0, s>>>(a.partial, a.nb, a.npart, energy);
  return check_launch("energy_sum_kernel");
}

template <class T>
int dispatch_grid_mech(cudaStream_t s, const GridMechArgs<T>& a, T* energy) {
  const bool diag = a.jinv[1] == (T)0 && a.jinv[2] == (T)0;
  const bool body = a.body[0] != (T)0 || a.body[1] != (T)0;
  constexpr int NS = sizeof(T) == 4 ? 2 : 1;     // float32: two samples per lane on the packed FP32 instructions
  if (diag) return body ? launch_grid_mech<T, NS, true, true>(s, a, energy) : launch_grid_mech<T, NS, true, false>(s, a, energy);
  return body ? launch_grid_mech<T, NS, false, true>(s, a, energy) : launch_grid_mech<T, NS, false, false>(s, a, energy);
}

}  // namespace
}  // namespace fol

using namespace fol;

extern "C" {

int64_t fol_energy_grid_mech_work_size(int64_t nx, int64_t ny, int64_t nb) {
  if (nx < 1 || ny < 1 || nb < 0) return 0;
  const GridShape g = grid_shape(nx);
  return nb * (cdiv(ny + 1, kMinRows) + 1) * g.npanels * g.W + 16;
}

int fol_energy_and_grads_grid_mech(fol_stream_t s, int dtype, int64_t nx, int64_t ny, int64_t nb, const double* jinv_host,
                                   double w_detj, const void* ctrl, const void* u, const void* dir_values,
                                   const uint8_t* dir_flag, const uint8_t* col_dir, double out_scale,
                                   const double* params_host, void* grad_u, void* energy, void* work) {
  FOL_REQUIRE(nx >= 1 && ny >= 1 && nb >= 0 && 2 * (nx + 1) * (ny + 1) < (1LL << 31), "fol_energy_and_grads_grid_mech: bad grid size");
  FOL_REQUIRE(jinv_host && params_host && ctrl && u && grad_u && energy && work, "fol_energy_and_grads_grid_mech: null pointer");
  FOL_REQUIRE(dtype == FOL_F64 || dtype == FOL_F32, "fol_energy_and_grads_grid_mech: unknown dtype");
  FOL_REQUIRE(((uintptr_t)ctrl | (uintptr_t)u | (uintptr_t)dir_values | (uintptr_t)grad_u) % 16 == 0,
              "fol_energy_and_grads_grid_mech: ctrl, u, dir_values and grad_u must be 16-byte aligned");
  if (nb == 0) return FOL_OK;
  auto run = [&](auto* tag) {
    using T = std::remove_pointer_t<decltype(tag)>;
    GridMechArgs<T> a;
    a.ctrl = (const T*)ctrl;
    a.u = (const T*)u;
    a.grad_u = (T*)grad_u;
    a.partial = (T*)work;
    a.dir_values = (const T*)dir_values;
    a.dir_flag = dir_flag;
    a.col_dir = col_dir;
    a.out_scale = (T)out_scale;
    a.wd = (T)w_detj;
    // plane stress (mechanical.py:60-70) for the unit control; the control field scales it point by point
    const double E = params_host[0], nu = params_host[1], f = E / (1.0 - nu * nu);
    a.d11 = (T)f;
    a.d12 = (T)(f * nu);
    a.d33 = (T)(f * (1.0 - nu) * 0.5);
    a.body[0] = (T)params_host[2];
    a.body[1] = (T)params_host[3];
    for (int i = 0; i < 4; ++i) a.jinv[i] = (T)jinv_host[i];
    a.nx = (int)nx;
    a.ny = (int)ny;
    a.nn = (nx + 1) * (ny + 1);
    a.nb = nb;
    a.rows = a.nchunks = a.npanels = a.npart = a.W = a.row_bytes = 0;
    return dispatch_grid_mech<T>((cudaStream_t)s, a, (T*)energy);
  };
  if (dtype == FOL_F64) return run((double*)nullptr);
  return run((float*)nullptr);
}

}  // extern "C"
