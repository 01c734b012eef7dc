// One-launch BiCGSTAB: the whole iteration loop of folax_b200/linalg.py::bicgstab (the recurrences, start, stopping
// rule and break-down codes of jax.scipy.sparse.linalg.bicgstab, which fol/solvers/fe_solver.py:62-67 calls) as ONE
// persistent kernel.  At the sizes of BASELINE.json configs[0] / configs[3] (5 k / 1 M dofs) the multi-launch loop is
// launch- and host-read-latency-bound (0.2-0.4 ms per iteration for ~0.1 ms of streaming, profiles/r2/newton_config4.md):
// here an iteration is two SELL products, four fused vector passes and five grid barriers, the recurrence scalars live
// in shared memory of every CTA (each CTA reduces the same per-CTA partial sums in the same fixed order, so all CTAs
// take the same decisions without a broadcast), and the host reads one number when the solve is over.
// Same iterates as the multi-launch loop up to the summation order of the dot products (per-CTA partials here, 592
// fixed blocks there); deterministic run to run for a given grid.
#include <type_traits>

#include "common.cuh"
#include "krylov_threads.cuh"

namespace fol {

template <class T>
struct FusedBicgArgs {
  const long long* slice_ptr;   // SELL matrix (krylov_threads.cuh)
  const int32_t* cols;          // node columns (D = 2, 3) or scalar columns (D = 0)
  const T* vals;
  int D;
  long long n;
  const T* b;
  T* x;                         // in: x0, out: solution
  const T* mdiag;               // Jacobi diagonal or null
  T *r, *rhat, *p, *phat, *q, *s, *shat, *t;
  T* partial;                   // [2][gridDim.x][2]
  unsigned long long* sync;     // [0] barrier counter (zeroed by the launcher), [1] spin waits that gave up
  T* result;                    // [0] iterations (or -10 / -11), [1] final |r|^2, [2] barrier waits that gave up (0)
  T tol, atol;
  long long maxiter;
};

namespace {

constexpr int kFusedThreads = 256;
#ifndef FOL_FUSED_UNROLL
#define FOL_FUSED_UNROLL 4
#endif
#ifndef FOL_FUSED_CTAS
#define FOL_FUSED_CTAS 5
#endif
// CTAs of 256 threads per SM.  Measured at 1.07 M dofs (ms per iteration): 3 CTAs / 80 registers 0.255, 4 / 64 0.258,
// 5 / 48 0.235, 6 / 40 0.248 -- the products want warps in flight to hide their gather latency, up to the point where
// the register cap spills the row loop
constexpr int kFusedCtasPerSm = FOL_FUSED_CTAS;
constexpr int kFusedUnroll = FOL_FUSED_UNROLL;   // runs of a SELL row in flight per thread

__device__ __forceinline__ void grid_sync(unsigned long long* sync, unsigned long long& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    __threadfence();
    atomicAdd(sync, 1ULL);
    unsigned long long seen;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(sync) : "memory");
      if (seen < epoch && clock64() - t0 > 4000000000LL) {   // a CTA that never arrives must not hang the GPU
        atomicAdd(sync + 1, 1ULL);
        break;
      }
    } while (seen < epoch);
    __threadfence();   // acquire side: orders the other CTAs' released writes before this CTA's later (plain) loads
  }
  __syncthreads();
}

// Grid-wide sums of (a, b): block tree -> per-CTA partials -> barrier -> every CTA adds all partials in the same fixed
// order (lane-strided, then a shuffle tree), so every CTA holds bit-identical totals.
template <class T>
__device__ __forceinline__ void grid_sum2(T a, T b, T* partial, int& parity, unsigned long long* sync,
                                          unsigned long long& epoch, T* sh, T& out_a, T& out_b) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    sh[warp * 2] = a;
    sh[warp * 2 + 1] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    T sa = (T)0, sb = (T)0;
    for (int w = 0; w < kFusedThreads / 32; ++w) {
      sa += sh[w * 2];
      sb += sh[w * 2 + 1];
    }
    T* mine = partial + ((size_t)parity * gridDim.x + blockIdx.x) * 2;
    mine[0] = sa;
    mine[1] = sb;
  }
  grid_sync(sync, epoch);
  if (warp == 0) {
    const T* all = partial + (size_t)parity * gridDim.x * 2;
    T sa = (T)0, sb = (T)0;
    for (unsigned j = lane; j < gridDim.x; j += 32) {
      sa += __ldcg(all + j * 2);
      sb += __ldcg(all + j * 2 + 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sa += __shfl_xor_sync(0xffffffffu, sa, o);
      sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    if (lane == 0) {
      sh[32] = sa;
      sh[33] = sb;
    }
  }
  __syncthreads();
  out_a = sh[32];
  out_b = sh[33];
  __syncthreads();
  parity ^= 1;
}

// One SELL row (krylov_threads.cuh layout; same fixed summation order as sell_spmv_thread / sell_spmv_block_thread).
// x is read with PLAIN loads: the neighbour rows of a slice share most of their columns, so the gathers hit in L1; the
// grid barrier before every product (release by the writers, acquire + fence by thread 0, bar.sync) is what makes the
// other CTAs' writes visible to them, as for cooperative_groups::grid_group::sync().
template <class T, int D>
__device__ __forceinline__ T sell_row(const FusedBicgArgs<T>& a, long long row, const T* x) {
  const long long s = row >> 5;
  const int lane = (int)(row & 31);
  const long long base = a.slice_ptr[s];
  const int width = (int)((a.slice_ptr[s + 1] - base) >> 5);
  T acc = (T)0;
  if (D > 0) {
    const int runs = width / D;
    const int32_t* c = a.cols + base / D + lane;
    const T* v = a.vals + base + lane;
#pragma unroll(kFusedUnroll)
    for (int q = 0; q < runs; ++q) {
      const T* xm = x + (long long)c[(long long)q * 32] * D;
#pragma unroll
      for (int j = 0; j < D; ++j) acc = fol_fma(v[((long long)q * D + j) * 32], xm[j], acc);
    }
  } else {
    const int32_t* c = a.cols + base + lane;
    const T* v = a.vals + base + lane;
#pragma unroll 8
    for (int k = 0; k < width; ++k) acc = fol_fma(v[(long long)k * 32], x[c[(long long)k * 32]], acc);
  }
  return acc;
}

}  // namespace

template <class T, int D>
__global__ void __launch_bounds__(kFusedThreads, kFusedCtasPerSm) bicgstab_fused_kernel(const FusedBicgArgs<T> a) {
  __shared__ T sh[40];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long epoch = 0;
  int parity = 0;
  const long long n = a.n;

  // r0 = b - A x0, rhat = r0, p = q = 0
  T pbb = (T)0, prs = (T)0;
  for (long long i = i0; i < n; i += stride) {
    const T ri = fol_fma((T)1, a.b[i], (T)-1 * sell_row<T, D>(a, i, a.x));
    a.r[i] = ri;
    a.rhat[i] = ri;
    a.p[i] = (T)0;
    a.q[i] = (T)0;
    pbb += a.b[i] * a.b[i];
    prs += ri * ri;
  }
  T bb, rs;
  grid_sum2<T>(pbb, prs, a.partial, parity, a.sync, epoch, sh, bb, rs);
  T rho_new = rs;                                    // <rhat, r> with rhat = r
  const T atol2 = fmax(a.tol * a.tol * bb, a.atol * a.atol);
  T rho = (T)1, alpha = (T)1, omega = (T)1;
  long long k = 0;
  while (rs > atol2 && k >= 0 && k < a.maxiter) {
    if (rho_new == (T)0) {
      k = -10;
      break;
    }
    const T beta = rho_new / rho * alpha / omega;
    for (long long i = i0; i < n; i += stride) {     // p = r + beta (p - omega q);  phat = M^-1 p
      const T tmp = fol_fma((T)1, a.p[i], -omega * a.q[i]);
      const T pi = fol_fma((T)1, a.r[i], beta * tmp);
      a.p[i] = pi;
      a.phat[i] = a.mdiag ? (T)1 * pi / a.mdiag[i] : pi;
    }
    grid_sync(a.sync, epoch);
    T prq = (T)0, dummy = (T)0;
    for (long long i = i0; i < n; i += stride) {     // q = A phat,  <rhat, q>
      const T qi = sell_row<T, D>(a, i, a.phat);
      a.q[i] = qi;
      prq += a.rhat[i] * qi;
    }
    T rq, unused;
    grid_sum2<T>(prq, dummy, a.partial, parity, a.sync, epoch, sh, rq, unused);
    if (rq == (T)0) {
      k = -11;
      break;
    }
    alpha = rho_new / rq;
    T pss = (T)0;
    for (long long i = i0; i < n; i += stride) {     // s = r - alpha q;  shat = M^-1 s
      const T si = fol_fma((T)1, a.r[i], -alpha * a.q[i]);
      a.s[i] = si;
      a.shat[i] = a.mdiag ? (T)1 * si / a.mdiag[i] : si;
      pss += si * si;
    }
    T ss;
    grid_sum2<T>(pss, dummy, a.partial, parity, a.sync, epoch, sh, ss, unused);
    if (ss < atol2) {                                // converged on the half step
      for (long long i = i0; i < n; i += stride) a.x[i] = fol_fma((T)1, a.x[i], alpha * a.phat[i]);
      rs = ss;
      rho = rho_new;
      k += 1;
      break;
    }
    T pts = (T)0, ptt = (T)0;
    for (long long i = i0; i < n; i += stride) {     // t = A shat,  <t, s>, <t, t>
      const T ti = sell_row<T, D>(a, i, a.shat);
      a.t[i] = ti;
      pts += ti * a.s[i];
      ptt += ti * ti;
    }
    T ts, tt;
    grid_sum2<T>(pts, ptt, a.partial, parity, a.sync, epoch, sh, ts, tt);
    omega = (tt != (T)0) ? ts / tt : (T)0;
    T prs2 = (T)0, prho = (T)0;
    for (long long i = i0; i < n; i += stride) {     // x += alpha phat + omega shat;  r = s - omega t
      T xi = fol_fma((T)1, a.x[i], alpha * a.phat[i]);
      xi = fol_fma((T)1, xi, omega * a.shat[i]);
      a.x[i] = xi;
      const T ri = fol_fma((T)1, a.s[i], -omega * a.t[i]);
      a.r[i] = ri;
      prs2 += ri * ri;
      prho += a.rhat[i] * ri;
    }
    rho = rho_new;
    grid_sum2<T>(prs2, prho, a.partial, parity, a.sync, epoch, sh, rs, rho_new);
    if (omega == (T)0 || alpha == (T)0) {
      k = -11;
      break;
    }
    k += 1;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.result[0] = (T)k;
    a.result[1] = rs;
    a.result[2] = (T)a.sync[1];
  }
}

template <class T, int D>
static int launch_fused_d(cudaStream_t s, const FusedBicgArgs<T>& a) {
  static PerDeviceGrid per_device;
  int grid = 0;
  FOL_CUDA(per_device.get(bicgstab_fused_kernel<T, D>, kFusedThreads, 0, &grid));
  const long long want = cdiv(a.n, kFusedThreads);
  const unsigned blocks = (unsigned)(want < grid ? (want < 1 ? 1 : want) : grid);
  FOL_REQUIRE(blocks <= 4096, "fol_bicgstab_fused: more resident CTAs than the partial-sum scratch holds");
  FOL_CUDA(cudaMemsetAsync(a.sync, 0, 2 * sizeof(unsigned long long), s));
  // cooperative launch: the runtime refuses the launch unless every CTA can be resident at once, which is what the
  // grid barriers need (the spin-wait limit in grid_sync stays as a second line of defence)
  void* params[] = {const_cast<FusedBicgArgs<T>*>(&a)};
  FOL_CUDA(cudaLaunchCooperativeKernel((const void*)bicgstab_fused_kernel<T, D>, dim3(blocks), dim3(kFusedThreads),
                                       params, 0, s));
  return check_launch("bicgstab_fused_kernel");
}

template <class T>
static int launch_fused(cudaStream_t s, const FusedBicgArgs<T>& a) {
  if (a.D == 3) return launch_fused_d<T, 3>(s, a);
  if (a.D == 2) return launch_fused_d<T, 2>(s, a);
  return launch_fused_d<T, 0>(s, a);
}

}  // namespace fol

using namespace fol;

extern "C" {

/* scratch, in VALUES of the call's dtype: 8 work vectors + per-CTA partial sums + 2 results + (as bytes) 2 counters */
int64_t fol_bicgstab_fused_work_size(int64_t n) { return 8 * n + 2 * 2 * 4096 + 4 + 8; }

int fol_bicgstab_fused(fol_stream_t s, int dtype, int dofs_per_node, int64_t n, const int64_t* slice_ptr,
                       const int32_t* cols, const void* vals, const void* b, void* x, const void* m_diagonal, double tol,
                       double atol, int64_t maxiter, void* work) {
  FOL_REQUIRE(n > 0 && slice_ptr && cols && vals && b && x && work, "fol_bicgstab_fused: null pointer / bad size");
  FOL_REQUIRE(dofs_per_node == 0 || dofs_per_node == 2 || dofs_per_node == 3, "fol_bicgstab_fused: dofs_per_node must be 0 (scalar columns), 2 or 3");
  FOL_REQUIRE(dtype == FOL_F64 || dtype == FOL_F32, "fol_bicgstab_fused: unknown dtype");
  auto fill = [&](auto* w) {
    using T = std::remove_pointer_t<decltype(w)>;
    FusedBicgArgs<T> a;
    a.slice_ptr = (const long long*)slice_ptr;
    a.cols = cols;
    a.vals = (const T*)vals;
    a.D = dofs_per_node;
    a.n = n;
    a.b = (const T*)b;
    a.x = (T*)x;
    a.mdiag = (const T*)m_diagonal;
    T* v = w;
    a.r = v; a.rhat = v + n; a.p = v + 2 * n; a.phat = v + 3 * n; a.q = v + 4 * n; a.s = v + 5 * n; a.shat = v + 6 * n;
    a.t = v + 7 * n;
    a.partial = v + 8 * n;
    a.result = a.partial + 2 * 2 * 4096;
    // the counters sit behind the results, 8-byte aligned
    uintptr_t c = reinterpret_cast<uintptr_t>(a.result + 4);
    c = (c + 7) & ~(uintptr_t)7;
    a.sync = reinterpret_cast<unsigned long long*>(c);
    a.tol = (T)tol;
    a.atol = (T)atol;
    a.maxiter = maxiter;
    return launch_fused<T>((cudaStream_t)s, a);
  };
  if (dtype == FOL_F64) return fill((double*)work);
  return fill((float*)work);
}

}  // extern "C"
