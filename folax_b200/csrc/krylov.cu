// Kernels + C ABI of the GPU linear algebra behind folax_b200/solvers: what replaces the host round trip of
// fe_solver.py:60-103 (BCOO -> scipy CSR -> solve) once the Jacobian is assembled on the device.
//   fol_sell_spmv      y = A x on the sliced-ELLPACK layout (one thread per row, coalesced, deterministic)
//   fol_sell_spmv_block  the same with one column index per run of D dofs of a neighbour node (8 + 4/D B per entry)
//   fol_gather_values  value permutation (CSR -> SELL, CSR -> diagonal)
//   fol_vec_op         a x + b y | a x*y | a x/y
//   fol_dot            x . y with a fixed two-stage reduction tree (deterministic, no atomics)
// All four are HBM-bound streams; the SpMV's algorithmic traffic is 12 B per stored entry (8 value + 4 column) in
// float64 plus the vectors (x stays L2-resident: 51 MB at 6.4 M dofs).
#include "krylov_threads.cuh"

namespace fol {

template <class T>
__global__ void __launch_bounds__(128) sell_spmv_kernel(const SellArgs<T> a) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row < a.nrows) sell_spmv_thread<T>(row, a);
}

template <class T, int D>
__global__ void __launch_bounds__(128) sell_spmv_block_kernel(const BlockSellArgs<T> a) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row < a.nrows) sell_spmv_block_thread<T, D>(row, a);
}

template <class T>
__global__ void __launch_bounds__(256) gather_values_kernel(long long n, const int32_t* __restrict__ src_index,
                                                            const T* __restrict__ src, T* __restrict__ dst) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    gather_values_thread<T>(i, src_index, src, dst);
}

template <class T>
__global__ void __launch_bounds__(256) vec_op_kernel(long long n, int op, T a, const T* x, T b, const T* y, T* out) {
  // no __restrict__: the Krylov loops update in place (out aliases x or y)
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    vec_op_thread<T>(i, op, a, x, b, y, out);
}

template <class T>
__global__ void bicg_scalar_kernel(int stage, T* sc) {
  if (blockIdx.x == 0 && threadIdx.x == 0) bicg_scalar_stage<T>(stage, sc);
}

template <class T>
__global__ void __launch_bounds__(256) vec_op_dev_kernel(long long n, const T* __restrict__ sc, int mask, int ia, T sa,
                                                         const T* x, int ib, T sb, const T* y, T* out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    vec_op_dev_thread<T>(i, sc, mask, ia, sa, x, ib, sb, y, out);
}

constexpr int kDotBlocks = 592;   // 4 per SM on 148 SMs; FIXED, so the reduction tree never depends on the device state
constexpr int kDotThreads = 256;

// stage 1: block b sums the products of its grid-strided elements (thread-strided partials, shared-memory tree)
template <class T>
__global__ void __launch_bounds__(kDotThreads) dot_partial_kernel(long long n, const T* __restrict__ x,
                                                                  const T* __restrict__ y, T* __restrict__ partial) {
  __shared__ T part[kDotThreads];
  T acc = (T)0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += x[i] * y[i];
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int s = kDotThreads / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = part[0];
}

// stage 2: one block adds the kDotBlocks partials in a fixed order
template <class T>
__global__ void __launch_bounds__(1024) dot_final_kernel(const T* __restrict__ partial, int m, T* __restrict__ out) {
  __shared__ T part[1024];
  part[threadIdx.x] = ((int)threadIdx.x < m) ? partial[threadIdx.x] : (T)0;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = part[0];
}

static unsigned stream_grid(long long n) {
  const long long g = cdiv(n, 256);
  return (unsigned)(g < 1 ? 1 : (g > 148LL * 32 ? 148LL * 32 : g));
}

}  // namespace fol

using namespace fol;

extern "C" {

int fol_sell_spmv(fol_stream_t s, int dtype, int64_t nrows, const int64_t* slice_ptr, const int32_t* cols,
                  const void* vals, const void* x, void* y) {
  FOL_REQUIRE(nrows >= 0 && slice_ptr && cols && vals && x && y, "fol_sell_spmv: null pointer / negative size");
  FOL_REQUIRE(x != y, "fol_sell_spmv: x and y must not alias");
  if (nrows == 0) return FOL_OK;
  const unsigned grid = (unsigned)cdiv(nrows, 128);
  static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
  if (dtype == FOL_F64) {
    SellArgs<double> a{(const long long*)slice_ptr, cols, (const double*)vals, (const double*)x, (double*)y, nrows};
    sell_spmv_kernel<double><<<grid, 128, 0, (cudaStream_t)s>>>(a);
  } else if (dtype == FOL_F32) {
    SellArgs<float> a{(const long long*)slice_ptr, cols, (const float*)vals, (const float*)x, (float*)y, nrows};
    sell_spmv_kernel<float><<<grid, 128, 0, (cudaStream_t)s>>>(a);
  } else {
    return fail(FOL_ERR_INVALID, "fol_sell_spmv: unknown dtype");
  }
  return check_launch("sell_spmv_kernel");
}

int fol_sell_spmv_block(fol_stream_t s, int dtype, int dofs_per_node, int64_t nrows, const int64_t* slice_ptr,
                        const int32_t* node_cols, const void* vals, const void* x, void* y) {
  FOL_REQUIRE(nrows >= 0 && slice_ptr && node_cols && vals && x && y, "fol_sell_spmv_block: null pointer / negative size");
  FOL_REQUIRE(x != y, "fol_sell_spmv_block: x and y must not alias");
  FOL_REQUIRE(dofs_per_node == 2 || dofs_per_node == 3, "fol_sell_spmv_block: dofs_per_node must be 2 or 3");
  FOL_REQUIRE(dtype == FOL_F64 || dtype == FOL_F32, "fol_sell_spmv_block: unknown dtype");
  if (nrows == 0) return FOL_OK;
  const unsigned grid = (unsigned)cdiv(nrows, 128);
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == FOL_F64) {
    BlockSellArgs<double> a{(const long long*)slice_ptr, node_cols, (const double*)vals, (const double*)x, (double*)y, nrows};
    if (dofs_per_node == 3) sell_spmv_block_kernel<double, 3><<<grid, 128, 0, st>>>(a);
    else sell_spmv_block_kernel<double, 2><<<grid, 128, 0, st>>>(a);
  } else {
    BlockSellArgs<float> a{(const long long*)slice_ptr, node_cols, (const float*)vals, (const float*)x, (float*)y, nrows};
    if (dofs_per_node == 3) sell_spmv_block_kernel<float, 3><<<grid, 128, 0, st>>>(a);
    else sell_spmv_block_kernel<float, 2><<<grid, 128, 0, st>>>(a);
  }
  return check_launch("sell_spmv_block_kernel");
}

int fol_gather_values(fol_stream_t s, int dtype, int64_t n, const int32_t* src_index, const void* src, void* dst) {
  FOL_REQUIRE(n >= 0 && src_index && src && dst, "fol_gather_values: null pointer / negative size");
  if (n == 0) return FOL_OK;
  if (dtype == FOL_F64)
    gather_values_kernel<double><<<stream_grid(n), 256, 0, (cudaStream_t)s>>>(n, src_index, (const double*)src,
                                                                             (double*)dst);
  else if (dtype == FOL_F32)
    gather_values_kernel<float><<<stream_grid(n), 256, 0, (cudaStream_t)s>>>(n, src_index, (const float*)src,
                                                                            (float*)dst);
  else
    return fail(FOL_ERR_INVALID, "fol_gather_values: unknown dtype");
  return check_launch("gather_values_kernel");
}

int fol_vec_op(fol_stream_t s, int dtype, int op, int64_t n, double a, const void* x, double b, const void* y,
               void* out) {
  FOL_REQUIRE(n >= 0 && x && out, "fol_vec_op: null pointer / negative size");
  FOL_REQUIRE(op == VEC_AXPBY || op == VEC_AXY || op == VEC_AX_OVER_Y, "fol_vec_op: unknown op");
  FOL_REQUIRE(y || (op == VEC_AXPBY && b == 0.0), "fol_vec_op: y is required");
  if (n == 0) return FOL_OK;
  if (dtype == FOL_F64)
    vec_op_kernel<double><<<stream_grid(n), 256, 0, (cudaStream_t)s>>>(n, op, a, (const double*)x, b,
                                                                      (const double*)y, (double*)out);
  else if (dtype == FOL_F32)
    vec_op_kernel<float><<<stream_grid(n), 256, 0, (cudaStream_t)s>>>(n, op, (float)a, (const float*)x, (float)b,
                                                                     (const float*)y, (float*)out);
  else
    return fail(FOL_ERR_INVALID, "fol_vec_op: unknown dtype");
  return check_launch("vec_op_kernel");
}

int fol_bicg_scalar_count(void) { return BS_COUNT; }

int fol_bicg_scalars(fol_stream_t s, int dtype, int stage, void* scalars) {
  FOL_REQUIRE(scalars && stage >= BSTAGE_TOP && stage <= BSTAGE_END, "fol_bicg_scalars: bad arguments");
  if (dtype == FOL_F64) bicg_scalar_kernel<double><<<1, 32, 0, (cudaStream_t)s>>>(stage, (double*)scalars);
  else if (dtype == FOL_F32) bicg_scalar_kernel<float><<<1, 32, 0, (cudaStream_t)s>>>(stage, (float*)scalars);
  else return fail(FOL_ERR_INVALID, "fol_bicg_scalars: unknown dtype");
  return check_launch("bicg_scalar_kernel");
}

int fol_vec_op_dev(fol_stream_t s, int dtype, int64_t n, const void* scalars, int state_mask, int ia, double sa,
                   const void* x, int ib, double sb, const void* y, void* out) {
  FOL_REQUIRE(n >= 0 && scalars && x && out, "fol_vec_op_dev: null pointer / negative size");
  FOL_REQUIRE(ia < BS_COUNT && ib < BS_COUNT, "fol_vec_op_dev: scalar index out of range");
  if (n == 0) return FOL_OK;
  if (dtype == FOL_F64)
    vec_op_dev_kernel<double><<<stream_grid(n), 256, 0, (cudaStream_t)s>>>(n, (const double*)scalars, state_mask, ia, sa,
                                                                          (const double*)x, ib, sb, (const double*)y,
                                                                          (double*)out);
  else if (dtype == FOL_F32)
    vec_op_dev_kernel<float><<<stream_grid(n), 256, 0, (cudaStream_t)s>>>(n, (const float*)scalars, state_mask, ia,
                                                                         (float)sa, (const float*)x, ib, (float)sb,
                                                                         (const float*)y, (float*)out);
  else
    return fail(FOL_ERR_INVALID, "fol_vec_op_dev: unknown dtype");
  return check_launch("vec_op_dev_kernel");
}

int64_t fol_dot_work_size(void) { return kDotBlocks; }

int fol_dot(fol_stream_t s, int dtype, int64_t n, const void* x, const void* y, void* work, void* out) {
  FOL_REQUIRE(n >= 0 && x && y && work && out, "fol_dot: null pointer / negative size");
  if (dtype == FOL_F64) {
    dot_partial_kernel<double><<<kDotBlocks, kDotThreads, 0, (cudaStream_t)s>>>(n, (const double*)x, (const double*)y,
                                                                               (double*)work);
    dot_final_kernel<double><<<1, 1024, 0, (cudaStream_t)s>>>((const double*)work, kDotBlocks, (double*)out);
  } else if (dtype == FOL_F32) {
    dot_partial_kernel<float><<<kDotBlocks, kDotThreads, 0, (cudaStream_t)s>>>(n, (const float*)x, (const float*)y,
                                                                              (float*)work);
    dot_final_kernel<float><<<1, 1024, 0, (cudaStream_t)s>>>((const float*)work, kDotBlocks, (float*)out);
  } else {
    return fail(FOL_ERR_INVALID, "fol_dot: unknown dtype");
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);   // two kernels, check_launch counts one
  return check_launch("dot kernels");
}

}  // extern "C"
