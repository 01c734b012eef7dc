// TEST INFRASTRUCTURE: the __host__ __device__ per-thread bodies of folax_b200/csrc/krylov_threads.cuh looped on the
// CPU (float64, host pointers) -- see adjoint_host.cu.  Never linked into libfolax_b200, never used by the product.
#include "../../folax_b200/csrc/krylov_threads.cuh"

using namespace fol;

extern "C" {

int host_sell_spmv(long long nrows, const long long* slice_ptr, const int32_t* cols, const double* vals,
                   const double* x, double* y) {
  SellArgs<double> a{slice_ptr, cols, vals, x, y, nrows};
  for (long long r = 0; r < nrows; ++r) sell_spmv_thread<double>(r, a);
  return 0;
}

int host_sell_spmv_block(int d, long long nrows, const long long* slice_ptr, const int32_t* node_cols, const double* vals,
                         const double* x, double* y) {
  BlockSellArgs<double> a{slice_ptr, node_cols, vals, x, y, nrows};
  for (long long r = 0; r < nrows; ++r) {
    if (d == 3) sell_spmv_block_thread<double, 3>(r, a);
    else sell_spmv_block_thread<double, 2>(r, a);
  }
  return 0;
}

int host_gather_values(long long n, const int32_t* src_index, const double* src, double* dst) {
  for (long long i = 0; i < n; ++i) gather_values_thread<double>(i, src_index, src, dst);
  return 0;
}

int host_vec_op(int op, long long n, double a, const double* x, double b, const double* y, double* out) {
  for (long long i = 0; i < n; ++i) vec_op_thread<double>(i, op, a, x, b, y, out);
  return 0;
}

int host_bicg_scalar_count(void) { return BS_COUNT; }

int host_bicg_scalars(int stage, double* sc) {
  bicg_scalar_stage<double>(stage, sc);
  return 0;
}

int host_vec_op_dev(long long n, const double* sc, int mask, int ia, double sa, const double* x, int ib, double sb,
                    const double* y, double* out) {
  for (long long i = 0; i < n; ++i) vec_op_dev_thread<double>(i, sc, mask, ia, sa, x, ib, sb, y, out);
  return 0;
}

}  // extern "C"
