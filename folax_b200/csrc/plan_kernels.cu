// One-off integer "plan" kernels (bit-exact) and the small streaming kernels around the element
// stage: BCOO indices, Dirichlet flags, node->element adjacency, deterministic residual gather,
// Dirichlet overwrite of dof vectors.
#include "common.cuh"

namespace fol {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

// indices[(e*nd*nd + i*nd + j)] = (gdof(e,i), gdof(e,j)), gdof(e,i) = d*conn[e,i/d] + i%d
// (ComputeElementJacobianIndices, fe_loss.py:178-184).  One int2 (8 B) per thread, coalesced.
__global__ void bcoo_indices_kernel(const int32_t* __restrict__ conn, long long ne, int nnode, int d,
                                    int2* __restrict__ out) {
  const int nd = nnode * d;
  const long long total = ne * nd * nd;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const long long e = t / (nd * nd);
    const int ij = (int)(t - e * nd * nd);
    const int i = ij / nd, j = ij - i * nd;
    const int gi = d * __ldg(conn + e * nnode + i / d) + i % d;
    const int gj = d * __ldg(conn + e * nnode + j / d) + j % d;
    out[t] = make_int2(gi, gj);
  }
}

__global__ void dirichlet_flags_kernel(const int32_t* __restrict__ idx, long long n, uint8_t* __restrict__ flag) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) flag[idx[t]] = 1;
}

// ---- node -> (element, local node) adjacency: count, scan, fill, per-node sort -------------
__global__ void adj_count_kernel(const int32_t* __restrict__ conn, long long total, int32_t* __restrict__ deg) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < total) atomicAdd(deg + conn[t], 1);
}

// single-block exclusive scan (one-off; n up to a few 10^7): each thread scans a contiguous chunk
__global__ void exclusive_scan_kernel(const int32_t* __restrict__ in, long long n, int32_t* __restrict__ out) {
  __shared__ long long part[1024];
  const int t = threadIdx.x, nt = blockDim.x;
  const long long chunk = (n + nt - 1) / nt;
  const long long lo = (long long)t * chunk, hi = (lo + chunk < n) ? lo + chunk : n;
  long long s = 0;
  for (long long i = lo; i < hi; ++i) s += in[i];
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    long long run = 0;
    for (int k = 0; k < nt; ++k) {
      const long long v = part[k];
      part[k] = run;
      run += v;
    }
  }
  __syncthreads();
  long long run = part[t];
  for (long long i = lo; i < hi; ++i) {
    const int32_t v = in[i];
    out[i] = (int32_t)run;
    run += v;
  }
  if (t == nt - 1) out[n] = (int32_t)(part[t] + s);
}

__global__ void adj_fill_kernel(const int32_t* __restrict__ conn, long long total, const int32_t* __restrict__ ptr,
                                int32_t* __restrict__ cursor, int32_t* __restrict__ adj) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < total) {
    const int n = conn[t];
    const int pos = atomicAdd(cursor + n, 1);
    adj[ptr[n] + pos] = (int32_t)t;  // t = e*nnode + a
  }
}

// the atomic fill order is arbitrary: sort each node's short list so the later float sums have a
// fixed order (ascending element id, then local node)
__global__ void adj_sort_kernel(const int32_t* __restrict__ ptr, long long nn, int32_t* __restrict__ adj) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nn) return;
  const int lo = ptr[n], hi = ptr[n + 1];
  for (int i = lo + 1; i < hi; ++i) {
    const int32_t v = adj[i];
    int j = i - 1;
    while (j >= lo && adj[j] > v) {
      adj[j + 1] = adj[j];
      --j;
    }
    adj[j + 1] = v;
  }
}

// R[d*n+k] = sum_{(e,a) in adj(n)} re_elem[(e*A + a)*d + k]   (replaces fe_loss.py:301-306)
template <class T>
__global__ void residual_gather_kernel(long long nn, int d, const int32_t* __restrict__ ptr,
                                       const int32_t* __restrict__ adj, const T* __restrict__ re, T* __restrict__ R) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nn * d) return;
  const long long n = t / d;
  const int k = (int)(t - n * d);
  T acc = (T)0;
  const int lo = ptr[n], hi = ptr[n + 1];
  for (int i = lo; i < hi; ++i) acc += __ldg(re + (long long)adj[i] * d + k);
  R[t] = acc;
}

// the same gather for a batch of samples (grid.y = sample): re is (nb, ne*A*d), R is (nb, nn*d)
template <class T>
__global__ void residual_gather_batched_kernel(long long nn, int d, long long re_stride, const int32_t* __restrict__ ptr,
                                               const int32_t* __restrict__ adj, const T* __restrict__ re,
                                               T* __restrict__ R) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nn * d) return;
  const long long n = t / d;
  const int k = (int)(t - n * d);
  const T* src = re + (long long)blockIdx.y * re_stride;
  T acc = (T)0;
  const int lo = ptr[n], hi = ptr[n + 1];
  for (int i = lo; i < hi; ++i) acc += __ldg(src + (long long)adj[i] * d + k);
  R[(long long)blockIdx.y * nn * d + t] = acc;
}

template <class T>
__global__ void apply_dirichlet_kernel(long long nb, long long ndof, const int32_t* __restrict__ idx, long long nd,
                                       const T* __restrict__ values, int per_sample, T load, T* __restrict__ u) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long b = blockIdx.y;
  if (t < nd) u[b * ndof + idx[t]] = load * values[per_sample ? b * nd + t : t];
}

}  // namespace fol

using namespace fol;

extern "C" {

const char* fol_last_error(void) { return g_last_error.c_str(); }
int fol_version(void) { return 100; }
int64_t fol_launch_count(void) { return (int64_t)g_launches.load(); }

int fol_bcoo_indices(fol_stream_t s, const int32_t* conn, int64_t ne, int nnode, int d, int32_t* indices) {
  FOL_REQUIRE(conn && indices && ne >= 0 && nnode > 0 && d > 0, "fol_bcoo_indices: bad arguments");
  if (ne == 0) return FOL_OK;
  const long long total = (long long)ne * nnode * d * nnode * d;
  const long long blocks = cdiv(total, 256);
  const unsigned grid = (unsigned)(blocks < 148LL * 64 ? blocks : 148LL * 64);
  bcoo_indices_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(conn, ne, nnode, d, reinterpret_cast<int2*>(indices));
  return check_launch("bcoo_indices_kernel");
}

int fol_dirichlet_flags(fol_stream_t s, const int32_t* idx, int64_t n, int64_t ndof, uint8_t* flag) {
  FOL_REQUIRE(flag && ndof >= 0 && n >= 0, "fol_dirichlet_flags: bad arguments");
  FOL_CUDA(cudaMemsetAsync(flag, 0, (size_t)ndof, (cudaStream_t)s));
  if (n == 0) return FOL_OK;
  dirichlet_flags_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)s>>>(idx, n, flag);
  return check_launch("dirichlet_flags_kernel");
}

int fol_node_adjacency(fol_stream_t s, const int32_t* conn, int64_t ne, int nnode, int64_t nn, int32_t* adj_ptr,
                       int32_t* adj, int32_t* work) {
  FOL_REQUIRE(conn && adj_ptr && adj && work && nn > 0, "fol_node_adjacency: bad arguments");
  cudaStream_t st = (cudaStream_t)s;
  const long long total = (long long)ne * nnode;
  FOL_CUDA(cudaMemsetAsync(work, 0, sizeof(int32_t) * (size_t)nn, st));
  if (total > 0) {
    adj_count_kernel<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(conn, total, work);
    if (int rc = check_launch("adj_count_kernel")) return rc;
  }
  exclusive_scan_kernel<<<1, 1024, 0, st>>>(work, nn, adj_ptr);
  if (int rc = check_launch("exclusive_scan_kernel")) return rc;
  FOL_CUDA(cudaMemsetAsync(work, 0, sizeof(int32_t) * (size_t)nn, st));
  if (total > 0) {
    adj_fill_kernel<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(conn, total, adj_ptr, work, adj);
    if (int rc = check_launch("adj_fill_kernel")) return rc;
  }
  adj_sort_kernel<<<(unsigned)cdiv(nn, 128), 128, 0, st>>>(adj_ptr, nn, adj);
  return check_launch("adj_sort_kernel");
}

int fol_residual_gather(fol_stream_t s, int dtype, int64_t nn, int nnode, int d, const int32_t* adj_ptr,
                        const int32_t* adj, const void* re_elem, void* residual) {
  (void)nnode;
  FOL_REQUIRE(adj_ptr && adj && re_elem && residual, "fol_residual_gather: null pointer");
  const long long total = (long long)nn * d;
  if (total == 0) return FOL_OK;
  const unsigned grid = (unsigned)cdiv(total, 256);
  if (dtype == FOL_F64)
    residual_gather_kernel<double><<<grid, 256, 0, (cudaStream_t)s>>>(nn, d, adj_ptr, adj, (const double*)re_elem,
                                                                      (double*)residual);
  else
    residual_gather_kernel<float><<<grid, 256, 0, (cudaStream_t)s>>>(nn, d, adj_ptr, adj, (const float*)re_elem,
                                                                     (float*)residual);
  return check_launch("residual_gather_kernel");
}

int fol_residual_gather_batched(fol_stream_t s, int dtype, int64_t nn, int nnode, int d, int64_t nb, int64_t ne,
                                const int32_t* adj_ptr, const int32_t* adj, const void* re_elem, void* residual) {
  FOL_REQUIRE(adj_ptr && adj && re_elem && residual && nb >= 1 && nb <= 65535, "fol_residual_gather_batched: bad arguments");
  const long long total = (long long)nn * d;
  if (total == 0) return FOL_OK;
  const dim3 grid((unsigned)cdiv(total, 256), (unsigned)nb);
  const long long stride = (long long)ne * nnode * d;
  if (dtype == FOL_F64)
    residual_gather_batched_kernel<double><<<grid, 256, 0, (cudaStream_t)s>>>(nn, d, stride, adj_ptr, adj,
                                                                              (const double*)re_elem, (double*)residual);
  else
    residual_gather_batched_kernel<float><<<grid, 256, 0, (cudaStream_t)s>>>(nn, d, stride, adj_ptr, adj,
                                                                             (const float*)re_elem, (float*)residual);
  return check_launch("residual_gather_batched_kernel");
}

int fol_apply_dirichlet(fol_stream_t s, int dtype, int64_t nb, int64_t ndof, const int32_t* idx, int64_t nd,
                        const void* values, int per_sample, double load, void* u) {
  FOL_REQUIRE(u && nb >= 0, "fol_apply_dirichlet: bad arguments");
  if (nd == 0 || nb == 0) return FOL_OK;
  dim3 grid((unsigned)cdiv(nd, 256), (unsigned)nb);
  if (dtype == FOL_F64)
    apply_dirichlet_kernel<double><<<grid, 256, 0, (cudaStream_t)s>>>(nb, ndof, idx, nd, (const double*)values,
                                                                      per_sample, load, (double*)u);
  else
    apply_dirichlet_kernel<float><<<grid, 256, 0, (cudaStream_t)s>>>(nb, ndof, idx, nd, (const float*)values,
                                                                     per_sample, (float)load, (float*)u);
  return check_launch("apply_dirichlet_kernel");
}

}  // extern "C"

// ---- duplicate-free CSR values from the BCOO data (hand-off to fol/solvers: fe_solver.py:71-72, 82) ----
// One thread per (node pair, i, j): fixed-order sum of the contributing element entries (ascending
// e*A*A + a*A + b), so the CSR values are deterministic.  The integer plan (pair_ptr, contrib,
// out_base, row_stride) is built once per mesh by folax_b200/csr_plan.py.
namespace fol {
template <class T>
__global__ void csr_values_kernel(long long npairs, int d, int A, const int32_t* __restrict__ pair_ptr,
                                  const int32_t* __restrict__ contrib, const int32_t* __restrict__ out_base,
                                  const int32_t* __restrict__ row_stride, const T* __restrict__ data,
                                  T* __restrict__ vals) {
  const int dd = d * d, nd = A * d;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npairs * dd) return;
  const long long p = t / dd;
  const int ij = (int)(t - p * dd), i = ij / d, j = ij - i * d;
  T acc = (T)0;
  const int lo = __ldg(pair_ptr + p), hi = __ldg(pair_ptr + p + 1);
  for (int k = lo; k < hi; ++k) {
    const int c = __ldg(contrib + k);
    const long long e = c / (A * A);
    const int ab = c - (int)e * (A * A), a = ab / A, b = ab - a * A;
    acc += __ldg(data + e * (long long)(nd * nd) + (a * d + i) * nd + (b * d + j));
  }
  vals[(long long)__ldg(out_base + p) + (long long)i * __ldg(row_stride + p) + j] = acc;
}
}  // namespace fol

extern "C" int fol_csr_values(fol_stream_t s, int dtype, int64_t npairs, int d, int nnode, const int32_t* pair_ptr,
                              const int32_t* contrib, const int32_t* out_base, const int32_t* row_stride,
                              const void* data, void* vals) {
  FOL_REQUIRE(pair_ptr && contrib && out_base && row_stride && data && vals, "fol_csr_values: null pointer");
  const long long total = (long long)npairs * d * d;
  if (total == 0) return FOL_OK;
  const unsigned grid = (unsigned)cdiv(total, 256);
  if (dtype == FOL_F64)
    csr_values_kernel<double><<<grid, 256, 0, (cudaStream_t)s>>>(npairs, d, nnode, pair_ptr, contrib, out_base,
                                                                row_stride, (const double*)data, (double*)vals);
  else
    csr_values_kernel<float><<<grid, 256, 0, (cudaStream_t)s>>>(npairs, d, nnode, pair_ptr, contrib, out_base,
                                                               row_stride, (const float*)data, (float*)vals);
  return check_launch("csr_values_kernel");
}
