// Kernels + C ABI of the adjoint-sensitivity element routines (adjoint.cuh): one element per thread.
// The per-thread bodies are __host__ __device__ functions (adjoint_threads.cuh) so that tests/host_shim runs
// exactly the code of a kernel thread -- indexing included -- on the CPU.
#include "adjoint_threads.cuh"

namespace fol {

template <class T, int ELEM, int ORDER>
__global__ void __launch_bounds__(128) gauss_interpolate_kernel(const InterpArgs<T> a) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < a.ne) gauss_interpolate_thread<T, ELEM, ORDER>(e, a);
}

template <class T, int ELEM, int ORDER>
__global__ void __launch_bounds__(128) response_kernel(const ResponseArgs<T> a) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < a.ne) response_thread<T, ELEM, ORDER>(e, a);
}

template <class T, int ELEM, int ORDER, int PHYS>
__global__ void __launch_bounds__(128) residual_adjoint_kernel(const AdjointArgs<T> a_in) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a_in.ne) return;
  if (a_in.batch_count > 0) {   // sample blockIdx.y of a batch: its own slices of ctrl / u / lam / dk
    AdjointArgs<T> a = a_in;
    const long long b = blockIdx.y;
    a.ctrl += b * a_in.batch_node;
    a.u += b * a_in.batch_dof;
    a.lam += b * a_in.batch_dof;
    if (a.dk) a.dk += b * a_in.batch_dk;
    residual_adjoint_thread<T, ELEM, ORDER, PHYS>(e, a);
  } else {
    residual_adjoint_thread<T, ELEM, ORDER, PHYS>(e, a_in);
  }
}

template <class T, int ELEM, int ORDER, int PHYS>
__global__ void __launch_bounds__(128) element_energy_kernel(const AdjointArgs<T> a) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < a.ne) element_energy_thread<T, ELEM, ORDER, PHYS>(e, a);
}

// out[0] = sum x[0..n): one block, fixed order (thread-strided partial sums, then a shared-memory tree)
template <class T>
__global__ void __launch_bounds__(1024) sum_kernel(const T* __restrict__ x, long long n, T* __restrict__ out) {
  __shared__ T part[1024];
  T acc = (T)0;
  for (long long i = threadIdx.x; i < n; i += 1024) acc += x[i];
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = part[0];
}

#define FOL_ELEM_ORDER_CASES(X)                                                                    \
  X(HEX, 1) X(HEX, 2) X(HEX, 3) X(QUAD, 1) X(QUAD, 2) X(QUAD, 3) X(TET, 1) X(TET, 2) X(TET, 3)     \
  X(TRI, 1) X(TRI, 2) X(TRI, 3)

template <class T>
static int launch_interp(cudaStream_t s, int element, int num_gp, const InterpArgs<T>& a) {
  if (a.ne == 0) return FOL_OK;
  const unsigned grid = (unsigned)cdiv(a.ne, 128);
#define X(E, O)                                                       \
  if (element == E && num_gp == O) {                                  \
    gauss_interpolate_kernel<T, E, O><<<grid, 128, 0, s>>>(a);        \
    return check_launch("gauss_interpolate_kernel");                  \
  }
  FOL_ELEM_ORDER_CASES(X)
#undef X
  return fail(FOL_ERR_UNSUPPORTED, "fol_gauss_interpolate: unsupported element / num_gp");
}

template <class T>
static int launch_response(cudaStream_t s, int element, int num_gp, const ResponseArgs<T>& a) {
  if (a.ne == 0) return FOL_OK;
  const unsigned grid = (unsigned)cdiv(a.ne, 128);
#define X(E, O)                                              \
  if (element == E && num_gp == O) {                         \
    response_kernel<T, E, O><<<grid, 128, 0, s>>>(a);        \
    return check_launch("response_kernel");                  \
  }
  FOL_ELEM_ORDER_CASES(X)
#undef X
  return fail(FOL_ERR_UNSUPPORTED, "fol_response_elements: unsupported element / num_gp");
}

template <class T, int PHYS>
static int launch_adjoint(cudaStream_t s, int element, int num_gp, const AdjointArgs<T>& a) {
  if (a.ne == 0) return FOL_OK;
  const unsigned grid = (unsigned)cdiv(a.ne, 128);
#define X(E, O)                                                            \
  if (element == E && num_gp == O) {                                       \
    residual_adjoint_kernel<T, E, O, PHYS><<<dim3(grid, (unsigned)(a.batch_count > 0 ? a.batch_count : 1)), 128, 0, s>>>(a); \
    return check_launch("residual_adjoint_kernel");                        \
  }
  FOL_ELEM_ORDER_CASES(X)
#undef X
  return fail(FOL_ERR_UNSUPPORTED, "fol_residual_adjoint_elements: unsupported element / num_gp");
}

template <class T, int PHYS>
static int launch_energy(cudaStream_t s, int element, int num_gp, const AdjointArgs<T>& a) {
  if (a.ne == 0) return FOL_OK;
  const unsigned grid = (unsigned)cdiv(a.ne, 128);
#define X(E, O)                                                          \
  if (element == E && num_gp == O) {                                     \
    element_energy_kernel<T, E, O, PHYS><<<grid, 128, 0, s>>>(a);        \
    return check_launch("element_energy_kernel");                        \
  }
  FOL_ELEM_ORDER_CASES(X)
#undef X
  return fail(FOL_ERR_UNSUPPORTED, "fol_element_energies: unsupported element / num_gp");
}

template <class T>
static int energies_typed(cudaStream_t s, int physics, int element, int num_gp, long long ne, const void* xyz,
                          const int32_t* conn, const void* ctrl, const void* u, const void* aux, const double* params,
                          void* energy) {
  AdjointArgs<T> a;
  a.xyz = (const T*)xyz;
  a.conn = conn;
  a.ctrl = (const T*)ctrl;
  a.u = (const T*)u;
  a.lam = nullptr;
  a.aux = (const T*)aux;
  a.dk = (T*)energy;
  a.dx = nullptr;
  a.ne = ne;
  a.accumulate = 0;
  a.p = make_params<T>(params);
  switch (physics) {
    case FOL_MECHANICAL: return launch_energy<T, ADJ_MECH>(s, element, num_gp, a);
    case FOL_THERMAL: return launch_energy<T, ADJ_THERMAL>(s, element, num_gp, a);
    case FOL_NEOHOOKE: return launch_energy<T, ADJ_NEOHOOKE>(s, element, num_gp, a);
    case FOL_STVENANT: return launch_energy<T, ADJ_STVK>(s, element, num_gp, a);
    case FOL_ALLEN_CAHN: return launch_energy<T, ADJ_ALLENCAHN>(s, element, num_gp, a);
    case FOL_TRANSIENT_THERMAL:
      if (!aux) return fail(FOL_ERR_INVALID, "fol_element_energies: transient thermal needs the nodal k0 in aux");
      return launch_energy<T, ADJ_TTHERMAL>(s, element, num_gp, a);
  }
  return fail(FOL_ERR_UNSUPPORTED, "fol_element_energies: no stateless element energy for this physics");
}

template <class T>
static int interp_typed(cudaStream_t s, int element, int num_gp, int dpn, long long ne, const int32_t* conn,
                        const void* ctrl, const void* u, void* kg, void* ug) {
  InterpArgs<T> a;
  a.conn = conn;
  a.ctrl = (const T*)ctrl;
  a.u = (const T*)u;
  a.kg = (T*)kg;
  a.ug = (T*)ug;
  a.ne = ne;
  a.dpn = dpn;
  return launch_interp<T>(s, element, num_gp, a);
}

template <class T>
static int response_typed(cudaStream_t s, int element, int num_gp, int dpn, long long ne, const void* xyz,
                          const int32_t* conn, const void* f, const void* fk, const void* fu, void* val, void* du,
                          void* dk, void* dx) {
  ResponseArgs<T> a;
  a.xyz = (const T*)xyz;
  a.conn = conn;
  a.f = (const T*)f;
  a.fk = (const T*)fk;
  a.fu = (const T*)fu;
  a.val = (T*)val;
  a.du = (T*)du;
  a.dk = (T*)dk;
  a.dx = (T*)dx;
  a.ne = ne;
  a.dpn = dpn;
  return launch_response<T>(s, element, num_gp, a);
}

template <class T>
static int adjoint_typed(cudaStream_t s, int physics, int element, int num_gp, int accumulate, long long ne,
                         const void* xyz, const int32_t* conn, const void* ctrl, const void* u, const void* lam,
                         const void* aux, const double* params, void* dk, void* dx, long long nb = 0, long long nn = 0) {
  AdjointArgs<T> a;
  if (nb > 0) {   // batched: per-sample strides of ctrl / (u, lam) / dk
    const int dpn = (physics == FOL_THERMAL || physics == FOL_TRANSIENT_THERMAL || physics == FOL_ALLEN_CAHN) ? 1 : elem_dim(element);
    a.batch_count = (int)nb;
    a.batch_node = nn;
    a.batch_dof = nn * dpn;
    a.batch_dk = ne * (long long)elem_nnode(element);
  }
  a.xyz = (const T*)xyz;
  a.conn = conn;
  a.ctrl = (const T*)ctrl;
  a.u = (const T*)u;
  a.lam = (const T*)lam;
  a.aux = (const T*)aux;
  a.dk = (T*)dk;
  a.dx = (T*)dx;
  a.ne = ne;
  a.accumulate = accumulate;
  a.p = make_params<T>(params);
  if (physics == FOL_MECHANICAL) return launch_adjoint<T, ADJ_MECH>(s, element, num_gp, a);
  if (physics == FOL_THERMAL) return launch_adjoint<T, ADJ_THERMAL>(s, element, num_gp, a);
  if (physics == FOL_NEOHOOKE) return launch_adjoint<T, ADJ_NEOHOOKE>(s, element, num_gp, a);
  if (physics == FOL_STVENANT) return launch_adjoint<T, ADJ_STVK>(s, element, num_gp, a);
  if (physics == FOL_ALLEN_CAHN) return launch_adjoint<T, ADJ_ALLENCAHN>(s, element, num_gp, a);
  if (physics == FOL_TRANSIENT_THERMAL) {
    if (!aux) return fail(FOL_ERR_INVALID, "fol_residual_adjoint_elements: transient thermal needs the nodal k0 in aux");
    return launch_adjoint<T, ADJ_TTHERMAL>(s, element, num_gp, a);
  }
  return fail(FOL_ERR_UNSUPPORTED,
              "fol_residual_adjoint_elements: no residual sensitivities for history-dependent (J2) elements");
}

}  // namespace fol

using namespace fol;

extern "C" {

int fol_gauss_interpolate(fol_stream_t s, int dtype, int element, int num_gp, int dofs_per_node, int64_t ne,
                          const int32_t* conn, const void* ctrl, const void* u, void* k_gp, void* u_gp) {
  FOL_REQUIRE(element >= 0 && element <= 3 && num_gp >= 1 && num_gp <= 3, "fol_gauss_interpolate: bad element / num_gp");
  FOL_REQUIRE(dofs_per_node >= 1 && dofs_per_node <= 3, "fol_gauss_interpolate: dofs_per_node must be 1..3");
  FOL_REQUIRE(ne >= 0 && conn && ctrl && u && k_gp && u_gp, "fol_gauss_interpolate: null pointer / negative size");
  if (dtype == FOL_F64)
    return interp_typed<double>((cudaStream_t)s, element, num_gp, dofs_per_node, ne, conn, ctrl, u, k_gp, u_gp);
  if (dtype == FOL_F32)
    return interp_typed<float>((cudaStream_t)s, element, num_gp, dofs_per_node, ne, conn, ctrl, u, k_gp, u_gp);
  return fail(FOL_ERR_INVALID, "fol_gauss_interpolate: unknown dtype");
}

int fol_response_elements(fol_stream_t s, int dtype, int element, int num_gp, int dofs_per_node, int64_t ne,
                          const void* xyz, const int32_t* conn, const void* f_gp, const void* fk_gp,
                          const void* fu_gp, void* value_elem, void* du_elem, void* dk_elem, void* dx_elem) {
  FOL_REQUIRE(element >= 0 && element <= 3 && num_gp >= 1 && num_gp <= 3, "fol_response_elements: bad element / num_gp");
  FOL_REQUIRE(dofs_per_node >= 1 && dofs_per_node <= 3, "fol_response_elements: dofs_per_node must be 1..3");
  FOL_REQUIRE(ne >= 0 && xyz && conn && f_gp, "fol_response_elements: null pointer / negative size");
  FOL_REQUIRE(!du_elem || fu_gp, "fol_response_elements: du_elem needs fu_gp");
  FOL_REQUIRE(!dk_elem || fk_gp, "fol_response_elements: dk_elem needs fk_gp");
  if (dtype == FOL_F64)
    return response_typed<double>((cudaStream_t)s, element, num_gp, dofs_per_node, ne, xyz, conn, f_gp, fk_gp, fu_gp,
                                  value_elem, du_elem, dk_elem, dx_elem);
  if (dtype == FOL_F32)
    return response_typed<float>((cudaStream_t)s, element, num_gp, dofs_per_node, ne, xyz, conn, f_gp, fk_gp, fu_gp,
                                 value_elem, du_elem, dk_elem, dx_elem);
  return fail(FOL_ERR_INVALID, "fol_response_elements: unknown dtype");
}

int fol_residual_adjoint_elements(fol_stream_t s, int dtype, int physics, int element, int num_gp, int accumulate,
                                  int64_t ne, const void* xyz, const int32_t* conn, const void* ctrl, const void* u,
                                  const void* adj, const void* aux, const double* params_host, void* dk_elem,
                                  void* dx_elem) {
  FOL_REQUIRE(element >= 0 && element <= 3 && num_gp >= 1 && num_gp <= 3,
              "fol_residual_adjoint_elements: bad element / num_gp");
  FOL_REQUIRE(ne >= 0 && xyz && conn && ctrl && u && adj && params_host && (dk_elem || dx_elem),
              "fol_residual_adjoint_elements: null pointer / negative size");
  if (dtype == FOL_F64)
    return adjoint_typed<double>((cudaStream_t)s, physics, element, num_gp, accumulate, ne, xyz, conn, ctrl, u, adj,
                                 aux, params_host, dk_elem, dx_elem);
  if (dtype == FOL_F32)
    return adjoint_typed<float>((cudaStream_t)s, physics, element, num_gp, accumulate, ne, xyz, conn, ctrl, u, adj,
                                aux, params_host, dk_elem, dx_elem);
  return fail(FOL_ERR_INVALID, "fol_residual_adjoint_elements: unknown dtype");
}

/* lam^T d(re)/dK of a BATCH of samples in one launch (grid.y = sample): ctrl (nb, nn), u and adj (nb, ndof) ->
 * dk_elem (nb, ne*A).  The nested VJP of the batched loss (fe_loss.py _SecondOrderFn) needs it per sample. */
int fol_residual_adjoint_elements_batched(fol_stream_t s, int dtype, int physics, int element, int num_gp, int64_t ne,
                                          int64_t nn, int64_t nb, const void* xyz, const int32_t* conn,
                                          const void* ctrl, const void* u, const void* adj, const void* aux,
                                          const double* params_host, void* dk_elem) {
  FOL_REQUIRE(element >= 0 && element <= 3 && num_gp >= 1 && num_gp <= 3,
              "fol_residual_adjoint_elements_batched: bad element / num_gp");
  FOL_REQUIRE(ne >= 0 && nn >= 1 && nb >= 1 && nb <= 65535 && xyz && conn && ctrl && u && adj && params_host && dk_elem,
              "fol_residual_adjoint_elements_batched: null pointer / bad size (1 <= nb <= 65535)");
  if (dtype == FOL_F64)
    return adjoint_typed<double>((cudaStream_t)s, physics, element, num_gp, 0, ne, xyz, conn, ctrl, u, adj, aux,
                                 params_host, dk_elem, nullptr, nb, nn);
  if (dtype == FOL_F32)
    return adjoint_typed<float>((cudaStream_t)s, physics, element, num_gp, 0, ne, xyz, conn, ctrl, u, adj, aux,
                                params_host, dk_elem, nullptr, nb, nn);
  return fail(FOL_ERR_INVALID, "fol_residual_adjoint_elements_batched: unknown dtype");
}

int fol_element_energies(fol_stream_t s, int dtype, int physics, int element, int num_gp, int64_t ne, const void* xyz,
                         const int32_t* conn, const void* ctrl, const void* u, const void* aux,
                         const double* params_host, void* energy_elem) {
  FOL_REQUIRE(element >= 0 && element <= 3 && num_gp >= 1 && num_gp <= 3, "fol_element_energies: bad element / num_gp");
  FOL_REQUIRE(ne >= 0 && xyz && conn && ctrl && u && params_host && energy_elem,
              "fol_element_energies: null pointer / negative size");
  if (dtype == FOL_F64)
    return energies_typed<double>((cudaStream_t)s, physics, element, num_gp, ne, xyz, conn, ctrl, u, aux, params_host,
                                  energy_elem);
  if (dtype == FOL_F32)
    return energies_typed<float>((cudaStream_t)s, physics, element, num_gp, ne, xyz, conn, ctrl, u, aux, params_host,
                                 energy_elem);
  return fail(FOL_ERR_INVALID, "fol_element_energies: unknown dtype");
}

int fol_sum(fol_stream_t s, int dtype, int64_t n, const void* x, void* out) {
  FOL_REQUIRE(n >= 0 && out && (x || n == 0), "fol_sum: null pointer / negative size");
  if (dtype == FOL_F64) sum_kernel<double><<<1, 1024, 0, (cudaStream_t)s>>>((const double*)x, n, (double*)out);
  else if (dtype == FOL_F32) sum_kernel<float><<<1, 1024, 0, (cudaStream_t)s>>>((const float*)x, n, (float*)out);
  else return fail(FOL_ERR_INVALID, "fol_sum: unknown dtype");
  return check_launch("sum_kernel");
}

}  // extern "C"
