// extern "C" surface of libfolax_b200 (see include/folax_b200.h) + the host-buffer plan object.
#include <vector>

#include "assemble.cuh"
#include "energy.cuh"
#include "assemble_ad_args.cuh"

namespace fol {
int assemble_mech_f64(cudaStream_t, int, int, const AsmArgs<double>&);
int assemble_mech_f32(cudaStream_t, int, int, const AsmArgs<float>&);
int assemble_thermal_f64(cudaStream_t, int, int, const AsmArgs<double>&);
int assemble_thermal_f32(cudaStream_t, int, int, const AsmArgs<float>&);
int assemble_neohooke_f64(cudaStream_t, int, int, const AsmArgs<double>&);
int assemble_neohooke_f32(cudaStream_t, int, int, const AsmArgs<float>&);
int assemble_tthermal_f64(cudaStream_t, int, int, const AsmArgs<double>&);
int assemble_tthermal_f32(cudaStream_t, int, int, const AsmArgs<float>&);
int assemble_allencahn_f64(cudaStream_t, int, int, const AsmArgs<double>&);
int assemble_allencahn_f32(cudaStream_t, int, int, const AsmArgs<float>&);
int assemble_stvk_f64(cudaStream_t, int, int, const AsmArgs<double>&);
int assemble_stvk_f32(cudaStream_t, int, int, const AsmArgs<float>&);
int assemble_j2_f64(cudaStream_t, int, int, const AsmArgs<double>&);
int assemble_j2_f32(cudaStream_t, int, int, const AsmArgs<float>&);
namespace hexk { struct HaloFuse; }
int assemble_hex_mech_f64(cudaStream_t, const AsmArgs<double>&, const hexk::HaloFuse* = nullptr);
int assemble_hex_j2_f64(cudaStream_t, const AsmArgs<double>&, const hexk::HaloFuse* = nullptr);
int assemble_hex_mech_f32(cudaStream_t, const AsmArgs<float>&);
int assemble_hex_thermal_f64(cudaStream_t, const AsmArgs<double>&);
int assemble_quad_mech_f64(cudaStream_t, const AsmArgs<double>&);
template <class T>
int assemble_ad(cudaStream_t, int, int, int, const AdAsmArgs<T>&);
extern std::atomic<int> g_grid_margin;

template <class T>
int energy_and_grads(cudaStream_t, int, int, int, const EnergyArgs<T>&, int, T*);
template <class T>
int geometry_cache(cudaStream_t, int, int, int, long long, const T*, const int32_t*, const T*, T*);
template <class T>
int loss_reduce(cudaStream_t, long long, double, const T*, T*, T*);
template <class T>
int scale_grads(cudaStream_t, long long, long long, long long, const T*, double, const T*, int, const uint8_t*, T*, T*);

static bool valid_element(int e) { return e >= 0 && e <= 3; }
static std::atomic<int> g_tuned{1};

template <class T>
static int assemble_typed(cudaStream_t s, int physics, int element, int num_gp, int transpose, long long ne,
                          const void* xyz, const int32_t* conn, const void* ctrl, const void* u, const uint8_t* dir,
                          const double* params, void* ke, void* re, const void* st_in, void* st_out,
                          const void* v = nullptr, long long nb = 0, long long nn = 0) {
  AsmArgs<T> a;
  a.v = (const T*)v;
  if (nb > 0) {   // batched matrix-free mode: per-sample strides of ctrl / (u, v) / ye
    const int dpn = (physics == FOL_THERMAL || physics == FOL_TRANSIENT_THERMAL || physics == FOL_ALLEN_CAHN) ? 1 : elem_dim(element);
    a.batch_count = (int)nb;
    a.batch_node = nn;
    a.batch_dof = nn * dpn;
    a.batch_elem = ne * (long long)elem_nnode(element) * dpn;
  }
  a.xyz = (const T*)xyz;
  a.conn = conn;
  a.ctrl = (const T*)ctrl;
  a.u = (const T*)u;
  a.dir = dir;
  a.ke = (T*)ke;
  a.re = (T*)re;
  a.state_in = (const T*)st_in;
  a.state_out = (T*)st_out;
  a.ne = ne;
  a.transpose = transpose;
  a.p = make_params<T>(params);
  constexpr bool f64 = sizeof(T) == 8;
  switch (physics) {
    case FOL_MECHANICAL:
      if constexpr (f64) {
        // (the tuned kernel writes Ke itself; the transpose switch of fe_loss.py:216-230 uses the generic kernel)
        if (element == HEX && num_gp == 2 && !transpose && !v && g_tuned.load()) return assemble_hex_mech_f64(s, a);
        if (element == QUAD && num_gp == 2 && !transpose && !v && g_tuned.load()) {
          const int rc = assemble_quad_mech_f64(s, a);    // 1 = not applicable (unaligned output): generic kernel
          if (rc != 1) return rc;
        }
        return assemble_mech_f64(s, element, num_gp, a);
      } else {
        if (element == HEX && num_gp == 2 && !transpose && !v && g_tuned.load()) {
          const int rc = assemble_hex_mech_f32(s, a);     // 1 = not applicable (unaligned output): generic kernel
          if (rc != 1) return rc;
        }
        return assemble_mech_f32(s, element, num_gp, a);
      }
    case FOL_THERMAL:
      if constexpr (f64) {
        // tuned Hex8 kernel (csrc/assemble_hex_thermal.cu); transpose and the matrix-free mode use the generic kernel
        if (element == HEX && num_gp == 2 && !transpose && !v && g_tuned.load()) {
          const int rc = assemble_hex_thermal_f64(s, a);  // 1 = not applicable (unaligned output): generic kernel
          if (rc != 1) return rc;
        }
        return assemble_thermal_f64(s, element, num_gp, a);
      } else {
        return assemble_thermal_f32(s, element, num_gp, a);
      }
    case FOL_NEOHOOKE:
      if constexpr (f64) return assemble_neohooke_f64(s, element, num_gp, a);
      else return assemble_neohooke_f32(s, element, num_gp, a);
    case FOL_TRANSIENT_THERMAL:
      if (!st_in) return fail(FOL_ERR_INVALID, "transient thermal needs the nodal heterogeneity k0 in aux_in");
      if constexpr (f64) return assemble_tthermal_f64(s, element, num_gp, a);
      else return assemble_tthermal_f32(s, element, num_gp, a);
    case FOL_ALLEN_CAHN:
      if constexpr (f64) return assemble_allencahn_f64(s, element, num_gp, a);
      else return assemble_allencahn_f32(s, element, num_gp, a);
    case FOL_STVENANT:
      if constexpr (f64) return assemble_stvk_f64(s, element, num_gp, a);
      else return assemble_stvk_f32(s, element, num_gp, a);
    case FOL_NEOHOOKE_AD:
    case FOL_STVENANT_AD: {
      if (v) return fail(FOL_ERR_UNSUPPORTED, "fol_apply_jacobian_elements: not available for the AD loss variants");
      AdAsmArgs<T> ad{a.xyz, a.conn, a.ctrl, a.u, a.dir, a.ke, a.re, (T*)st_out, ne, transpose, a.p};
      return assemble_ad<T>(s, physics, element, num_gp, ad);
    }
#ifdef FOL_HAVE_J2
    case FOL_J2PLASTICITY:
      if (!st_in || (!st_out && !v)) return fail(FOL_ERR_INVALID, "J2 plasticity needs state_in/state_out");
      if constexpr (f64) {
        // tuned Hex8 kernel (csrc/assemble_hex_j2.cu); transpose and the matrix-free mode use the generic kernel
        if (element == HEX && num_gp == 2 && !transpose && !v && st_out && g_tuned.load()) return assemble_hex_j2_f64(s, a);
        return assemble_j2_f64(s, element, num_gp, a);
      } else {
        return assemble_j2_f32(s, element, num_gp, a);
      }
#endif
  }
  return fail(FOL_ERR_UNSUPPORTED, "fol_assemble_elements: unsupported physics");
}

}  // namespace fol

using namespace fol;

// ---- plan object (host-buffer entry point) -------------------------------------------------
struct fol_plan {
  int dtype, physics, element, num_gp, nnode, dim, dpn, nd;
  long long ne, nn, ndof;
  size_t esz;
  double params[FOL_NUM_PARAMS];
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t chunk_done[16] = {};
  void *xyz = nullptr, *ctrl = nullptr, *u = nullptr, *ke = nullptr, *re = nullptr, *R = nullptr;
  int32_t *conn = nullptr, *adj_ptr = nullptr, *adj = nullptr, *work = nullptr, *dir_idx = nullptr;
  uint8_t* dir = nullptr;
  // duplicate-free CSR hand-off (fol_plan_set_csr): integer plan of folax_b200/csr_plan.py + the value buffer
  int32_t *pair_ptr = nullptr, *contrib = nullptr, *out_base = nullptr, *row_stride = nullptr;
  void* vals = nullptr;
  long long npairs = 0, nnz = 0;
  std::vector<long long> pair_cut;   // chunk boundaries in pairs, each on a node boundary (contiguous value ranges)
};

extern "C" {

int fol_element_info(int element, int num_gp, int* nnode, int* dim, int* ngauss) {
  FOL_REQUIRE(valid_element(element) && num_gp >= 1 && num_gp <= 3, "fol_element_info: bad element / num_gp");
  if (nnode) *nnode = elem_nnode(element);
  if (dim) *dim = elem_dim(element);
  if (ngauss) *ngauss = elem_ngauss(element, num_gp);
  return FOL_OK;
}

int fol_set_tuned_kernels(int enable) {
  return g_tuned.exchange(enable ? 1 : 0);
}

int fol_set_grid_margin(int ctas) { return g_grid_margin.exchange(ctas < 0 ? 0 : ctas); }

int fol_dofs_per_node(int physics, int element) {
  if (!valid_element(element)) return FOL_ERR_INVALID;
  return (physics == FOL_THERMAL || physics == FOL_TRANSIENT_THERMAL || physics == FOL_ALLEN_CAHN) ? 1
                                                                                                    : elem_dim(element);
}

int fol_assemble_elements(fol_stream_t s, int dtype, int physics, int element, int num_gp, int transpose, int64_t ne,
                          int64_t nn, const void* xyz, const int32_t* conn, const void* ctrl, const void* u,
                          const uint8_t* dir_flag, const double* params_host, void* ke_data, void* re_elem,
                          const void* state_in, void* state_out) {
  (void)nn;
  FOL_REQUIRE(valid_element(element), "fol_assemble_elements: unknown element");
  FOL_REQUIRE(num_gp >= 1 && num_gp <= 3, "fol_assemble_elements: num_gp must be 1, 2 or 3");
  FOL_REQUIRE(xyz && conn && ctrl && u && dir_flag && ke_data && re_elem && params_host,
              "fol_assemble_elements: null pointer");
  FOL_REQUIRE(ne >= 0, "fol_assemble_elements: negative element count");
  if (dtype == FOL_F64)
    return assemble_typed<double>((cudaStream_t)s, physics, element, num_gp, transpose, ne, xyz, conn, ctrl, u,
                                  dir_flag, params_host, ke_data, re_elem, state_in, state_out);
  if (dtype == FOL_F32)
    return assemble_typed<float>((cudaStream_t)s, physics, element, num_gp, transpose, ne, xyz, conn, ctrl, u,
                                 dir_flag, params_host, ke_data, re_elem, state_in, state_out);
  return fail(FOL_ERR_INVALID, "fol_assemble_elements: dtype must be FOL_F32 or FOL_F64");
}

int fol_apply_jacobian_elements(fol_stream_t s, int dtype, int physics, int element, int num_gp, int transpose,
                                int64_t ne, int64_t nn, const void* xyz, const int32_t* conn, const void* ctrl,
                                const void* u, const uint8_t* dir_flag, const double* params_host, const void* v,
                                void* ye_elem, const void* state_in) {
  (void)nn;
  FOL_REQUIRE(valid_element(element), "fol_apply_jacobian_elements: unknown element");
  FOL_REQUIRE(num_gp >= 1 && num_gp <= 3, "fol_apply_jacobian_elements: num_gp must be 1, 2 or 3");
  FOL_REQUIRE(xyz && conn && ctrl && u && dir_flag && v && ye_elem && params_host,
              "fol_apply_jacobian_elements: null pointer");
  FOL_REQUIRE(ne >= 0, "fol_apply_jacobian_elements: negative element count");
  if (dtype == FOL_F64)
    return assemble_typed<double>((cudaStream_t)s, physics, element, num_gp, transpose, ne, xyz, conn, ctrl, u,
                                  dir_flag, params_host, nullptr, ye_elem, state_in, nullptr, v);
  if (dtype == FOL_F32)
    return assemble_typed<float>((cudaStream_t)s, physics, element, num_gp, transpose, ne, xyz, conn, ctrl, u,
                                 dir_flag, params_host, nullptr, ye_elem, state_in, nullptr, v);
  return fail(FOL_ERR_INVALID, "fol_apply_jacobian_elements: dtype must be FOL_F32 or FOL_F64");
}

int fol_apply_jacobian_elements_batched(fol_stream_t s, int dtype, int physics, int element, int num_gp, int transpose,
                                        int64_t ne, int64_t nn, int64_t nb, const void* xyz, const int32_t* conn,
                                        const void* ctrl, const void* u, const uint8_t* dir_flag,
                                        const double* params_host, const void* v, void* ye_elem) {
  FOL_REQUIRE(valid_element(element), "fol_apply_jacobian_elements_batched: unknown element");
  FOL_REQUIRE(num_gp >= 1 && num_gp <= 3, "fol_apply_jacobian_elements_batched: num_gp must be 1, 2 or 3");
  FOL_REQUIRE(xyz && conn && ctrl && u && dir_flag && v && ye_elem && params_host,
              "fol_apply_jacobian_elements_batched: null pointer");
  FOL_REQUIRE(ne >= 0 && nb >= 1 && nb <= 65535 && nn >= 1, "fol_apply_jacobian_elements_batched: bad sizes (1 <= nb <= 65535)");
  FOL_REQUIRE(physics != FOL_J2PLASTICITY, "fol_apply_jacobian_elements_batched: history-dependent elements are not batched");
  if (dtype == FOL_F64)
    return assemble_typed<double>((cudaStream_t)s, physics, element, num_gp, transpose, ne, xyz, conn, ctrl, u,
                                  dir_flag, params_host, nullptr, ye_elem, nullptr, nullptr, v, nb, nn);
  if (dtype == FOL_F32)
    return assemble_typed<float>((cudaStream_t)s, physics, element, num_gp, transpose, ne, xyz, conn, ctrl, u,
                                 dir_flag, params_host, nullptr, ye_elem, nullptr, nullptr, v, nb, nn);
  return fail(FOL_ERR_INVALID, "fol_apply_jacobian_elements_batched: dtype must be FOL_F32 or FOL_F64");
}

int fol_geometry_width(int physics, int element) {
  if (!valid_element(element)) return FOL_ERR_INVALID;
  return geom_width(physics, element);
}

int fol_geometry_cache_physics(fol_stream_t s, int dtype, int physics, int element, int num_gp, int64_t ne,
                               const void* xyz, const int32_t* conn, const void* aux, void* geom) {
  FOL_REQUIRE(valid_element(element) && xyz && conn && geom, "fol_geometry_cache: bad arguments");
  if (dtype == FOL_F64)
    return geometry_cache<double>((cudaStream_t)s, physics, element, num_gp, ne, (const double*)xyz, conn,
                                  (const double*)aux, (double*)geom);
  return geometry_cache<float>((cudaStream_t)s, physics, element, num_gp, ne, (const float*)xyz, conn,
                               (const float*)aux, (float*)geom);
}

int fol_geometry_cache(fol_stream_t s, int dtype, int element, int num_gp, int64_t ne, const void* xyz,
                       const int32_t* conn, void* geom) {
  return fol_geometry_cache_physics(s, dtype, FOL_MECHANICAL, element, num_gp, ne, xyz, conn, nullptr, geom);
}

int64_t fol_energy_work_size(int64_t ntiles, int64_t nb) { return ntiles * nb * 10 + 16; }  // <= 10 warps per tile

int fol_energy_and_grads_flags(fol_stream_t s, int dtype, int physics, int element, int num_gp, int64_t ne, int64_t nn,
                         int64_t nb, const void* geom, const int32_t* conn, const int32_t* adj_ptr,
                         const int32_t* adj_local, const int32_t* tile_node_ptr, const int32_t* tile_nodes,
                         const int32_t* tile_elem_ptr, const int32_t* tile_elems, const int32_t* tile_conn,
                         const int32_t* tile_lnode_ptr, const int32_t* tile_lnodes, int64_t ntiles, int64_t ecap,
                         int64_t lcap, int64_t ncap, const void* ctrl, const void* u, const void* dir_values,
                         const uint8_t* dir_flag, double out_scale, const double* params_host, void* grad_u,
                         void* grad_k,
                         void* energy, void* work, int64_t mesh_flags) {
  FOL_REQUIRE(valid_element(element), "fol_energy_and_grads: unknown element");
  FOL_REQUIRE(geom && conn && adj_ptr && adj_local && tile_node_ptr && tile_nodes && tile_elem_ptr && tile_elems &&
                  tile_conn && tile_lnode_ptr && tile_lnodes &&
                  ctrl && u && grad_u && energy && work && params_host,
              "fol_energy_and_grads: null pointer");
  FOL_REQUIRE(nb >= 1 && ntiles >= 1 && ecap >= 1 && lcap >= 1 && ncap >= 1, "fol_energy_and_grads: bad sizes");
  if (dtype == FOL_F64) {
    EnergyArgs<double> a{(const double*)geom, conn, adj_ptr, adj_local, tile_node_ptr, tile_nodes, tile_elem_ptr,
                         tile_elems, tile_conn, tile_lnode_ptr, tile_lnodes, (const double*)ctrl, (const double*)u,
                         (double*)grad_u, (double*)grad_k, (double*)work, (const double*)dir_values, dir_flag, out_scale,
                         ne, nn, nb, (int)ntiles, (int)ecap, (int)lcap,
                         make_params<double>(params_host), (long long)mesh_flags};
    return energy_and_grads<double>((cudaStream_t)s, physics, element, num_gp, a, (int)ncap, (double*)energy);
  }
  FOL_REQUIRE(dtype == FOL_F32, "fol_energy_and_grads: bad dtype");
  EnergyArgs<float> a{(const float*)geom, conn, adj_ptr, adj_local, tile_node_ptr, tile_nodes, tile_elem_ptr,
                      tile_elems, tile_conn, tile_lnode_ptr, tile_lnodes, (const float*)ctrl, (const float*)u,
                      (float*)grad_u, (float*)grad_k, (float*)work, (const float*)dir_values, dir_flag, (float)out_scale,
                      ne, nn, nb, (int)ntiles, (int)ecap, (int)lcap,
                      make_params<float>(params_host), (long long)mesh_flags};
  return energy_and_grads<float>((cudaStream_t)s, physics, element, num_gp, a, (int)ncap, (float*)energy);
}

int fol_energy_and_grads(fol_stream_t s, int dtype, int physics, int element, int num_gp, int64_t ne, int64_t nn,
                         int64_t nb, const void* geom, const int32_t* conn, const int32_t* adj_ptr,
                         const int32_t* adj_local, const int32_t* tile_node_ptr, const int32_t* tile_nodes,
                         const int32_t* tile_elem_ptr, const int32_t* tile_elems, const int32_t* tile_conn,
                         const int32_t* tile_lnode_ptr, const int32_t* tile_lnodes, int64_t ntiles, int64_t ecap,
                         int64_t lcap, int64_t ncap, const void* ctrl, const void* u, const void* dir_values,
                         const uint8_t* dir_flag, double out_scale, const double* params_host, void* grad_u,
                         void* grad_k, void* energy, void* work) {
  return fol_energy_and_grads_flags(s, dtype, physics, element, num_gp, ne, nn, nb, geom, conn, adj_ptr, adj_local,
                                    tile_node_ptr, tile_nodes, tile_elem_ptr, tile_elems, tile_conn, tile_lnode_ptr,
                                    tile_lnodes, ntiles, ecap, lcap, ncap, ctrl, u, dir_values, dir_flag, out_scale,
                                    params_host, grad_u, grad_k, energy, work, 0);
}

int fol_loss_reduce(fol_stream_t s, int dtype, int64_t nb, double exponent, const void* energy, void* out4,
                    void* scale) {
  FOL_REQUIRE(energy && out4 && scale && nb >= 1, "fol_loss_reduce: bad arguments");
  if (dtype == FOL_F64)
    return loss_reduce<double>((cudaStream_t)s, nb, exponent, (const double*)energy, (double*)out4, (double*)scale);
  return loss_reduce<float>((cudaStream_t)s, nb, exponent, (const float*)energy, (float*)out4, (float*)scale);
}

int fol_scale_grads(fol_stream_t s, int dtype, int64_t nb, int64_t ndof, int64_t nn, const void* scale,
                    double upstream, const void* upstream_dev, int prescaled, const uint8_t* dir_flag, void* grad_u,
                    void* grad_k) {
  FOL_REQUIRE(scale && dir_flag && grad_u, "fol_scale_grads: null pointer");
  if (dtype == FOL_F64)
    return scale_grads<double>((cudaStream_t)s, nb, ndof, nn, (const double*)scale, upstream,
                               (const double*)upstream_dev, prescaled, dir_flag, (double*)grad_u, (double*)grad_k);
  return scale_grads<float>((cudaStream_t)s, nb, ndof, nn, (const float*)scale, upstream, (const float*)upstream_dev,
                            prescaled, dir_flag, (float*)grad_u, (float*)grad_k);
}

// ---- plan ------------------------------------------------------------------------------------
void fol_plan_destroy(fol_plan* p) {
  if (!p) return;
  void* bufs[] = {p->xyz, p->ctrl, p->u, p->ke, p->re, p->R, p->conn, p->adj_ptr, p->adj, p->work, p->dir_idx, p->dir,
                  p->pair_ptr, p->contrib, p->out_base, p->row_stride, p->vals};
  for (void* b : bufs)
    if (b) cudaFree(b);
  for (cudaEvent_t ev : p->chunk_done)
    if (ev) cudaEventDestroy(ev);
  if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
}

#define FOL_PLAN_CUDA(call)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      fol_plan_destroy(p);                                                                   \
      return ::fol::fail(FOL_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));  \
    }                                                                                        \
  } while (0)

int fol_plan_create(fol_plan** plan, int dtype, int physics, int element, int num_gp, int64_t ne, int64_t nn,
                    const void* xyz_host, const int32_t* conn_host, const int32_t* dir_idx_host, int64_t n_dir,
                    const double* params_host) {
  FOL_REQUIRE(plan && xyz_host && conn_host && params_host, "fol_plan_create: null pointer");
  FOL_REQUIRE(valid_element(element) && num_gp >= 1 && num_gp <= 3, "fol_plan_create: bad element / num_gp");
  FOL_REQUIRE(dtype == FOL_F32 || dtype == FOL_F64, "fol_plan_create: bad dtype");
  FOL_REQUIRE(physics == FOL_MECHANICAL || physics == FOL_THERMAL || physics == FOL_NEOHOOKE ||
                  physics == FOL_STVENANT,
              "fol_plan_create: physics without history only");
  fol_plan* p = new fol_plan();
  p->dtype = dtype; p->physics = physics; p->element = element; p->num_gp = num_gp;
  p->nnode = elem_nnode(element); p->dim = elem_dim(element);
  p->dpn = physics == FOL_THERMAL ? 1 : p->dim;
  p->nd = p->nnode * p->dpn;
  p->ne = ne; p->nn = nn; p->ndof = nn * p->dpn;
  p->esz = dtype == FOL_F64 ? 8 : 4;
  for (int i = 0; i < FOL_NUM_PARAMS; ++i) p->params[i] = params_host[i];
  FOL_PLAN_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  FOL_PLAN_CUDA(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
  for (cudaEvent_t& ev : p->chunk_done) FOL_PLAN_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  FOL_PLAN_CUDA(cudaMalloc(&p->xyz, p->esz * 3 * nn));
  FOL_PLAN_CUDA(cudaMalloc(&p->ctrl, p->esz * nn));
  FOL_PLAN_CUDA(cudaMalloc(&p->u, p->esz * p->ndof));
  FOL_PLAN_CUDA(cudaMalloc(&p->R, p->esz * p->ndof));
  FOL_PLAN_CUDA(cudaMalloc(&p->ke, p->esz * (size_t)ne * p->nd * p->nd));
  FOL_PLAN_CUDA(cudaMalloc(&p->re, p->esz * (size_t)ne * p->nd));
  FOL_PLAN_CUDA(cudaMalloc((void**)&p->conn, sizeof(int32_t) * (size_t)ne * p->nnode));
  FOL_PLAN_CUDA(cudaMalloc((void**)&p->adj_ptr, sizeof(int32_t) * (size_t)(nn + 1)));
  FOL_PLAN_CUDA(cudaMalloc((void**)&p->adj, sizeof(int32_t) * (size_t)ne * p->nnode));
  FOL_PLAN_CUDA(cudaMalloc((void**)&p->work, sizeof(int32_t) * (size_t)nn));
  FOL_PLAN_CUDA(cudaMalloc((void**)&p->dir, (size_t)p->ndof));
  FOL_PLAN_CUDA(cudaMemcpyAsync(p->xyz, xyz_host, p->esz * 3 * nn, cudaMemcpyHostToDevice, p->stream));
  FOL_PLAN_CUDA(cudaMemcpyAsync(p->conn, conn_host, sizeof(int32_t) * (size_t)ne * p->nnode, cudaMemcpyHostToDevice,
                                p->stream));
  if (n_dir > 0) {
    FOL_PLAN_CUDA(cudaMalloc((void**)&p->dir_idx, sizeof(int32_t) * (size_t)n_dir));
    FOL_PLAN_CUDA(cudaMemcpyAsync(p->dir_idx, dir_idx_host, sizeof(int32_t) * (size_t)n_dir, cudaMemcpyHostToDevice,
                                  p->stream));
  }
  int rc = fol_dirichlet_flags(p->stream, p->dir_idx, n_dir, p->ndof, p->dir);
  if (!rc) rc = fol_node_adjacency(p->stream, p->conn, ne, p->nnode, nn, p->adj_ptr, p->adj, p->work);
  if (rc) {
    fol_plan_destroy(p);
    return rc;
  }
  FOL_PLAN_CUDA(cudaStreamSynchronize(p->stream));
  *plan = p;
  return FOL_OK;
}

fol_stream_t fol_plan_stream(fol_plan* p) { return p ? (fol_stream_t)p->stream : nullptr; }

static int plan_run(fol_plan* p, int transpose, const void* ctrl, const void* u) {
  int rc = fol_assemble_elements(p->stream, p->dtype, p->physics, p->element, p->num_gp, transpose, p->ne, p->nn,
                                 p->xyz, p->conn, ctrl, u, p->dir, p->params, p->ke, p->re, nullptr, nullptr);
  if (rc) return rc;
  return fol_residual_gather(p->stream, p->dtype, p->nn, p->nnode, p->dpn, p->adj_ptr, p->adj, p->re, p->R);
}

int fol_plan_assemble_device(fol_plan* p, int transpose, const void* ctrl_dev, const void* u_dev, void** ke_dev,
                             void** R_dev) {
  FOL_REQUIRE(p && ctrl_dev && u_dev, "fol_plan_assemble_device: null pointer");
  if (int rc = plan_run(p, transpose, ctrl_dev, u_dev)) return rc;
  if (ke_dev) *ke_dev = p->ke;
  if (R_dev) *R_dev = p->R;
  return FOL_OK;
}

int fol_plan_assemble_host(fol_plan* p, int transpose, const void* ctrl_host, const void* u_host, void* ke_host,
                           void* R_host) {
  FOL_REQUIRE(p && ctrl_host && u_host && ke_host && R_host, "fol_plan_assemble_host: null pointer");
  FOL_CUDA(cudaMemcpyAsync(p->ctrl, ctrl_host, p->esz * p->nn, cudaMemcpyHostToDevice, p->stream));
  FOL_CUDA(cudaMemcpyAsync(p->u, u_host, p->esz * p->ndof, cudaMemcpyHostToDevice, p->stream));
  // The BCOO data (nd^2 values per element) dominates the transfer: the element stage runs in chunks and each
  // chunk's matrices start their way to the host as soon as they exist, so the link never waits for the kernels.
  constexpr int kChunks = 16;
  const long long per = (p->ne + kChunks - 1) / kChunks;
  const size_t ke_elem = p->esz * (size_t)p->nd * p->nd, re_elem = p->esz * (size_t)p->nd;
  for (int c = 0; c < kChunks; ++c) {
    const long long e0 = c * per, cnt = (e0 + per <= p->ne ? per : p->ne - e0);
    if (cnt <= 0) break;
    int rc = fol_assemble_elements(p->stream, p->dtype, p->physics, p->element, p->num_gp, transpose, cnt, p->nn, p->xyz,
                                   p->conn + e0 * p->nnode, p->ctrl, p->u, p->dir, p->params,
                                   (char*)p->ke + e0 * ke_elem, (char*)p->re + e0 * re_elem, nullptr, nullptr);
    if (rc) return rc;
    FOL_CUDA(cudaEventRecord(p->chunk_done[c], p->stream));
    FOL_CUDA(cudaStreamWaitEvent(p->copy_stream, p->chunk_done[c], 0));
    FOL_CUDA(cudaMemcpyAsync((char*)ke_host + e0 * ke_elem, (char*)p->ke + e0 * ke_elem, cnt * ke_elem,
                             cudaMemcpyDeviceToHost, p->copy_stream));
  }
  if (int rc = fol_residual_gather(p->stream, p->dtype, p->nn, p->nnode, p->dpn, p->adj_ptr, p->adj, p->re, p->R)) return rc;
  FOL_CUDA(cudaMemcpyAsync(R_host, p->R, p->esz * p->ndof, cudaMemcpyDeviceToHost, p->stream));
  FOL_CUDA(cudaStreamSynchronize(p->stream));
  FOL_CUDA(cudaStreamSynchronize(p->copy_stream));
  return FOL_OK;
}

/* Uploads the integer plan of the duplicate-free CSR (host arrays of folax_b200/csr_plan.py: pairs sorted by (row
 * node, column node), contributors per pair in ascending (element, a, b) order) once per mesh. */
int fol_plan_set_csr(fol_plan* p, int64_t npairs, int64_t nnz, const int32_t* pair_ptr_host, const int32_t* contrib_host,
                     const int32_t* out_base_host, const int32_t* row_stride_host) {
  FOL_REQUIRE(p && pair_ptr_host && contrib_host && out_base_host && row_stride_host && npairs > 0 && nnz > 0,
              "fol_plan_set_csr: bad arguments");
  const size_t ncontrib = (size_t)p->ne * p->nnode * p->nnode;
  FOL_REQUIRE((size_t)pair_ptr_host[npairs] == ncontrib, "fol_plan_set_csr: the plan does not belong to this mesh");
  void** slots[] = {(void**)&p->pair_ptr, (void**)&p->contrib, (void**)&p->out_base, (void**)&p->row_stride, &p->vals};
  for (void** sl : slots) {
    if (*sl) cudaFree(*sl);
    *sl = nullptr;
  }
  FOL_CUDA(cudaMalloc((void**)&p->pair_ptr, sizeof(int32_t) * (size_t)(npairs + 1)));
  FOL_CUDA(cudaMalloc((void**)&p->contrib, sizeof(int32_t) * ncontrib));
  FOL_CUDA(cudaMalloc((void**)&p->out_base, sizeof(int32_t) * (size_t)npairs));
  FOL_CUDA(cudaMalloc((void**)&p->row_stride, sizeof(int32_t) * (size_t)npairs));
  FOL_CUDA(cudaMalloc(&p->vals, p->esz * (size_t)nnz));
  FOL_CUDA(cudaMemcpyAsync(p->pair_ptr, pair_ptr_host, sizeof(int32_t) * (size_t)(npairs + 1), cudaMemcpyHostToDevice, p->stream));
  FOL_CUDA(cudaMemcpyAsync(p->contrib, contrib_host, sizeof(int32_t) * ncontrib, cudaMemcpyHostToDevice, p->stream));
  FOL_CUDA(cudaMemcpyAsync(p->out_base, out_base_host, sizeof(int32_t) * (size_t)npairs, cudaMemcpyHostToDevice, p->stream));
  FOL_CUDA(cudaMemcpyAsync(p->row_stride, row_stride_host, sizeof(int32_t) * (size_t)npairs, cudaMemcpyHostToDevice, p->stream));
  p->npairs = npairs;
  p->nnz = nnz;
  // chunk boundaries: a pair P opens the rows of a node exactly where out_base[P] == d*d*P (its index within the
  // node's pairs is 0), and the values of whole nodes are one contiguous range of the CSR value array
  constexpr int kChunks = 16;
  const long long dd = (long long)p->dpn * p->dpn;
  p->pair_cut.assign(1, 0);
  for (int c = 1; c < kChunks; ++c) {
    long long P = npairs * c / kChunks;
    while (P < npairs && (long long)out_base_host[P] != dd * P) ++P;
    if (P > p->pair_cut.back() && P < npairs) p->pair_cut.push_back(P);
  }
  p->pair_cut.push_back(npairs);
  FOL_CUDA(cudaStreamSynchronize(p->stream));
  return FOL_OK;
}

/* Host hand-off of what the reference's solvers actually consume (fe_solver.py:71-72: the duplicates of the BCOO
 * summed into a CSR): H2D of the inputs, element stage, de-duplication on the device in chunks of whole node rows,
 * each chunk's values starting their way to the host as soon as they exist.  vals_host: nnz values in the order of
 * the plan's (indptr, indices); 1/2.4 of the bytes of the duplicate-keeping hand-off for Hex8. */
int fol_plan_assemble_host_csr(fol_plan* p, int transpose, const void* ctrl_host, const void* u_host, void* vals_host,
                               void* R_host) {
  FOL_REQUIRE(p && ctrl_host && u_host && vals_host && R_host, "fol_plan_assemble_host_csr: null pointer");
  FOL_REQUIRE(p->vals && p->npairs > 0, "fol_plan_assemble_host_csr: call fol_plan_set_csr first");
  FOL_CUDA(cudaMemcpyAsync(p->ctrl, ctrl_host, p->esz * p->nn, cudaMemcpyHostToDevice, p->stream));
  FOL_CUDA(cudaMemcpyAsync(p->u, u_host, p->esz * p->ndof, cudaMemcpyHostToDevice, p->stream));
  if (int rc = plan_run(p, transpose, p->ctrl, p->u)) return rc;
  FOL_CUDA(cudaMemcpyAsync(R_host, p->R, p->esz * p->ndof, cudaMemcpyDeviceToHost, p->stream));
  const long long dd = (long long)p->dpn * p->dpn;
  for (size_t c = 0; c + 1 < p->pair_cut.size(); ++c) {
    const long long p0 = p->pair_cut[c], p1 = p->pair_cut[c + 1];
    // the kernel indexes pairs from 0: pass the sub-range through offset pointers; out_base stays absolute
    int rc = fol_csr_values(p->stream, p->dtype, p1 - p0, p->dpn, p->nnode, p->pair_ptr + p0, p->contrib,
                            p->out_base + p0, p->row_stride + p0, p->ke, p->vals);
    if (rc) return rc;
    FOL_CUDA(cudaEventRecord(p->chunk_done[c], p->stream));
    FOL_CUDA(cudaStreamWaitEvent(p->copy_stream, p->chunk_done[c], 0));
    const size_t v0 = (size_t)(dd * p0) * p->esz, v1 = (p1 == p->npairs ? (size_t)p->nnz : (size_t)(dd * p1)) * p->esz;
    FOL_CUDA(cudaMemcpyAsync((char*)vals_host + v0, (char*)p->vals + v0, v1 - v0, cudaMemcpyDeviceToHost, p->copy_stream));
  }
  FOL_CUDA(cudaStreamSynchronize(p->stream));
  FOL_CUDA(cudaStreamSynchronize(p->copy_stream));
  return FOL_OK;
}

}  // extern "C"

// ---- FMA peak microbenchmark ------------------------------------------------------------------
namespace fol {
template <class T>
__global__ void fma_peak_kernel(T* out, int iters) {
  T a0 = (T)threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const T x = (T)1.0000001, y = (T)1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = a0 * x + y; a1 = a1 * x + y; a2 = a2 * x + y; a3 = a3 * x + y;
    a4 = a4 * x + y; a5 = a5 * x + y; a6 = a6 * x + y; a7 = a7 * x + y;
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <class T>
static int measure_peak(double* tflops) {
  const int blocks = 148 * 8, threads = 256, iters = 1 << 14;
  T* out = nullptr;
  FOL_CUDA(cudaMalloc(&out, sizeof(T) * blocks * threads));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  fma_peak_kernel<T><<<blocks, threads>>>(out, iters);  // warm-up
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    fma_peak_kernel<T><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  g_launches.fetch_add(6);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  FOL_CUDA(cudaGetLastError());
  *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
  return FOL_OK;
}
}  // namespace fol

namespace fol {
// write-only streaming kernel: the ceiling of a kernel whose traffic is a pure store stream
__global__ void write_stream_kernel(double2* __restrict__ out, long long n2, double v) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride)
    __stcs(out + i, make_double2(v, v));
}
}  // namespace fol

// Measured HBM bandwidth of a pure WRITE stream of `bytes` (best of 5), GB/s.
extern "C" int fol_measure_write_bandwidth(int64_t bytes, double* gbs) {
  FOL_REQUIRE(gbs && bytes >= (1 << 20), "fol_measure_write_bandwidth: bad arguments");
  double2* buf = nullptr;
  FOL_CUDA(cudaMalloc(&buf, (size_t)bytes));
  const long long n2 = bytes / 16;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  write_stream_kernel<<<148 * 16, 256>>>(buf, n2, 1.0);
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    write_stream_kernel<<<148 * 16, 256>>>(buf, n2, (double)r);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  g_launches.fetch_add(6);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  FOL_CUDA(cudaGetLastError());
  *gbs = (double)(n2 * 16) / (best * 1e-3) / 1e9;
  return FOL_OK;
}

extern "C" int fol_measure_fma_peak(int dtype, double* tflops) {
  FOL_REQUIRE(tflops, "fol_measure_fma_peak: null pointer");
  return dtype == FOL_F64 ? measure_peak<double>(tflops) : measure_peak<float>(tflops);
}
