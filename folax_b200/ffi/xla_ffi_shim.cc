// XLA typed-FFI handlers over the C ABI of libfolax_b200 -- the jax.ffi side of the drop-in boundary.
//
// Binding style mirrors the reference's only FFI precedent,
//   fol/loss_functions/ffi_functions/kr_small_displacement_element.cc:293-337
// (Ffi::Bind().Ctx<PlatformStream<cudaStream_t>>().Arg<AnyBuffer>()...Ret<AnyBuffer>(), F32/F64
// dispatch on the buffer element type :49-57, S32 connectivity :102, 320), but instead of copying
// to the host and looping over CPU elements the handlers enqueue the sm_100a kernels on XLA's
// stream and return immediately (no synchronisation, no retained state).
//
// JAX / jaxlib (and therefore xla/ffi/api/ffi.h) are NOT available in this build image, so this
// translation unit is compiled only where the header exists (see INTEGRATION.md for the build line
// and the Python registration stub).  It is untested under XLA here and says so; tests/test_cabi.py
// compiles it against a stand-in header (tests/xla_stub) to keep it valid C++ that matches the C ABI and whose handler
// signatures match their Ffi::Bind() chains.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define FOLAX_HAVE_XLA_FFI 1
#endif
#endif

#ifdef FOLAX_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include <string>

#include "../../include/folax_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error FromRc(int rc) {
  if (rc == FOL_OK) return ffi::Error::Success();
  if (rc == FOL_ERR_INVALID) return ffi::Error::InvalidArgument(fol_last_error());
  return ffi::Error::Internal(fol_last_error());
}

int DtypeOf(ffi::DataType t) { return t == ffi::F64 ? FOL_F64 : (t == ffi::F32 ? FOL_F32 : -1); }

// compute_elements: (coords (nn,3), conn (ne,a) S32, ctrl (nn), u (ndof), dir_flag (ndof) U8, params (12) F64 host
// attribute) -> ke_data (ne*nd*nd), re_elem (ne*nd).   Replaces compute_elements of ...cc:226-291.
ffi::Error AssembleElementsImpl(cudaStream_t stream, ffi::AnyBuffer xyz, ffi::Buffer<ffi::S32> conn,
                                ffi::AnyBuffer ctrl, ffi::AnyBuffer u, ffi::Buffer<ffi::U8> dir_flag,
                                ffi::Result<ffi::AnyBuffer> ke, ffi::Result<ffi::AnyBuffer> re, int32_t physics,
                                int32_t element, int32_t num_gp, int32_t transpose,
                                ffi::Span<const double> params) {
  const int dt = DtypeOf(xyz.element_type());
  if (dt < 0) return ffi::Error::InvalidArgument("Unsupported data type for FFI call.");
  if (params.size() != FOL_NUM_PARAMS) return ffi::Error::InvalidArgument("params must have 12 entries");
  const auto cd = conn.dimensions();
  const int64_t ne = cd[0], nn = xyz.dimensions()[0];
  return FromRc(fol_assemble_elements(stream, dt, physics, element, num_gp, transpose, ne, nn, xyz.untyped_data(),
                                      conn.typed_data(), ctrl.untyped_data(), u.untyped_data(),
                                      dir_flag.typed_data(), params.begin(), ke->untyped_data(), re->untyped_data(),
                                      nullptr, nullptr));
}

// residual_gather: (adj_ptr, adj, re_elem) -> residual (ndof).  Replaces the scatter-add of fe_loss.py:301-306.
ffi::Error ResidualGatherImpl(cudaStream_t stream, ffi::Buffer<ffi::S32> adj_ptr, ffi::Buffer<ffi::S32> adj,
                              ffi::AnyBuffer re, ffi::Result<ffi::AnyBuffer> residual, int32_t nnode,
                              int32_t dofs_per_node) {
  const int dt = DtypeOf(re.element_type());
  if (dt < 0) return ffi::Error::InvalidArgument("Unsupported data type for FFI call.");
  const int64_t nn = adj_ptr.dimensions()[0] - 1;
  return FromRc(fol_residual_gather(stream, dt, nn, nnode, dofs_per_node, adj_ptr.typed_data(), adj.typed_data(),
                                    re.untyped_data(), residual->untyped_data()));
}

// energy_and_grads: the forward of the custom_vjp'd batched loss (fe_loss.py:250-262).
ffi::Error EnergyAndGradsImpl(cudaStream_t stream, ffi::AnyBuffer geom, ffi::Buffer<ffi::S32> conn,
                              ffi::Buffer<ffi::S32> adj_ptr, ffi::Buffer<ffi::S32> adj_local,
                              ffi::Buffer<ffi::S32> tile_node_ptr, ffi::Buffer<ffi::S32> tile_nodes,
                              ffi::Buffer<ffi::S32> tile_elem_ptr, ffi::Buffer<ffi::S32> tile_elems,
                              ffi::Buffer<ffi::S32> tile_conn, ffi::Buffer<ffi::S32> tile_lnode_ptr,
                              ffi::Buffer<ffi::S32> tile_lnodes,
                              ffi::AnyBuffer ctrl, ffi::AnyBuffer u, ffi::Result<ffi::AnyBuffer> grad_u,
                              ffi::Result<ffi::AnyBuffer> grad_k, ffi::Result<ffi::AnyBuffer> energy,
                              ffi::Result<ffi::AnyBuffer> work, int32_t physics, int32_t element, int32_t num_gp,
                              int64_t ecap, int64_t lcap, int64_t ncap, ffi::Span<const double> params) {
  const int dt = DtypeOf(u.element_type());
  if (dt < 0) return ffi::Error::InvalidArgument("Unsupported data type for FFI call.");
  const int64_t ne = conn.dimensions()[0], nb = u.dimensions()[0], nn = ctrl.dimensions()[1];
  const int64_t ntiles = tile_node_ptr.dimensions()[0] - 1;
  return FromRc(fol_energy_and_grads(stream, dt, physics, element, num_gp, ne, nn, nb, geom.untyped_data(),
                                     conn.typed_data(), adj_ptr.typed_data(), adj_local.typed_data(),
                                     tile_node_ptr.typed_data(), tile_nodes.typed_data(), tile_elem_ptr.typed_data(),
                                     tile_elems.typed_data(), tile_conn.typed_data(), tile_lnode_ptr.typed_data(),
                                     tile_lnodes.typed_data(), ntiles, ecap, lcap, ncap, ctrl.untyped_data(), u.untyped_data(),
                                     /*dir_values=*/nullptr, /*dir_flag=*/nullptr, /*out_scale=*/1.0, params.begin(), grad_u->untyped_data(), grad_k->untyped_data(),
                                     energy->untyped_data(), work->untyped_data()));
}

// geometry_cache: per-element, per-Gauss-point geometry factors shared by all samples of the batched loss.
ffi::Error GeometryCacheImpl(cudaStream_t stream, ffi::AnyBuffer xyz, ffi::Buffer<ffi::S32> conn,
                             ffi::Result<ffi::AnyBuffer> geom, int32_t physics, int32_t element, int32_t num_gp) {
  const int dt = DtypeOf(xyz.element_type());
  if (dt < 0) return ffi::Error::InvalidArgument("Unsupported data type for FFI call.");
  return FromRc(fol_geometry_cache_physics(stream, dt, physics, element, num_gp, conn.dimensions()[0],
                                           xyz.untyped_data(), conn.typed_data(), /*aux=*/nullptr,
                                           geom->untyped_data()));
}

// apply_jacobian_elements: ye_elem = Ke'(ctrl, u) v_e without forming Ke (what `BCOO @ v` does, fe_solver.py:61).
ffi::Error ApplyJacobianElementsImpl(cudaStream_t stream, ffi::AnyBuffer xyz, ffi::Buffer<ffi::S32> conn,
                                     ffi::AnyBuffer ctrl, ffi::AnyBuffer u, ffi::Buffer<ffi::U8> dir_flag,
                                     ffi::AnyBuffer v, ffi::Result<ffi::AnyBuffer> ye, int32_t physics, int32_t element,
                                     int32_t num_gp, int32_t transpose, ffi::Span<const double> params) {
  const int dt = DtypeOf(xyz.element_type());
  if (dt < 0) return ffi::Error::InvalidArgument("Unsupported data type for FFI call.");
  if (params.size() != FOL_NUM_PARAMS) return ffi::Error::InvalidArgument("params must have 12 entries");
  return FromRc(fol_apply_jacobian_elements(stream, dt, physics, element, num_gp, transpose, conn.dimensions()[0],
                                            xyz.dimensions()[0], xyz.untyped_data(), conn.typed_data(),
                                            ctrl.untyped_data(), u.untyped_data(), dir_flag.typed_data(),
                                            params.begin(), v.untyped_data(), ye->untyped_data(), nullptr));
}

// the three kernels behind FiniteElementResponse (fe_response.py:91-524); the formula and its jax.grad stay in Python
ffi::Error GaussInterpolateImpl(cudaStream_t stream, ffi::Buffer<ffi::S32> conn, ffi::AnyBuffer ctrl, ffi::AnyBuffer u,
                                ffi::Result<ffi::AnyBuffer> k_gp, ffi::Result<ffi::AnyBuffer> u_gp, int32_t element,
                                int32_t num_gp, int32_t dofs_per_node) {
  const int dt = DtypeOf(u.element_type());
  if (dt < 0) return ffi::Error::InvalidArgument("Unsupported data type for FFI call.");
  return FromRc(fol_gauss_interpolate(stream, dt, element, num_gp, dofs_per_node, conn.dimensions()[0],
                                      conn.typed_data(), ctrl.untyped_data(), u.untyped_data(), k_gp->untyped_data(),
                                      u_gp->untyped_data()));
}

ffi::Error ResponseElementsImpl(cudaStream_t stream, ffi::AnyBuffer xyz, ffi::Buffer<ffi::S32> conn, ffi::AnyBuffer f_gp,
                                ffi::AnyBuffer fk_gp, ffi::AnyBuffer fu_gp, ffi::Result<ffi::AnyBuffer> value_elem,
                                ffi::Result<ffi::AnyBuffer> du_elem, ffi::Result<ffi::AnyBuffer> dk_elem,
                                ffi::Result<ffi::AnyBuffer> dx_elem, int32_t element, int32_t num_gp,
                                int32_t dofs_per_node) {
  const int dt = DtypeOf(xyz.element_type());
  if (dt < 0) return ffi::Error::InvalidArgument("Unsupported data type for FFI call.");
  return FromRc(fol_response_elements(stream, dt, element, num_gp, dofs_per_node, conn.dimensions()[0],
                                      xyz.untyped_data(), conn.typed_data(), f_gp.untyped_data(), fk_gp.untyped_data(),
                                      fu_gp.untyped_data(), value_elem->untyped_data(), du_elem->untyped_data(),
                                      dk_elem->untyped_data(), dx_elem->untyped_data()));
}

ffi::Error ResidualAdjointElementsImpl(cudaStream_t stream, ffi::AnyBuffer xyz, ffi::Buffer<ffi::S32> conn,
                                       ffi::AnyBuffer ctrl, ffi::AnyBuffer u, ffi::AnyBuffer adj,
                                       ffi::Result<ffi::AnyBuffer> dk_elem, ffi::Result<ffi::AnyBuffer> dx_elem,
                                       int32_t physics, int32_t element, int32_t num_gp,
                                       ffi::Span<const double> params) {
  const int dt = DtypeOf(xyz.element_type());
  if (dt < 0) return ffi::Error::InvalidArgument("Unsupported data type for FFI call.");
  if (params.size() != FOL_NUM_PARAMS) return ffi::Error::InvalidArgument("params must have 12 entries");
  return FromRc(fol_residual_adjoint_elements(stream, dt, physics, element, num_gp, /*accumulate=*/0,
                                              conn.dimensions()[0], xyz.untyped_data(), conn.typed_data(),
                                              ctrl.untyped_data(), u.untyped_data(), adj.untyped_data(),
                                              /*aux=*/nullptr, params.begin(), dk_elem->untyped_data(),
                                              dx_elem->untyped_data()));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(FolGeometryCache, GeometryCacheImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int32_t>("physics")
                                  .Attr<int32_t>("element")
                                  .Attr<int32_t>("num_gp"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(FolApplyJacobianElements, ApplyJacobianElementsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()          // xyz
                                  .Arg<ffi::Buffer<ffi::S32>>()   // conn
                                  .Arg<ffi::AnyBuffer>()          // ctrl
                                  .Arg<ffi::AnyBuffer>()          // u
                                  .Arg<ffi::Buffer<ffi::U8>>()    // dir_flag
                                  .Arg<ffi::AnyBuffer>()          // v
                                  .Ret<ffi::AnyBuffer>()          // ye_elem
                                  .Attr<int32_t>("physics")
                                  .Attr<int32_t>("element")
                                  .Attr<int32_t>("num_gp")
                                  .Attr<int32_t>("transpose")
                                  .Attr<ffi::Span<const double>>("params"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(FolGaussInterpolate, GaussInterpolateImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int32_t>("element")
                                  .Attr<int32_t>("num_gp")
                                  .Attr<int32_t>("dofs_per_node"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(FolResponseElements, ResponseElementsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int32_t>("element")
                                  .Attr<int32_t>("num_gp")
                                  .Attr<int32_t>("dofs_per_node"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(FolResidualAdjointElements, ResidualAdjointElementsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int32_t>("physics")
                                  .Attr<int32_t>("element")
                                  .Attr<int32_t>("num_gp")
                                  .Attr<ffi::Span<const double>>("params"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(FolAssembleElements, AssembleElementsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()          // xyz
                                  .Arg<ffi::Buffer<ffi::S32>>()   // conn
                                  .Arg<ffi::AnyBuffer>()          // ctrl
                                  .Arg<ffi::AnyBuffer>()          // u
                                  .Arg<ffi::Buffer<ffi::U8>>()    // dir_flag
                                  .Ret<ffi::AnyBuffer>()          // ke_data
                                  .Ret<ffi::AnyBuffer>()          // re_elem
                                  .Attr<int32_t>("physics")
                                  .Attr<int32_t>("element")
                                  .Attr<int32_t>("num_gp")
                                  .Attr<int32_t>("transpose")
                                  .Attr<ffi::Span<const double>>("params"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(FolResidualGather, ResidualGatherImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int32_t>("nnode")
                                  .Attr<int32_t>("dofs_per_node"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(FolEnergyAndGrads, EnergyAndGradsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int32_t>("physics")
                                  .Attr<int32_t>("element")
                                  .Attr<int32_t>("num_gp")
                                  .Attr<int64_t>("ecap")
                                  .Attr<int64_t>("lcap")
                                  .Attr<int64_t>("ncap")
                                  .Attr<ffi::Span<const double>>("params"));
#endif  // FOLAX_HAVE_XLA_FFI
