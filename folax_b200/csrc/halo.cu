// Halo-DOF exchange of the slab-partitioned assembly over NVLink peer memory (north_star:
// "single huge meshes shard by element partition with a halo-DOF exchange").
//
// The residual gather of an interface node plane and the transfer to the neighbour are ONE kernel: every value of
// the plane is summed in the fixed adjacency order, written to the local residual and stored straight into the
// neighbour's receive buffer (a peer pointer opened from its CUDA IPC handle); the last thread of each CTA then
// bumps an arrival counter in the neighbour's memory (system-scope release).  After its interior element stage the
// neighbour runs halo_add: it waits (system-scope acquire) until all CTAs of the step have arrived -- by then long
// true -- and adds the received partial sums to its copy of the plane.  No NCCL kernel competes with the persistent
// element-stage kernel for SMs, no side stream, no pack / unpack.  Receive buffers are double-buffered by step
// parity: a rank can only produce step s+2 after consuming s+1, which its neighbour produced after consuming s.
#include <cuda_runtime.h>
#include <string.h>

#include "assemble.cuh"
#include "assemble_hex_common.cuh"
#include "common.cuh"

namespace fol {

int assemble_hex_mech_f64(cudaStream_t, const AsmArgs<double>&, const hexk::HaloFuse*);
int assemble_hex_j2_f64(cudaStream_t, const AsmArgs<double>&, const hexk::HaloFuse*);

template <class T>
__global__ void halo_gather_push_kernel(long long n0, long long count, int d, const int32_t* __restrict__ ptr,
                                        const int32_t* __restrict__ adj, const T* __restrict__ re,
                                        T* __restrict__ R, T* __restrict__ peer_recv,
                                        unsigned long long* __restrict__ peer_arrive) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count * d) {
    const long long n = n0 + t / d;
    const int k = (int)(t % d);
    T acc = (T)0;
    const int lo = ptr[n], hi = ptr[n + 1];
    for (int i = lo; i < hi; ++i) acc += __ldg(re + (long long)adj[i] * d + k);
    R[n0 * d + t] = acc;
    if (peer_recv) peer_recv[t] = acc;          // NVLink peer store
  }
  if (peer_arrive) {
    __threadfence_system();                      // this thread's peer stores are visible system-wide ...
    __syncthreads();                             // ... for every thread of the CTA, before the CTA reports in
    if (threadIdx.x == 0) {
      // the reporting thread fences AFTER the barrier too: under the PTX memory model only a fence that follows
      // the barrier (which ordered the other threads' stores before it) is cumulative over the whole CTA's stores
      __threadfence_system();
      atomicAdd_system(peer_arrive, 1ULL);
    }
  }
}

template <class T>
__global__ void halo_add_kernel(long long n0, long long count, int d, const T* __restrict__ recv,
                                const unsigned long long* __restrict__ arrive, unsigned long long target,
                                unsigned long long* __restrict__ timeouts, T* __restrict__ R) {
  if (threadIdx.x == 0) {
    unsigned long long seen;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(arrive) : "memory");
      // a neighbour that died must not hang this GPU: give up after ~2 s and count it (fol_halo_timeouts)
      if (seen < target && clock64() - t0 > 4000000000LL) {
        atomicAdd(timeouts, 1ULL);
        break;
      }
    } while (seen < target);
  }
  __syncthreads();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count * d) R[n0 * d + t] += __ldcv(recv + t);
}


// Completes a step of the fused path (assemble_hex_common.cuh): blocks [0, nb_int) gather the interior nodes in the
// fixed adjacency order; the following 2 * nb_plane blocks wait (system-scope acquire) for the neighbour's chunk
// arrivals -- long there by then -- and add what it pushed to the plane sums the element-stage launch left in R.
// IEEE addition commutes, so both copies of an interface plane end up bit-identical.
template <class T>
__global__ void gather_halo_kernel(long long plane, long long nn, int d, const int32_t* __restrict__ ptr,
                                   const int32_t* __restrict__ adj, const T* __restrict__ re, T* __restrict__ R,
                                   long long nb_int, long long nb_plane, const T* recv_lo, const T* recv_hi,
                                   const unsigned long long* arrive_lo, const unsigned long long* arrive_hi,
                                   unsigned long long target, unsigned long long* timeouts,
                                   unsigned long long* iface_done, unsigned long long* chunk_next) {
  const long long b = blockIdx.x;
  if (b == 0 && threadIdx.x == 0) {      // the element-stage launch of this step is complete (stream order)
    *iface_done = 0ULL;
    *chunk_next = 0ULL;
  }
  if (b < nb_int) {
    const long long t = b * blockDim.x + threadIdx.x;
    if (t < (nn - 2 * plane) * d) {
      const long long n = plane + t / d;
      const int k = (int)(t % d);
      T acc = (T)0;
      const int lo = ptr[n], hi = ptr[n + 1];
      for (int i = lo; i < hi; ++i) acc += __ldg(re + (long long)adj[i] * d + k);
      R[n * d + k] = acc;
    }
    return;
  }
  const int side = (b - nb_int) >= nb_plane ? 1 : 0;
  const T* recv = side ? recv_hi : recv_lo;
  if (!recv) return;
  const unsigned long long* arrive = side ? arrive_hi : arrive_lo;
  if (threadIdx.x == 0) {
    unsigned long long seen;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(arrive) : "memory");
      if (seen < target && clock64() - t0 > 4000000000LL) {
        atomicAdd(timeouts, 1ULL);
        break;
      }
    } while (seen < target);
  }
  __syncthreads();
  const long long t = (b - nb_int - side * nb_plane) * blockDim.x + threadIdx.x;
  const long long n0 = side ? nn - plane : 0;
  if (t < plane * d) R[n0 * d + t] += __ldcv(recv + t);
}

}  // namespace fol

using namespace fol;

// one per rank: receive buffers + arrival counters in ONE cudaMalloc allocation (so one IPC handle exports it)
struct fol_halo {
  int dtype = 0;
  size_t esz = 8;
  long long plane_dofs = 0;
  unsigned ctas = 0;                 // CTAs of one push launch = arrivals per step
  unsigned char* base = nullptr;     // [from-lower: 2 x plane][from-upper: 2 x plane][counters: 4 x u64]
  unsigned char* peer[2] = {nullptr, nullptr};   // opened allocations of the lower / upper neighbour
  size_t recv_off(int side, int parity) const { return ((size_t)side * 2 + parity) * plane_dofs * esz; }
  size_t arrive_off(int side, int parity) const { return (size_t)4 * plane_dofs * esz + ((size_t)side * 2 + parity) * 8; }
  size_t timeout_off() const { return (size_t)4 * plane_dofs * esz + 32; }
  // fused path (element stage pushes the planes itself): its own arrival counters + the two work counters
  size_t fused_arrive_off(int side, int parity) const { return (size_t)4 * plane_dofs * esz + 64 + ((size_t)side * 2 + parity) * 8; }
  size_t iface_done_off() const { return (size_t)4 * plane_dofs * esz + 96; }
  size_t chunk_next_off() const { return (size_t)4 * plane_dofs * esz + 104; }
};

extern "C" {

int fol_halo_create(fol_halo** out, int dtype, int64_t plane_dofs) {
  FOL_REQUIRE(out && plane_dofs > 0 && (dtype == FOL_F32 || dtype == FOL_F64), "fol_halo_create: bad arguments");
  fol_halo* h = new fol_halo();
  h->dtype = dtype;
  h->esz = dtype == FOL_F64 ? 8 : 4;
  h->plane_dofs = plane_dofs;
  h->ctas = (unsigned)cdiv(plane_dofs, 256);
  const size_t bytes = (size_t)4 * plane_dofs * h->esz + 128;
  cudaError_t e = cudaMalloc(&h->base, bytes);
  if (e == cudaSuccess) e = cudaMemset(h->base, 0, bytes);
  if (e != cudaSuccess) {
    delete h;
    return fail(FOL_ERR_CUDA, std::string("fol_halo_create: ") + cudaGetErrorString(e));
  }
  *out = h;
  return FOL_OK;
}

void fol_halo_destroy(fol_halo* h) {
  if (!h) return;
  for (unsigned char* p : h->peer)
    if (p) cudaIpcCloseMemHandle(p);
  if (h->base) cudaFree(h->base);
  delete h;
}

/* 64-byte CUDA IPC handle of this rank's buffers, to be sent to both neighbours */
int fol_halo_export(fol_halo* h, void* handle64) {
  FOL_REQUIRE(h && handle64, "fol_halo_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  FOL_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), h->base));
  return FOL_OK;
}

/* side 0: handle of the lower neighbour (rank - 1), side 1: of the upper neighbour (rank + 1) */
int fol_halo_connect(fol_halo* h, int side, const void* handle64) {
  FOL_REQUIRE(h && handle64 && (side == 0 || side == 1), "fol_halo_connect: bad arguments");
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle64, sizeof(hd));
  void* p = nullptr;
  FOL_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
  h->peer[side] = static_cast<unsigned char*>(p);
  return FOL_OK;
}

/* Fused gather + push of one interface plane (side 0 = this rank's lower plane -> lower neighbour's from-upper
 * buffer, side 1 = upper plane -> upper neighbour's from-lower buffer).  n0 / count: node range of the plane;
 * without a connected neighbour on that side it is a plain residual gather. */
int fol_halo_gather_push(fol_stream_t s, fol_halo* h, int side, int64_t step, int64_t n0, int64_t count, int d,
                         const int32_t* adj_ptr, const int32_t* adj, const void* re_elem, void* residual) {
  FOL_REQUIRE(h && adj_ptr && adj && re_elem && residual && (side == 0 || side == 1), "fol_halo_gather_push: bad arguments");
  FOL_REQUIRE(count * d == h->plane_dofs, "fol_halo_gather_push: plane size differs from the halo object");
  const int parity = (int)(step & 1);
  unsigned char* peer = h->peer[side];
  // the neighbour receives in its buffer of the OPPOSITE side (my lower plane is its upper plane)
  void* peer_recv = peer ? peer + h->recv_off(1 - side, parity) : nullptr;
  auto* peer_arrive = peer ? reinterpret_cast<unsigned long long*>(peer + h->arrive_off(1 - side, parity)) : nullptr;
  if (h->dtype == FOL_F64)
    halo_gather_push_kernel<double><<<h->ctas, 256, 0, (cudaStream_t)s>>>(n0, count, d, adj_ptr, adj, (const double*)re_elem,
                                                                         (double*)residual, (double*)peer_recv, peer_arrive);
  else
    halo_gather_push_kernel<float><<<h->ctas, 256, 0, (cudaStream_t)s>>>(n0, count, d, adj_ptr, adj, (const float*)re_elem,
                                                                        (float*)residual, (float*)peer_recv, peer_arrive);
  return check_launch("halo_gather_push_kernel");
}

/* Adds what the neighbour on `side` pushed for `step` to the plane (waits for its arrival on the device). */
int fol_halo_add(fol_stream_t s, fol_halo* h, int side, int64_t step, int64_t n0, int64_t count, int d, void* residual) {
  FOL_REQUIRE(h && residual && (side == 0 || side == 1), "fol_halo_add: bad arguments");
  FOL_REQUIRE(count * d == h->plane_dofs, "fol_halo_add: plane size differs from the halo object");
  if (!h->peer[side]) return FOL_OK;
  const int parity = (int)(step & 1);
  const unsigned long long target = (unsigned long long)(step / 2 + 1) * h->ctas;
  const void* recv = h->base + h->recv_off(side, parity);
  const auto* arrive = reinterpret_cast<const unsigned long long*>(h->base + h->arrive_off(side, parity));
  auto* timeouts = reinterpret_cast<unsigned long long*>(h->base + h->timeout_off());
  if (h->dtype == FOL_F64)
    halo_add_kernel<double><<<h->ctas, 256, 0, (cudaStream_t)s>>>(n0, count, d, (const double*)recv, arrive, target,
                                                                 timeouts, (double*)residual);
  else
    halo_add_kernel<float><<<h->ctas, 256, 0, (cudaStream_t)s>>>(n0, count, d, (const float*)recv, arrive, target,
                                                                timeouts, (float*)residual);
  return check_launch("halo_add_kernel");
}

/* Element stage of a slab with its two interface element layers FIRST and the plane gather + NVLink push riding inside
 * the same launch (csrc/assemble_hex_common.cuh).  Tuned Hex8 float64 kernels only (FOL_MECHANICAL / FOL_J2PLASTICITY,
 * num_gp = 2); FOL_ERR_UNSUPPORTED otherwise -- callers then use the layered path (fol_halo_gather_push / fol_halo_add).
 * layer_elems: elements per z-layer (the bottom layer is [0, layer_elems), the top one the last layer_elems). */
int fol_assemble_elements_halo(fol_stream_t s, int dtype, int physics, int element, int num_gp, int64_t ne, int64_t nn,
                               const void* xyz, const int32_t* conn, const void* ctrl, const void* u, const uint8_t* dir,
                               const double* params, void* ke, void* re, const void* state_in, void* state_out,
                               fol_halo* h, int64_t step, int64_t layer_elems, int64_t plane_nodes,
                               const int32_t* adj_ptr, const int32_t* adj, void* residual) {
  FOL_REQUIRE(h && xyz && conn && ctrl && u && dir && ke && re && adj_ptr && adj && residual, "fol_assemble_elements_halo: null pointer");
  if (dtype != FOL_F64 || element != HEX || num_gp != 2 || (physics != FOL_MECHANICAL && physics != FOL_J2PLASTICITY))
    return fail(FOL_ERR_UNSUPPORTED, "fol_assemble_elements_halo: tuned Hex8 float64 kernels only");
  FOL_REQUIRE(h->dtype == FOL_F64 && plane_nodes * 3 == h->plane_dofs, "fol_assemble_elements_halo: plane size differs from the halo object");
  FOL_REQUIRE(layer_elems > 0 && ne % layer_elems == 0 && nn >= 2 * plane_nodes, "fol_assemble_elements_halo: bad layer / plane sizes");
  if (physics == FOL_J2PLASTICITY) FOL_REQUIRE(state_in && state_out, "fol_assemble_elements_halo: J2 needs state_in / state_out");
  AsmArgs<double> a;
  a.xyz = (const double*)xyz; a.conn = conn; a.ctrl = (const double*)ctrl; a.u = (const double*)u; a.dir = dir;
  a.ke = (double*)ke; a.re = (double*)re; a.state_in = (const double*)state_in; a.state_out = (double*)state_out;
  a.ne = ne; a.transpose = 0; a.p = make_params<double>(params);
  const long long ntiles = cdiv(ne, hexk::kTile);
  hexk::HaloFuse hf;
  hf.tiles_lo = cdiv(layer_elems, hexk::kTile);
  hf.tiles_hi = ntiles - (ne - layer_elems) / hexk::kTile;
  if (hf.tiles_lo + hf.tiles_hi >= ntiles) {   // one or two layers: every tile is an interface tile
    hf.tiles_lo = ntiles;
    hf.tiles_hi = 0;
  }
  hf.plane_dofs = h->plane_dofs;
  hf.n0[0] = 0;
  hf.n0[1] = nn - plane_nodes;
  hf.adj_ptr = adj_ptr; hf.adj = adj; hf.re = (const double*)re; hf.R = (double*)residual;
  const int parity = (int)(step & 1);
  for (int side = 0; side < 2; ++side) {
    unsigned char* peer = h->peer[side];
    hf.peer_recv[side] = peer ? reinterpret_cast<double*>(peer + h->recv_off(1 - side, parity)) : nullptr;
    hf.peer_arrive[side] = peer ? reinterpret_cast<unsigned long long*>(peer + h->fused_arrive_off(1 - side, parity)) : nullptr;
  }
  hf.iface_done = reinterpret_cast<unsigned long long*>(h->base + h->iface_done_off());
  hf.chunk_next = reinterpret_cast<unsigned long long*>(h->base + h->chunk_next_off());
  hf.timeouts = reinterpret_cast<unsigned long long*>(h->base + h->timeout_off());
  if (physics == FOL_MECHANICAL) return assemble_hex_mech_f64((cudaStream_t)s, a, &hf);
  return assemble_hex_j2_f64((cudaStream_t)s, a, &hf);
}

/* Completes the step started by fol_assemble_elements_halo: gather of the interior nodes + add of what the neighbours
 * pushed into the two interface planes (device-side wait for their arrivals), one launch. */
int fol_residual_gather_halo(fol_stream_t s, fol_halo* h, int64_t step, int64_t nn, int64_t plane_nodes,
                             const int32_t* adj_ptr, const int32_t* adj, const void* re, void* residual) {
  FOL_REQUIRE(h && adj_ptr && adj && re && residual, "fol_residual_gather_halo: null pointer");
  FOL_REQUIRE(h->dtype == FOL_F64 && plane_nodes * 3 == h->plane_dofs && nn >= 2 * plane_nodes, "fol_residual_gather_halo: bad sizes");
  const int parity = (int)(step & 1);
  const long long per_side = cdiv(h->plane_dofs, hexk::kHaloChunk);
  const unsigned long long target = (unsigned long long)(step / 2 + 1) * (unsigned long long)per_side;
  const long long nb_int = cdiv((nn - 2 * plane_nodes) * 3, 256), nb_plane = cdiv(h->plane_dofs, 256);
  const double* recv[2];
  const unsigned long long* arrive[2];
  for (int side = 0; side < 2; ++side) {
    recv[side] = h->peer[side] ? reinterpret_cast<const double*>(h->base + h->recv_off(side, parity)) : nullptr;
    arrive[side] = reinterpret_cast<const unsigned long long*>(h->base + h->fused_arrive_off(side, parity));
  }
  gather_halo_kernel<double><<<(unsigned)(nb_int + 2 * nb_plane), 256, 0, (cudaStream_t)s>>>(
      plane_nodes, nn, 3, adj_ptr, adj, (const double*)re, (double*)residual, nb_int, nb_plane, recv[0], recv[1],
      arrive[0], arrive[1], target, reinterpret_cast<unsigned long long*>(h->base + h->timeout_off()),
      reinterpret_cast<unsigned long long*>(h->base + h->iface_done_off()),
      reinterpret_cast<unsigned long long*>(h->base + h->chunk_next_off()));
  return check_launch("gather_halo_kernel");
}

/* number of arrival waits that gave up (a neighbour never pushed): synchronising read, 0 in a healthy run */
int64_t fol_halo_timeouts(fol_halo* h) {
  if (!h) return -1;
  unsigned long long v = 0;
  if (cudaMemcpy(&v, h->base + h->timeout_off(), sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int64_t)v;
}

}  // extern "C"
