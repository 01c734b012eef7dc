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

#include "common.cuh"

namespace fol {

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
};

extern "C" {

int fol_halo_create(fol_halo** out, int dtype, int64_t plane_dofs) {
  FOL_REQUIRE(out && plane_dofs > 0 && (dtype == FOL_F32 || dtype == FOL_F64), "fol_halo_create: bad arguments");
  fol_halo* h = new fol_halo();
  h->dtype = dtype;
  h->esz = dtype == FOL_F64 ? 8 : 4;
  h->plane_dofs = plane_dofs;
  h->ctas = (unsigned)cdiv(plane_dofs, 256);
  const size_t bytes = (size_t)4 * plane_dofs * h->esz + 64;
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

/* number of arrival waits that gave up (a neighbour never pushed): synchronising read, 0 in a healthy run */
int64_t fol_halo_timeouts(fol_halo* h) {
  if (!h) return -1;
  unsigned long long v = 0;
  if (cudaMemcpy(&v, h->base + h->timeout_off(), sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int64_t)v;
}

}  // extern "C"
