// Building blocks shared by the tuned Hex8 float64 element-stage kernels (assemble_hex.cu: small-strain elasticity,
// assemble_hex_j2.cu: J2 elastoplasticity): FP64 tensor-path MMA, bulk (TMA-engine) stores, 8-byte cp.async gathers.
#pragma once
#include <cuda_runtime.h>

namespace fol {
namespace hexk {

constexpr int kTile = 4;             // elements per warp iteration

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, unsigned bytes) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(saddr), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void cp_async8(void* sdst, const void* gsrc) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(saddr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }


// ---- interface-first tile order + in-kernel halo push (slab-partitioned meshes) ------------------------
// The persistent element-stage kernel of a slab visits the tiles of its two interface element layers FIRST.  When the
// last of them has stored its element residuals (device-scope counter), the plane work -- the fixed-order residual
// gather of the two interface node planes, fused with the NVLink peer store into the neighbour's receive buffer and a
// system-scope arrival count per chunk -- is handed out in chunks of kHaloChunk dofs to whichever warp finishes a tile
// next.  The exchange therefore rides inside the one element-stage launch, hidden behind the interior tiles: no
// separate interface launches, no plane kernels, no side stream (round 1 needed 7 launches per step; this needs 2).
struct HaloFuse {
  long long tiles_lo, tiles_hi;         // virtual tiles [0, tiles_lo) = bottom layer, [tiles_lo, tiles_lo + tiles_hi) = top layer
  long long plane_dofs;                 // dofs of one interface node plane
  long long n0[2];                      // first node of the lower / upper plane
  const int32_t* adj_ptr;               // node -> (element, local node) adjacency of the slab (ascending, fixed order)
  const int32_t* adj;
  const double* re;                     // element residuals written by this very launch
  double* R;                            // local residual
  double* peer_recv[2];                 // neighbour's receive buffer of this step's parity (nullptr: no neighbour)
  unsigned long long* peer_arrive[2];   // neighbour's arrival counter (system scope)
  unsigned long long* iface_done;       // interface tiles finished (zeroed by the kernel that completes the step)
  unsigned long long* chunk_next;       // next plane chunk to hand out (zeroed likewise)
  unsigned long long* timeouts;         // spin waits that gave up
};
constexpr int kHaloChunk = 64;

__device__ __forceinline__ long long halo_real_tile(const HaloFuse& hf, long long v, long long ntiles) {
  if (v < hf.tiles_lo) return v;
  if (v < hf.tiles_lo + hf.tiles_hi) return ntiles - hf.tiles_hi + (v - hf.tiles_lo);
  return v - hf.tiles_hi;               // interior tiles follow in their natural order
}

// after the element residuals of an interface tile were stored by this warp
__device__ __forceinline__ void halo_tile_done(const HaloFuse& hf, int lane) {
  __threadfence();
  __syncwarp();
  if (lane == 0) {
    __threadfence();                    // cumulative over the warp's stores ordered by the barrier above
    atomicAdd(hf.iface_done, 1ULL);
  }
}

// One attempt to take plane work; returns true when nothing is left to hand out.
__device__ __forceinline__ bool halo_try_push(const HaloFuse& hf, int lane) {
  unsigned long long seen = 0;
  if (lane == 0) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(hf.iface_done) : "memory");
  seen = __shfl_sync(0xffffffffu, seen, 0);
  if (seen < (unsigned long long)(hf.tiles_lo + hf.tiles_hi)) return false;
  const long long per_side = (hf.plane_dofs + kHaloChunk - 1) / kHaloChunk;
  unsigned long long c = 0;
  if (lane == 0) c = atomicAdd(hf.chunk_next, 1ULL);
  c = __shfl_sync(0xffffffffu, c, 0);
  if (c >= (unsigned long long)(2 * per_side)) return true;
  const int side = c >= (unsigned long long)per_side ? 1 : 0;
  const long long t0 = ((long long)c - side * per_side) * kHaloChunk;
  // (selects, not indexing: a run-time index into the kernel-parameter struct would copy it to local memory)
  const long long n_first = side ? hf.n0[1] : hf.n0[0];
  double* const peer_recv = side ? hf.peer_recv[1] : hf.peer_recv[0];
  unsigned long long* const peer_arrive = side ? hf.peer_arrive[1] : hf.peer_arrive[0];
#pragma unroll
  for (int j = 0; j < kHaloChunk / 32; ++j) {
    const long long t = t0 + j * 32 + lane;
    if (t < hf.plane_dofs) {
      const long long n = n_first + t / 3;
      const int k = (int)(t % 3);
      double acc = 0.0;
      const int lo = __ldg(hf.adj_ptr + n), hi = __ldg(hf.adj_ptr + n + 1);
      for (int i = lo; i < hi; ++i) acc += __ldcg(hf.re + (long long)__ldg(hf.adj + i) * 3 + k);   // L2: written by this launch
      hf.R[n * 3 + k] = acc;
      if (peer_recv) peer_recv[t] = acc;      // NVLink peer store
    }
  }
  if (peer_arrive) {
    __threadfence_system();
    __syncwarp();
    if (lane == 0) {
      __threadfence_system();
      atomicAdd_system(peer_arrive, 1ULL);
    }
  }
  return false;
}

// after the tile loop: whatever plane work is left is done by the warps as they finish (bounded spin: ~2 s)
__device__ __forceinline__ void halo_drain(const HaloFuse& hf, int lane) {
  const long long t0 = clock64();
  while (!halo_try_push(hf, lane)) {
    if (clock64() - t0 > 4000000000LL) {
      if (lane == 0) atomicAdd(hf.timeouts, 1ULL);
      break;
    }
  }
}

}  // namespace hexk
}  // namespace fol
