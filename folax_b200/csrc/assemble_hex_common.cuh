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
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }


}  // namespace hexk
}  // namespace fol
