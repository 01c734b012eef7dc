// Microbenchmark of the element-stage kernel's STORE PATH alone: persistent warps, each owning one shared-memory staging
// slot per `slots`, writing a 9.66 GB stream as `chunk`-byte cp.async.bulk copies in the visiting order of
// assemble_hex_mech_f64_kernel (chunk index = global warp, stride = all warps), optionally with a busy-wait of `spin`
// clocks per chunk standing in for the element arithmetic.  Separates what the HBM write path sustains for this access
// pattern from what the kernel's own instruction stream costs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_path_bench store_path_bench.cu
//   ./store_path_bench            (prints one JSON line per variant)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(saddr), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }

// mode 0: bulk copies from a per-warp slot; mode 1: 16-byte st.global straight from registers (coalesced 512 B per
// instruction); mode 2: bulk copies, the slot refilled by the warp (STS.128) before every copy like the real kernel;
// mode 3: slot refilled, then read back and written with coalesced 16-byte st.global (no bulk copy, no fence);
// mode 4: mode 2 WITHOUT the async-proxy fence (wrong data, diagnostic: what the fence costs);
// mode 5: mode 2 with the warps' first chunk delayed by warp * spin / 16 (no convoy at the copy engine)
template <int SLOTS>
__global__ void store_kernel(double* out, long long nchunks, int chunk, int mode, int spin, int warps) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* slot0 = smem + (size_t)warp * SLOTS * chunk;
  const long long nwarps = (long long)gridDim.x * warps;
  int it = 0;
  if (mode == 5) {
    const long long t0 = clock64();
    while (clock64() - t0 < (long long)warp * spin / warps) {}
  }
  for (long long c = (long long)blockIdx.x * warps + warp; c < nchunks; c += nwarps, ++it) {
    if (spin) {
      const long long t0 = clock64();
      while (clock64() - t0 < spin) {}
    }
    unsigned char* dst = reinterpret_cast<unsigned char*>(out) + c * chunk;
    if (mode == 1) {
      const double2 v = make_double2((double)c, (double)lane);
      for (int o = lane * 16; o < chunk; o += 512) *reinterpret_cast<double2*>(dst + o) = v;
    } else if (mode == 3) {
      const double2 v = make_double2((double)c, (double)lane);
      __syncwarp();
      for (int o = lane * 16; o < chunk; o += 512) *reinterpret_cast<double2*>(slot0 + o) = v;
      __syncwarp();
      for (int o = lane * 16; o < chunk; o += 512)
        *reinterpret_cast<double2*>(dst + o) = *reinterpret_cast<const double2*>(slot0 + o);
    } else {
      unsigned char* slot = slot0 + (SLOTS > 1 ? (it % SLOTS) * chunk : 0);
      if (lane == 0) bulk_wait_read<SLOTS - 1>();
      __syncwarp();
      if (mode >= 2) {
        const double2 v = make_double2((double)c, (double)lane);
        for (int o = lane * 16; o < chunk; o += 512) *reinterpret_cast<double2*>(slot + o) = v;
      }
      if (mode != 4) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) bulk_store(dst, slot, chunk);
    }
  }
  if (lane == 0) bulk_wait_read<0>();
}

template <int SLOTS>
static void run(double* out, long long total, int chunk, int mode, int spin, int warps, int ctas_per_sm, int sms) {
  const long long nchunks = total / chunk;
  const size_t smem = (size_t)warps * SLOTS * chunk;
  CK(cudaFuncSetAttribute(store_kernel<SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, store_kernel<SLOTS>, warps * 32, smem));
  if (occ < ctas_per_sm) { printf("{\"skip\": \"occupancy %d < %d\", \"chunk\": %d, \"warps\": %d, \"slots\": %d}\n", occ, ctas_per_sm, chunk, warps, SLOTS); return; }
  const int grid = sms * ctas_per_sm;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) store_kernel<SLOTS><<<grid, warps * 32, smem>>>(out, nchunks, chunk, mode, spin, warps);
  CK(cudaDeviceSynchronize());
  const int reps = 10;
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) store_kernel<SLOTS><<<grid, warps * 32, smem>>>(out, nchunks, chunk, mode, spin, warps);
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  ms /= reps;
  printf("{\"mode\": %d, \"chunk\": %d, \"slots\": %d, \"warps_per_sm\": %d, \"spin\": %d, \"ms\": %.4f, \"gbs\": %.1f}\n", mode, chunk, SLOTS,
         warps * ctas_per_sm, spin, ms, (double)nchunks * chunk / (ms * 1e-3) / 1e9);
  fflush(stdout);
}

int main() {
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const long long total = 2097152LL * 4608;   // the Ke stream of the 128^3 mesh
  double* out;
  CK(cudaMalloc(&out, total));
  // store path alone
  run<1>(out, total, 4608, 1, 0, 8, 2, sms);    // plain 16-byte stores
  run<1>(out, total, 4608, 0, 0, 8, 2, sms);    // bulk, 16 warps / SM, one slot
  run<1>(out, total, 4608, 2, 0, 8, 2, sms);    // bulk + slot refill
  run<1>(out, total, 4608, 3, 0, 8, 2, sms);    // slot refill + LDS / STG
  // with a stand-in for the element arithmetic: 16 warps / SM, 2.1 M chunks -> 886 chunks per warp; a spin of S clocks
  // per chunk alone takes 886 S / f: S = 3000 -> 1.35 ms at 1.97 GHz
  for (int spin : {2400, 3000, 3400}) {
    run<1>(out, total, 4608, 2, spin, 8, 2, sms);   // the kernel's protocol
    run<1>(out, total, 4608, 0, spin, 8, 2, sms);   // no refill
    run<1>(out, total, 4608, 4, spin, 8, 2, sms);   // no fence
    run<1>(out, total, 4608, 5, spin, 8, 2, sms);   // staggered warps
    run<1>(out, total, 4608, 3, spin, 8, 2, sms);   // LDS / STG instead of the bulk copy
    run<1>(out, total, 4608, 1, spin, 8, 2, sms);   // plain stores from registers
  }
  run<2>(out, total, 4608, 2, 2250, 6, 2, sms);     // 12 warps, two slots
  run<1>(out, total, 4608, 2, 2250, 6, 2, sms);     // 12 warps, one slot
  run<1>(out, total, 4608, 3, 2250, 6, 2, sms);     // 12 warps, LDS / STG
  CK(cudaFree(out));
  return 0;
}
