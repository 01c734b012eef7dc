// Shared pieces of the structured-grid batched loss + VJP kernels (energy_grid.cu: thermal Quad4; energy_grid_mech.cu:
// plane-stress elasticity Quad4): mbarrier / bulk-copy wrappers, explicit-rounding lane arithmetic (one sample per lane,
// or two float32 samples on the packed FP32 instructions), the row staging of the producer warp, the launch shape.
#pragma once
#include <type_traits>

#include "energy2.cuh"
#include "energy2_launch.cuh"

namespace fol {
namespace {

#ifndef FOL_GRID_RING
#define FOL_GRID_RING 8
#endif
#ifndef FOL_GRID_PREFETCH
#define FOL_GRID_PREFETCH 0
#endif
#ifndef FOL_GRID_REGS64
#define FOL_GRID_REGS64 96
#endif
#ifndef FOL_GRID_REGS32P
#define FOL_GRID_REGS32P 72
#endif
constexpr int kRing = FOL_GRID_RING;      // node rows in flight per CTA (a power of two)

template <class T>
struct alignas(2 * sizeof(T)) NodePair {
  T t, k;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "GRID_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra GRID_DONE_%=;\n\t"
      "bra GRID_WAIT_%=;\n\t"
      "GRID_DONE_%=:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <class T>
__device__ __forceinline__ T shfl_up1(T v) {
  return __shfl_up_sync(0xffffffffu, v, 1);
}

// Arithmetic with the rounding written out: every product, sum and fused multiply-add below is the instruction it
// names (no compiler contraction), so the element vectors are bit-identical in every inlined copy of grid_element --
// the gradients do not depend on the chunk height or on which code path (recomputed row or owned row) evaluated an
// element.
__device__ __forceinline__ double op_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float op_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double op_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float op_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double op_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float op_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double op_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float op_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double op_neg(double a) { return -a; }
__device__ __forceinline__ float op_neg(float a) { return -a; }
// Two float32 samples per lane: the packed instructions of sm_100 (FMUL2 / FADD2 / FFMA2: one issue slot for the two
// samples' operations, each rounded like the scalar instruction) -- the float32 kernel is bound by issue slots.
__device__ __forceinline__ float2 op_mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 op_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 op_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 op_sub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }   // a - b, one rounding
__device__ __forceinline__ float2 op_neg(float2 a) { return make_float2(-a.x, -a.y); }

// lane type: S (one sample per lane) or float2 (two float32 samples per lane)
template <class S, int NS> struct LaneT { using type = S; };
template <> struct LaneT<float, 2> { using type = float2; };
template <class V> struct ScalarOf { using type = V; };
template <> struct ScalarOf<float2> { using type = float; };
template <class V>
__device__ __forceinline__ V bc(typename ScalarOf<V>::type x) {
  if constexpr (std::is_same<V, float2>::value) return make_float2(x, x);
  else return x;
}
template <class V>
__device__ __forceinline__ V bcd(double x) { return bc<V>((typename ScalarOf<V>::type)x); }

// a x + b y
template <class T>
__device__ __forceinline__ T lin2(T a, T x, T b, T y) {
  return op_fma(a, x, op_mul(b, y));
}


// One row of `n` values starting at element `first` of an array of `total` values (base 16-byte aligned) goes into a
// ring row as ONE bulk copy from the 16-byte block that holds the first value to the last WHOLE block of the array;
// the (at most 16 / sizeof(T) - 1) values of a partial last block of the array follow by plain stores.  The lanes read
// value j of the row at dst[shift + j], shift = first % (16 / sizeof(T)).
struct RowCopy {
  long long a0;          // first value of the first block
  long long tail_beg, tail_end;   // values copied by plain stores
  uint32_t bytes;        // bulk bytes
};
template <class T>
__device__ __forceinline__ RowCopy plan_row(long long total, long long first, int n) {
  constexpr long long PER = 16 / sizeof(T);
  RowCopy r;
  r.a0 = first & ~(PER - 1);
  long long a1 = (first + n + PER - 1) & ~(PER - 1);                         // one past the last block
  const long long whole = total & ~(PER - 1);                                // one past the last whole block of the array
  r.tail_beg = r.tail_end = 0;
  if (a1 > whole) {
    r.tail_beg = whole > first ? whole : first;
    r.tail_end = first + n;
    a1 = whole;
  }
  r.bytes = a1 > r.a0 ? (uint32_t)((a1 - r.a0) * sizeof(T)) : 0u;
  return r;
}
template <class T>
__device__ __forceinline__ void copy_tail(const RowCopy& r, const T* base, T* dst) {
  for (long long j = r.tail_beg; j < r.tail_end; ++j) dst[j - r.a0] = base[j];
}
template <class T>
__device__ __forceinline__ void copy_bulk(const RowCopy& r, const T* base, T* dst, uint32_t bar) {
  if (r.bytes) bulk_g2s(smem_u32(dst), base + r.a0, r.bytes, bar);
}


struct GridShape {
  int W, npanels;
};
inline GridShape grid_shape(long long nx) {
  GridShape g;
  g.W = (int)(nx <= 256 ? cdiv(nx, 32) : 8);
  g.npanels = (int)(nx <= 256 ? 1 : cdiv(nx - 1, 32 * g.W - 1));
  return g;
}
constexpr int kMinRows = 8;

template <class T>
int grid_row_bytes(int W, long long nx) {
  const long long ncols = (32LL * W + 1 < nx + 1) ? 32LL * W + 1 : nx + 1;
  return (int)(((ncols + 16 / sizeof(T)) * sizeof(T) + 15) / 16 * 16);   // + shift + one entry read past the last column
}

}  // namespace
}  // namespace fol
