// Shared host-side plumbing of libfolax_b200: error strings, launch accounting, dtype dispatch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

#include "../../include/folax_b200.h"

namespace fol {

extern thread_local std::string g_last_error;
extern std::atomic<long long> g_launches;

inline int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(FOL_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return FOL_OK;
}

#define FOL_CUDA(call)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      (void)cudaGetLastError(); /* clear the non-sticky error for the next call */         \
      return ::fol::fail(FOL_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); \
    }                                                                                      \
  } while (0)

#define FOL_REQUIRE(cond, msg) \
  do {                         \
    if (!(cond)) return ::fol::fail(FOL_ERR_INVALID, msg); \
  } while (0)

inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

// One-time kernel configuration (cudaFuncSetAttribute) is PER DEVICE: remembered per call site and device, so a
// process that drives several GPUs configures each of them; safe to race (the attribute call is idempotent).
struct PerDeviceOnce {
  std::atomic<unsigned long long> mask{0};   // devices 0..63; others are configured on every call
  bool need() const {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    return ((mask.load(std::memory_order_relaxed) >> d) & 1ull) == 0ull;
  }
  void done() {
    int d = 0;
    if (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 64) mask.fetch_or(1ull << d, std::memory_order_relaxed);
  }
};

// Persistent-kernel grid (SMs x resident CTAs per SM) of one kernel, configured and remembered PER DEVICE (a process
// may drive several GPUs with different SM counts); devices >= 64 are re-queried on every call.
struct PerDeviceGrid {
  std::atomic<int> grid[64] = {};
  template <class Kernel>
  cudaError_t get(Kernel kernel, int block, size_t smem, int* out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) {
      const int g = grid[dev].load(std::memory_order_relaxed);
      if (g > 0) {
        *out = g;
        return cudaSuccess;
      }
    }
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int sms = 148, per_sm = 1;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem);
    if (e != cudaSuccess) return e;
    *out = sms * (per_sm > 0 ? per_sm : 1);
    if (dev >= 0 && dev < 64) grid[dev].store(*out, std::memory_order_relaxed);
    return cudaSuccess;
  }
};

// material / loss parameters in the arithmetic type of the call (see FOL_NUM_PARAMS)
template <class T>
struct Params {
  T v[FOL_NUM_PARAMS];
};
template <class T>
inline Params<T> make_params(const double* host) {
  Params<T> p;
  for (int i = 0; i < FOL_NUM_PARAMS; ++i) p.v[i] = host ? (T)host[i] : (T)0;
  return p;
}

}  // namespace fol
