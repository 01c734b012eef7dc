// Launcher of the pipelined batched loss / VJP kernel (energy2.cuh), shared by the per-physics instantiation files.
#pragma once
#include <cstdlib>

#include "energy2.cuh"

namespace fol {

// block size of the pipelined kernel = energy_plan.MAX_ELEMS: 160-node tiles of 2-D meshes touch <= 192 elements
constexpr int ENERGY2_BLOCK = 192;
constexpr int ENERGY2_LCAP = ENERGY2_BLOCK + 64;

inline int energy2_env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

// samples per pass: as many as keep two CTAs (~110 KB each) resident per SM
template <class T, int ELEM, int PHYS, int BLOCK, int LCAP>
constexpr int energy2_samples() {
  constexpr int KW = energy_kw(PHYS, ELEM), C = phys_dpn(PHYS, ELEM) + 1;
  for (int s = 4; s > 1; --s)
    if (sizeof(T) * (2 * s * KW * BLOCK + 2 * s * C * LCAP) <= 110 * 1024) return s;
  return 1;
}

// returns 0 launched, 1 not applicable (the caller falls back to energy_tile_kernel), <0 error
template <class T, int ELEM, int ORDER, int PHYS, int NL, int S, int BLOCK, int MINB, int LCAP = BLOCK + 64>
int launch_energy2(cudaStream_t s, const EnergyArgs<T>& args, int ncap, int* parts) {
  constexpr int KW = energy_kw(PHYS, ELEM), C = phys_dpn(PHYS, ELEM) + 1;
  constexpr size_t smem = sizeof(T) * ((size_t)2 * S * KW * BLOCK + (size_t)2 * S * C * LCAP);
  if (smem > 220 * 1024 || args.lcap > LCAP || args.ecap > BLOCK || ncap > BLOCK) return 1;
  *parts = BLOCK / 32;
  auto kern = energy_tile2_kernel<T, ELEM, ORDER, PHYS, NL, S, BLOCK, MINB, LCAP>;
  static PerDeviceOnce configured;
  if (configured.need()) {
    FOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured.done();
  }
  // sample chunks per tile: the grid should fill whole rounds of the 148 * MINB resident CTAs (a fractional
  // last round idles SMs) while every CTA keeps >= 16 passes to amortise its tile set-up
  static const int forced = energy2_env_int("FOL_ENERGY_YCHUNKS", 0);
  const long long slots = 148LL * MINB, passes = cdiv(args.nb, S);
  const long long ymax = passes / 16 < 1 ? 1 : passes / 16;
  long long y = 1;
  double best = 0.0;
  for (long long c = 1; c <= ymax && c <= 64; ++c) {
    const long long items = c * args.ntiles, rounds = cdiv(items, slots);
    const double eff = (double)items / (double)(rounds * slots) * (rounds >= 2 ? 1.0 : 0.9);
    if (eff > best + 1e-9) { best = eff; y = c; }
  }
  if (forced > 0) y = forced;
  dim3 grid((unsigned)args.ntiles, (unsigned)y);
  kern<<<grid, BLOCK, smem, s>>>(args);
  return check_launch("energy_tile2_kernel");
}

template <class T, int ELEM, int ORDER, int PHYS, int NL = -1>
int launch_energy2_default(cudaStream_t s, const EnergyArgs<T>& args, int ncap, int* parts) {
  constexpr int S = energy2_samples<T, ELEM, PHYS, ENERGY2_BLOCK, ENERGY2_LCAP>();
  // float32 halves the registers of the geometry factors and the shared memory: three CTAs per SM
  constexpr int MINB = (sizeof(T) == 4 && (PHYS == THERMAL || PHYS == MECH)) ? 3 : 2;
  return launch_energy2<T, ELEM, ORDER, PHYS, NL, S, ENERGY2_BLOCK, MINB, ENERGY2_LCAP>(s, args, ncap, parts);
}

}  // namespace fol
