// Instantiations + launcher of the sample-vectorised batched loss / VJP kernel (energy_qt.cuh): ThermalLoss2DQuad,
// 2x2 rule, float32 and float64.
#include "energy2_launch.cuh"
#include "energy_qt.cuh"

namespace fol {

// returns 0 launched, 1 not applicable (the caller falls back to energy_tile2_kernel), <0 error
template <class T, int NL, int BLOCK, int MINB, bool AFFINE>
int launch_energy_qt(cudaStream_t s, const EnergyArgs<T>& args, int ncap, int* parts) {
  constexpr int LCAP = BLOCK + 64, S = 16 / (int)sizeof(T);
  constexpr size_t smem = 16 * ((size_t)2 * 8 * BLOCK + (size_t)2 * 2 * LCAP);
  if (args.lcap > LCAP || args.ecap > BLOCK || ncap > BLOCK) return 1;
  *parts = BLOCK / 32;
  auto kern = energy_qt_kernel<T, NL, BLOCK, MINB, LCAP, AFFINE>;
  static PerDeviceOnce configured;
  if (configured.need()) {
    FOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // MINB CTAs of this size only fit with the largest shared-memory carve-out
    FOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured.done();
  }
  // sample chunks per tile: whole rounds of the 148 * MINB resident CTAs, >= 16 passes per CTA (energy2_launch.cuh)
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
  return check_launch("energy_qt_kernel");
}

template <class T>
int energy_qt_thermal(cudaStream_t s, const EnergyArgs<T>& args, int ncap, int* parts) {
  // float32: the sample-major kernel with three CTAs per SM (energy2.cuh) measured 9 % faster than this one
  // (profiles/r2/energy_kernels_ab.md), so this kernel serves float64 unless FOL_ENERGY_QT=2 forces it
  static const int enabled = energy2_env_int("FOL_ENERGY_QT", 1);
  if (!enabled || (sizeof(T) == 4 && enabled != 2)) return 1;
  // affine meshes (all elements parallelograms; the host plan checked it) keep 5 geometry values per element in
  // registers instead of 36, which is what lets 256-thread CTAs (tiles of up to 256 elements, 16 warps / SM) fit
  const bool affine = (args.mesh_flags & FOL_MESH_AFFINE) != 0 && energy2_env_int("FOL_ENERGY_AFFINE", 1) != 0;
  const bool wide = affine && (args.ecap > 192 || ncap > 192 || energy2_env_int("FOL_ENERGY_QT_BLOCK", 0) == 256);
  const T beta = args.p.v[5], c = args.p.v[6];
  const int ci = (int)c;
  const int nl = (beta == (T)0) ? 0 : (((T)ci == c && ci >= 1 && ci <= 4) ? ci : -1);
#define FOL_QT(NLV)                                                                         \
  if (nl == NLV) {                                                                          \
    if (wide) return launch_energy_qt<T, NLV, 256, 2, true>(s, args, ncap, parts);          \
    if (affine) return launch_energy_qt<T, NLV, 192, 2, true>(s, args, ncap, parts);        \
    return launch_energy_qt<T, NLV, 192, 2, false>(s, args, ncap, parts);                   \
  }
  FOL_QT(0) FOL_QT(1) FOL_QT(2) FOL_QT(3) FOL_QT(4) FOL_QT(-1)
#undef FOL_QT
  return 1;
}
template int energy_qt_thermal<double>(cudaStream_t, const EnergyArgs<double>&, int, int*);
template int energy_qt_thermal<float>(cudaStream_t, const EnergyArgs<float>&, int, int*);

}  // namespace fol
