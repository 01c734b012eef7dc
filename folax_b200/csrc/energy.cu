// Instantiations + launchers of the batched energy / VJP kernels and the loss tail.
#include "energy.cuh"

namespace fol {

template <class T, int ELEM, int ORDER, int PHYS>
int launch_energy(cudaStream_t s, const EnergyArgs<T>& args) {
  constexpr int BLOCK = 192;           // >= nodes per tile (128) and, on the fast path, elements per tile
  constexpr int ND = elem_nnode(ELEM) * phys_dpn(PHYS, ELEM);
  constexpr int S = ND <= 8 ? 4 : 2;   // samples per CTA pass (shared-memory rows)
  constexpr int KW = energy_kw(PHYS, ELEM);
  constexpr int C = phys_dpn(PHYS, ELEM) + 1;
  const size_t smem = sizeof(T) * ((size_t)S * KW * args.ecap + (size_t)2 * S * C * args.lcap);
  if (smem > 200 * 1024) return fail(FOL_ERR_INVALID, "fol_energy_and_grads: tile lists too long for shared memory");
  auto kern = energy_tile_kernel<T, ELEM, ORDER, PHYS, S, BLOCK>;
  static size_t configured[64] = {};   // largest size set so far, per device (the attribute is per device)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = -1;
  if (dev < 0 || smem > configured[dev]) {
    FOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0) configured[dev] = smem;
  }
  if (args.ntiles == 0 || args.nb == 0) return FOL_OK;
  long long y = cdiv(148LL * 16, args.ntiles);
  const long long ymax = cdiv(args.nb, S);
  y = y < 1 ? 1 : (y > ymax ? ymax : y);
  dim3 grid((unsigned)args.ntiles, (unsigned)y);
  kern<<<grid, BLOCK, smem, s>>>(args);
  return check_launch("energy_tile_kernel");
}

template <class T, int PHYS>
int dispatch_energy(cudaStream_t s, int element, int num_gp, const EnergyArgs<T>& args) {
#define FOL_CASE(E, O) \
  if (element == E && num_gp == O) return launch_energy<T, E, O, PHYS>(s, args);
  FOL_CASE(HEX, 1) FOL_CASE(HEX, 2) FOL_CASE(HEX, 3)
  FOL_CASE(QUAD, 1) FOL_CASE(QUAD, 2) FOL_CASE(QUAD, 3)
  FOL_CASE(TET, 1) FOL_CASE(TET, 2) FOL_CASE(TET, 3)
  FOL_CASE(TRI, 1) FOL_CASE(TRI, 2) FOL_CASE(TRI, 3)
#undef FOL_CASE
  return fail(FOL_ERR_UNSUPPORTED, "unsupported element / num_gp");
}

template <class T>
int energy2_thermal(cudaStream_t, int, int, const EnergyArgs<T>&, int, int*);
template <class T>
int energy2_mech(cudaStream_t, int, int, int, const EnergyArgs<T>&, int, int*);
template <class T>
int energy2_scalar(cudaStream_t, int, int, int, const EnergyArgs<T>&, int, int*);

// The pipelined kernel (energy2.cuh) runs when the plan fits it, energy_tile_kernel otherwise.
// npart = energy shares per sample written by the kernel that ran (pipelined kernel: one per warp)
template <class T>
int energy_and_grads(cudaStream_t s, int physics, int element, int num_gp, const EnergyArgs<T>& a, int ncap, T* energy) {
  if (a.ntiles == 0 || a.nb == 0) return FOL_OK;
  int parts = 1;
  const bool scalar_implicit = physics == FOL_TRANSIENT_THERMAL || physics == FOL_ALLEN_CAHN;
  int rc = (physics == FOL_THERMAL) ? energy2_thermal<T>(s, element, num_gp, a, ncap, &parts)
           : scalar_implicit        ? energy2_scalar<T>(s, physics, element, num_gp, a, ncap, &parts)
                                    : energy2_mech<T>(s, physics, element, num_gp, a, ncap, &parts);
  if (rc == 1) {
    parts = 1;
    // energy_tile_kernel gives every owned node of a tile its own thread (192 per CTA)
    if (ncap > 192) return fail(FOL_ERR_INVALID, "fol_energy_and_grads: tiles own more than 192 nodes and no pipelined kernel "
                                                 "covers this (physics, element, rule): rebuild the plan with smaller tiles");
    if (physics == FOL_MECHANICAL) rc = dispatch_energy<T, MECH>(s, element, num_gp, a);
    else if (physics == FOL_THERMAL) rc = dispatch_energy<T, THERMAL>(s, element, num_gp, a);
    else if (physics == FOL_NEOHOOKE) rc = dispatch_energy<T, NEOHOOKE>(s, element, num_gp, a);
    else if (physics == FOL_STVENANT) rc = dispatch_energy<T, STVK>(s, element, num_gp, a);
    else if (physics == FOL_TRANSIENT_THERMAL) rc = dispatch_energy<T, TTHERMAL>(s, element, num_gp, a);
    else if (physics == FOL_ALLEN_CAHN) rc = dispatch_energy<T, ALLENCAHN>(s, element, num_gp, a);
    else return fail(FOL_ERR_UNSUPPORTED, "fol_energy_and_grads: physics not supported");
  }
  const int npart = a.ntiles * parts;
  if (rc) return rc;
  energy_sum_kernel<T><<<(unsigned)cdiv(a.nb, 8), 256, 0, s>>>(a.partial, a.nb, npart, energy);
  return check_launch("energy_sum_kernel");
}
template int energy_and_grads<double>(cudaStream_t, int, int, int, const EnergyArgs<double>&, int, double*);
template int energy_and_grads<float>(cudaStream_t, int, int, int, const EnergyArgs<float>&, int, float*);

template <class T, int ELEM, bool TR, bool AUX>
int launch_geom(cudaStream_t s, int num_gp, long long ne, const T* xyz, const int32_t* conn, const T* aux, T* geom) {
#define FOL_CASE(O)                                                                                         \
  if (num_gp == O) {                                                                                        \
    if (ne == 0) return FOL_OK;                                                                             \
    dim3 grid((unsigned)cdiv(ne, 128), (unsigned)elem_ngauss(ELEM, O));                                     \
    geometry_cache_kernel<T, ELEM, O, TR, AUX><<<grid, 128, 0, s>>>(xyz, conn, ne, aux, geom);              \
    return check_launch("geometry_cache_kernel");                                                           \
  }
  FOL_CASE(1) FOL_CASE(2) FOL_CASE(3)
#undef FOL_CASE
  return fail(FOL_ERR_UNSUPPORTED, "unsupported num_gp");
}

template <class T, bool TR, bool AUX>
int geometry_cache_variant(cudaStream_t s, int element, int num_gp, long long ne, const T* xyz, const int32_t* conn,
                           const T* aux, T* geom) {
  switch (element) {
    case HEX: return launch_geom<T, HEX, TR, AUX>(s, num_gp, ne, xyz, conn, aux, geom);
    case QUAD: return launch_geom<T, QUAD, TR, AUX>(s, num_gp, ne, xyz, conn, aux, geom);
    case TET: return launch_geom<T, TET, TR, AUX>(s, num_gp, ne, xyz, conn, aux, geom);
    case TRI: return launch_geom<T, TRI, TR, AUX>(s, num_gp, ne, xyz, conn, aux, geom);
  }
  return fail(FOL_ERR_UNSUPPORTED, "unsupported element");
}

// geometry factors in the layout / gradient convention of `physics` (energy.cuh: geom_width)
template <class T>
int geometry_cache(cudaStream_t s, int physics, int element, int num_gp, long long ne, const T* xyz,
                   const int32_t* conn, const T* aux, T* geom) {
  if (physics == FOL_TRANSIENT_THERMAL) {
    if (!aux) return fail(FOL_ERR_INVALID, "fol_geometry_cache: transient thermal needs the nodal heterogeneity k0");
    return geometry_cache_variant<T, true, true>(s, element, num_gp, ne, xyz, conn, aux, geom);
  }
  if (physics == FOL_ALLEN_CAHN) return geometry_cache_variant<T, true, false>(s, element, num_gp, ne, xyz, conn, aux, geom);
  return geometry_cache_variant<T, false, false>(s, element, num_gp, ne, xyz, conn, aux, geom);
}
template int geometry_cache<double>(cudaStream_t, int, int, int, long long, const double*, const int32_t*, const double*,
                                    double*);
template int geometry_cache<float>(cudaStream_t, int, int, int, long long, const float*, const int32_t*, const float*,
                                   float*);

template <class T>
int loss_reduce(cudaStream_t s, long long nb, double exponent, const T* energy, T* out4, T* scale) {
  loss_reduce_kernel<T><<<1, 256, 0, s>>>(nb, exponent, energy, out4, scale);
  return check_launch("loss_reduce_kernel");
}
template int loss_reduce<double>(cudaStream_t, long long, double, const double*, double*, double*);
template int loss_reduce<float>(cudaStream_t, long long, double, const float*, float*, float*);

template <class T>
int scale_grads(cudaStream_t s, long long nb, long long ndof, long long nn, const T* scale, double up, const T* up_dev,
                int prescaled, const uint8_t* dir, T* gu, T* gk) {
  if (nb == 0 || (ndof == 0 && nn == 0)) return FOL_OK;
  scale_grads_kernel<T><<<148 * 8, 256, 0, s>>>(nb, ndof, nn, scale, (T)up, up_dev, prescaled, dir, gu, gk);
  return check_launch("scale_grads_kernel");
}
template int scale_grads<double>(cudaStream_t, long long, long long, long long, const double*, double, const double*, int,
                                 const uint8_t*, double*, double*);
template int scale_grads<float>(cudaStream_t, long long, long long, long long, const float*, double, const float*, int,
                                const uint8_t*, float*, float*);

}  // namespace fol
