// Instantiations of the pipelined batched loss / VJP kernel (energy2.cuh) for the implicit-Euler scalar
// losses (transient thermal, Allen-Cahn).
#include "energy2_launch.cuh"

namespace fol {

template <class T, int PHYS>
int energy2_scalar_typed(cudaStream_t s, int element, int num_gp, const EnergyArgs<T>& args, int ncap, int* parts) {
#define FOL_CASE(E, O) \
  if (element == E && num_gp == O) return launch_energy2_default<T, E, O, PHYS>(s, args, ncap, parts);
  FOL_CASE(QUAD, 1) FOL_CASE(QUAD, 2) FOL_CASE(TRI, 1) FOL_CASE(TRI, 2) FOL_CASE(TET, 1) FOL_CASE(HEX, 1)
#undef FOL_CASE
  return 1;
}

template <class T>
int energy2_scalar(cudaStream_t s, int physics, int element, int num_gp, const EnergyArgs<T>& args, int ncap, int* parts) {
  if (energy2_env_int("FOL_ENERGY_V1", 0)) return 1;
  if (physics == FOL_TRANSIENT_THERMAL) return energy2_scalar_typed<T, TTHERMAL>(s, element, num_gp, args, ncap, parts);
  if (physics == FOL_ALLEN_CAHN) return energy2_scalar_typed<T, ALLENCAHN>(s, element, num_gp, args, ncap, parts);
  return 1;
}
template int energy2_scalar<double>(cudaStream_t, int, int, int, const EnergyArgs<double>&, int, int*);
template int energy2_scalar<float>(cudaStream_t, int, int, int, const EnergyArgs<float>&, int, int*);

}  // namespace fol
