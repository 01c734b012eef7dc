// Instantiations of the pipelined batched loss / VJP kernel (energy2.cuh) for THERMAL.
#include "energy2_launch.cuh"

namespace fol {

template <class T, int ELEM, int ORDER, int NL>
int launch_thermal2(cudaStream_t s, const EnergyArgs<T>& args, int ncap, int* parts) {
  if constexpr (ELEM == QUAD && ORDER == 2 && sizeof(T) == 8 && NL == 4) {
    // tuning variants of the north-star configuration (FOL_ENERGY_VARIANT; scripts/energy_sweep.sh only)
    static const int variant = energy2_env_int("FOL_ENERGY_VARIANT", 0);
#define FOL_VARIANT(ID, S_, B_, M_) \
  if (variant == ID) return launch_energy2<T, ELEM, ORDER, THERMAL, NL, S_, B_, M_>(s, args, ncap, parts);
    FOL_VARIANT(1, 2, 192, 2)
    FOL_VARIANT(2, 2, 160, 2)
    FOL_VARIANT(3, 3, 160, 2)
    FOL_VARIANT(4, 2, 128, 3)
#undef FOL_VARIANT
  }
  return launch_energy2_default<T, ELEM, ORDER, THERMAL, NL>(s, args, ncap, parts);
}

template <class T, int ELEM, int ORDER>
int dispatch_thermal2(cudaStream_t s, const EnergyArgs<T>& args, int ncap, int* parts) {
  const T beta = args.p.v[5], c = args.p.v[6];
  const int ci = (int)c;
  const int nl = (beta == (T)0) ? 0 : (((T)ci == c && ci >= 1 && ci <= 4) ? ci : -1);
  switch (nl) {
    case 0: return launch_thermal2<T, ELEM, ORDER, 0>(s, args, ncap, parts);
    case 1: return launch_thermal2<T, ELEM, ORDER, 1>(s, args, ncap, parts);
    case 2: return launch_thermal2<T, ELEM, ORDER, 2>(s, args, ncap, parts);
    case 3: return launch_thermal2<T, ELEM, ORDER, 3>(s, args, ncap, parts);
    case 4: return launch_thermal2<T, ELEM, ORDER, 4>(s, args, ncap, parts);
    default: return launch_thermal2<T, ELEM, ORDER, -1>(s, args, ncap, parts);
  }
}

template <class T>
int energy_qt_thermal(cudaStream_t, const EnergyArgs<T>&, int, int*);

template <class T>
int energy2_thermal(cudaStream_t s, int element, int num_gp, const EnergyArgs<T>& args, int ncap, int* parts) {
  if (energy2_env_int("FOL_ENERGY_V1", 0)) return 1;
  if (element == QUAD && num_gp == 2) {   // north-star configuration: sample-vectorised kernel (energy_qt.cuh)
    const int rc = energy_qt_thermal<T>(s, args, ncap, parts);
    if (rc != 1) return rc;
  }
#define FOL_CASE(E, O) \
  if (element == E && num_gp == O) return dispatch_thermal2<T, E, O>(s, args, ncap, parts);
  FOL_CASE(QUAD, 1) FOL_CASE(QUAD, 2) FOL_CASE(TRI, 1) FOL_CASE(TRI, 2) FOL_CASE(TET, 1) FOL_CASE(HEX, 1)
#undef FOL_CASE
  return 1;
}
template int energy2_thermal<double>(cudaStream_t, int, int, const EnergyArgs<double>&, int, int*);
template int energy2_thermal<float>(cudaStream_t, int, int, const EnergyArgs<float>&, int, int*);

}  // namespace fol
