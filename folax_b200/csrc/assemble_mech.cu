// Instantiations of the generic element-stage kernel for physics MECH (all elements, orders 1-3).
#include "assemble.cuh"

namespace fol {
int assemble_mech_f64(cudaStream_t s, int element, int num_gp, const AsmArgs<double>& a) {
  return dispatch_assemble<double, MECH>(s, element, num_gp, a);
}
int assemble_mech_f32(cudaStream_t s, int element, int num_gp, const AsmArgs<float>& a) {
  return dispatch_assemble<float, MECH>(s, element, num_gp, a);
}
}  // namespace fol
